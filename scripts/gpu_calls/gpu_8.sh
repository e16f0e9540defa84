mkdir -p gpurun_out
nvidia-smi -L | wc -l
run() { # nproc port args...
  n=$1; port=$2; shift 2
  timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n "$@" 2> gpurun_out/err_$port.log | tail -1
}
run 8 29601 --workload tp70b --steps 100 --warmup 5 > gpurun_out/bench_tp70b_n8.json; cut -c1-1500 gpurun_out/bench_tp70b_n8.json; grep -i "error\|Traceback" -A3 gpurun_out/err_29601.log | head -10
run 4 29602 --workload tp70b --steps 100 --warmup 5 > gpurun_out/bench_tp70b_n4.json; cut -c1-200 gpurun_out/bench_tp70b_n4.json; python -c "import json;d=json.load(open('gpurun_out/bench_tp70b_n4.json'));print(d['comm'])"
run 8 29603 --steps 200 --warmup 5 > gpurun_out/bench_llama3_n8.json; cut -c1-330 gpurun_out/bench_llama3_n8.json
