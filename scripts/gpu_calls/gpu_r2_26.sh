# round 2, call 26: hold the producer's first lap back (B200AWQ_DECODE_DELAY_NS) so that the predecessor's activation loads do not queue behind it
mkdir -p gpurun_out
for d in 0 1000 1500 2200; do
  echo "delay $d ns"
  B200AWQ_DECODE_DELAY_NS=$d timeout -s KILL 200 python scripts/microbench.py --m 1 --tag _r26_d$d 2>&1 | cut -c1-140
done
B200AWQ_DECODE_DELAY_NS=1500 timeout -s KILL 100 python scripts/ring_probe.py 4096 4096 1 48 > gpurun_out/r2_26_probe_4096x4096.txt 2>&1; grep -A3 "^== decode kernel (16" gpurun_out/r2_26_probe_4096x4096.txt | cut -c1-170
