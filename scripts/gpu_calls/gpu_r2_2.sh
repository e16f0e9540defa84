# round 2, call 2: in-kernel timeline of the ring kernel in a launch chain (trace build)
mkdir -p gpurun_out
for shp in "4096 4096" "14336 4096" "4096 14336"; do
  timeout -s KILL 200 python scripts/trace_chain.py $shp 1 48 2>&1 | grep -A14 "graph pdl=1\|plain launches pdl=1" > gpurun_out/r2_2_trace_$(echo $shp | tr ' ' 'x').txt
done
cat gpurun_out/r2_2_trace_*.txt | cut -c1-150
