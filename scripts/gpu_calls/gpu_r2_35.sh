# round 2, call 35: where the 0.6 us between the dependency wait and the activations go (cycles: wait -> loads issued -> data)
mkdir -p gpurun_out
for shp in "4096 4096" "14336 4096"; do timeout -s KILL 100 python scripts/ring_probe.py $shp 1 48 > gpurun_out/r2_35_probe_$(echo $shp | tr ' ' 'x').txt 2>&1; grep -A2 "^== decode kernel (16" gpurun_out/r2_35_probe_$(echo $shp | tr ' ' 'x').txt | cut -c1-190; grep "SM clock" gpurun_out/r2_35_probe_$(echo $shp | tr ' ' 'x').txt | head -3; done
