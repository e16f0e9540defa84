echo "labels: start | post-alloc | first-w | x-wait-ret | deq-loop-end | drained | end"
B200AWQ_FLAT_MIN_M=1 timeout -s KILL 300 python scripts/trace_chain.py 4096 4096 1 48 2>&1 | grep -A8 "graph pdl" | head -40
B200AWQ_FLAT_MIN_M=1 timeout -s KILL 300 python scripts/trace_chain.py 4096 4096 16 48 2>&1 | grep -A8 "graph pdl=1" | head -12
