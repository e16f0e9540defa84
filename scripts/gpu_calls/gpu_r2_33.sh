# round 2, call 33: skinny-batch dispatch sweep (which kernel serves 8 / 16 / 32 / 64 tokens best today)
mkdir -p gpurun_out
run() { echo "== $1"; shift; env "$@" timeout -s KILL 200 python scripts/microbench.py --m 8,16,32,64 --tag _r33 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('   %5dx%-5d M=%-3d %7.2f us  frac %.3f' % (d['N'],d['K'],d['M'],d['us'],d['hbm_frac']))"; }
run "default" A=1
run "stream kernel up to 16 tokens (B200AWQ_SKINNY=stream)" B200AWQ_SKINNY=stream
run "flat kernel up to 64 tokens" B200AWQ_FLAT_MAX_M=64
run "tile kernel from 8 tokens (flat off)" B200AWQ_FLAT_MAX_M=4 B200AWQ_STREAM_MAX_M=4
run "tile kernel, split 8" B200AWQ_FLAT_MAX_M=4 B200AWQ_STREAM_MAX_M=4 B200AWQ_UMMA_SPLIT=8
run "tile kernel, token tile 64" B200AWQ_FLAT_MAX_M=4 B200AWQ_STREAM_MAX_M=4 B200AWQ_UMMA_TN=64
