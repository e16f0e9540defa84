#!/bin/bash
# which kernel/knob is best at M = 3, 4 (batch sweep bs=4)
cd /root/repo
echo "== kc 4096"; B200AWQ_STREAM_KC=4096 timeout 200 python scripts/microbench.py --m 3,4,7 --tag _m4e 2>&1 | cut -c1-75
echo "== kc 4096 rbs 2"; B200AWQ_STREAM_KC=4096 B200AWQ_STREAM_RBS=2 timeout 200 python scripts/microbench.py --m 3,4,7 --tag _m4f 2>&1 | cut -c1-75
echo "== flat from 3, m 7"; B200AWQ_FLAT_MIN_M=3 timeout 200 python scripts/microbench.py --m 7 --tag _m4g 2>&1 | cut -c1-75
