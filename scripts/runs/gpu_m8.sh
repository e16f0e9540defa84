#!/bin/bash
# crossovers around M = 5..8 (stream vs flat) and 32..64 (flat vs umma split-k)
cd /root/repo
echo "== default"; timeout 200 python scripts/microbench.py --m 3,4,5,8,32,64 --tag _m8a 2>&1 | cut -c1-75
echo "== stream kc 4096 up to 8"; B200AWQ_FLAT_MIN_M=9 B200AWQ_STREAM_KC=4096 timeout 200 python scripts/microbench.py --m 5,8 --tag _m8b 2>&1 | cut -c1-75
echo "== flat up to 64"; B200AWQ_FLAT_MAX_M=64 timeout 200 python scripts/microbench.py --m 32,64 --tag _m8c 2>&1 | cut -c1-75
