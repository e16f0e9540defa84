#!/bin/bash
cd /root/repo
timeout 200 python scripts/umma_stage_trace.py 2048 4096 4096 0 63 39 2>&1 | tail -150
