timeout -s KILL 400 python -m pytest tests/ -q -m gpu -x 2>&1 | tail -2
timeout -s KILL 300 python scripts/microbench.py --m 5,8,16 --tag _v10 2>&1 | cut -c1-100
