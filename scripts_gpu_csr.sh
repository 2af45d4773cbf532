#!/bin/bash
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -6
python tools/bench_k1.py 200000 2000
python tools/bench_k1.py 100000 3000
