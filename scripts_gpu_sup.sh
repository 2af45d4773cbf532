#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_supervised.py 100 2>&1 | tail -2
timeout 900 python tools/bench_supervised.py 200 --fast 2>&1 | tail -2
