#!/bin/bash
mkdir -p gpurun_out
for cfg in "16 16" "16 8" "32 8" "32 16" "16 4" "32 32"; do
set -- $cfg
echo "== chunks $1 dev $2"; MN_STREAM_CHUNKS=$1 MN_STREAM_DEV_CHUNKS=$2 timeout 600 python tools/profile_e2e.py 2>&1 | grep -E "e2e calls" | head -12
done
