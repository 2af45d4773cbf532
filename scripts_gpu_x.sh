#!/bin/bash
mkdir -p gpurun_out
for ch in 8 1 4 1 8; do
echo "== chunks $ch"; MN_STREAM_CHUNKS=$ch timeout 600 python tools/profile_e2e.py 2>&1 | grep -E "e2e calls|us_default|label_layout|cpu'|MetaNeighborUS.py" | head -12
done
