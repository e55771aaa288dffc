#!/bin/bash
# Build + kbench a few compile-time variants of the partitioned dedup (developer tool).
for v in "-DDYD_PT_THREADS=256 -DDYD_PT_FILL=128" "-DDYD_PT_THREADS=128 -DDYD_PT_FILL=128" "-DDYD_PT_THREADS=64 -DDYD_PT_FILL=128" \
         "-DDYD_PT_THREADS=128 -DDYD_PT_FILL=192" "-DDYD_PT_THREADS=256 -DDYD_PT_FILL=192" "-DDYD_PT_THREADS=128 -DDYD_PT_FILL=96"; do
  echo "=== $v"
  DYD_NVCC_FLAGS="$v" python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 200 python tools/kbench.py --images 10000000 --reps 10 --which dedup,antijoin 2>&1 | grep -E "K4|K5"
done
python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1
