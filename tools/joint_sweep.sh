#!/bin/bash
# Build + kbench a few compile-time variants of the joint K4+K5 kernel (developer tool).
variants=("-DDYD_PT_THREADS=128 -DDYD_JOINT_FILL=192" "-DDYD_PT_THREADS=128 -DDYD_JOINT_FILL=384" "-DDYD_PT_THREADS=64 -DDYD_JOINT_FILL=384" "-DDYD_PT_THREADS=64 -DDYD_JOINT_FILL=192")
for v in "${variants[@]}"; do
  echo "=== $v"
  DYD_NVCC_FLAGS="$v" python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 200 python tools/kbench.py --images 10000000 --reps 10 --which dedup,antijoin 2>&1 | grep -E "joint|K4 dedup" | grep -v "^\["
done
python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1
