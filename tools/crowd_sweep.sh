#!/bin/bash
# Build + time the binned crowd kernel with different residency targets (developer tool).
for mb in 8 7 6 5; do
  echo "=== DYD_CROWD_MIN_BLOCKS=$mb"
  DYD_NVCC_FLAGS="-DDYD_CROWD_MIN_BLOCKS=$mb" python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 200 python tools/crowd_bench.py 1000000 2>&1 | grep -E "\"ms\"|equal_to_oracle" | tr '\n' ' '; echo
done
python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1
timeout 300 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_scale.py -x -q -m gpu -k "iou or crowd or threshold or c4 or tile or fused" 2>&1 | tail -3
