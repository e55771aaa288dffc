#!/bin/bash
# Build + kbench a few compile-time variants of the fused kernel on the GPU box (developer tool).
# Stage size per warp = 16 * DYD_TILE_CAP_V + 1744 bytes; DYD_NW stages must fit 227 KB.
for v in "-DDYD_NW=20 -DDYD_TILE_CAP_V=600" \
         "-DDYD_NW=20 -DDYD_TILE_CAP_V=616" \
         "-DDYD_NW=21 -DDYD_TILE_CAP_V=576" \
         "-DDYD_NW=19 -DDYD_TILE_CAP_V=648" \
         "-DDYD_NW=20 -DDYD_TILE_CAP_V=600 -DDYD_K1_PIPE=2"; do
  echo "=== $v"
  DYD_NVCC_FLAGS="$v" python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 200 python tools/kbench.py --images 10000000 --reps 8 --which fused 2>&1 | sed -n 2,3p
done
python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1
