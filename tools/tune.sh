#!/bin/bash
# Build + kbench a few compile-time variants of the fused kernel on the GPU box (developer tool).
for v in "-DDYD_K1_PERM=1" "-DDYD_K1_PERM=0" \
         "-DDYD_K1_PERM=0 -DDYD_NW=21 -DDYD_TILE_CAP_V=576" \
         "-DDYD_K1_PERM=1 -DDYD_NW=21 -DDYD_TILE_CAP_V=576" \
         "-DDYD_K1_PERM=0 -DDYD_NW=22 -DDYD_TILE_CAP_V=544" \
         "-DDYD_K1_PERM=0 -DDYD_NW=20 -DDYD_TILE_CAP_V=608"; do
  echo "=== $v"
  DYD_NVCC_FLAGS="$v" python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 200 python tools/kbench.py --images 10000000 --reps 8 --which fused 2>&1 | sed -n 2,3p
done
