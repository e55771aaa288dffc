#!/bin/bash
# Build + kbench a few compile-time variants of the fused kernel on the GPU box (developer tool).
for v in "-DDYD_NW=20 -DDYD_TILE_CAP_V=600 -DDYD_SEG_IMAGES=32 -DDYD_TILE_MAX_IMAGES=6" \
         "-DDYD_NW=20 -DDYD_TILE_CAP_V=600 -DDYD_SEG_IMAGES=32 -DDYD_TILE_MAX_IMAGES=5" \
         "-DDYD_NW=21 -DDYD_TILE_CAP_V=568 -DDYD_SEG_IMAGES=32 -DDYD_TILE_MAX_IMAGES=6" \
         "-DDYD_NW=19 -DDYD_TILE_CAP_V=640 -DDYD_SEG_IMAGES=32 -DDYD_TILE_MAX_IMAGES=6" \
         "-DDYD_NW=20 -DDYD_TILE_CAP_V=600 -DDYD_SEG_IMAGES=32 -DDYD_TILE_MAX_IMAGES=6"; do
  echo "=== $v"
  DYD_NVCC_FLAGS="$v" python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 200 python tools/kbench.py --images 4000000 --reps 8 --which fused 2>&1 | sed -n 2,2p
done
