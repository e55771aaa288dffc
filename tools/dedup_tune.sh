#!/bin/bash
# K4 with different table sizes (load factors) and pass counts (developer tool)
for x in 8 6 5; do
  echo "=== DYD_TABLE_X4=$x"
  DYD_TABLE_X4=$x python tools/dedup_tune.py ${1:-10000000} 2>&1 | head -3
done
