#!/bin/bash
# Overlap experiments of the URL chain with the fused kernel at N=1 (developer tool): one stream, SM carve-out, co-residency.
out=gpurun_out; mkdir -p $out
B="python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-legs"
pick='import json,sys; d=json.loads(sys.stdin.read()); s=d["streams"]; print(sys.argv[1], "ms_per_step %.3f fused %.3f url %.3f value %.3f G" % (d["ms_per_step"], s["fused_ms"], s["url_chain_ms"], d["value"]/1e9))'
$B | python -c "$pick" "A one stream (616)" | tee $out/overlap_exp.txt
$B --url-sms 8 | python -c "$pick" "B url-sms 8 (616)" | tee -a $out/overlap_exp.txt
for cap in 610 560; do
  DYD_NVCC_FLAGS="-DDYD_TILE_CAP_V=$cap" python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  $B | python -c "$pick" "A one stream ($cap)" | tee -a $out/overlap_exp.txt
  $B --coresident | python -c "$pick" "C coresident ($cap), partitioned dedup" | tee -a $out/overlap_exp.txt
  DYD_DEDUP_PARTITION=0 $B --coresident | python -c "$pick" "C coresident ($cap), global-table dedup" | tee -a $out/overlap_exp.txt
  DYD_DEDUP_PARTITION=0 $B | python -c "$pick" "A one stream ($cap), global-table dedup" | tee -a $out/overlap_exp.txt
done
python -m deal_yolo_daya_b200.build --force > /dev/null 2>&1
