#!/bin/bash
# One gpurun call's worth of evidence (developer tool): GPU tests, the default bench line, kernel micro-benchmarks, the ncu
# launch list of the bench command and --set full captures of the fused kernel and the C4 crowd kernel.  Everything lands
# in gpurun_out/; the summaries worth keeping are copied to profiles/ by hand.
#   gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > $out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -3 $out/${tag}_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $out/${tag}_smoke.log 2>&1; tail -1 $out/${tag}_smoke.log
timeout 900 python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench exit $?"
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu --no-legs > $out/${tag}_bench_n1_steps20.json 2>> $out/${tag}_bench_n1.err; echo "bench 20+5 exit $?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err; echo "ref arm exit $?"
timeout 600 python tools/kbench.py --images 10000000 --reps 10 > $out/${tag}_kbench_10M.txt 2>&1; echo "kbench exit $?"
# ncu: launch list of the bench step (no CPU legs), then full captures
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-legs > $out/${tag}_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fused_tma -c 1 --launch-skip 1 -f -o $out/${tag}_fused \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-legs > $out/${tag}_ncu_fused.log 2>&1; echo "ncu fused exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:iou_crowd -c 2 --launch-skip 3 -f -o $out/${tag}_crowd \
    python tools/crowd_bench.py 1000000 > $out/${tag}_ncu_crowd.log 2>&1; echo "ncu crowd exit $?"
ls -la $out
