#!/bin/bash
# Multi-GPU evidence (developer tool): the two-rank exchange test and the bench at N GPUs, as the driver launches it.
#   gpurun --gpus N --timeout 1500 -- 'bash tools/gpu_scale.sh N [tag]'
N=${1:-2}; tag=${2:-r2}
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/${tag}_topo_n$N.txt 2>&1
lscpu | head -30 > $out/${tag}_lscpu.txt; grep -m1 flags /proc/cpuinfo | tr ' ' '\n' | grep -E "avx|bmi|pclmul|vpclmul" | tr '\n' ' ' >> $out/${tag}_lscpu.txt
df -h /tmp /dev/shm >> $out/${tag}_lscpu.txt; mount | grep -E " / | /tmp " >> $out/${tag}_lscpu.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q -m gpu > $out/${tag}_pytest_multi_n$N.log 2>&1; tail -2 $out/${tag}_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 \
    > $out/${tag}_bench_n$N.json 2> $out/${tag}_bench_n$N.err; echo "bench N=$N exit $?"
tail -c 300 $out/${tag}_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29614 tools/pcie_all.py > $out/${tag}_pcie_all_n$N.txt 2>&1
cat $out/${tag}_pcie_all_n$N.txt | tail -4
