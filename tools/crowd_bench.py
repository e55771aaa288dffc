"""C4 micro-benchmark (developer tool): the block-per-image K2 kernel on 1 M dense-crowd images, worst case and natural mix.
python tools/crowd_bench.py [images]"""
import sys, json, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench_legs
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
r = bench_legs.c4_leg(torch.device("cuda", 0), 0, 1, 6542.1, n)
print(json.dumps(r, indent=1))
