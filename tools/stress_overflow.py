"""Robustness: run the bench workload with a deliberately small neighbour capacity; overflow must surface as an error
return, never as a device fault (run under compute-sanitizer on the GPU box)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from upside_md_b200 import upside_engine as ue
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 20
pos = bench.workload_positions(B, 0)
eng = ue.BatchEngine(bench.CONFIG, B)
eng.set_pos(pos)
eng.md_init(0.8, seed=42)
n_err = 0
for i in range(rounds):
    try:
        eng.md_run(1)
    except RuntimeError as e:
        n_err += 1
print('rounds', rounds, 'errors reported', n_err, 'finite', np.isfinite(eng.get_pos()).all() if n_err == 0 else '-')
