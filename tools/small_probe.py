"""Per-kernel-group times of ONE replica (config 2, 76 residues) and the latency of a whole evaluation through the graph."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import parity
from parity import ue
cfg = parity.CONFIGS[int(sys.argv[1]) if len(sys.argv) > 1 else 2]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
be = ue.BatchEngine(cfg, B)
p0 = parity.initial_pos(cfg)
be.set_pos(np.repeat(p0[None], B, 0)); be.md_init(0.8, seed=1); be.md_run(30)
be.sync(); t0 = time.perf_counter(); be.md_run(100); be.sync(); dt = time.perf_counter() - t0
print('B=%d: %.1f us per force evaluation inside md_run (graph replay)' % (B, dt * 1e6 / 300))
acc = {}
for rep in range(5):
    for label, ms in be.profile_eval():
        acc[label] = acc.get(label, 0.0) + ms / 5
print('sum of serial kernel groups %.0f us' % (1e3 * sum(acc.values())))
print('  '.join('%s=%.0f' % (k, 1e3 * v) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:14]))
