"""config 5 (300 residues, membrane) at the bench state: per-group times and pair-list rebuild statistics for B replicas."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from parity import ue
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 150
eng = ue.BatchEngine(bench.CONFIG5, B)
n_res = eng.n_atom // 3
eng.set_pos(bench.workload_positions(B, 0, n_res))
eng.md_init_seeds(np.full(B, bench.TEMPERATURE, dtype='f4'), bench.SEED + np.arange(B), dt=bench.DT)
eng.md_run(warm)
eng.sync(); t0 = time.perf_counter(); eng.md_run(20); eng.sync(); dt = time.perf_counter() - t0
print('B=%d after %d rounds: %.1f us per force evaluation' % (B, warm, dt * 1e6 / 60))
acc = {}
for rep in range(5):
    for label, ms in eng.profile_eval():
        acc[label] = acc.get(label, 0.0) + ms / 5
print('sum of serial groups %.0f us' % (1e3 * sum(acc.values())))
print('  '.join('%s=%.0f' % (k, 1e3 * v) for k, v in sorted(acc.items(), key=lambda kv: -kv[1])[:16]))
try:
    print('build_stats', eng.get_value_by_name('rotamer', 'build_stats', 0)[:8])
except Exception as e:
    print('no build stats:', e)
print('potential mean', float(np.mean(eng.evaluate(want_deriv=False))))
eng.sync(); t0 = time.perf_counter()
for k in range(20): eng.evaluate(want_deriv=True)
eng.sync(); print('fixed state: %.1f us per evaluate() call' % ((time.perf_counter() - t0) * 1e6 / 20))
eng.sync(); t0 = time.perf_counter(); eng.md_run(20); eng.sync(); print('md_run again: %.1f us per force evaluation' % ((time.perf_counter() - t0) * 1e6 / 60))
