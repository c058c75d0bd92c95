"""config 5 at the bench state: successive timed md_run(20) calls, with and without evaluate() calls in between"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from parity import ue
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
CFG = bench.CONFIG4 if len(sys.argv) > 2 and sys.argv[2] == "4" else bench.CONFIG5
eng = ue.BatchEngine(CFG, B)
n_res = eng.n_atom // 3
eng.set_pos(bench.workload_positions(B, 0, n_res))
eng.md_init_seeds(np.full(B, bench.TEMPERATURE, dtype='f4'), bench.SEED + np.arange(B), dt=bench.DT)
eng.md_run(150)
def timed(n=20):
    eng.sync(); t0 = time.perf_counter(); eng.md_run(n); eng.sync(); return (time.perf_counter() - t0) * 1e6 / (3 * n)
print('md_run x4:', ' '.join('%.0f' % timed() for k in range(4)))
eng.evaluate(want_deriv=True)
print('after one evaluate():', ' '.join('%.0f' % timed() for k in range(3)))
for label, ms in eng.profile_eval(): pass
print('after profile_eval:', ' '.join('%.0f' % timed() for k in range(3)))
print('slow bp list / bad solves:', eng.get_value_by_name('rotamer', 'build_stats', 0)[:4])
