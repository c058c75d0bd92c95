"""Development probe: equilibrated config-3 batch, ms per MD round and the live per-kernel-group split (ub_profile_eval).
usage: quick_probe.py [tag] [equil_rounds]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from upside_md_b200 import upside_engine as ue
tag = sys.argv[1] if len(sys.argv) > 1 else ''
equil = int(sys.argv[2]) if len(sys.argv) > 2 else bench.EQUIL_ROUNDS
B = bench.N_REPLICA
eng = ue.BatchEngine(bench.CONFIG, B)
eng.set_pos(bench.workload_positions(B, 0)); eng.md_init(bench.TEMPERATURE, seed=bench.SEED, dt=bench.DT)
def bstats(label):
    try:
        st = np.array([eng.get_value_by_name('rotamer', 'build_stats', r) for r in range(0, B, max(1, B // 64))])
        print('   build_stats %s: survivors mean %.0f max %.0f | active pairs mean %.0f max %.0f' % (label, st[:, 0].mean(), st[:, 0].max(), st[:, 1].mean(), st[:, 1].max()))
    except Exception as e:
        print('   build_stats unavailable', e)
eng.evaluate(want_deriv=False); bstats('start')
eng.md_run(equil)
eng.sync(); t0 = time.perf_counter(); eng.md_run(20); eng.sync(); dt = time.perf_counter() - t0
acc = {}
for rep in range(3):
    for label, ms in eng.profile_eval():
        acc[label] = acc.get(label, 0.0) + ms / 3
pot = eng.evaluate(want_deriv=False)
bstats('equilibrated')
print('%s: %.3f ms/round  %.0f replica-timesteps/s  <V>=%.2f' % (tag, 1e3 * dt / 20, B * 3 * 20 / dt, float(np.mean(pot))))
print('   ' + '  '.join('%s=%.0f' % (k, 1e3 * v) for k, v in acc.items()))
