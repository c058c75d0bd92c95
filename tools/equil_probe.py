"""How the cost of one MD round drifts as the batch relaxes from random_initial_config: ms per round and live edge
counts in chunks of rounds (decides bench.py's EQUIL_ROUNDS).  usage: equil_probe.py [chunk] [n_chunks]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from upside_md_b200 import upside_engine as ue
chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n_chunks = int(sys.argv[2]) if len(sys.argv) > 2 else 20
B = bench.N_REPLICA
eng = ue.BatchEngine(bench.CONFIG, B)
eng.set_pos(bench.workload_positions(B, 0)); eng.md_init(bench.TEMPERATURE, seed=bench.SEED, dt=bench.DT)
done = 0
for c in range(n_chunks):
    eng.sync(); t0 = time.perf_counter()
    eng.md_run(chunk)
    eng.sync(); dt = time.perf_counter() - t0
    done += chunk
    pot = eng.evaluate(want_deriv=False)
    cnt = {}
    for r in range(0, B, B // 8):
        for k in ('rotamer', 'hbond_coverage', 'environment_coverage', 'protein_hbond'):
            cnt[k] = cnt.get(k, 0.0) + len(eng.pairlist(k, r)) / 8
        st = eng.get_value_by_name('rotamer', 'solve_stats', r)
        cnt['sweeps'] = cnt.get('sweeps', 0.0) + (st[0] + 1) / 8
        cnt['bp_pairs'] = cnt.get('bp_pairs', 0.0) + st[1] / 8
    print('rounds %4d  %.2f ms/round  <V> %.1f  %s' % (done, 1e3 * dt / chunk, float(np.mean(pot)),
          ' '.join('%s=%.0f' % kv for kv in cnt.items())), flush=True)
