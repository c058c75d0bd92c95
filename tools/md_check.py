"""Trajectory parity + crude throughput: GPU engine vs the reference MD loop (oracle/_ref) with the same seeds."""
import sys, time, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from parity import CONFIGS, test_positions, ue, ref_engine

cid = int(sys.argv[1]); n_rep = int(sys.argv[2]); n_round = int(sys.argv[3]); big = int(sys.argv[4]) if len(sys.argv) > 4 else 0
cfg = CONFIGS[cid]
pos = test_positions(cfg, n_rep + 1)[1:]
be = ue.BatchEngine(cfg, n_rep)
be.set_pos(pos)
be.md_init(0.8, seed=42)
m0 = be.get_mom()
be.md_run(n_round)
pg, mg = be.get_pos(), be.get_mom()
r = ref_engine.md_run(cfg, pos, 0.8, n_round, seed=42, n_thread=min(8, n_rep), flavour='pinned')
print('after %d rounds: max|dpos| %.3e  max|dmom| %.3e  (pos scale %.1f)' % (n_round, np.abs(pg - r['pos']).max(), np.abs(mg - r['mom']).max(), np.abs(pg).max()))
for k in (1, 3, 10, 30):
    if k <= n_round:
        be.set_pos(pos); be.md_init(0.8, seed=42); be.md_run(k)
        rr = ref_engine.md_run(cfg, pos, 0.8, k, seed=42, n_thread=min(8, n_rep), flavour='pinned')
        print('  %3d rounds: max|dpos| %.3e' % (k, np.abs(be.get_pos() - rr['pos']).max()))
print('kinetic/1.5T', be.kinetic_energy() / (1.5 * 0.8) * 1.0)
be.close()
if big:
    B = big
    be = ue.BatchEngine(cfg, B)
    allpos = np.repeat(pos[:1], B, 0)
    be.set_pos(allpos); be.md_init(0.8, seed=7)
    be.md_run(5)
    print('launches per eval', be.launches_per_eval())
    for nr in (10, 20):
        t = time.time(); be.md_run(nr); dt = time.time() - t
        print('B=%d: %d rounds in %.3fs -> %.1f us per force-eval batch, %.3g replica-timesteps/s' % (B, nr, dt, dt * 1e6 / (3 * nr), B * 3 * nr / dt))
    en = be.evaluate(want_deriv=False)
    print('energies', en[:4], 'finite', np.isfinite(en).all())
