"""Diagnose high-energy outliers of the bench workload: per-node potentials of the worst replica, and the reference engine
(oracle/_ref, test infrastructure) on the same coordinates."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import bench
from parity import ue
import ref_engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng = ue.BatchEngine(bench.CONFIG, B)
n_res = eng.n_atom // 3
eng.set_pos(bench.workload_positions(B, 0, n_res))
eng.md_init_seeds(np.full(B, bench.TEMPERATURE, dtype='f4'), bench.SEED + np.arange(B), dt=bench.DT)
eng.md_run(rounds)
en, dv = eng.evaluate(want_deriv=True)
order = np.argsort(en)[::-1]
print('energies: mean %.1f median %.1f; top five' % (en.mean(), np.median(en)), en[order[:5]], 'replicas', order[:5])
print('replicas above 1000:', int((en > 1000).sum()), 'above 500:', int((en > 500).sum()))
pos = eng.get_pos()
w = int(order[0])
ref = ref_engine.RefEngine(bench.CONFIG, eng.n_atom)
e_ref = ref.energy(pos[w]); d_ref = ref.deriv(pos[w])
print('worst replica %d: GPU %.3f  reference %.3f   max |force| GPU %.1f reference %.1f  max force diff %.3g' % (
    w, en[w], e_ref, np.abs(dv[w]).max(), np.abs(d_ref).max(), np.abs(dv[w] - d_ref).max()))
for name, is_pot in ref.node_names():
    if is_pot:
        print('   %-36s GPU %12.3f   reference %12.3f' % (name, eng.node_potential(name)[w], ref.node_potential(name)))
bonds = np.linalg.norm(pos[w, 1:] - pos[w, :-1], axis=1)
print('bond lengths of the worst replica: min %.2f max %.2f' % (bonds.min(), bonds.max()))
m = int(np.argsort(en)[B // 2])
bonds = np.linalg.norm(pos[m, 1:] - pos[m, :-1], axis=1)
print('bond lengths of the median replica: min %.2f max %.2f' % (bonds.min(), bonds.max()))
