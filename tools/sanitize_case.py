"""Small end-to-end case for compute-sanitizer: 3 replicas of config 1, evaluation with derivatives, accessors, 4 MD rounds."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from parity import CONFIGS, ue
g = np.load(os.path.join(ROOT, 'tests', 'golden', 'config1.npz'))
pos = g['pos'][:3]
be = ue.BatchEngine(CONFIGS[1], len(pos))
en, dv = be.evaluate(pos)
print('energies', en)
print('param deriv', float(np.abs(be.get_param_deriv('rotamer', 0)).sum()), float(np.abs(be.get_param_deriv('hbond_coverage', -1)).sum()))
print('free energy', be.get_value_by_name('rotamer', 'rotamer_free_energy', 1).sum())
be.md_init(0.8, seed=1)
be.md_run(4)
print('finite', bool(np.isfinite(be.get_pos()).all()))
be.close()
