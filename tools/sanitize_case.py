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
# round 2: the spill twin of k_rot_build (tiny shared-memory capacities), the device ladder, host-buffer evaluation, checkpoint
os.environ['UPSIDE_B200_BUILD_CAPS'] = '1.5,0.5'
be = ue.BatchEngine(CONFIGS[1], 4)
del os.environ['UPSIDE_B200_BUILD_CAPS']
en2, _ = be.evaluate(np.repeat(pos[:1], 4, 0))
print('spill-path energies', en2)
T = np.array([0.7, 0.8, 0.9, 1.0], dtype='f4')
be.md_init(T, seed=3)
lad = ue.Ladder(be, ['0-1,2-3', '1-2'], T, seed=3)
for k in range(3):
    be.md_run(2, sync=False)
    lad.attempt(2 * (k + 1))
print('ladder', lad.state()[0])
blob = be.checkpoint(); be.restore(blob); be.md_run(1)
lad.close(); be.close()
u = ue.Upside(CONFIGS[1])
print('host-buffer evaluation', u.energy(pos[0]), float(np.abs(u.deriv(pos[0])).max()))
