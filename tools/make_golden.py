#!/usr/bin/env python
"""Generate tests/golden/*.npz from the oracle (oracle/_ref = unmodified reference engine, pinned flavour).
Run in the build container (needs oracle/_ref); the fixtures are committed and travel to the GPU box.
For each configuration: relaxed coordinates (a few reference MD rounds from /input/pos), total and per-node
energies, dV/dx, pair lists of every interaction graph, per-bead BP marginals, and a short reference trajectory."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import parity
from oracle import ref_engine

def main():
    for cid, n_rep in ((1, 3), (2, 2), (3, 2), (5, 1)):
        cfg = parity.CONFIGS[cid]
        pos = parity.test_positions(cfg, n_rep + 1)[1:]
        n_atom = pos.shape[1]
        ref = ref_engine.RefEngine(cfg, n_atom, 'pinned')
        out = dict(pos=pos)
        en, dv, marg = [], [], []
        for r in range(n_rep):
            en.append(ref.energy(pos[r])); dv.append(ref.deriv(pos[r]))
            marg.append(ref.rotamer_bead_marginals())
            for name, is_pot in ref.node_names():
                if is_pot:
                    out.setdefault('pot_' + name, []).append(ref.node_potential(name))
            for name in parity.PAIRLIST_NODES:
                if name in dict(ref.node_names()):
                    out['pairs_%s_%d' % (name, r)] = ref.pairlist(name)
            if r == 0:
                out['beads_0'] = ref.get_output('placement_fixed_point_vector_only')
        out['energy'] = np.array(en, 'f4'); out['deriv'] = np.array(dv, 'f4'); out['marginal'] = np.array(marg, 'f4')
        for k in list(out):
            if k.startswith('pot_'): out[k] = np.array(out[k], 'f4')
        ref.close()
        if cid in (1, 3):
            tr = ref_engine.md_run(cfg, pos, 0.8, 10, seed=42, n_thread=2, flavour='pinned')
            out['traj_pos_10'] = tr['pos']; out['traj_mom_10'] = tr['mom']
            tr = ref_engine.md_run(cfg, pos, 0.8, 1, seed=42, n_thread=2, flavour='pinned')
            out['traj_pos_1'] = tr['pos']; out['traj_mom_1'] = tr['mom']
        path = os.path.join(ROOT, 'tests', 'golden', 'config%d.npz' % cid)
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path), 'E', out['energy'])
    # RNG known answers from the vendored Random123 header through the reference's RandomGenerator
    L = ref_engine.load('pinned')
    import ctypes as ct
    cases = [(42, 0, 0, 0), (42, 0, 7, 0), (43, 0, 299, 5), (42, 0, 1, 4294967299), (42, 1, 0, 10), (7, 0, 123, 99999)]
    bits, norm = [], []
    for s, st, a, t in cases:
        b = (ct.c_uint32 * 4)(); L.ref_rng_bits(s, st, a, t, b); bits.append(list(b))
        n = np.zeros(3, 'f4'); L.ref_rng_normal3(s, st, a, t, n.ctypes.data_as(ct.POINTER(ct.c_float))); norm.append(n)
    np.savez(os.path.join(ROOT, 'tests', 'golden', 'rng.npz'), cases=np.array(cases, dtype=np.uint64), bits=np.array(bits, dtype=np.uint32), normal3=np.array(norm))
    print('rng', bits[0], norm[0])

if __name__ == '__main__':
    main()
