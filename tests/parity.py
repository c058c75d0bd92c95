"""Shared helpers for the parity tests: evaluate the CUDA engine (through the C ABI) and the oracle (the unmodified
reference engine, oracle/_ref) on the same coordinates and compare node by node."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from upside_md_b200 import upside_engine as ue  # noqa: E402
from upside_md_b200 import h5lite  # noqa: E402
from oracle import ref_engine  # noqa: E402

CONFIGS = {
    1: os.path.join(ROOT, 'configs', 'config1_20res.up'),
    2: os.path.join(ROOT, 'configs', 'config2_76res.up'),
    3: os.path.join(ROOT, 'configs', 'config3_100res.up'),
    4: os.path.join(ROOT, 'configs', 'config4_150res.up'),
    5: os.path.join(ROOT, 'configs', 'config5_300res.up'),
}
PAIRLIST_NODES = ['rotamer', 'hbond_coverage', 'hbond_coverage_hydrophobe', 'environment_coverage', 'protein_hbond',
                  'backbone_pairs']


def initial_pos(cfg):
    return np.array(h5lite.load(cfg)['input/pos'].data[:, :, 0], dtype='f4')


def test_positions(cfg, n_rep, seed=0, relax_rounds=40):
    """n_rep coordinate sets: the configuration's own start, then snapshots of a short reference MD run from it
    (relaxed, clash-free structures), each slightly different"""
    p0 = initial_pos(cfg)
    out = [p0]
    if n_rep > 1:
        r = ref_engine.md_run(cfg, np.repeat(p0[None], n_rep - 1, 0), 0.8, relax_rounds, seed=100 + seed, n_thread=4,
                              flavour='pinned')
        out.extend(list(r['pos']))
    return np.array(out, dtype='f4')


def align_quat_sign(a, b):
    """affine outputs carry a unit quaternion whose overall sign is arbitrary: flip b's to match a's"""
    b = b.copy()
    s = np.sign((a[:, 3:7] * b[:, 3:7]).sum(axis=1))
    s[s == 0] = 1
    b[:, 3:7] *= s[:, None]
    return b


def compare_engines(cfg, pos, verbose=True):
    """returns dict of per-node max abs differences (output, sens), potentials, pair-list equality, for each replica"""
    n_rep, n_atom = pos.shape[0], pos.shape[1]
    be = ue.BatchEngine(cfg, n_rep)
    en, deriv = be.evaluate(pos)
    ref = ref_engine.RefEngine(cfg, n_atom, 'pinned')
    report = []
    for r in range(n_rep):
        e_ref = ref.energy(pos[r])
        d_ref = ref.deriv(pos[r])
        rep = dict(replica=r, energy=(float(en[r]), float(e_ref)), nodes={}, pairlists={})
        rep['deriv_maxabs'] = float(np.abs(deriv[r] - d_ref).max())
        rep['deriv_scale'] = float(np.abs(d_ref).max())
        for name, is_pot in ref.node_names():
            if is_pot:
                a, b = float(be.node_potential(name)[r]), ref.node_potential(name)
                rep['nodes'][name] = dict(pot=(a, b))
            else:
                o_g, o_r = be.get_output(name, r), ref.get_output(name)
                s_g, s_r = be.get_sens(name, r), ref.get_sens(name)
                if name.startswith('affine_alignment'):
                    o_g = align_quat_sign(o_r, o_g)
                    s_g, s_r = s_g[:, :6], s_r[:, :6]      # 7th sens component is unused padding
                rep['nodes'][name] = dict(out=float(np.abs(o_g - o_r).max()), out_scale=float(np.abs(o_r).max()),
                                          sens=float(np.abs(s_g - s_r).max()), sens_scale=float(np.abs(s_r).max()))
        for name in PAIRLIST_NODES:
            if name not in dict(ref.node_names()):
                continue
            pg, pr = be.pairlist(name, r), ref.pairlist(name)
            same = pg.shape == pr.shape and bool((pg == pr).all())
            sg, sr = set(map(tuple, pg)), set(map(tuple, pr))
            rep['pairlists'][name] = dict(n_gpu=len(pg), n_ref=len(pr), identical=same, only_gpu=len(sg - sr),
                                          only_ref=len(sr - sg))
        mg = be.get_value_by_name('rotamer', 'bead_marginal', r)
        mr = ref.rotamer_bead_marginals()
        rep['marginal_maxabs'] = float(np.abs(mg - mr).max())
        rep['bp_stats'] = (be.get_value_by_name('rotamer', 'solve_stats', r).tolist(), ref.rotamer_solve_stats())
        report.append(rep)
        if verbose:
            print_report(rep)
    be.close()
    ref.close()
    return report


def print_report(rep):
    e_g, e_r = rep['energy']
    print('replica %d: E gpu %.5f ref %.5f rel %.2e | dV max|diff| %.2e (scale %.1f) | marginals %.2e' % (
        rep['replica'], e_g, e_r, abs(e_g - e_r) / max(1e-6, abs(e_r)), rep['deriv_maxabs'], rep['deriv_scale'],
        rep['marginal_maxabs']))
    print('   bp', rep['bp_stats'])
    for name, d in rep['nodes'].items():
        if 'pot' in d:
            a, b = d['pot']
            print('   %-40s pot gpu % .5f ref % .5f diff %.2e' % (name, a, b, abs(a - b)))
        else:
            print('   %-40s out %.2e (/%.1f)  sens %.2e (/%.1f)' % (name, d['out'], d['out_scale'], d['sens'], d['sens_scale']))
    for name, d in rep['pairlists'].items():
        print('   pairlist %-32s gpu %d ref %d identical %s only_gpu %d only_ref %d' % (
            name, d['n_gpu'], d['n_ref'], d['identical'], d['only_gpu'], d['only_ref']))


if __name__ == '__main__':
    cid = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = CONFIGS[cid]
    compare_engines(cfg, test_positions(cfg, nrep))
