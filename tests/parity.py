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
    6: os.path.join(ROOT, 'configs', 'config6_restraints_20res.up'),
    9: os.path.join(ROOT, 'configs', 'config9_coupling_20res.up'),     # config 1 + uniform_transform / linear_coupling nodes
    8: os.path.join(ROOT, 'configs', 'config8_radial_20res.up'),       # config 1 + radial / hbond_sc_radial pair nodes
    10: os.path.join(ROOT, 'configs', 'config10_two_chains_40res.up'),  # two chains: chain_break group, bonded terms split, a jump move per chain
    7: os.path.join(ROOT, 'configs', 'config7_concat_20res.up'),       # config 1 + slice/constant/concat -> tether springs   # config 1 + restraint / plumbing nodes (tools/make_restraint_config.py)
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


def igraph_specs(cfg):
    """(node, group-1 node, group-2 node, index1, id1, index2, id2, cutoff, exclusion, symmetric) of every pair list,
    read from the configuration exactly as the reference constructors do"""
    pot = h5lite.load(cfg)['input/potential']
    f32 = np.float32
    specs = {}
    def two(name, kind, cutoff):
        g = pot[name]
        a = [_s(x) for x in g.attrs['arguments']]
        specs[name] = (a[0], a[1] if len(a) > 1 else a[0], g['index1'].data, g['id1'].data, g['index2'].data, g['id2'].data,
                       f32(cutoff), kind, False)
    g = pot['rotamer/pair_interaction']
    sc = _s(pot['rotamer'].attrs['arguments'][0])
    specs['rotamer'] = (sc, sc, g['index'].data, g['id'].data, g['index'].data, g['id'].data, f32((16 - 2 - 1e-6) / 2.0), 'rotamer', True)
    two('hbond_coverage', 'seq2', (12 - 2 - 1e-6) / 2.0)
    two('hbond_coverage_hydrophobe', 'seq2', (12 - 2 - 1e-6) / 2.0)
    prm = pot['environment_coverage/interaction_param'].data.astype('f4')
    two('environment_coverage', 'seq2', (prm[..., 0] + f32(1.) / prm[..., 1]).max())
    two('protein_hbond', 'none', np.sqrt(f32(3.5 * 3.5)))
    bb = pot['backbone_pairs']
    ref = bb['ref_pos'].data.astype('f4'); na = bb['n_atom'].data
    dev = max(np.sqrt((ref[i, a] ** 2).sum(dtype='f4')) for i in range(len(na)) for a in range(na[i]))
    cut = f32(2) * f32(dev) + np.sqrt(f32(3.) * f32(3.) + f32(0.1) * f32(3.))
    specs['backbone_pairs'] = ('affine_alignment', 'affine_alignment', bb['id'].data, bb['id'].data, bb['id'].data, bb['id'].data, f32(cut), 'seq1', True)
    return specs


def _s(x):
    return x.decode() if isinstance(x, bytes) else str(x)


def restated_pairlist(spec, out1, out2):
    """the reference predicate + emission order (oracle/restate.py) applied to given node outputs"""
    from oracle import restate
    n1, n2, i1, id1, i2, id2, cutoff, kind, sym = spec
    return restate.pairlist(out1[i1][:, :3], id1, out2[i2][:, :3], id2, cutoff, kind, sym)


def boundary_gap(spec, out1, out2, edges):
    """largest |dist - cutoff| over the given edges (edges that differ between two engines must sit on the cutoff)"""
    if not len(edges):
        return 0.0
    n1, n2, i1, id1, i2, id2, cutoff, kind, sym = spec
    e = np.array(sorted(edges))
    d = np.linalg.norm(out1[i1][e[:, 0], :3].astype('f8') - out2[i2][e[:, 1], :3].astype('f8'), axis=1)
    return float(np.abs(d - float(cutoff)).max())


def compare_engines(cfg, pos, verbose=True):
    """returns dict of per-node max abs differences (output, sens), potentials, pair-list equality, for each replica"""
    n_rep, n_atom = pos.shape[0], pos.shape[1]
    be = ue.BatchEngine(cfg, n_rep)
    en, deriv = be.evaluate(pos)
    ref = ref_engine.RefEngine(cfg, n_atom, 'pinned')
    report = []
    specs = igraph_specs(cfg)
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
            sp = specs[name]
            o1, o2 = be.get_output(sp[0], r), be.get_output(sp[1], r)
            restated = restated_pairlist(sp, o1, o2)
            rep['pairlists'][name] = dict(n_gpu=len(pg), n_ref=len(pr), identical=same, only_gpu=len(sg - sr),
                                          only_ref=len(sr - sg),
                                          # bit-exact given IDENTICAL coordinates: reference predicate on the GPU's own inputs
                                          exact_given_same_coords=restated.shape == pg.shape and bool((restated == pg).all()),
                                          boundary_gap=boundary_gap(sp, ref.get_output(sp[0]), ref.get_output(sp[1]), sg ^ sr))
        mg = be.get_value_by_name('rotamer', 'bead_marginal', r)
        mr = ref.rotamer_bead_marginals()
        rep['marginal_maxabs'] = float(np.abs(mg - mr).max())
        rep['bp_stats'] = (be.get_value_by_name('rotamer', 'solve_stats', r).tolist(), ref.rotamer_solve_stats())
        # tooling accessors of the rotamer node (reference rotamer.cpp:675-773)
        n_node = int(ref.get_value_by_name('rotamer', 'n_node', 1)[0])
        n_prob = len(h5lite.load(cfg)['input/potential/rotamer'].attrs['arguments']) - 1
        rep['accessors'] = {}
        for nm, n in (('rotamer_1body_energy', n_node * n_prob), ('node_energy', n_node * 6), ('rotamer_free_energy', n_node)):
            a, b = be.get_value_by_name('rotamer', nm, r), ref.get_value_by_name('rotamer', nm, n)
            rep['accessors'][nm] = (a.shape == b.shape, float(np.abs(a - b).max()) if a.shape == b.shape else np.inf,
                                    float(np.abs(b[b < 1e4]).max()))
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
        print('   pairlist %-32s gpu %d ref %d identical %s only_gpu %d only_ref %d exact_given_same_coords %s gap %.1e' % (
            name, d['n_gpu'], d['n_ref'], d['identical'], d['only_gpu'], d['only_ref'], d['exact_given_same_coords'], d['boundary_gap']))


if __name__ == '__main__':
    cid = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    nrep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = CONFIGS[cid]
    compare_engines(cfg, test_positions(cfg, nrep))
