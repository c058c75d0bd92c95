"""GPU parity tests (pytest -m gpu): the CUDA engine, called through the C ABI, against
  (1) the oracle = the unmodified reference engine (oracle/_ref, pinned flavour) on the same coordinates,
  (2) the committed golden fixtures (tests/golden, generated from the oracle by tools/make_golden.py),
  (3) size-independent invariants at BASELINE.json's full batch size.
Tolerances are BASELINE.json's: energy 1e-4 relative, forces 1e-3*max(1,|F|_inf), BP marginals 1e-3, pair lists
bit-exact (identical edges in identical order) given identical coordinates, trajectories 1e-3 A after 30 timesteps."""
import os

import numpy as np
import pytest

import parity
from parity import ue, h5lite
from oracle import ref_engine, restate

pytestmark = pytest.mark.gpu
GOLD = os.path.join(parity.ROOT, 'tests', 'golden')
E_RTOL, F_RTOL, MARG_ATOL = 1e-4, 1e-3, 1e-3


def _check_report(rep):
    for r in rep:
        e_g, e_r = r['energy']
        assert abs(e_g - e_r) <= E_RTOL * max(1.0, abs(e_r)), r['energy']
        assert r['deriv_maxabs'] <= F_RTOL * max(1.0, r['deriv_scale']), (r['deriv_maxabs'], r['deriv_scale'])
        assert r['marginal_maxabs'] <= MARG_ATOL
        # same stopping rule as the oracle (rotamer.cpp:1037-1049): equal sweep counts, at most one chunk apart on a borderline deviation
        assert abs(int(r['bp_stats'][0][0]) - r['bp_stats'][1]['n_iter']) <= 2, r['bp_stats']
        for name, d in r['pairlists'].items():
            # bit-exact (edges and order) given identical inputs to the pair-list stage: the reference predicate applied
            # to the GPU's own node outputs.  End to end the two engines' bead coordinates differ by ~1e-5 A of upstream
            # rounding, so a pair sitting exactly on the cutoff may flip; such edges must be on the cutoff.
            assert d['exact_given_same_coords'], (name, d)
            assert d['identical'] or d['boundary_gap'] < 2e-4, (name, d)
        for name, (same_shape, diff, scale) in r.get('accessors', {}).items():
            assert same_shape and diff <= 2e-3 * max(1.0, scale), (name, diff, scale)
        for name, d in r['nodes'].items():
            if 'pot' in d:
                a, b = d['pot']
                # small terms: absolute floor 1e-3.  The rotamer free energy is evaluated at beliefs that both engines iterate
                # only to |delta belief| <= 1e-3 (same sweep count, marginals within 1e-3 of each other): floor 5e-3 there
                floor = 5e-3 if name.startswith('rotamer') else 1e-3
                assert abs(a - b) <= E_RTOL * max(1.0, abs(b)) + floor, (name, d)
            else:
                assert d['out'] <= 2e-3 * max(1.0, d['out_scale']), (name, d)
                assert d['sens'] <= 2e-3 * max(1.0, d['sens_scale']), (name, d)


@pytest.mark.skipif(not ref_engine.available('pinned'), reason='oracle/_ref not shipped')
@pytest.mark.parametrize('cid,n_rep', [(1, 4), (2, 2), (3, 3), (4, 2), (5, 2), (6, 3), (8, 3), (9, 3), (10, 3)])
def test_every_node_matches_oracle(cid, n_rep):
    cfg = parity.CONFIGS[cid]
    pos = parity.test_positions(cfg, n_rep + 1)[1:]     # relaxed structures (see DESIGN.md on /input/pos itself)
    _check_report(parity.compare_engines(cfg, pos, verbose=False))


def test_radial_pair_lists_are_the_reference_predicate():
    """radial / hbond_sc_radial (src/sidechain_radial.cpp:16-136; energies, forces and node potentials of config 8 are compared
    with the oracle in test_every_node_matches_oracle): their pair lists equal the reference's predicate and emission order
    (oracle/restate.py) applied to the engine's own node outputs; cutoff = the largest (16-2-1e-6)/inv_dx over the type pairs"""
    cfg = parity.CONFIGS[8]
    pot = h5lite.load(cfg)['input/potential']
    pos = parity.test_positions(cfg, 3)[1:]
    be = ue.BatchEngine(cfg, len(pos))
    be.evaluate(pos)
    for name, sym in (('radial', True), ('hbond_sc_radial', False)):
        g = pot[name]
        args = [parity._s(x) for x in g.attrs['arguments']]
        prm = np.asarray(g['interaction_param'].data, dtype='f4')
        cutoff = np.float32(((16 - 2 - 1e-6) / prm[..., 0].astype('f8')).max())
        i1, id1 = (g['index'].data, g['id'].data) if sym else (g['index1'].data, g['id1'].data)
        i2, id2 = (i1, id1) if sym else (g['index2'].data, g['id2'].data)
        n_edge = 0
        for r in range(len(pos)):
            o1, o2 = be.get_output(args[0], r), be.get_output(args[-1], r)
            want = restate.pairlist(o1[i1][:, :3], id1, o2[i2][:, :3], id2, cutoff, 'seq2', sym)
            got = be.pairlist(name, r)
            assert got.shape == want.shape and (got == want).all(), name
            n_edge += len(got)
        assert n_edge > 0, name
    be.close()


@pytest.mark.parametrize('cid', [1, 2, 3, 5])
def test_golden_fixtures(cid):
    g = np.load(os.path.join(GOLD, 'config%d.npz' % cid))
    pos = g['pos']
    be = ue.BatchEngine(parity.CONFIGS[cid], len(pos))
    en, dv = be.evaluate(pos)
    np.testing.assert_allclose(en, g['energy'], rtol=E_RTOL)
    for r in range(len(pos)):
        assert np.abs(dv[r] - g['deriv'][r]).max() <= F_RTOL * max(1.0, np.abs(g['deriv'][r]).max())
        assert np.abs(be.get_value_by_name('rotamer', 'bead_marginal', r) - g['marginal'][r]).max() <= MARG_ATOL
        for name in parity.PAIRLIST_NODES:
            key = 'pairs_%s_%d' % (name, r)
            if key in g:
                pl = be.pairlist(name, r)
                assert pl.shape == g[key].shape and (pl == g[key]).all(), name
    for k in g.files:
        if k.startswith('pot_'):
            np.testing.assert_allclose(be.node_potential(k[4:]), g[k], rtol=E_RTOL, atol=5e-3 if k[4:].startswith('rotamer') else 1e-3)
    be.close()


def test_single_replica_reference_abi():
    """the reference's own entry points (construct_deriv_engine/evaluate_energy/evaluate_deriv/get_output/...)"""
    g = np.load(os.path.join(GOLD, 'config1.npz'))
    u = ue.Upside(parity.CONFIGS[1])
    assert u.n_atom == 60
    e = u.energy(g['pos'][0])
    assert e == pytest.approx(float(g['energy'][0]), rel=E_RTOL)
    d = u.deriv(g['pos'][0])
    assert np.abs(d - g['deriv'][0]).max() <= F_RTOL * np.abs(g['deriv'][0]).max()
    beads = u.get_output('placement_fixed_point_vector_only')
    np.testing.assert_allclose(beads, g['beads_0'], atol=1e-3)
    assert u.get_output('rotamer').shape == (1, 1)            # potential nodes report (1,1)
    assert u.get_sens('hbond_coverage').shape[1] == 1
    p = u.get_param((20, 20, 62), 'rotamer')
    u.set_param(p, 'rotamer')
    assert u.energy(g['pos'][0]) == pytest.approx(float(e), rel=1e-6)
    with pytest.raises(RuntimeError):
        u.get_output('no_such_node')


def test_device_rng_known_answers():
    g = np.load(os.path.join(GOLD, 'rng.npz'))
    for (s, st, a, t), bits, n3 in zip(g['cases'], g['bits'], g['normal3']):
        b, n, u = ue.rng_probe(int(s), int(st), int(a), int(t))
        assert (b == bits).all()                                    # integer stream is bit-exact
        np.testing.assert_allclose(n, n3, rtol=1e-5, atol=1e-6)     # libm vs CUDA sinf/cosf/logf: ~1-2 ulp
        assert u == pytest.approx(float(restate.u01(int(bits[0]))), rel=1e-7)


@pytest.mark.skipif(not ref_engine.available('pinned'), reason='oracle/_ref not shipped')
def test_restraint_nodes_trajectory_and_afm_clock():
    """config 6 (restraints, AFM with a moving tip, slice): 10 MD rounds from the same start with the same seeds on both
    engines; the AFM tip advances once per DerivMode evaluation on both sides (bonds.cpp:151-154)"""
    cfg = parity.CONFIGS[6]
    pos = parity.test_positions(cfg, 3)[1:]
    ref = ref_engine.md_run(cfg, pos, 0.8, 10, seed=42, n_thread=2, flavour='pinned')
    be = ue.BatchEngine(cfg, len(pos))
    be.set_pos(pos); be.md_init(0.8, seed=42); be.md_run(10)
    assert np.abs(be.get_pos() - ref['pos']).max() <= 5e-3
    t = be.get_value_by_name('AFM', 'time_estimate', 0)
    assert t[0] == pytest.approx(0.009 * 3 * 30, rel=1e-5)            # 30 DerivMode evaluations
    tip = be.get_value_by_name('AFM', 'tip_pos', 0).reshape(-1, 3)
    np.testing.assert_allclose(tip[0], [-10. - 0.01 * t[0], 0., 0.], rtol=1e-5, atol=1e-5)
    be.close()


@pytest.mark.skipif(not ref_engine.available('pinned'), reason='oracle/_ref not shipped')
@pytest.mark.parametrize('cid', [1, 3])
def test_monte_carlo_pivot_moves_match_reference(cid):
    """batched pivot moves (csrc/monte_carlo.cu) against the reference sampler: same seeds, same rounds => the same proposals
    (same random stream), the same accept decisions and the same final coordinates"""
    cfg = parity.CONFIGS[cid]
    pos = parity.test_positions(cfg, 7)[1:]
    rounds = list(range(5, 17))
    ref_pos, ref_stats = ref_engine.mc_steps(cfg, pos, 0.8, 42, rounds[0], len(rounds))
    be = ue.BatchEngine(cfg, len(pos))
    assert be.mc_samplers() == ['pivot']
    be.set_pos(pos); be.md_init(0.8, seed=42)
    for nr in rounds:
        be.mc_execute(nr)
    ok, tr = be.mc_stats(0)
    assert (tr == len(rounds)).all() and (ref_stats[:, 0, 1] == len(rounds)).all()
    assert ref_stats[:, 0, 0].sum() > 0                                  # some moves were accepted: the comparison is not vacuous
    same = ok == ref_stats[:, 0, 0]
    assert same.sum() >= len(pos) - 1                                   # a Metropolis test within rounding of its variate may flip
    got = be.get_pos()
    assert np.abs(got[same] - ref_pos[same]).max() <= 2e-3
    be.close()


@pytest.mark.skipif(not ref_engine.available('pinned'), reason='oracle/_ref not shipped')
@pytest.mark.parametrize('cid', [5, 10])
def test_monte_carlo_jump_moves_match_reference(tmp_path, cid):
    """rigid-body jump moves (JumpSampler, src/monte_carlo_sampler.cpp:203-251) next to the pivot moves: config 5 (the membrane
    potential makes the energy depend on where the molecule sits, so the Metropolis test of a jump is not vacuous) with a
    /input/jump_moves group over the whole chain, and the two-chain config 10 with its own move per chain (the chains see
    each other through the pair potentials); same seeds and rounds on both engines => the same proposals (translation
    or rotation from the reference's random stream), the same decisions, the same coordinates"""
    from upside_md_b200 import config
    if cid == 5:
        cfg = str(tmp_path / 'config5_jump.up')
        w = config.ConfigWriter.from_file(parity.CONFIGS[5])
        w.write_jump_moves([[0, w.n_atom]], 1.5, 0.4)
        w.save(cfg)
    else:
        cfg = parity.CONFIGS[cid]
    pos = parity.test_positions(parity.CONFIGS[cid], 7, relax_rounds=20)[1:]
    rounds = list(range(5, 15))
    ref_pos, ref_stats = ref_engine.mc_steps(cfg, pos, 0.8, 42, rounds[0], len(rounds))
    be = ue.BatchEngine(cfg, len(pos))
    names = be.mc_samplers()
    assert names == ['pivot', 'jump']
    be.set_pos(pos); be.md_init(0.8, seed=42)
    for nr in rounds:
        be.mc_execute(nr)
    same = np.ones(len(pos), dtype=bool)
    for m, nm in enumerate(names):
        ok, tr = be.mc_stats(m)
        assert (tr == len(rounds)).all() and (ref_stats[:, m, 1] == len(rounds)).all()
        same &= ok == ref_stats[:, m, 0]
    n_jump_ok = ref_stats[:, 1, 0].sum()
    assert 0 < n_jump_ok < len(pos) * len(rounds)                        # jumps were both accepted and rejected
    assert same.sum() >= len(pos) - 1                                   # a Metropolis test within rounding of its variate may flip
    got = be.get_pos()
    assert np.abs(got[same] - ref_pos[same]).max() <= 5e-3              # coordinates after up to 20 rigid moves of a 300-residue chain
    be.close()


@pytest.mark.parametrize('cid', [1, 3])
def test_trajectory_matches_golden(cid):
    g = np.load(os.path.join(GOLD, 'config%d.npz' % cid))
    be = ue.BatchEngine(parity.CONFIGS[cid], len(g['pos']))
    be.set_pos(g['pos']); be.md_init(0.8, seed=42); be.md_run(1)
    assert np.abs(be.get_pos() - g['traj_pos_1']).max() <= 1e-4
    assert np.abs(be.get_mom() - g['traj_mom_1']).max() <= 2e-3
    be.set_pos(g['pos']); be.md_init(0.8, seed=42); be.md_run(10)       # 30 timesteps
    assert np.abs(be.get_pos() - g['traj_pos_10']).max() <= 1e-3      # 1e-3 A after 30 deterministic-noise steps
    be.close()


def test_constant_and_concat_nodes():
    """slice + constant -> concat -> springs (config 7): the tether energy is what numpy computes from the positions, and
    dV/dx agrees with central differences (the reference cannot construct its Concat node)"""
    cfg = parity.CONFIGS[7]
    g = h5lite.load(cfg)['input/potential']
    p0 = parity.initial_pos(cfg)
    pts = np.concatenate((p0[np.array(g['slice_some/id'].data)], np.array(g['constant_anchor/value'].data, dtype='f4')))
    ids, eq, k = (np.array(g['dist_spring_tethers/' + n].data) for n in ('id', 'equil_dist', 'spring_const'))
    d = np.sqrt(((pts[ids[:, 0]] - pts[ids[:, 1]]) ** 2).sum(axis=1))
    be = ue.BatchEngine(cfg, 1)
    be.evaluate(p0[None])
    assert float(be.node_potential('dist_spring_tethers')[0]) == pytest.approx(float((0.5 * k * (d - eq) ** 2).sum()), rel=1e-5)
    np.testing.assert_allclose(be.get_output('concat_points', 0), pts, rtol=1e-6, atol=1e-6)
    be.close()
    atoms = np.array(g['slice_some/id'].data)
    idx = np.concatenate([3 * atoms + c for c in range(3)])
    eps = 2e-3
    pos = np.repeat(p0.reshape(1, -1), 2 * len(idx), 0)
    for kk, i in enumerate(idx):
        pos[2 * kk, i] += eps
        pos[2 * kk + 1, i] -= eps
    be = ue.BatchEngine(cfg, len(pos))
    be.evaluate(pos.reshape(len(pos), -1, 3))
    tether = be.node_potential('dist_spring_tethers')
    _, dv = ue.BatchEngine(cfg, 1).evaluate(p0[None])
    be.close()
    # derivative of the tether term alone: total derivative minus that of plain config 1 (same nodes otherwise)
    _, dv1 = ue.BatchEngine(parity.CONFIGS[1], 1).evaluate(p0[None])
    fd = (tether[0::2] - tether[1::2]) / (2 * eps)
    np.testing.assert_allclose((dv[0] - dv1[0]).ravel()[idx], fd, rtol=2e-2, atol=2e-3)


def test_finite_difference_agreement():
    """the reference's own self-check (main.cpp:279-315): central differences of the potential vs dV/dx"""
    g = np.load(os.path.join(GOLD, 'config1.npz'))
    p0 = g['pos'][0]
    rng = np.random.default_rng(0)
    idx = rng.choice(p0.size, 24, replace=False)
    eps = 2e-3
    pos = np.repeat(p0[None], 2 * len(idx), 0).reshape(2 * len(idx), -1)
    for k, i in enumerate(idx):
        pos[2 * k, i] += eps
        pos[2 * k + 1, i] -= eps
    be = ue.BatchEngine(parity.CONFIGS[1], 2 * len(idx))
    en = be.evaluate(pos.reshape(-1, 60, 3), want_deriv=False).astype('f8')
    be.close()
    fd = (en[0::2] - en[1::2]) / (2 * eps)
    u = ue.Upside(parity.CONFIGS[1])
    d = u.deriv(p0).ravel()[idx]
    assert np.sqrt(((fd - d) ** 2).sum() / (d ** 2).sum()) < 2e-2


def test_full_batch_invariants():
    """BASELINE config 3 at full size (4096 replicas): identical replicas give bit-identical results (the engine is
    deterministic and replica-independent), distinct replicas match their single-replica evaluation, internal forces sum
    to zero, and a short MD run keeps every replica finite with a sane kinetic temperature."""
    g = np.load(os.path.join(GOLD, 'config3.npz'))
    B = 4096
    pos = np.empty((B,) + g['pos'].shape[1:], dtype='f4')
    pos[0::2], pos[1::2] = g['pos'][0], g['pos'][1]
    be = ue.BatchEngine(parity.CONFIGS[3], B)
    en, dv = be.evaluate(pos)
    assert (en[0::2] == en[0]).all() and (en[1::2] == en[1]).all()
    # pair terms are gather-form (fixed summation order); the small element-wise nodes scatter with float atomics, so
    # forces of identical replicas agree to rounding, not bit for bit
    assert np.abs(dv[0::2] - dv[0]).max() < 1e-3 and np.abs(dv[1::2] - dv[1]).max() < 1e-3
    np.testing.assert_allclose(en[:2], g['energy'], rtol=E_RTOL)
    net = dv.sum(axis=1)
    assert np.abs(net).max() < 2e-2          # no external field in config 3: net force vanishes (fp32 accumulation)
    be.md_init(0.8, seed=11)
    be.md_run(20)
    p = be.get_pos()
    assert np.isfinite(p).all()
    assert not (p[0] == p[2]).all()            # different seeds => different trajectories
    ke = be.kinetic_energy() / (1.5 * 0.8)
    assert 0.5 < np.median(ke) < 4.0
    be.close()


@pytest.mark.skipif(not ref_engine.available('pinned'), reason='oracle/_ref not shipped')
@pytest.mark.parametrize('cid', [1, 3])
def test_param_deriv_matches_oracle(cid):
    """get_param_deriv (reference PARAM_DERIV build: interaction_graph.h:404-415, bead_interaction.h:86-129,
    hbond.cpp:278-283,447-449, environment.cpp:62-65,375-390, placement.cpp:138-161) after an evaluation with
    derivatives, node by node, through the single-system C ABI and through the batched one"""
    cfg = parity.CONFIGS[cid]
    pos = parity.test_positions(cfg, 3)[1:]
    n_atom = pos.shape[1]
    ref = ref_engine.RefEngine(cfg, n_atom)
    up = ue.Upside(cfg)
    be = ue.BatchEngine(cfg, len(pos))
    be.evaluate(pos)
    names = [n for n, _ in ref.node_names()]
    checked = 0
    for node in names:
        size = be.get_param(node).size
        got_batch = [be.get_param_deriv(node, r) for r in range(len(pos))]
        if not got_batch[0].size:
            continue
        assert got_batch[0].size == size, node
        for r in range(len(pos)):
            ref.deriv(pos[r])                      # every query starts from a fresh backward pass of the reference
            want = ref.get_param_deriv(node, size)
            scale = max(1.0, float(np.abs(want).max()))
            assert np.abs(got_batch[r] - want).max() <= 2e-3 * scale, (node, r, np.abs(got_batch[r] - want).max(), scale)
            if r == 0:
                up.deriv(pos[r])
                one = up.get_param_deriv((size,), node)
                assert np.abs(one - want).max() <= 2e-3 * scale, (node, 'C ABI')
        total = be.get_param_deriv(node, -1)
        assert np.abs(total - np.sum(got_batch, axis=0)).max() <= 1e-3 * max(1.0, float(np.abs(total).max())), node
        checked += 1
    # rotamer pair table, both coverage tables, the environment spline, E_hb and the three fixed placements
    assert checked >= 7, checked
    assert np.abs(be.get_param_deriv('rotamer', 0)).max() > 0
    be.close()
    ref.close()


@pytest.mark.parametrize('env,cids', [({'UPSIDE_B200_NO_FAST_BUILD': '1'}, (1, 3, 5)),        # Verlet cache + k_refine + k_rot_prep
                                      ({'UPSIDE_B200_BUILD_CAPS': '1.5,0.5'}, (1, 3)),          # k_rot_build with its arrays spilling to global memory
                                      ({'UPSIDE_B200_FUSED_REBUILD': '0'}, (2, 3)),             # cache check / tiled rebuild / refine as three launches
                                      ({'UPSIDE_B200_FUSED_REBUILD': '1'}, (5,)),               # in-CTA rebuild on a graph above its size threshold
                                      ({'UPSIDE_B200_NO_VERLET_CACHE': '1'}, (1, 3))])          # all-pairs exact rows every evaluation
def test_alternative_list_paths_give_identical_lists(env, cids, monkeypatch):
    """every way the engine can build its pair structures - the fast rotamer build, its global-memory spill twin, the general
    Verlet / refine / prep path, the fused and the three-launch Verlet cache of the sparse graphs, uncached all-pairs rows - must produce the same pair lists bit for
    bit (they apply the same predicate to the same node outputs) and energies / forces equal to summation-order rounding"""
    for cid in cids:
        cfg = parity.CONFIGS[cid]
        g = np.load(os.path.join(GOLD, 'config%d.npz' % cid)) if cid != 4 else None
        pos = g['pos']
        base = ue.BatchEngine(cfg, len(pos))
        e0, d0 = base.evaluate(pos)
        lists0 = {n: [base.pairlist(n, r) for r in range(len(pos))] for n in parity.PAIRLIST_NODES if n != 'backbone_pairs'}
        m0 = [base.get_value_by_name('rotamer', 'bead_marginal', r) for r in range(len(pos))]
        base.close()
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        alt = ue.BatchEngine(cfg, len(pos))
        e1, d1 = alt.evaluate(pos)
        for n, ls in lists0.items():
            for r, l0 in enumerate(ls):
                l1 = alt.pairlist(n, r)
                assert l1.shape == l0.shape and (l1 == l0).all(), (env, cid, n, r)
        for r in range(len(pos)):
            assert np.abs(alt.get_value_by_name('rotamer', 'bead_marginal', r) - m0[r]).max() <= 1e-5
        np.testing.assert_allclose(e1, e0, rtol=2e-6, atol=1e-4)
        assert np.abs(d1 - d0).max() <= 2e-3
        alt.close()
        for k in env:
            monkeypatch.delenv(k)
