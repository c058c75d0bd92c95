"""GPU tests of the `upside` command line (upside_main of libupside_b200.so) and of replica exchange on the batched engine.
Golden data: tests/golden/cli_replex.npz = the reference binary (unmodified src/main.cpp, oracle/_ref/upside_ref) run with
the same flags on the same inputs (tools/make_golden_cli.py)."""
import os
import shutil
import sys

import numpy as np
import pytest

import parity
from parity import ue
from upside_md_b200 import h5lite
from upside_md_b200 import replica_exchange as rx

pytestmark = pytest.mark.gpu
GOLD = os.path.join(parity.ROOT, 'tests', 'golden')
sys.path.insert(0, os.path.join(parity.ROOT, 'tools'))


def _write_inputs(tmp_path, start):
    import make_golden_cli
    return make_golden_cli.write_inputs(str(tmp_path), start)


def test_cli_replica_exchange_matches_reference_binary(tmp_path):
    g = np.load(os.path.join(GOLD, 'cli_replex.npz'))
    paths = _write_inputs(tmp_path, g['start'])
    ue.in_process_upside([str(a) for a in g['args']] + paths, verbose=False)
    for i, p in enumerate(paths):
        o = h5lite.load(p)['output']
        assert o['pos'].data.shape == g['pos_%d' % i].shape                      # same datasets, same shapes
        np.testing.assert_allclose(o['time'].data, g['time_%d' % i], rtol=1e-6)
        np.testing.assert_allclose(o['temperature'].data, g['temperature_%d' % i], rtol=1e-6)
        # identical exchange history: same decisions (same random stream, energies equal to ~1e-6), same bookkeeping
        assert (o['replica_index'].data == g['replica_index_%d' % i]).all(), i
        assert (o['replica_swap_partner'].data == g['replica_swap_partner_%d' % i]).all()
        assert (o['replica_cumulative_swaps'].data == g['replica_cumulative_swaps_%d' % i]).all()
        # trajectories: 10 rounds = 30 timesteps of deterministic-noise Langevin dynamics (+ exchanges)
        assert np.abs(o['pos'].data[:4] - g['pos_%d' % i][:4]).max() < 2e-3
        assert np.abs(o['pos'].data - g['pos_%d' % i]).max() < 3e-2
        np.testing.assert_allclose(o['potential'].data[:4], g['potential_%d' % i][:4], rtol=2e-4, atol=2e-3)
        np.testing.assert_allclose(o['kinetic'].data[:4], g['kinetic_%d' % i][:4], rtol=2e-3)
        assert 'invocation' in o.attrs


@pytest.mark.skipif(not os.path.exists(os.path.join(parity.ROOT, 'oracle', '_ref', 'upside_ref')), reason='oracle/_ref not shipped')
def test_cli_monte_carlo_matches_reference_binary(tmp_path):
    """--monte-carlo-interval: the reference binary and upside_main run the same command line on the same inputs; pivot moves
    are proposed at the same rounds with the same random streams, so the per-frame pivot_stats and the trajectories agree"""
    import shutil
    import subprocess
    g = np.load(os.path.join(GOLD, 'config1.npz'))
    args = ['--duration', '0.27', '--frame-interval', '0.054', '--monte-carlo-interval', '0.027', '--temperature', '0.8,0.9',
            '--seed', '11']
    mine = _write_inputs(tmp_path, g['start'][:2] if 'start' in g.files else g['pos'][:2])
    (tmp_path / 'ref').mkdir()
    theirs = []
    for p in mine:
        theirs.append(str(tmp_path / 'ref' / os.path.basename(p)))
        shutil.copy(p, theirs[-1])
    exe = os.path.join(parity.ROOT, 'oracle', '_ref', 'upside_ref')
    r = subprocess.run([exe] + args + theirs, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS='1'))
    assert r.returncode == 0, r.stderr
    ue.in_process_upside(args + mine, verbose=False)
    for a, b in zip(mine, theirs):
        oa, ob = h5lite.load(a)['output'], h5lite.load(b)['output']
        sa, sb = np.array(oa['pivot_stats'].data), np.array(ob['pivot_stats'].data)
        assert sa.shape == sb.shape and (sa[:, 1] == sb[:, 1]).all()            # attempts per frame
        assert sb[:, 1].sum() == 8                                              # rounds 1..8 precede the last frame (none at t = 0)
        assert np.abs(sa[:, 0] - sb[:, 0]).sum() <= 1                           # acceptances (a borderline test may flip)
        if (sa == sb).all():
            assert np.abs(np.array(oa['pos'].data) - np.array(ob['pos'].data)).max() < 3e-2


@pytest.mark.skipif(not os.path.exists(os.path.join(parity.ROOT, 'oracle', '_ref', 'upside_ref')), reason='oracle/_ref not shipped')
def test_cli_detailed_loggers_match_reference_binary(tmp_path):
    """--log-level detailed: every dataset the reference's node loggers write (hbond, rama, rama_map_potential,
    nonlinear_coupling, nonbonded_spring_energy, rotamer_free_energy, rotamer_1body_energy*, rotamer_bad_solves_cumulative;
    state_logger.h, hbond.cpp:306, bonds.cpp:199,280, rotamer.cpp:657-672 ...) exists with the same shape, type and, on
    the early frames, the same values"""
    import shutil
    import subprocess
    g = np.load(os.path.join(GOLD, 'config1.npz'))
    args = ['--duration', '0.2', '--frame-interval', '0.054', '--temperature', '0.8', '--seed', '7', '--log-level', 'detailed']
    mine = _write_inputs(tmp_path, g['start'][:2] if 'start' in g.files else g['pos'][:2])
    (tmp_path / 'ref').mkdir()
    theirs = []
    for p in mine:
        theirs.append(str(tmp_path / 'ref' / os.path.basename(p)))
        shutil.copy(p, theirs[-1])
    exe = os.path.join(parity.ROOT, 'oracle', '_ref', 'upside_ref')
    r = subprocess.run([exe] + args + theirs, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS='1'))
    assert r.returncode == 0, r.stderr
    ue.in_process_upside(args + mine, verbose=False)
    for a, b in zip(mine, theirs):
        oa, ob = h5lite.load(a)['output'], h5lite.load(b)['output']
        assert sorted(oa.children) == sorted(ob.children)
        for name in ob.children:
            da, db = np.array(oa[name].data), np.array(ob[name].data)
            assert da.shape == db.shape and da.dtype == db.dtype, (name, da.shape, db.shape, da.dtype, db.dtype)
        for name in ('hbond', 'rama', 'rama_map_potential', 'nonlinear_coupling', 'nonbonded_spring_energy', 'rotamer_free_energy',
                     'rotamer_1body_energy0', 'rotamer_1body_energy1', 'rotamer_1body_energy2', 'rotamer_bad_solves_cumulative'):
            da, db = np.array(oa[name].data, dtype='f8'), np.array(ob[name].data, dtype='f8')
            # frame 0 = the shared start, frame 1 = two rounds later (trajectories agree to ~1e-4 A there)
            assert np.abs(da[:2] - db[:2]).max() <= 5e-3 * max(1.0, np.abs(db[:2]).max()), (name, np.abs(da[:2] - db[:2]).max())
        # the per-residue free energies add up to the rotamer potential
        assert np.isfinite(np.array(oa['rotamer_free_energy'].data)).all()


def test_cli_extensive_loggers_match_reference_binary(tmp_path):
    """--log-level extensive adds `virtual` (hbond.cpp:48-56), `environment_coverage` (environment.cpp:77-82) and
    `placement_pos` (placement.cpp:254-261).  Every placement node asks for the same dataset name, so the reference cannot
    run a full ff_1 configuration at this level (its second H5Dcreate fails) - and neither does the CLI here; a backbone-only
    configuration runs on both and the logged H / O site positions agree."""
    import subprocess
    g = np.load(os.path.join(GOLD, 'config1.npz'))
    full = _write_inputs(tmp_path, g['start'][:1] if 'start' in g.files else g['pos'][:1])[0]
    args = ['--duration', '0.2', '--frame-interval', '0.054', '--temperature', '0.8', '--seed', '7', '--log-level', 'extensive']
    exe = os.path.join(parity.ROOT, 'oracle', '_ref', 'upside_ref')
    env = dict(os.environ, OMP_NUM_THREADS='1')
    ref_copy = str(tmp_path / 'full_ref.up')
    shutil.copy(full, ref_copy)
    assert subprocess.run([exe] + args + [ref_copy], capture_output=True, text=True, env=env).returncode != 0
    with pytest.raises(RuntimeError):      # "while adding '...', logger placement_pos exists already", return code 2
        ue.in_process_upside(args + [full], verbose=False)
    # backbone only: springs, Rama map, backbone pairs, H-bond inference and energy - no placement node
    t = h5lite.load(full)
    pot = t['input/potential']
    keep = {'dist_spring', 'angle_spring', 'dihedral_spring', 'rama_coord', 'rama_map_pot', 'rama_map_pot_ref', 'affine_alignment',
            'backbone_pairs', 'infer_H_O', 'protein_hbond', 'hbond_energy'}
    for name in list(pot.children):
        if name not in keep:
            del pot.children[name]
    if 'output' in t.children:
        del t.children['output']
    mine, theirs = str(tmp_path / 'bb.up'), str(tmp_path / 'bb_ref.up')
    h5lite.save(t, mine)
    shutil.copy(mine, theirs)
    r = subprocess.run([exe] + args + [theirs], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    ue.in_process_upside(args + [mine], verbose=False)
    oa, ob = h5lite.load(mine)['output'], h5lite.load(theirs)['output']
    assert sorted(oa.children) == sorted(ob.children) and 'virtual' in oa.children
    da, db = np.array(oa['virtual'].data), np.array(ob['virtual'].data)
    assert da.shape == db.shape and da.dtype == db.dtype
    assert np.abs(da[:2] - db[:2]).max() <= 2e-3


def test_cli_errors_and_flags(tmp_path):
    g = np.load(os.path.join(GOLD, 'cli_replex.npz'))
    paths = _write_inputs(tmp_path, g['start'][:2])
    with pytest.raises(RuntimeError):     # missing required --frame-interval
        ue.in_process_upside(['--duration', '0.1'] + paths, verbose=False)
    with pytest.raises(RuntimeError):     # 3 temperatures for 2 systems
        ue.in_process_upside(['--duration', '0.1', '--frame-interval', '0.05', '--temperature', '0.7,0.8,0.9'] + paths, verbose=False)
    with pytest.raises(RuntimeError):     # overlapping swap set
        ue.in_process_upside(['--duration', '0.1', '--frame-interval', '0.05', '--replica-interval', '0.05', '--swap-set', '0-1,1-0'] + paths, verbose=False)
    with pytest.raises(RuntimeError):     # unknown flag
        ue.in_process_upside(['--duration', '0.1', '--frame-interval', '0.05', '--no-such-flag'] + paths, verbose=False)
    # annealing + thermostat interval + no recentering run through and log the annealed temperature
    ue.in_process_upside(['--duration', '0.2', '--frame-interval', '0.054', '--temperature', '0.9', '--anneal-factor', '0.5',
                          '--thermostat-interval', '0.054', '--disable-recentering', '--seed', '3'] + paths, verbose=False)
    o = h5lite.load(paths[0])['output']
    T = o['temperature'].data.ravel()
    assert T[0] == pytest.approx(0.9, rel=1e-6) and T[-1] < T[0] and np.isfinite(o['pos'].data).all()


def test_sharded_ladder_on_the_engine_single_rank():
    """ShardedLadder driving a real BatchEngine: exchanges permute coordinates between replicas and nothing else"""
    cfg = parity.CONFIGS[1]
    g = np.load(os.path.join(GOLD, 'config1.npz'))
    n = 6
    pos = np.array([g['pos'][i % 3] for i in range(n)], dtype='f4')
    pos += (np.arange(n, dtype='f4') * 1e-3)[:, None, None]          # make the six configurations distinguishable
    T = np.geomspace(0.7, 1.0, n).astype('f4')
    be = ue.BatchEngine(cfg, n)
    be.set_pos(pos)
    be.md_init_seeds(T, 100 + np.arange(n))
    lad = rx.ShardedLadder(rx.batch_engine_adapter(be), T, ['0-1,2-3,4-5', '1-2,3-4'], seed=5)
    e0 = be.evaluate(want_deriv=False)
    acc = lad.attempt_swaps(3)
    p1 = be.get_pos()
    perm = list(range(n))
    for s in acc:
        for a, b in s:
            perm[a], perm[b] = perm[b], perm[a]
    assert sum(len(s) for s in acc) > 0
    for i in range(n):
        assert (p1[i] == pos[perm[i]]).all()
    e1 = be.evaluate(want_deriv=False)
    assert np.allclose(e1, e0[perm], rtol=1e-6)
    lad.run(4, 2)
    assert np.isfinite(be.get_pos()).all()
    be.close()


def test_checkpoint_resumes_the_random_stream():
    """pos + mom + RNG keys + (thermostat invocations, round number): a run restored into a FRESH engine continues like the
    uninterrupted one - same thermostat noise - to within the rounding of the force evaluations (float atomics in the
    element-wise nodes), far inside the 1e-3 A trajectory gate; without the counters the noise differs and so do the paths"""
    g = np.load(os.path.join(parity.ROOT, 'tests', 'golden', 'config1.npz'))
    pos = np.repeat(g['pos'][:1], 4, 0)
    T = np.array([0.7, 0.8, 0.9, 1.0], dtype='f4')
    a = ue.BatchEngine(parity.CONFIGS[1], 4)
    a.set_pos(pos); a.md_init(T, seed=5); a.md_run(7)
    blob = a.checkpoint()
    mom_at_ckpt = a.get_mom()
    a.md_run(9)
    want_pos, want_mom = a.get_pos(), a.get_mom()
    a.close()
    b = ue.BatchEngine(parity.CONFIGS[1], 4)
    b.restore(blob)
    assert (b.get_mom() == mom_at_ckpt).all()
    b.md_run(9)
    assert np.abs(b.get_pos() - want_pos).max() <= 2e-4
    assert np.abs(b.get_mom() - want_mom).max() <= 5e-3
    # a restart that forgets the counters (fresh md_init on the same coordinates) takes a different path
    c = ue.BatchEngine(parity.CONFIGS[1], 4)
    c.restore(blob); c.md_init(T, seed=5); c.md_run(9)
    assert np.abs(c.get_pos() - want_pos).max() > 1e-2
    with pytest.raises(RuntimeError):
        ue.BatchEngine(parity.CONFIGS[1], 3).restore(blob)          # other batch size
    b.close(); c.close()


def test_continue_config_like_the_reference(tmp_path):
    """continue_sim of py/run_upside.py:231-257: last frame -> /input/pos, /output -> /output_previous_0, run again"""
    cfg = str(tmp_path / 'c1.up')
    shutil.copy(parity.CONFIGS[1], cfg)
    flags = ['--duration', '0.3', '--frame-interval', '0.1', '--temperature', '0.8', '--seed', '3']
    ue.in_process_upside(flags + [cfg], verbose=False)
    last = np.array(h5lite.load(cfg)['output/pos'].data)[-1, 0]
    T = ue.continue_config(cfg)
    t = h5lite.load(cfg)
    assert T == pytest.approx(0.8) and 'output' not in t and 'output_previous_0' in t
    np.testing.assert_array_equal(np.array(t['input/pos'].data)[:, :, 0], last)
    ue.in_process_upside(flags + [cfg], verbose=False)
    t = h5lite.load(cfg)
    assert 'output' in t and 'output_previous_0' in t
    np.testing.assert_allclose(np.array(t['output/pos'].data)[0, 0], last, atol=1e-5)   # the new run starts where the old one stopped
