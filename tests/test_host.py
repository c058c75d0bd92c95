"""CPU tests of the host-side pieces: HDF5 reader/writer, configuration generator, C-ABI surface."""
import os
import re
import subprocess

import numpy as np
import pytest

import parity
from upside_md_b200 import config, h5lite
from upside_md_b200 import upside_engine as ue


def test_h5lite_roundtrip(tmp_path):
    g = h5lite.File()
    g.attrs['x'] = 3
    p = g.create_group('input/potential/foo')
    p.attrs['arguments'] = np.array(['pos', 'bar'])
    p.attrs['f'] = np.float32(1.5)
    g.create_dataset('input/pos', np.random.rand(30, 3, 1).astype('f4'))
    big = np.arange(5000).reshape(50, 100)
    g.create_dataset('input/big', big, chunks=(16, 100), compress=True)
    g.create_dataset('input/empty', np.zeros((0, 2)))
    for i in range(30):
        g.create_dataset('input/many/d%02d' % i, np.arange(i))
    path = str(tmp_path / 't.h5')
    h5lite.save(g, path)
    q = h5lite.load(path)
    assert int(q.attrs['x']) == 3
    assert [a.decode() for a in q['input/potential/foo'].attrs['arguments']] == ['pos', 'bar']
    assert (q['input/big'].data == big).all()
    assert q['input/empty'].data.shape == (0, 2)
    assert q['input/many'].keys() == ['d%02d' % i for i in range(30)]
    assert (q['input/many/d07'].data == np.arange(7)).all()


def test_config_files_have_reference_schema():
    for cid, cfg in parity.CONFIGS.items():
        t = h5lite.load(cfg)
        pot = t['input/potential']
        n_res = len(t['input/sequence'].data)
        assert t['input/pos'].data.shape == (3 * n_res, 3, 1)
        for name in ('dist_spring', 'angle_spring', 'dihedral_spring', 'rama_coord', 'affine_alignment', 'infer_H_O',
                     'protein_hbond', 'hbond_energy', 'hbond_coverage', 'hbond_coverage_hydrophobe', 'rotamer',
                     'placement_scalar', 'placement_fixed_point_vector_only', 'environment_coverage',
                     'nonlinear_coupling_environment', 'weighted_pos', 'rama_map_pot', 'rama_map_pot_ref', 'backbone_pairs'):
            assert name in pot, (cid, name)
            assert 'arguments' in pot[name].attrs
        assert pot['rotamer/pair_interaction/interaction_param'].data.shape == (20, 20, 62)
        assert pot['hbond_coverage/interaction_param'].data.shape == (2, 20, 54)
        assert ('membrane_potential' in pot) == (cid == 5)
        ids = pot['rotamer/pair_interaction/id'].data
        assert ((ids & 15) < ((ids >> 4) & 15)).all()


def test_chain_break_writer():
    """two-chain configuration (config 10): the chain_break group of upside_config.py:1440-1441, no donor / acceptor on the
    residues next to the break (:1445-1449), bonded terms and Rama coordinates confined to one chain, one jump move per chain"""
    t = h5lite.load(parity.CONFIGS[10])
    cfr = t['input/chain_break/chain_first_residue'].data
    assert cfr.tolist() == [20]
    pot = t['input/potential']
    for site in ('donors', 'acceptors'):
        assert not set(pot['infer_H_O/%s/residue' % site].data.tolist()) & {19, 20}
    for node in ('dist_spring', 'angle_spring', 'dihedral_spring'):
        ids = pot[node + '/id'].data
        assert ((ids // 3 >= 20) == (ids[:, :1] // 3 >= 20)).all(), node
    assert len(pot['dist_spring/id'].data) == 3 * 40 - 2
    rc = pot['rama_coord/id'].data
    assert rc[19, 4] == -1 and rc[20, 0] == -1 and (rc[19, :4] >= 0).all() and (rc[20, 1:] >= 0).all()
    assert t['input/jump_moves/atom_range'].data.tolist() == [[0, 60], [60, 120]]
    w = config.ConfigWriter(['ALA'] * 6, np.zeros((18, 3)))
    with pytest.raises(ValueError):
        w.write_chain_break([4, 2])


def test_random_initial_config_geometry():
    pos = config.random_initial_config(30, np.random.default_rng(1))
    d = np.linalg.norm(pos[1:] - pos[:-1], axis=1)
    # the reference assigns lengths[:,k] to the bond INTO atom k (upside_config.py:435-476), so the random start has
    # N-CA = 1.526, CA-C = 1.300, C-N = 1.453 (strained relative to dist_spring's 1.453/1.526/1.300); restated as is
    np.testing.assert_allclose(d[0::3], 1.526, atol=1e-6)
    np.testing.assert_allclose(d[1::3], 1.300, atol=1e-6)
    np.testing.assert_allclose(d[2::3], 1.453, atol=1e-6)
    assert (config.random_sequence(50, 7) == config.random_sequence(50, 7)).all()


def _declared_functions(header):
    txt = open(os.path.join(parity.ROOT, 'include', header)).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b([a-z_][a-z0-9_]*)\s*\([^;{]*\)\s*;', txt)))


def test_library_exports_every_declared_symbol():
    L = ue.lib()
    names = _declared_functions('engine_c_library.h') + _declared_functions('upside_b200.h')
    assert len(names) > 40
    for nm in names:
        assert hasattr(L, nm), 'libupside_b200.so does not export %s' % nm


def test_no_cpu_fallback():
    """without a CUDA device the product path must fail loudly instead of computing on the CPU"""
    if ue.lib().ub_device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(RuntimeError):
        ue.BatchEngine(parity.CONFIGS[1], 1)
    with pytest.raises(RuntimeError):
        ue.Upside(parity.CONFIGS[1])


def test_product_does_not_link_the_oracle():
    """nothing under upside-md_b200/ or include/ links, imports, loads or even names the oracle (tier rule 3)"""
    out = subprocess.run(['ldd', ue.LIB_PATH], capture_output=True, text=True).stdout
    assert 'upside_ref' not in out
    n_checked = 0
    for top in ('upside-md_b200', 'include'):
        for dp, _, fs in os.walk(os.path.join(parity.ROOT, top)):
            for f in fs:
                if f.endswith(('.py', '.cu', '.cpp', '.h', '.cuh')) or f == 'Makefile':
                    text = open(os.path.join(dp, f), errors='ignore').read().lower()
                    for word in ('oracle', 'upside_ref', 'ref_engine', '/root/reference'):
                        assert word not in text, (os.path.join(dp, f), word)
                    n_checked += 1
    assert n_checked > 20


def test_clamped_spline_helpers():
    vals = np.array([0., 1., 4., 9., 16., 25.], dtype='f4')
    c = ue.clamped_spline_solve(vals)
    x = np.arange(1, 7, dtype='f4')
    got = ue.clamped_spline_value(c, x)
    np.testing.assert_allclose(got, vals, atol=1e-4)     # the spline interpolates its data at integer knots
    vd = ue.clamped_value_and_deriv(c, np.array([3.5], dtype='f4'))
    eps = 1e-2
    fd = (ue.clamped_spline_value(c, np.array([3.5 + eps], 'f4')) - ue.clamped_spline_value(c, np.array([3.5 - eps], 'f4'))) / (2 * eps)
    assert vd[0, 1] == pytest.approx(float(fd[0]), rel=1e-2)


def test_restraint_config_writers_against_reference_potentials():
    """config 6 (tools/make_restraint_config.py, written with config.py's restraint writers): the reference engine reads it,
    and its restraint energies are what numpy computes from the written tables (bonds.cpp:36-47,77-88,355-372,406-424)"""
    from oracle import ref_engine
    if not ref_engine.available('pinned'):
        pytest.skip('oracle/_ref not built')
    import parity
    cfg = parity.CONFIGS[6]
    pot = h5lite.load(cfg)['input/potential']
    pos = parity.initial_pos(cfg).astype('f8')
    ref = ref_engine.RefEngine(cfg, len(pos), 'pinned')
    ref.energy(pos.astype('f4'))
    g = pot['atom_pos_spring']
    d = pos[np.array(g['id'].data)] - np.array(g['x0'].data)
    assert ref.node_potential('atom_pos_spring') == pytest.approx(float((0.5 * np.array(g['spring_const'].data) * (d ** 2).sum(1)).sum()), rel=1e-5)
    g = pot['tension']
    assert ref.node_potential('tension') == pytest.approx(float(-(pos[np.array(g['atom'].data)] * np.array(g['tension_coeff'].data)).sum()), rel=1e-5)
    g = pot['z_flat_bottom']
    dz = pos[np.array(g['atom'].data), 2] - np.array(g['z0'].data)
    rad = np.array(g['radius'].data)
    ex = np.where(dz > rad, dz - rad, np.where(dz < -rad, dz + rad, 0.))
    assert ref.node_potential('z_flat_bottom') == pytest.approx(float((0.5 * np.array(g['spring_constant'].data) * ex ** 2).sum()), rel=1e-5, abs=1e-6)
    g = pot['cavity_radial']
    rr = np.sqrt((pos[np.array(g['id'].data)] ** 2).sum(1))
    ex = np.maximum(rr - np.array(g['radius'].data), 0.)
    assert ref.node_potential('cavity_radial') == pytest.approx(float((0.5 * np.array(g['spring_constant'].data) * ex ** 2).sum()), rel=1e-4, abs=1e-5)
    # the restraint group appended unbonded springs to dist_spring
    ds = pot['dist_spring']
    assert (np.array(ds['bonded_atoms'].data) == 0).sum() > 0 and len(ds['id'].data) == len(ds['equil_dist'].data)
    ref.close()
