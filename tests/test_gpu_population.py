"""Population-scale parity (pytest -m gpu): >= 256 replicas per BASELINE config, drawn from the state the bench times -
random_initial_config starts equilibrated for 300 rounds - plus the extended-strand starts of SURVEY.md section 8(d), through
both engines (CUDA via the C ABI, oracle = the unmodified reference, pinned flavour).  BASELINE.json: ">= 90 % of replicas'
energies and forces within tolerance of the reference on every config" - here the pass fraction is asserted per config with
energy 1e-4 relative, force 1e-3 * max(1,|F|_inf), BP marginals 1e-3; replicas whose rotamer pair list differs from the
reference's (a bead pair within rounding of the cutoff) are counted separately; belief-propagation
sweep counts must equal the oracle's (the stopping rule is checked every chunk of 2 sweeps, so a deviation that sits on the
tolerance may move the stop by one chunk)."""
import numpy as np
import pytest

import parity
from parity import ue
from oracle import ref_engine
from upside_md_b200 import config

pytestmark = pytest.mark.gpu
N_POP = 256
E_RTOL, F_RTOL, MARG_ATOL = 1e-4, 1e-3, 1e-3


def population(cfg, n_rep, seed):
    """half: random coils equilibrated 300 rounds at T = 0.8 on the GPU engine (the bench state); half: extended strands
    after 20 rounds (structures after >= 1 MD round: DESIGN.md section 4 on the reference's eigen-solver at /input/pos)"""
    n_res = parity.initial_pos(cfg).shape[0] // 3
    half = n_rep // 2
    coil = np.array([config.random_initial_config(n_res, np.random.default_rng(seed + r)) for r in range(half)], dtype='f4')
    ext = np.array([config.extended_initial_config(n_res, np.random.default_rng(seed + 10000 + r)) for r in range(n_rep - half)], dtype='f4')
    out = []
    for pos, rounds in ((coil, 300), (ext, 20)):
        be = ue.BatchEngine(cfg, len(pos))
        be.set_pos(pos)
        be.md_init(0.8, seed=seed)
        be.md_run(rounds)
        out.append(be.get_pos())
        be.close()
    return np.concatenate(out)


@pytest.mark.skipif(not ref_engine.available('pinned'), reason='oracle/_ref not shipped')
@pytest.mark.parametrize('cid', [1, 2, 3, 4, 5])
def test_population_pass_fraction(cid):
    cfg = parity.CONFIGS[cid]
    pos = population(cfg, N_POP, 900 + cid)
    be = ue.BatchEngine(cfg, len(pos))
    en, dv = be.evaluate(pos)
    ref = ref_engine.RefEngine(cfg, pos.shape[1], 'pinned')
    ok = np.zeros(len(pos), dtype=bool)
    flips, sweeps_equal, sweeps_within_chunk = 0, 0, 0
    worst = dict(e=0.0, f=0.0, m=0.0)
    for r in range(len(pos)):
        e_ref = ref.energy(pos[r])
        d_ref = ref.deriv(pos[r])
        e_err = abs(float(en[r]) - e_ref) / max(1.0, abs(e_ref))
        f_err = float(np.abs(dv[r] - d_ref).max()) / max(1.0, float(np.abs(d_ref).max()))
        m_err = float(np.abs(be.get_value_by_name('rotamer', 'bead_marginal', r) - ref.rotamer_bead_marginals()).max())
        pg, pr = be.pairlist('rotamer', r), ref.pairlist('rotamer')
        same_list = pg.shape == pr.shape and bool((pg == pr).all())
        flips += not same_list
        if not same_list:   # a flip is a bead pair or two on the cutoff (the two engines' bead coordinates differ by ~1e-5 A), not a different list
            n_diff = len(set(map(tuple, pg)) ^ set(map(tuple, pr)))
            assert n_diff <= 3, (r, n_diff)
        it_g = int(be.get_value_by_name('rotamer', 'solve_stats', r)[0])
        it_r = ref.rotamer_solve_stats()['n_iter']
        sweeps_equal += it_g == it_r
        sweeps_within_chunk += abs(it_g - it_r) <= 2
        ok[r] = e_err <= E_RTOL and f_err <= F_RTOL and m_err <= MARG_ATOL
        if same_list:
            worst['e'], worst['f'], worst['m'] = max(worst['e'], e_err), max(worst['f'], f_err), max(worst['m'], m_err)
    be.close()
    ref.close()
    frac = ok.mean()
    print('config %d: %d replicas, pass fraction %.4f, rotamer pair-list flips %d, BP sweep counts equal %d / within one chunk %d, '
          'worst (identical lists) energy %.2e force %.2e marginal %.2e' % (cid, len(pos), frac, flips, sweeps_equal, sweeps_within_chunk,
                                                                             worst['e'], worst['f'], worst['m']))
    assert frac >= 0.90, frac
    assert sweeps_equal >= 0.90 * len(pos) and sweeps_within_chunk >= 0.98 * len(pos), (sweeps_equal, sweeps_within_chunk)
    # about one bead pair in 1e5 lies within rounding of the cutoff: ~1 % of replicas at 100 residues, ~5 % at 300 (12 k pairs each)
    assert flips <= 0.10 * len(pos), flips


@pytest.mark.skipif(not ref_engine.available('pinned'), reason='oracle/_ref not shipped')
def test_outliers_of_the_bench_batch_match_the_reference():
    """The bench batch itself (4096 random coils of config 3, 300 rounds): a dozen of its replicas sit in trapped, clashing
    states with stretched bonds and energies of 10^3-10^4 (median ~10^2).  Those are where a rare-event defect of a kernel
    would hide, so the ten highest-energy replicas go through the reference engine on the same coordinates: total energy,
    every node potential and the forces must agree there as they do on relaxed structures."""
    import bench
    B = 4096
    be = ue.BatchEngine(bench.CONFIG, B)
    n_res = be.n_atom // 3
    be.set_pos(bench.workload_positions(B, 0, n_res))
    be.md_init_seeds(np.full(B, bench.TEMPERATURE, dtype='f4'), bench.SEED + np.arange(B), dt=bench.DT)
    be.md_run(300)
    en, dv = be.evaluate(want_deriv=True)
    pos = be.get_pos()
    assert np.isfinite(en).all() and np.isfinite(dv).all()
    worst = np.argsort(en)[::-1][:10]
    node_pot = {name: be.node_potential(name) for name, is_pot in be.node_names() if is_pot}
    ref = ref_engine.RefEngine(bench.CONFIG, be.n_atom)
    for w in worst:
        e_ref = ref.energy(pos[w])
        d_ref = ref.deriv(pos[w])
        assert abs(en[w] - e_ref) <= E_RTOL * max(1.0, abs(e_ref)), (int(w), float(en[w]), e_ref)
        assert np.abs(dv[w] - d_ref).max() <= F_RTOL * max(1.0, np.abs(d_ref).max()), (int(w), float(np.abs(dv[w] - d_ref).max()))
        for name, v in node_pot.items():
            r = ref.node_potential(name)
            assert abs(v[w] - r) <= 5e-3 + E_RTOL * abs(r), (int(w), name, float(v[w]), r)
    assert en[worst[0]] > 10 * np.median(en)          # the sample does contain outliers
    ref.close()
    be.close()
