"""GPU tests of the device-resident ladder exchange (csrc/ladder_nccl.cu) on one rank: its decisions, counters, replica
indices and coordinate moves against the host plan (ub_replex_*, the code path that tests/test_gpu_cli.py pins to the
reference binary's swap history), from the same energies.  The multi-rank path is exercised by tools/ladder_nccl.py under
torchrun on 2-8 GPUs (profiles/)."""
import numpy as np
import pytest

import parity
from parity import ue
from upside_md_b200 import replica_exchange as rx

pytestmark = pytest.mark.gpu


def ladder_sets(n):
    return [','.join('%d-%d' % (i, i + 1) for i in range(0, n - 1, 2)), ','.join('%d-%d' % (i, i + 1) for i in range(1, n - 1, 2))]


@pytest.mark.parametrize('cid,n_rung', [(1, 8), (3, 12)])
def test_device_ladder_matches_host_plan(cid, n_rung):
    cfg = parity.CONFIGS[cid]
    T = np.geomspace(0.70, 1.00, n_rung).astype('f4')
    sets = ladder_sets(n_rung)
    p0 = parity.initial_pos(cfg)
    be = ue.BatchEngine(cfg, n_rung)
    be.set_pos(np.repeat(p0[None], n_rung, 0))
    be.md_init(T, seed=42)
    be.md_run(30)
    lad = ue.Ladder(be, sets, T, seed=42)
    plan = rx.ReplexPlan(n_rung, sets)
    beta = (1.0 / T).astype('f4')
    n_accept = 0
    for it in range(12):
        rnd = 10 * (it + 1)
        pos_before = be.get_pos()
        energy = be.evaluate(want_deriv=False).astype('f4')
        lad.attempt(rnd)
        ri, acc, n_att, n_suc, en_dev = lad.state()
        # the attempt decided on the energies of these coordinates (two evaluations agree to rounding: potentials of the
        # element-wise nodes are summed with float atomics); the host plan below gets exactly the device's numbers
        np.testing.assert_allclose(en_dev, energy, rtol=2e-6, atol=1e-4)
        # host plan on the same energies: set after set, energies follow the accepted configurations
        plan.begin(42, rnd)
        e = en_dev.copy()
        perm = np.arange(n_rung)                                # perm[slot] = slot whose coordinates end up here
        acc_host = []
        for s, pairs in enumerate(plan.sets):
            a = plan.decide_same_hamiltonian(s, beta, e)
            acc_host.extend(a.tolist())
            for (s1, s2), ok in zip(pairs, a):
                if ok:
                    e[[s1, s2]] = e[[s2, s1]]
                    perm[[s1, s2]] = perm[[s2, s1]]
        assert acc.astype(bool).tolist() == acc_host
        assert (be.get_pos() == pos_before[perm]).all()
        assert (ri == plan.replica_indices()).all()
        n_accept += int(np.sum(acc_host))
        be.md_run(5)
    tot_att = np.concatenate([plan.counts(s)[0] for s in range(len(plan.sets))])
    tot_suc = np.concatenate([plan.counts(s)[1] for s in range(len(plan.sets))])
    _, _, n_att, n_suc, _ = lad.state()
    assert (n_att == tot_att).all() and (n_suc == tot_suc).all()
    assert 0 < n_accept < 12 * (n_rung - 1)                     # both outcomes occurred: the comparison is not vacuous
    lad.close()
    be.close()


def test_ladder_rejects_bad_sharding():
    be = ue.BatchEngine(parity.CONFIGS[1], 4)
    with pytest.raises(RuntimeError):
        ue.Ladder(be, ladder_sets(6), np.ones(6, dtype='f4'), rank=0, world=1)     # engine holds 4 replicas, ladder has 6 rungs
    with pytest.raises(RuntimeError):
        ue.Ladder(be, ['0-1,1-2'], np.ones(4, dtype='f4'))                          # overlapping pairs in a swap set
    be.close()
