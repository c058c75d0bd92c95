"""CPU tests of the replica-exchange host logic (no GPU): the host random stream, swap-set parsing, Metropolis decisions
against the numpy restatement of ReplicaExchange::attempt_swaps (oracle/restate.py), and the sharded ladder over
torch.distributed (gloo, world_size 2) with a numpy stand-in for the engine."""
import os
import socket
import sys

import numpy as np
import pytest

import parity  # noqa: F401  (sets sys.path)
from oracle import restate
from upside_md_b200 import upside_engine as ue
from upside_md_b200 import replica_exchange as rx

GOLD = os.path.join(parity.ROOT, 'tests', 'golden')


def test_host_rng_matches_restatement_and_device_kats():
    g = np.load(os.path.join(GOLD, 'rng.npz'))
    for (s, st, a, t), bits in zip(g['cases'], g['bits']):
        u, b = ue.host_rng_uniform(int(s), int(st), int(a), int(t), 3)
        assert (b == bits).all()                                    # same known answers as the device generator
        for call in range(3):
            assert u[call] == restate.u01(restate.random_bits(int(s), int(st), int(a), int(t), call)[0])


def test_swap_set_parsing_errors():
    rx.ReplexPlan(4, ['0-1,2-3', '1-2'])
    for bad in (['0-1,1-2'], ['0-0'], ['0-7'], ['0-1-2'], ['a-b']):
        with pytest.raises(RuntimeError):
            rx.ReplexPlan(4, bad)


class NumpyEngine(object):
    """stand-in for BatchEngine: coordinates are (n,1,3) arrays tagging each configuration, the 'potential' is a fixed
    function of the configuration (one Hamiltonian for all rungs)"""

    def __init__(self, pos):
        self.pos = np.array(pos, dtype='f4')
        self.n_replica = len(self.pos)
        self.rounds = 0

    def energies(self):
        return energy_of(self.pos)

    def get_pos_range(self, first, n):
        return self.pos[first:first + n].copy()

    def set_pos_range(self, pos, first):
        self.pos[first:first + len(pos)] = pos

    def swap_pos(self, pairs):
        for a, b in np.asarray(pairs).reshape(-1, 2):
            self.pos[[a, b]] = self.pos[[b, a]]

    def md_run(self, n):
        self.rounds += n
        self.pos[:, 0, 1] += np.float32(0.25) * n     # "dynamics": the energy drifts, the tag in [:,0,0] stays


def energy_of(pos):
    return (np.float32(50.) * np.sin(pos[:, 0, 0] * np.float32(1.7)) + pos[:, 0, 1] * pos[:, 0, 0] * np.float32(0.3)).astype('f4')


def ladder_inputs(n):
    T = np.geomspace(0.7, 1.0, n).astype('f4')
    sets = [','.join('%d-%d' % (i, i + 1) for i in range(0, n - 1, 2)), ','.join('%d-%d' % (i, i + 1) for i in range(1, n - 1, 2))]
    pos = np.zeros((n, 1, 3), dtype='f4')
    pos[:, 0, 0] = np.arange(n) + 1           # tag = configuration id + 1
    return T, sets, pos


def reference_run(n, n_round, interval, seed):
    """the numpy restatement driven the same way: returns the configuration tag in every slot after the run"""
    T, sets, pos = ladder_inputs(n)
    pairs = restate.parse_swap_sets(sets)
    state = {i: pos[i].copy() for i in range(n)}      # configuration id -> coordinates
    slots = list(range(n))
    stats = {}
    for done in range(interval, n_round + 1, interval):
        for c in state.values():
            c[0, 1] += np.float32(0.25) * interval
        slots = restate.attempt_swaps(seed, done, pairs, T, lambda i, conf: energy_of(state[conf][None])[0], slots, stats)
    return [int(state[c][0, 0]) for c in slots], stats


@pytest.mark.parametrize('n', [4, 7, 12])
def test_single_process_ladder_matches_restatement(n):
    T, sets, pos = ladder_inputs(n)
    eng = NumpyEngine(pos)
    lad = rx.ShardedLadder(eng, T, sets, seed=11)
    lad.run(24, 3)
    tags, stats = reference_run(n, 24, 3, 11)
    assert [int(t) for t in eng.pos[:, 0, 0]] == tags
    assert sorted(tags) == list(range(1, n + 1))
    assert tags != list(range(1, n + 1))                  # the test ladder really exchanges
    for si in range(len(lad.plan.sets)):
        att, suc = lad.plan.counts(si)
        for pi in range(len(att)):
            assert [int(suc[pi]), int(att[pi])] == stats[(si, pi)]
    # replica_indices follows the configurations (configuration c sits where replica_indices == c)
    assert [int(i) + 1 for i in lad.plan.replica_indices()] == tags


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    T, sets, pos = ladder_inputs(n)
    lo, hi = rx.block_bounds(n, world, rank)
    eng = NumpyEngine(pos[lo:hi])
    lad = rx.ShardedLadder(eng, T, sets, seed=11, group=dist.group.WORLD, device='cpu')
    lad.run(24, 3)
    np.savez(os.path.join(out_dir, 'rank%d.npz' % rank), tags=eng.pos[:, 0, 0], lo=lo, hi=hi, cross=lad.n_cross_rank_swaps,
             idx=lad.plan.replica_indices())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n', [7, 12])
def test_sharded_ladder_world_size_2_gloo(n, tmp_path):
    """two processes, each owning a contiguous block of rungs: the all-gathered energies give both ranks the same decisions
    and the boundary pair travels by send/recv; the result equals the single-process run"""
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), n, str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(os.path.join(str(tmp_path), 'rank%d.npz' % r)) for r in range(2)]
    tags = [int(t) for p in parts for t in p['tags']]
    ref_tags, _ = reference_run(n, 24, 3, 11)
    assert tags == ref_tags
    assert (parts[0]['idx'] == parts[1]['idx']).all()                 # identical bookkeeping on every rank
    assert int(parts[0]['cross']) == int(parts[1]['cross']) > 0      # the boundary pair did exchange across ranks
    assert parts[0]['hi'] == parts[1]['lo']


def test_library_exports_every_declared_symbol():
    """every function declared in include/*.h is exported by libupside_b200.so (no compute calls here)"""
    import re
    L = ue.lib()
    names = set()
    for h in ('engine_c_library.h', 'upside_b200.h'):
        src = open(os.path.join(parity.ROOT, 'include', h)).read()
        src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
        names |= set(re.findall(r'\b(\w+)\s*\([^;{]*\)\s*;', src))
    names -= {'defined'}
    assert len(names) > 40
    for n in sorted(names):
        assert hasattr(L, n), n
