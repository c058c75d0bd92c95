"""Replica exchange over batched engines: one process per GPU, rungs sharded in contiguous blocks.

Reference behaviour (src/main.cpp:120-275, README.md:207-218): every `replica_interval` rounds, for each swap set, all
systems evaluate their energy, the coordinates of every pair of the set are exchanged, the energies are evaluated again
and each pair is kept or reverted by the Metropolis rule with a counter-based random number that depends only on
(seed, round, draw index).  The reference does this serially on one core.

Here a ladder of n_system rungs over ONE configuration (a temperature ladder, the reference's use case) is split over
`world_size` processes, rank r owning the contiguous block of rungs [lo_r, hi_r) as the replicas of its own batched
engine.  Per swap set:
  1. every rank evaluates the energies of its rungs (one batched evaluation),
  2. one all-gather of n_system floats,
  3. the Hamiltonian is the same for all rungs, so the trial energies are a permutation of the gathered ones
     (E_i(x_j) = E_j): no second evaluation,
  4. every rank runs the same Metropolis pass (libupside_b200's ub_replex_*, the same code the `upside` CLI uses) and gets
     identical decisions - nothing is broadcast,
  5. accepted pairs exchange coordinates: inside a rank on the device, across ranks with one send/recv per pair
     (3*n_atom floats; with contiguous blocks only the block-boundary pairs cross GPUs).
Only steps 2 and 5 touch the interconnect (NCCL over NVLink on a GPU box, gloo in the CPU tests).

The engine is passed in as an object with `n_replica`, `energies()`, `get_pos_range(first,n)`, `set_pos_range(pos,first)`
and `swap_pos(pairs)`; `upside_engine.BatchEngine` provides them, and the CPU tests use a numpy stand-in with the same
interface so that the host logic (sharding, gather, decisions, routing of coordinates) runs without a GPU.
"""
import ctypes as ct

import numpy as np

from . import upside_engine as ue


def block_bounds(n_system, world_size, rank):
    """contiguous block [lo, hi) of rungs owned by `rank` (blocks differ in size by at most one)"""
    base, extra = divmod(n_system, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def owner_of(n_system, world_size, system):
    for r in range(world_size):
        lo, hi = block_bounds(n_system, world_size, r)
        if lo <= system < hi:
            return r
    raise ValueError(system)


class ReplexPlan(object):
    """ctypes wrapper of the host-side plan (include/upside_b200.h: ub_replex_*)"""

    def __init__(self, n_system, swap_sets):
        L = ue.lib()
        arr = (ct.c_char_p * len(swap_sets))(*[s.encode() for s in swap_sets])
        self.h = L.ub_replex_create(int(n_system), len(swap_sets), arr)
        if not self.h:
            raise RuntimeError('replica exchange: %s' % (L.ub_last_error() or b'').decode())
        self.L, self.n_system = L, int(n_system)
        self.sets = []
        for s in range(L.ub_replex_n_sets(self.h)):
            n = L.ub_replex_set_size(self.h, s)
            p = np.zeros(2 * n, dtype='i4')
            L.ub_replex_pairs(self.h, s, p.ctypes.data_as(ct.POINTER(ct.c_int)))
            self.sets.append(p.reshape(n, 2))
        if not self.sets:
            raise RuntimeError('replica exchange requested but no swap sets proposed')

    def close(self):
        if getattr(self, 'h', None):
            self.L.ub_replex_destroy(self.h)
            self.h = None

    __del__ = close

    def begin(self, seed, round_num):
        self.L.ub_replex_begin(self.h, int(seed) & 0xffffffff, int(round_num))

    def decide_same_hamiltonian(self, set_index, beta, energy):
        beta = np.require(beta, dtype='f4', requirements='C')
        energy = np.require(energy, dtype='f4', requirements='C')
        acc = np.zeros(len(self.sets[set_index]), dtype='i4')
        fp = ct.POINTER(ct.c_float)
        if self.L.ub_replex_decide_same_hamiltonian(self.h, set_index, beta.ctypes.data_as(fp), energy.ctypes.data_as(fp),
                                                    acc.ctypes.data_as(ct.POINTER(ct.c_int))):
            raise RuntimeError('replica exchange decide failed')
        return acc.astype(bool)

    def decide(self, set_index, old_lboltz, new_lboltz):
        o = np.require(old_lboltz, dtype='f4', requirements='C')
        n = np.require(new_lboltz, dtype='f4', requirements='C')
        acc = np.zeros(len(self.sets[set_index]), dtype='i4')
        fp = ct.POINTER(ct.c_float)
        if self.L.ub_replex_decide(self.h, set_index, o.ctypes.data_as(fp), n.ctypes.data_as(fp), acc.ctypes.data_as(ct.POINTER(ct.c_int))):
            raise RuntimeError('replica exchange decide failed')
        return acc.astype(bool)

    def replica_indices(self):
        out = np.zeros(self.n_system, dtype='i4')
        self.L.ub_replex_replica_indices(self.h, out.ctypes.data_as(ct.POINTER(ct.c_int)))
        return out

    def counts(self, set_index):
        n = len(self.sets[set_index])
        a, s = np.zeros(n, dtype='u8'), np.zeros(n, dtype='u8')
        p = ct.POINTER(ct.c_uint64)
        self.L.ub_replex_counts(self.h, set_index, a.ctypes.data_as(p), s.ctypes.data_as(p))
        return a, s


class ShardedLadder(object):
    """Temperature ladder sharded over the ranks of a torch.distributed process group (or a single process)."""

    def __init__(self, engine, temperatures, swap_sets, seed, group=None, device=None):
        self.engine = engine
        self.T = np.asarray(temperatures, dtype='f4')
        self.n_system = len(self.T)
        self.seed = int(seed)
        self.plan = ReplexPlan(self.n_system, swap_sets)
        self.dist = None
        self.rank, self.world = 0, 1
        if group is not None:
            import torch.distributed as dist
            self.dist, self.group = dist, group
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.device = device
        self.lo, self.hi = block_bounds(self.n_system, self.world, self.rank)
        if engine.n_replica != self.hi - self.lo:
            raise ValueError('engine holds %d replicas but rank %d owns rungs [%d,%d)' % (engine.n_replica, self.rank, self.lo, self.hi))
        self.beta = (np.float32(1.) / self.T).astype('f4')
        self.n_cross_rank_swaps = 0

    def local_temperatures(self):
        return self.T[self.lo:self.hi]

    def _gather_energies(self):
        local = np.asarray(self.engine.energies(), dtype='f4')
        if self.dist is None:
            return local
        import torch
        counts = [block_bounds(self.n_system, self.world, r) for r in range(self.world)]
        width = max(hi - lo for lo, hi in counts)
        buf = torch.zeros(width, dtype=torch.float32, device=self.device)
        buf[:len(local)] = torch.from_numpy(local).to(self.device)
        out = [torch.zeros(width, dtype=torch.float32, device=self.device) for _ in range(self.world)]
        self.dist.all_gather(out, buf, group=self.group)
        return np.concatenate([o.cpu().numpy()[:hi - lo] for o, (lo, hi) in zip(out, counts)])

    def _exchange_coordinates(self, pairs):
        """apply the accepted exchanges: device swaps inside the block, send/recv across blocks"""
        local = [(a - self.lo, b - self.lo) for a, b in pairs if self.lo <= a < self.hi and self.lo <= b < self.hi]
        if local:
            self.engine.swap_pos(np.array(local, dtype='i4'))
        if self.dist is None:
            return
        import torch
        ops, recv = [], []
        for a, b in pairs:
            mine = [s for s in (a, b) if self.lo <= s < self.hi]
            if len(mine) != 1:
                continue
            me, other = (a, b) if mine[0] == a else (b, a)
            peer = owner_of(self.n_system, self.world, other)
            out = torch.from_numpy(self.engine.get_pos_range(me - self.lo, 1)).to(self.device)
            inp = torch.empty_like(out)
            ops.append(self.dist.P2POp(self.dist.isend, out, peer, group=self.group))
            ops.append(self.dist.P2POp(self.dist.irecv, inp, peer, group=self.group))
            recv.append((me, inp))
            self.n_cross_rank_swaps += 1
        if ops:
            for w in self.dist.batch_isend_irecv(ops):
                w.wait()
        for me, inp in recv:
            self.engine.set_pos_range(inp.cpu().numpy(), me - self.lo)

    def attempt_swaps(self, round_num):
        """one ReplicaExchange::attempt_swaps (main.cpp:227-275) for the whole ladder; returns accepted pairs per set"""
        self.plan.begin(self.seed, round_num)
        result = []
        for s, pairs in enumerate(self.plan.sets):
            energy = self._gather_energies()
            acc = self.plan.decide_same_hamiltonian(s, self.beta, energy)
            accepted = [tuple(int(x) for x in p) for p, ok in zip(pairs, acc) if ok]
            self._exchange_coordinates(accepted)
            result.append(accepted)
        return result

    def run(self, n_round, replica_interval):
        """n_round MD rounds with an exchange attempt every `replica_interval` rounds (as the CLI's main loop)"""
        done = 0
        while done < n_round:
            n = min(replica_interval - done % replica_interval, n_round - done)
            self.engine.md_run(n)
            done += n
            if done % replica_interval == 0:
                self.attempt_swaps(done)
        return done


def batch_engine_adapter(be):
    """give a BatchEngine the small interface ShardedLadder needs"""

    class _A(object):
        n_replica = be.n_replica

        def energies(self):
            return be.evaluate(want_deriv=False)

        def get_pos_range(self, first, n):
            return be.get_pos_range(first, n)

        def set_pos_range(self, pos, first):
            be.set_pos_range(pos, first)

        def swap_pos(self, pairs):
            be.swap_pos(pairs)

        def md_run(self, n):
            be.md_run(n)

    return _A()
