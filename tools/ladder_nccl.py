#!/usr/bin/env python
"""BASELINE config 4: 48-rung temperature ladder of the 150-residue chain, rungs sharded in contiguous blocks over the ranks
of a torchrun job (one process per GPU, NCCL), exchanges every 10 rounds.  Prints one JSON line from rank 0.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/ladder_nccl.py [rounds]
Also runs on one GPU without torchrun."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.distributed as dist
from upside_md_b200 import upside_engine as ue, replica_exchange as rx, h5lite, config

CFG = os.path.join(ROOT, 'configs', 'config4_150res.up')
N_RUNG, INTERVAL, SEED = 48, 10, 42
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
group = None
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    group = dist.group.WORLD
torch.cuda.set_device(local)
T = np.geomspace(0.70, 1.00, N_RUNG).astype('f4')
sets = [','.join('%d-%d' % (i, i + 1) for i in range(0, N_RUNG - 1, 2)), ','.join('%d-%d' % (i, i + 1) for i in range(1, N_RUNG - 1, 2))]
lo, hi = rx.block_bounds(N_RUNG, world, rank)
be = ue.BatchEngine(CFG, hi - lo, device=local)
n_res = be.n_atom // 3
pos = np.array([config.random_initial_config(n_res, np.random.default_rng(7000 + g)) for g in range(lo, hi)], dtype='f4')
be.set_pos(pos)
be.md_init_seeds(T[lo:hi], SEED + np.arange(lo, hi))          # seed of system ns = base + ns (main.cpp:459)
be.md_run(30)                                                 # relax the random starts
lad = rx.ShardedLadder(rx.batch_engine_adapter(be), T, sets, seed=SEED, group=group, device=torch.device('cuda', local))
lad.run(2 * INTERVAL, INTERVAL)                                # untimed: NCCL opens its all-gather and send/recv channels lazily
torch.cuda.synchronize()
if world > 1: dist.barrier()
t0 = time.perf_counter()
lad.run(rounds, INTERVAL)
be.sync(); torch.cuda.synchronize()
if world > 1: dist.barrier()
dt = time.perf_counter() - t0
idx = torch.from_numpy(lad.plan.replica_indices().astype('i4')).cuda()
same = True
if world > 1:
    ref = idx.clone(); dist.broadcast(ref, 0)
    flag = torch.tensor([int((ref == idx).all())], device='cuda'); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    same = bool(flag.item())
en = be.evaluate(want_deriv=False)
fin = torch.tensor([int(np.isfinite(en).all())], device='cuda')
if world > 1: dist.all_reduce(fin, op=dist.ReduceOp.MIN)
att = sum(int(lad.plan.counts(s)[0].sum()) for s in range(2)); suc = sum(int(lad.plan.counts(s)[1].sum()) for s in range(2))
if rank == 0:
    print(json.dumps(dict(workload='config4: 48-rung ladder 0.70-1.00, 150 residues, exchange every 10 rounds', n_gpus=world, rounds=rounds,
                          seconds=dt, replica_timesteps_per_s=N_RUNG * 3 * rounds / dt, swap_attempts=att, swap_accepted=suc,
                          cross_rank_swaps_rank0=lad.n_cross_rank_swaps, decisions_identical_on_all_ranks=same, energies_finite=bool(fin.item()),
                          replica_indices=lad.plan.replica_indices().tolist())))
be.close()
if world > 1: dist.destroy_process_group()
