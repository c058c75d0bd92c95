#!/usr/bin/env python
"""Benchmark of the Upside MD inner loop on B200 (BASELINE.json metric: replica-timesteps/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one MD round of the whole batch = thermostat + integration_cycle = 3 force evaluations + 3 integration
stages (reference src/main.cpp:657-663, src/deriv_engine.cpp:172-192) for every replica, i.e. 3 timesteps per replica.
Workload (config 3 of BASELINE.json): 4096 independent replicas of a 100-residue chain with the full ff_1 force field,
4096 IN TOTAL batched across the N GPUs (4096/N per rank; replicas never interact => no data-path collective,
"scaling": "strong").  `value` = 4096 * 3 * K / seconds with state resident in HBM; `e2e` = the same metric driven
through the batched C ABI with HOST buffers (positions uploaded from pinned host memory and read back every step).
With N > 1 a `weak_scaling` block adds the 4096-replicas-PER-GPU number.  Further blocks on the same line: `config5`
(256 x 300-residue membrane system), `config4` (48-rung ladder, replica exchange over NCCL inside the timed region),
`single_replica_latency` (config 2) and `config1` (CLI wall time), each with the reference's CPU number beside it.

--impl reference times the UNMODIFIED reference engine (oracle/_ref, built from /root/reference by oracle/Makefile)
on this box's host cores, one OpenMP thread per replica as the reference intends, on a bounded sample of the same
workload.  The default arm also reports that number as `cpu_baseline`.
"""
import argparse
import ctypes as ct
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIG = os.path.join(ROOT, 'configs', 'config3_100res.up')
CONFIG4 = os.path.join(ROOT, 'configs', 'config4_150res.up')
CONFIG5 = os.path.join(ROOT, 'configs', 'config5_300res.up')
N_TOTAL = int(os.environ.get('UPSIDE_BENCH_REPLICAS', 4096))   # BASELINE config 3: 4096 replicas IN TOTAL, batched across the GPUs
N_REPLICA = N_TOTAL                                             # (tools/ use this name for the single-GPU batch)
N_TOTAL5 = 256                                                  # BASELINE config 5: 256 replicas of the 300-residue membrane system
N_RUNG, SWAP_INTERVAL = 48, 10                                  # BASELINE config 4: 48-rung ladder, exchange every 10 rounds
TEMPERATURE = 0.80
DT = 0.009
SEED = 42
N_RES = 100
EQUIL_ROUNDS = 300   # the cost of a round drifts until ~250 rounds after random_initial_config (tools/equil_probe.py)
METRIC = 'replica-timesteps/sec, 100-residue protein batch'
UNIT = 'replica-timesteps/s'


def workload_positions(n_replica, first, n_res=N_RES):
    """random_initial_config starts, one per replica (seed = base + global replica index), SURVEY.md section 8(d)"""
    from upside_md_b200 import config
    pos = np.zeros((n_replica, 3 * n_res, 3), dtype='f4')
    for r in range(n_replica):
        pos[r] = config.random_initial_config(n_res, np.random.default_rng(5000 + first + r))
    return pos


def _median(v):
    return float(np.median(np.asarray(v, dtype='f8')))


def cpu_md_sample(cfg, n_res, per_core, warm_rounds, rounds, n_rep, temperature=TEMPERATURE, what=''):
    """reference engine on the host cores (oracle/_ref, unmodified reference sources): `per_core` replicas per core, one
    OpenMP thread per core (the reference's one-thread-per-system scheme, README.md:198-200, main.cpp:618); engines built
    and warmed up (warm_rounds, thread team running) OUTSIDE the timed region, then n_rep repetitions of `rounds` rounds
    are timed and the median is quoted"""
    from oracle import ref_engine
    if not ref_engine.available('fast'):
        return None
    cores = os.cpu_count() or 1
    n_sys = per_core * cores
    pos = workload_positions(n_sys, 0, n_res)
    r = ref_engine.md_bench(cfg, pos, temperature, warm_rounds, rounds, n_rep=n_rep, seed=SEED, dt=DT, n_thread=cores, flavour='fast')
    sec = _median(r['seconds'])
    return dict(value=n_sys * 3 * rounds / sec, unit=UNIT, cores=cores, kind='reference', seconds=sec, rounds=rounds, n_sys=n_sys,
                spread=[float(min(r['seconds'])), float(max(r['seconds']))],
                sample='%s: %d replicas (%d per host core, one OpenMP thread per core), %d warm-up rounds untimed, median of %d x %d '
                       'rounds (%.2f s each); reference sources built -O3 -ffast-math -march=x86-64-v3 -DPARAM_7A_CUTOFF'
                       % (what, n_sys, per_core, warm_rounds, n_rep, rounds, sec),
                us_per_force_eval=sec * 1e6 / (3 * rounds))


def cpu_reference_arm(steps, warmup, target_seconds=None):
    """config 3 on the host cores.  target_seconds: size the repetitions to that much wall time (cpu_baseline leg of the
    default arm); otherwise every repetition is `steps` rounds as the driver passed them (--impl reference)"""
    rounds = steps
    if target_seconds:
        probe = cpu_md_sample(CONFIG, N_RES, 4, 30, 5, 1)
        if probe is None:
            return None
        rounds = int(max(5, target_seconds / 3 / max(probe['seconds'] / 5, 1e-6)))
    return cpu_md_sample(CONFIG, N_RES, 8 if not target_seconds else 4, EQUIL_ROUNDS + max(10, warmup), rounds, 5 if not target_seconds else 3,
                         what='config3 (100 res, ff_1)')


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(',')
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith('active'):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        return dict(sm_mhz=float(np.median(self.samples)) if self.samples else None, sm_max_mhz=self.max_mhz,
                    reasons=sorted(self.reasons), samples=len(self.samples))


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return json.load(open(p)), 'measured'
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0), 'fallback'


def _timed_md(torch, dist, world, eng, stream, rounds, body=None):
    """milliseconds (max over ranks) of `rounds` MD rounds - or of body() - enqueued on the engine's stream, CUDA events on
    that stream, barrier + synchronize on both sides"""
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        if body is None:
            eng.md_run(rounds, sync=False)
        else:
            body()
        ev1.record(stream)
    eng.sync()
    barrier()
    t = torch.tensor([ev0.elapsed_time(ev1)], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _shard(total, world, rank):
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, base + (1 if rank < extra else 0)


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from upside_md_b200 import upside_engine as ue

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    torch.cuda.set_device(local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- headline: BASELINE config 3 as written - 4096 replicas IN TOTAL, sharded over the ranks (no data-path collective) ---
    first, B = _shard(N_TOTAL, world, rank)
    eng = ue.BatchEngine(CONFIG, B, device=local)
    n_atom = eng.n_atom
    eng.set_pos(workload_positions(B, first))
    eng.md_init_seeds(np.full(B, TEMPERATURE, dtype='f4'), SEED + first + np.arange(B), dt=DT)   # seed of replica = base + global index
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)
    # random_initial_config chains start self-intersecting (potential ~ +1e3): EQUIL_ROUNDS untimed rounds bring the batch
    # to the state the metric is quoted on (timing starts after warm-up rounds)
    eng.md_run(EQUIL_ROUNDS)
    eng.md_run(args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    ms = _timed_md(torch, dist, world, eng, stream, args.steps)
    sampler.stop_flag = True
    value = N_TOTAL * 3 * args.steps / (ms * 1e-3)

    # ---- end-to-end arm: host buffers through the C ABI every step -----------------------------------------------
    host_in = torch.from_numpy(eng.get_pos()).pin_memory()
    host_out = torch.empty_like(host_in).pin_memory()
    L, e = eng.L, eng.e
    fp = ct.POINTER(ct.c_float)
    pin, pout = ct.cast(host_in.data_ptr(), fp), ct.cast(host_out.data_ptr(), fp)
    e2e_steps = max(1, min(args.steps, 20))
    for _ in range(2):
        L.ub_set_pos(e, pin); L.ub_md_run(e, 1); L.ub_get_pos(e, pout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        L.ub_set_pos(e, pin)
        L.ub_md_run(e, 1)
        L.ub_get_pos(e, pout)
        pin, pout = pout, pin
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = N_TOTAL * 3 * e2e_steps / float(t.item())
    bytes_io = B * n_atom * 3 * 4

    # ---- roofline of the dominant kernel, measured live with CUDA events on the engine's stream ----------------------
    roof = None
    launches = eng.launches_per_eval()
    if rank == 0:
        roof = kernel_rooflines(eng, stream, torch)
    eng.close()
    del eng

    # ---- weak-scaling companion (N > 1): 4096 replicas PER GPU, the round-1 measurement ---------------------------------
    weak = None
    if world > 1:
        engw = ue.BatchEngine(CONFIG, N_TOTAL, device=local)
        engw.set_pos(workload_positions(N_TOTAL, rank * N_TOTAL))
        engw.md_init(TEMPERATURE, seed=SEED + 100000 * rank, dt=DT)
        sw = torch.cuda.ExternalStream(engw.stream(), device=local)
        engw.md_run(EQUIL_ROUNDS + args.warmup)
        msw = _timed_md(torch, dist, world, engw, sw, args.steps)
        weak = dict(value=world * N_TOTAL * 3 * args.steps / (msw * 1e-3), unit=UNIT, replicas_per_gpu=N_TOTAL, ms_per_step=msw / args.steps,
                    scaling='weak')
        engw.close()
        del engw

    blocks = {}
    if not args.headline_only:
        blocks['config5'] = bench_config5(args, torch, dist, ue, rank, world, local)
        blocks['config4'] = bench_config4(args, torch, dist, ue, rank, world, local)
        if rank == 0:
            blocks['config2'] = bench_config2(ue)
            blocks['config1'] = bench_config1(ue)
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_reference_arm(args.steps, 3, target_seconds=12.0)
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='strong', vs_baseline=None, dtype='f32',
                    data='synthetic',
                    config=dict(workload='config3: %d replicas in total (%d per GPU on %d GPU) x 100-residue chain, ff_1 + side-chain BP, '
                                         'T=0.8, dt=0.009' % (N_TOTAL, N_TOTAL // world, world),
                                replicas_total=N_TOTAL, replicas_per_gpu=N_TOTAL // world, n_res=N_RES, n_atom=n_atom,
                                l2='per-replica state ~1 MB: %.0f MB per GPU >> 126 MB L2' % (B * 1.0),
                                step='1 MD round = thermostat + 3 x (force evaluation + integration stage)',
                                equilibration='%d untimed rounds from random_initial_config before warm-up' % EQUIL_ROUNDS),
                    us_per_force_eval=ms * 1e3 / (3 * args.steps), clocks=sampler.summary(),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=bytes_io, d2h_bytes_per_step=bytes_io, steps=e2e_steps),
                    gpu_launches=int((launches * 3 + 3 + 2) * args.steps), roofline=roof, weak_scaling=weak,
                    single_replica_latency=blocks.get('config2'), config1=blocks.get('config1'), config4=blocks.get('config4'),
                    config5=blocks.get('config5'),
                    cpu_baseline=({k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'spread')} if cpu else None))
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def bench_config5(args, torch, dist, ue, rank, world, local):
    """BASELINE config 5: 256 replicas of the 300-residue chain with membrane_potential + hbond terms, sharded over the ranks;
    CPU sample of the same system beside it (rank 0)"""
    first, B = _shard(N_TOTAL5, world, rank)
    eng = ue.BatchEngine(CONFIG5, B, device=local)
    n_res = eng.n_atom // 3
    eng.set_pos(workload_positions(B, first, n_res))
    eng.md_init_seeds(np.full(B, TEMPERATURE, dtype='f4'), SEED + first + np.arange(B), dt=DT)
    st = torch.cuda.ExternalStream(eng.stream(), device=local)
    # the same protocol as the headline: EQUIL_ROUNDS untimed rounds from the random starts.  (A 300-residue chain is still
    # shedding the clashes of its start after 150 rounds: successive 20-round windows there cost 2.06, 1.47, 1.10, 1.09 ms per
    # evaluation at 256 replicas.)
    eng.md_run(EQUIL_ROUNDS + args.warmup)
    steps = max(10, args.steps)
    ms = _timed_md(torch, dist, world, eng, st, steps)
    out = None
    if rank == 0:
        acc = {}
        for label, m in eng.profile_eval():
            acc[label] = acc.get(label, 0.0) + m
        top = sorted(acc.items(), key=lambda kv: -kv[1])[:6]
        cpu = None if args.no_cpu_baseline else cpu_md_sample(CONFIG5, n_res, 1, EQUIL_ROUNDS, 6, 3, what='config5 (300 res, membrane)')
        out = dict(workload='config5: %d replicas in total (%d per GPU) x 300-residue chain, ff_1 + membrane_potential, T=0.8' % (N_TOTAL5, N_TOTAL5 // world),
                   value=N_TOTAL5 * 3 * steps / (ms * 1e-3), unit=UNIT, ms_per_step=ms / steps, us_per_force_eval=ms * 1e3 / (3 * steps), steps=steps,
                   equilibration='%d untimed rounds' % EQUIL_ROUNDS, top_kernel_groups_ms=dict(top),
                   cpu_baseline=({k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'spread')} if cpu else None))
    eng.close()
    return out


def bench_config4(args, torch, dist, ue, rank, world, local):
    """BASELINE config 4: 48-rung temperature ladder of the 150-residue chain, rungs sharded in contiguous blocks over the ranks,
    replica exchange every SWAP_INTERVAL rounds ON THE DEVICE (csrc/ladder_nccl.cu): energies all-gathered and boundary
    coordinates exchanged over NCCL inside the timed region, no host synchronisation.  CPU: the reference binary on the same ladder."""
    if N_RUNG % world:
        return dict(skipped='48 rungs do not divide over %d ranks' % world) if rank == 0 else None
    n_local = N_RUNG // world
    lo = rank * n_local
    T = np.geomspace(0.70, 1.00, N_RUNG).astype('f4')
    sets = [','.join('%d-%d' % (i, i + 1) for i in range(0, N_RUNG - 1, 2)), ','.join('%d-%d' % (i, i + 1) for i in range(1, N_RUNG - 1, 2))]
    eng = ue.BatchEngine(CONFIG4, n_local, device=local)
    n_res = eng.n_atom // 3
    eng.set_pos(workload_positions(n_local, 7000 + lo, n_res))
    eng.md_init_seeds(T[lo:lo + n_local], SEED + lo + np.arange(n_local), dt=DT)     # seed of system ns = base + ns (main.cpp:459)
    st = torch.cuda.ExternalStream(eng.stream(), device=local)
    comm = ue.nccl_comm_from_torch(dist, local) if world > 1 else None
    lad = ue.Ladder(eng, sets, T, seed=SEED, rank=rank, world=world, nccl_comm=comm)
    rounds = 20 * SWAP_INTERVAL

    def run(n_round, first_round):
        for k in range(n_round // SWAP_INTERVAL):
            eng.md_run(SWAP_INTERVAL, sync=False)
            lad.attempt(first_round + (k + 1) * SWAP_INTERVAL)
    run(150, 0)                                # untimed: relax the random starts; NCCL opens its channels lazily
    eng.sync()
    ms = _timed_md(torch, dist, world, eng, st, rounds, body=lambda: run(rounds, 150))
    ri, acc, n_att, n_suc, _ = lad.state()
    same = True
    if world > 1:
        ref = torch.from_numpy(ri.astype('i4')).cuda()
        mine = ref.clone()
        dist.broadcast(ref, 0)
        flag = torch.tensor([int((ref == mine).all())], device='cuda')
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
    gather_b, coord_b = lad.comm_bytes()
    out = None
    if rank == 0:
        cpu = None if args.no_cpu_baseline else cpu_ladder_reference(T, sets)
        out = dict(workload='config4: 48-rung ladder T=0.70..1.00 of the 150-residue chain (%d rungs per GPU), exchange every %d rounds, '
                            'swap sets 0-1,2-3,.. / 1-2,3-4,..' % (n_local, SWAP_INTERVAL),
                   value=N_RUNG * 3 * rounds / (ms * 1e-3), unit=UNIT, ms_per_step=ms / rounds, rounds=rounds, exchange_attempts=rounds // SWAP_INTERVAL,
                   swap_attempts=int(n_att.sum()), swap_accepted=int(n_suc.sum()), decisions_identical_on_all_ranks=same,
                   comm=dict(backend='nccl' if world > 1 else 'none (one GPU)', allgather_bytes_per_attempt_per_rank=gather_b,
                             boundary_coordinate_bytes_sent_per_attempt_rank0=coord_b, host_syncs_per_attempt=0),
                   cpu_baseline=cpu)
    lad.close()
    eng.close()
    if comm:
        ue.lib().ub_nccl_comm_destroy(comm)
    return out


def _ref_binary():
    p = os.path.join(ROOT, 'oracle', '_ref', 'upside_ref')
    return p if os.path.exists(p) else None


def cpu_ladder_reference(T, sets, rounds=60):
    """the reference binary on the same 48-rung ladder, one OpenMP thread per rung up to the host's cores (README.md:198-218);
    reads its own timing line"""
    import re, shutil, tempfile
    exe = _ref_binary()
    if not exe:
        return None
    cores = os.cpu_count() or 1
    from upside_md_b200 import h5lite
    with tempfile.TemporaryDirectory() as tmp:
        files = []
        for i in range(N_RUNG):
            f = os.path.join(tmp, 'r%02d.up' % i)
            shutil.copy(CONFIG4, f)
            files.append(f)
        dur = rounds * 3 * DT
        cmd = [exe, '--duration', '%.6f' % dur, '--frame-interval', '%.6f' % (dur * 2), '--temperature', ','.join('%.6f' % t for t in T),
               '--seed', str(SEED), '--time-step', str(DT), '--replica-interval', '%.6f' % (SWAP_INTERVAL * 3 * DT), '--log-level', 'basic']
        for s_ in sets:
            cmd += ['--swap-set', s_]
        env = dict(os.environ, OMP_NUM_THREADS=str(min(cores, N_RUNG)))
        t0 = time.perf_counter()
        r = subprocess.run(cmd + files, capture_output=True, text=True, env=env, timeout=600)
        wall = time.perf_counter() - t0
    m = re.search(r'finished in ([0-9.]+) seconds', r.stdout)
    if r.returncode or not m:
        return dict(error='upside_ref failed', stderr=r.stderr[-300:])
    sec = float(m.group(1))
    return dict(value=N_RUNG * 3 * rounds / sec, unit=UNIT, cores=min(cores, N_RUNG), kind='reference', seconds=sec, wall_seconds=wall,
                sample='the reference binary (oracle/_ref/upside_ref) on the same 48-rung ladder, %d rounds from /input/pos, %d OpenMP '
                       'threads; its own "finished in" timing (construction excluded)' % (rounds, min(cores, N_RUNG)))


def bench_config2(ue):
    """BASELINE config 2: force-evaluation latency of ONE replica of the 76-residue chain through the reference's own C ABI
    (evaluate_deriv of include/engine_c_library.h: host positions in, host derivatives out, synchronous), next to the
    reference engine's own latency on one host core"""
    cfg2 = os.path.join(ROOT, 'configs', 'config2_76res.up')
    up = ue.Upside(cfg2)
    p2 = np.ascontiguousarray(up.initial_pos, dtype='f4')
    for _ in range(20):
        up.deriv(p2)
    n_lat = 200
    t0 = time.perf_counter()
    for _ in range(n_lat):
        up.deriv(p2)
    us = (time.perf_counter() - t0) * 1e6 / n_lat
    del up
    cpu = None
    try:
        from oracle import ref_engine
        if ref_engine.available('fast'):
            cpu = dict(value=ref_engine.eval_latency(cfg2, p2, 20, n_lat), unit='us per force evaluation', cores=1, kind='reference',
                       sample='%d evaluate_deriv calls of the reference library on one host core, same coordinates' % n_lat)
    except Exception as ex:   # the CPU leg must never take the GPU numbers down with it
        cpu = dict(error=str(ex))
    return dict(us_per_force_eval=us, workload='config2: 76-residue chain, 1 replica, evaluate_deriv through the single-system C ABI '
                '(host buffers, synchronous)', evaluations=n_lat, cpu_baseline=cpu)


def bench_config1(ue):
    """BASELINE config 1: the `upside` command line on the 20-residue chain, 333 rounds = 1000 Langevin steps, wall time of
    upside_main (this library, in process) and of the reference binary on the same file"""
    import re, shutil, tempfile
    cfg1 = os.path.join(ROOT, 'configs', 'config1_20res.up')
    flags = ['--duration', '9', '--frame-interval', '0.9', '--temperature', '0.8', '--seed', str(SEED), '--time-step', str(DT)]
    out = dict(workload='config1: 20-residue chain, 1 replica, upside CLI, --duration 9 (333 rounds = 1000 force evaluations), 10 frames')
    with tempfile.TemporaryDirectory() as tmp:
        f = os.path.join(tmp, 'c1.up')
        shutil.copy(cfg1, f)
        ue.in_process_upside(flags + [f], verbose=False)       # first call: CUDA context, graph capture
        shutil.copy(cfg1, f)
        t0 = time.perf_counter()
        rc = ue.in_process_upside(flags + [f], verbose=False)
        out['seconds'] = time.perf_counter() - t0
        out['rc'] = rc
        out['us_per_step'] = out['seconds'] * 1e6 / 999
        exe = _ref_binary()
        if exe:
            shutil.copy(cfg1, f)
            t0 = time.perf_counter()
            r = subprocess.run([exe] + flags + [f], capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS='1'), timeout=600)
            wall = time.perf_counter() - t0
            m = re.search(r'finished in ([0-9.]+) seconds', r.stdout)
            out['cpu_baseline'] = dict(value=wall, unit='seconds', cores=1, kind='reference', reported=m.group(0) if m else None,
                                       sample='the reference binary on the same file and flags, one host core; wall time of the '
                                              'process (its own "finished in" line has 0.1 s resolution)')
    return out


def kernel_source_hash():
    """hash of the CUDA sources: stamps profiles/traffic.json so that ncu evidence of another build is never quoted"""
    import hashlib
    h = hashlib.sha1()
    d = os.path.join(ROOT, 'upside-md_b200', 'csrc')
    for f in sorted(os.listdir(d)):
        if f.endswith(('.cu', '.cuh', '.h', '.cpp')):
            h.update(f.encode())
            h.update(open(os.path.join(d, f), 'rb').read())
    return h.hexdigest()[:16]


def kernel_rooflines(eng, stream, torch):
    """Roofline of the dominant kernel group, measured live: ub_profile_eval times one evaluation kernel group by kernel
    group with CUDA events on the engine's stream (after warm-up, mean of 3); the algorithmic flops of SURVEY.md
    section 8(d) are attributed group by group from live counts (edges, BP sweeps, residue-pair classes)."""
    peaks, how = measured_peaks()
    fp32_nominal = 148 * 128 * 2 * peaks.get('sm_max_mhz', 1965.0) * 1e6 / 1e12
    from upside_md_b200 import upside_engine as ue
    fp32_peak, _ = ue.measure_fp32_peak(int(os.environ.get('LOCAL_RANK', 0)))   # FFMA microbenchmark on this GPU (csrc/peaks.cu)
    n = 5
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        eng.md_run(n, sync=False)
        ev1.record(stream)
    eng.sync()
    ms_eval = ev0.elapsed_time(ev1) / (3 * n)
    acc = {}
    order = []
    for rep in range(3):
        for label, ms in eng.profile_eval():
            if label not in acc:
                acc[label] = 0.0
                order.append(label)
            acc[label] += ms / 3
    # live counts from a sample of replicas
    sample = range(0, eng.n_replica, max(1, eng.n_replica // 16))
    eng.evaluate(want_deriv=False)
    e_rot = np.mean([len(eng.pairlist('rotamer', r)) for r in sample])
    e_cov = np.mean([len(eng.pairlist('hbond_coverage', r)) + len(eng.pairlist('hbond_coverage_hydrophobe', r)) for r in sample])
    e_env = np.mean([len(eng.pairlist('environment_coverage', r)) for r in sample])
    e_hb = np.mean([len(eng.pairlist('protein_hbond', r)) for r in sample])
    st = np.array([eng.get_value_by_name('rotamer', 'solve_stats', r) for r in sample])
    sweeps, pairs = st[:, 0].mean() + 1, st[:, 1].mean()
    n66, n36 = st[:, 3].mean(), st[:, 4].mean()
    n33 = pairs - n66 - n36
    nrot = np.array([1, 1, 3, 3, 3, 3, 3, 3, 3, 3, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6]).mean()   # per residue type, SURVEY section 8
    f_bp = sweeps * (275 * n66 + 190 * n36 + 110 * n33 + N_RES * (4 * nrot + 5))
    groups = {   # algorithmic flops per replica and evaluation, and the profile labels that do the work
        'rotamer pair term (k_rot_energy + k_rot_deriv)': (e_rot * 324, ['rotamer/energy', 'rotamer/deriv']),
        'rotamer belief propagation (k_rot_bp2)': (f_bp, ['rotamer/bp']),
        'hbond_coverage x2 (fwd + bwd)': (e_cov * 324, ['hbond_coverage:fwd', 'hbond_coverage_hydrophobe:fwd', 'hbond_coverage:bwd',
                                                         'hbond_coverage_hydrophobe:bwd']),
        'environment_coverage (fwd + bwd)': (e_env * 110, ['environment_coverage:fwd', 'environment_coverage:bwd']),
        'protein_hbond (fwd + bwd)': (e_hb * 314, ['protein_hbond:fwd', 'protein_hbond:bwd']),
    }
    table = []
    for name, (flops, labels) in groups.items():
        ms = sum(acc.get(l, 0.0) for l in labels)
        ach = flops * eng.n_replica / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
        table.append(dict(kernel=name, ms=ms, algorithmic_flops_per_replica=flops, achieved_tflops=ach, frac=ach / fp32_peak))
    pl_ms = acc.get('(pairlist)', 0.0) + acc.get('rotamer/pairlist', 0.0)
    table.append(dict(kernel='pair lists (k_cache_check + k_pairlist + k_refine, the four sparse graphs)', ms=pl_ms))
    table.append(dict(kernel='rotamer build (k_rot_build: residue spheres, bead masks, CSR rows)', ms=acc.get('rotamer/build', 0.0) + acc.get('rotamer/prep', 0.0)))
    top = max(table[:5], key=lambda t: t['ms'])
    flops_total = e_rot * 324 + e_cov * 324 + e_hb * 314 + e_env * 110 + f_bp + 2000 * N_RES
    # ncu evidence (DRAM bytes, issue-slot and FMA-pipe utilisation per launch) comes from profiles/traffic.json, captured by
    # tools/ncu_full.sh + tools/make_traffic.py and stamped with a hash of the kernel sources: a stale file is not quoted
    traffic, ncu, traffic_note = None, None, 'profiles/traffic.json absent'
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get('source_hash') == kernel_source_hash():
            for t in table:
                if t['kernel'] in tj:
                    t['ncu'] = tj[t['kernel']]
            ncu = tj.get(top['kernel'])
            traffic = ncu and ncu.get('dram_bytes_per_launch')
            traffic_note = 'ncu --set full of this build (source hash %s)' % tj.get('source_hash')
        else:
            traffic_note = 'profiles/traffic.json was captured from other kernel sources (hash %s, now %s): not quoted' % (tj.get('source_hash'), kernel_source_hash())
    return dict(bound='fp32', achieved=top['achieved_tflops'], peak=fp32_peak, unit='TFLOP/s', frac=top['frac'], traffic=traffic, ncu=ncu,
                peak_nominal=fp32_nominal,
                kernel=top['kernel'], ms_per_launch=top['ms'], algorithmic_flops_per_replica=top['algorithmic_flops_per_replica'],
                whole_evaluation=dict(ms=ms_eval, algorithmic_flops_per_replica=flops_total,
                                      achieved_tflops=flops_total * eng.n_replica / (ms_eval * 1e-3) / 1e12,
                                      frac=flops_total * eng.n_replica / (ms_eval * 1e-3) / 1e12 / fp32_peak,
                                      sum_of_serial_kernel_ms=sum(acc.values())),
                kernels=table,
                counts=dict(rotamer_edges=e_rot, coverage_edges=e_cov, env_edges=e_env, hbond_edges=e_hb, bp_sweeps=sweeps, bp_pairs=pairs,
                            bp_pairs_6x6=n66, bp_pairs_3x6=n36),
                peak_source='FP32 FMA pipe measured live by an FFMA microbenchmark (ub_measure_fp32_peak, csrc/peaks.cu); nominal = 148 SM x '
                            '128 lanes x 2 x sm_max_mhz (%s clocks).  MEASURED_PEAKS.json has no FP32 entry; the path has no dense '
                            'contraction (profiles/).  traffic/ncu: ' % how + traffic_note)


def reference_arm(args):
    """the reference's own CPU implementation (oracle/_ref) on the host cores, same metric/config; engines and thread team
    warmed up outside the timed region, median of 5 repetitions of `steps` rounds"""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cpu = cpu_reference_arm(args.steps, args.warmup)
    if cpu is None:
        emit(dict(impl='reference', unavailable='oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)'))
        return
    line = dict(impl='reference', metric=METRIC, value=cpu['value'], unit=UNIT, n_gpus=args.gpus, steps=cpu['rounds'], warmup=args.warmup,
                ms_per_step=cpu['seconds'] * 1e3 / cpu['rounds'], higher_is_better=True, scaling='strong', vs_baseline=None,
                dtype='f32', data='synthetic',
                config=dict(workload='config3: 100-residue chain, ff_1 + side-chain BP, T=0.8, dt=0.009; CPU sample of %d replicas '
                                     '(the host does not grow with --gpus)' % cpu['n_sys']),
                us_per_force_eval=cpu['us_per_force_eval'],
                cpu_baseline={k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample', 'spread')},
                e2e=dict(value=cpu['value'], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    emit(line)


_REAL_STDOUT = None


def emit(line):
    """the one JSON line goes to the real stdout; everything else a library prints (NCCL's version banner) went to stderr"""
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line, default=float) + '\n').encode())


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=30)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--headline-only', action='store_true', help='config 3 only (skip the config 1/2/4/5 blocks)')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
