#!/usr/bin/env python
"""Benchmark of the Upside MD inner loop on B200 (BASELINE.json metric: replica-timesteps/s).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

One "step" = one MD round of the whole batch = thermostat + integration_cycle = 3 force evaluations + 3 integration
stages (reference src/main.cpp:657-663, src/deriv_engine.cpp:172-192) for every replica, i.e. 3 timesteps per replica.
Workload (config 3 of BASELINE.json): 4096 independent replicas of a 100-residue chain with the full ff_1 force field
per GPU; with N GPUs every rank runs its own 4096 replicas (replicas never interact => no data-path collective,
"scaling": "weak").  `value` = replicas * 3 * K / seconds with state resident in HBM; `e2e` = the same metric driven
through the batched C ABI with HOST buffers (positions uploaded from pinned host memory and read back every step).

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
N_REPLICA = int(os.environ.get('UPSIDE_BENCH_REPLICAS', 4096))
TEMPERATURE = 0.80
DT = 0.009
SEED = 42
N_RES = 100
EQUIL_ROUNDS = 300   # the cost of a round drifts until ~250 rounds after random_initial_config (tools/equil_probe.py)
METRIC = 'replica-timesteps/sec, 100-residue protein batch'
UNIT = 'replica-timesteps/s'


def workload_positions(n_replica, rank):
    """random_initial_config starts, one per replica (seed = base + replica), SURVEY.md §8(d)"""
    from upside_md_b200 import config
    pos = np.zeros((n_replica, 3 * N_RES, 3), dtype='f4')
    for r in range(n_replica):
        pos[r] = config.random_initial_config(N_RES, np.random.default_rng(5000 + 100000 * rank + r))
    return pos


def cpu_reference_arm(steps, warmup, max_seconds=60.0, target_seconds=None):
    """reference engine on the host cores: a bounded sample of the workload = 4 replicas per core (one OpenMP thread per
    core, the reference's one-thread-per-system scheme, README.md:198-200); `steps` rounds, or - for the cpu_baseline leg -
    as many rounds as fill `target_seconds` of CPU time"""
    from oracle import ref_engine
    if not ref_engine.available('fast'):
        return None
    cores = os.cpu_count() or 1
    n_sys = 4 * cores
    pos = workload_positions(n_sys, 0)
    # warm-up rounds relax the clashing random starts exactly like the GPU arm's warm-up does
    w = ref_engine.md_run(CONFIG, pos, TEMPERATURE, EQUIL_ROUNDS + max(1, warmup), seed=SEED, dt=DT, n_thread=cores, flavour='fast')
    per_round = max(w['seconds'] / (EQUIL_ROUNDS + max(1, warmup)), 1e-6)
    rounds = int(max(1, target_seconds / per_round)) if target_seconds else int(max(1, min(steps, max_seconds / per_round)))
    r = ref_engine.md_run(CONFIG, w['pos'], TEMPERATURE, rounds, seed=SEED + 1, dt=DT, n_thread=cores, flavour='fast')
    value = n_sys * 3 * rounds / r['seconds']
    return dict(value=value, unit=UNIT, cores=cores, kind='reference', seconds=r['seconds'], rounds=rounds, n_sys=n_sys,
                sample='%d replicas (4 per host core, one OpenMP thread per core) x %d rounds = %.1f s of config3 (100 res, ff_1), '
                       'reference sources built -O3 -ffast-math -march=x86-64-v3 -DPARAM_7A_CUTOFF' % (n_sys, rounds, r['seconds']),
                us_per_force_eval=r['seconds'] * 1e6 / (3 * rounds))


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
    B = N_REPLICA
    eng = ue.BatchEngine(CONFIG, B, device=local)
    n_atom = eng.n_atom
    pos0 = workload_positions(B, rank)
    eng.set_pos(pos0)
    eng.md_init(TEMPERATURE, seed=SEED + 1000 * rank, dt=DT)
    stream = torch.cuda.ExternalStream(eng.stream(), device=local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -----------------------------------------------------------------------------------
    # random_initial_config chains start self-intersecting (potential ~ +1e3): EQUIL_ROUNDS untimed rounds bring the batch
    # to the state the metric is quoted on (timing starts after 30 warm-up rounds)
    eng.md_run(EQUIL_ROUNDS)
    eng.md_run(args.warmup)
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ev0.record(stream)
        eng.md_run(args.steps, sync=False)
        ev1.record(stream)
    eng.sync()
    barrier()
    ms = ev0.elapsed_time(ev1)
    sampler.stop_flag = True
    t = torch.tensor([ms], device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * 3 * args.steps / (ms * 1e-3)

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
    e2e_value = world * B * 3 * e2e_steps / float(t.item())
    bytes_io = B * n_atom * 3 * 4

    # ---- roofline of the dominant kernel (k_rotamer), measured live with CUDA events on the engine's stream --------
    roof = None
    launches = eng.launches_per_eval()
    if rank == 0:
        roof = kernel_rooflines(eng, stream, torch)
    # force-evaluation latency of BASELINE config 2 (76 residues, ONE replica) through the reference's own C ABI
    # (evaluate_deriv of include/engine_c_library.h: host positions in, host derivatives out, synchronous)
    latency = None
    if rank == 0:
        cfg2 = os.path.join(ROOT, 'configs', 'config2_76res.up')
        up = ue.Upside(cfg2)
        p2 = np.ascontiguousarray(up.initial_pos, dtype='f4')
        for _ in range(20):
            up.deriv(p2)
        t0 = time.perf_counter()
        n_lat = 200
        for _ in range(n_lat):
            up.deriv(p2)
        latency = dict(us_per_force_eval=(time.perf_counter() - t0) * 1e6 / n_lat, workload='config2: 76-residue chain, 1 replica, '
                       'evaluate_deriv through the single-system C ABI (host buffers, synchronous)', evaluations=n_lat)
        del up
    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        cpu = cpu_reference_arm(args.steps, 3, target_seconds=12.0)
    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=args.warmup,
                    ms_per_step=ms / args.steps, higher_is_better=True, scaling='weak', vs_baseline=None, dtype='f32',
                    data='synthetic',
                    config=dict(workload='config3: %d replicas/GPU x 100-residue chain, ff_1 + side-chain BP, T=0.8, dt=0.009' % B,
                                replicas_per_gpu=B, n_res=N_RES, n_atom=n_atom, l2='per-replica state %.0f MB total >> 126 MB L2' % (B * 1.2),
                                step='1 MD round = thermostat + 3 x (force evaluation + integration stage)',
                                equilibration='%d untimed rounds from random_initial_config before warm-up' % EQUIL_ROUNDS),
                    us_per_force_eval=ms * 1e3 / (3 * args.steps), clocks=sampler.summary(),
                    e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=bytes_io, d2h_bytes_per_step=bytes_io, steps=e2e_steps),
                    gpu_launches=int((launches * 3 + 3 + 2) * args.steps), roofline=roof, single_replica_latency=latency,
                    cpu_baseline=({k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')} if cpu else None))
        emit(line)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


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
    table.append(dict(kernel='pair lists (k_cache_check + k_pairlist + k_refine, all graphs)', ms=pl_ms))
    table.append(dict(kernel='rotamer prep (k_rot_prep)', ms=acc.get('rotamer/prep', 0.0)))
    top = max(table[:5], key=lambda t: t['ms'])
    flops_total = e_rot * 324 + e_cov * 324 + e_hb * 314 + e_env * 110 + f_bp + 2000 * N_RES
    traffic, ncu = None, None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        for t in table:
            if t['kernel'] in tj:
                t['ncu'] = tj[t['kernel']]
        ncu = tj.get(top['kernel'])
        traffic = ncu and ncu.get('dram_bytes_per_launch')
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
                            'contraction and moves ~1.5 TB/s of HBM (profiles/).  `traffic` and `ncu` come from profiles/traffic.json '
                            '(ncu --set full of the same workload), not from this run' % how)


def reference_arm(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cpu = cpu_reference_arm(args.steps, args.warmup, max_seconds=90.0)
    if cpu is None:
        emit(dict(impl='reference', unavailable='oracle/_ref not built (run __graft_entry__.build() where /root/reference exists)'))
        return
    line = dict(impl='reference', metric=METRIC, value=cpu['value'], unit=UNIT, n_gpus=args.gpus, steps=cpu['rounds'], warmup=args.warmup,
                ms_per_step=cpu['seconds'] * 1e3 / cpu['rounds'], higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic',
                config=dict(workload='config3: 100-residue chain, ff_1 + side-chain BP, T=0.8, dt=0.009; CPU sample of %d replicas' % cpu['n_sys']),
                us_per_force_eval=cpu['us_per_force_eval'],
                cpu_baseline={k: cpu[k] for k in ('value', 'unit', 'cores', 'kind', 'sample')},
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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == 'reference':
        reference_arm(args)
    else:
        gpu_arm(args)


if __name__ == '__main__':
    main()
