"""profiles/traffic.json from an `ncu --set full` report of tools/profile_run.py (B=4096, config 3): per kernel group of
bench.py's roofline table, the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) and issue-slot utilisation per
launch, averaged over the captured launches.  usage: make_traffic.py <report.ncu-rep> <out.json> [tag]"""
import csv, io, json, subprocess, sys, collections
rep, out = sys.argv[1], sys.argv[2]
tag = sys.argv[3] if len(sys.argv) > 3 else ''
txt = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
h, units = rows[0], rows[1]
def val(d, k):
    x = float(d[k].replace(',', ''))
    u = units[h.index(k)]
    if 'bytes' in k: x *= {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    if k.startswith('gpu__time'): x *= {'ns': 1e-3, 'nsecond': 1e-3, 'us': 1, 'usecond': 1, 'ms': 1e3, 'msecond': 1e3}.get(u, 1)
    return x
GROUPS = {   # bench.py group name -> kernel-name substrings
    'rotamer pair term (k_rot_energy + k_rot_deriv)': ['k_rot_energy', 'k_rot_deriv'],
    'rotamer belief propagation (k_rot_bp2)': ['k_rot_bp2'],
    'hbond_coverage x2 (fwd + bwd)': ['k_hbond_coverage'],
    'environment_coverage (fwd + bwd)': ['k_env_coverage'],
    'protein_hbond (fwd + bwd)': ['k_protein_hbond'],
    'pair lists (k_cache_check + k_pairlist + k_refine, the four sparse graphs)': ['k_refine', 'k_pairlist', 'k_cache_check'],
    'rotamer build (k_rot_build: residue spheres, bead masks, CSR rows)': ['k_rot_build'],
}
per = collections.defaultdict(lambda: collections.defaultdict(list))
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d.get('Kernel Name', '')
    short = next((s for g in GROUPS.values() for s in g if s in name), None)
    if not short: continue
    per[short]['us'].append(val(d, 'gpu__time_duration.sum'))
    per[short]['dram'].append(val(d, 'dram__bytes_read.sum') + val(d, 'dram__bytes_write.sum'))
    per[short]['issue'].append(val(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'))
    per[short]['fma'].append(val(d, 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'))
    # FP32 flops the kernel really executed (SURVEY.md section 8(d) names these counters as the roofline numerator)
    fl = 0.0
    for k, w in (('sm__sass_thread_inst_executed_op_fadd_pred_on.sum', 1), ('sm__sass_thread_inst_executed_op_fmul_pred_on.sum', 1),
                 ('sm__sass_thread_inst_executed_op_ffma_pred_on.sum', 2)):
        if k in d and d[k] not in ('', 'n/a'):
            fl += w * float(d[k].replace(',', ''))
    per[short]['flops'].append(fl)
    if 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio' in d:
        try: per[short]['barrier'].append(float(d['smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']))
        except ValueError: pass
res = {'_source': 'ncu --set full --clock-control none, tools/profile_run.py 4096 (config 3), capture %s' % tag, '_kernels': {}}
mean = lambda v: sum(v) / len(v)
for k, m in per.items():
    res['_kernels'][k] = dict(launches=len(m['us']), us_per_launch=mean(m['us']), dram_bytes_per_launch=mean(m['dram']),
                              issue_active_pct=mean(m['issue']), fma_pipe_active_pct=mean(m['fma']),
                              fp32_flops_executed_per_launch=mean(m['flops']) if m['flops'] else None,
                              barrier_stall_warps_per_issue=mean(m['barrier']) if m['barrier'] else None)
for g, subs in GROUPS.items():
    ks = [res['_kernels'][s] for s in subs if s in res['_kernels']]
    if not ks: continue
    # the group figure is per LAUNCH of its dominant kernel (launch counts per evaluation differ between kernels)
    top = max(ks, key=lambda k: k['us_per_launch'])
    res[g] = dict(dram_bytes_per_launch=top['dram_bytes_per_launch'], issue_active_pct=top['issue_active_pct'],
                  fma_pipe_active_pct=top['fma_pipe_active_pct'], us_per_launch_under_ncu=top['us_per_launch'],
                  fp32_flops_executed_per_launch=top['fp32_flops_executed_per_launch'],
                  barrier_stall_warps_per_issue=top['barrier_stall_warps_per_issue'])
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
res['source_hash'] = bench.kernel_source_hash()     # bench.py quotes this file only for the build it was captured from
json.dump(res, open(out, 'w'), indent=1)
print(json.dumps({k: v for k, v in res.items() if not k.startswith('_')}, indent=1))
