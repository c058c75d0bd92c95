import csv, subprocess, sys, io
rep = sys.argv[1]
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[0]
cols = [('gpu__time_duration.sum','us',1e-3),('launch__grid_size','grid',1),('launch__block_size','blk',1),('launch__registers_per_thread','reg',1),
 ('sm__warps_active.avg.pct_of_peak_sustained_active','occ%',1),('smsp__issue_active.avg.pct_of_peak_sustained_active','issue%',1),
 ('smsp__thread_inst_executed_per_inst_executed.ratio','thr/inst',1),('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','fma%',1),
 ('dram__bytes_read.sum','rdMB',1e-6),('dram__bytes_write.sum','wrMB',1e-6),('lts__t_bytes.sum','l2MB',1e-6),
 ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','longsb',1),
 ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','shortsb',1),
 ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio','barr',1),
 ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','mio',1),
 ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','lg',1),
 ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','wait',1),
 ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','br',1),
 ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','bankconf',1e-6),('l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smwave',1e-6)]
units = rows[1]
print('%-34s' % 'kernel' + ''.join('%9s' % c[1] for c in cols))
for r in rows[2:]:
    d = dict(zip(h, r))
    name = d.get('Kernel Name','?').replace('ub::','').replace('(anonymous namespace)::','')[:33]
    s = '%-34s' % name
    for k, lab, sc in cols:
        v = d.get(k, '')
        try:
            x = float(v.replace(',', ''))
            u = units[h.index(k)]
            if k.startswith('gpu__time') and u in ('ns','nsecond'): x *= 1e-3
            elif k.startswith('gpu__time') and u in ('us','usecond'): pass
            elif k.startswith('gpu__time') and u in ('ms','msecond'): x *= 1e3
            elif 'bytes' in k:
                m = {'byte':1e-6,'Kbyte':1e-3,'Mbyte':1,'Gbyte':1e3}.get(u, 1e-6); x *= m
            elif sc != 1: x *= sc
            s += '%9.1f' % x if abs(x) < 1e5 else '%9.2e' % x
        except Exception:
            s += '%9s' % v[:8]
    print(s)
