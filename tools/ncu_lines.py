"""Aggregate an ncu SASS source page by CUDA source line (needs -lineinfo): joins `ncu --page source --csv
--print-source sass` with `nvdisasm -g` line markers of the same cubin.
usage: ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top_n] [nth-match]"""
import csv, re, subprocess, sys, collections, io
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
nth = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dis = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
addr2line = {}
cur = None; inside = False
for ln in dis.splitlines():
    if ln.startswith('.text.'):
        inside = kern in ln
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', ln)
    if m and cur: addr2line[int(m.group(1), 16)] = cur
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# the page holds one section per profiled kernel: pick the one whose 'Kernel Name' row matches
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
match = [i for i in starts if kern in rows[i][1]]
sec = match[min(nth, len(match) - 1)] if match else (starts[0] if starts else 0)
end = next((i for i in starts if i > sec), len(rows))
rows = rows[sec:end]
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
ci = {k: h.index(k) for k in ('Address', '# Samples', 'Instructions Executed', 'Thread Instructions Executed', 'stall_long_sb', 'stall_barrier', 'stall_wait', 'stall_short_sb', 'stall_math')}
agg = collections.defaultdict(lambda: collections.Counter())
base = None
for r in rows[hi + 1:]:
    if len(r) < len(h): continue
    a = int(r[ci['Address']], 16) if r[ci['Address']].startswith('0x') else int(r[ci['Address']])
    if base is None: base = a
    line = addr2line.get(a - base, ('?', 0))
    for k in ci:
        if k == 'Address': continue
        try: agg[line][k] += float(r[ci[k]])
        except ValueError: pass
tot = sum(v['# Samples'] for v in agg.values()); toti = sum(v['Instructions Executed'] for v in agg.values())
print('total samples %d, warp instr %d' % (tot, toti))
for line, v in sorted(agg.items(), key=lambda kv: -kv[1]['# Samples'])[:top]:
    print('%-16s:%4d  samples %5.1f%%  instr %5.1f%%  thr/inst %4.1f  long_sb %5.0f barrier %5.0f wait %5.0f short %5.0f math %5.0f' % (
        line[0], line[1], 100 * v['# Samples'] / tot, 100 * v['Instructions Executed'] / toti,
        v['Thread Instructions Executed'] / max(1, v['Instructions Executed']), v['stall_long_sb'], v['stall_barrier'], v['stall_wait'], v['stall_short_sb'], v['stall_math']))
