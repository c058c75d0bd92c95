#!/usr/bin/env python
"""upside_main with a temperature ladder dealt to several GPUs of one process (UPSIDE_B200_DEVICES=D, replica exchange on the
devices over NCCL) against the same run on one GPU and against the host exchange path: identical swap history, trajectories
equal to rounding.  usage: cli_multidevice_check.py [n_device]   (run under gpurun --gpus N)"""
import json, os, shutil, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from upside_md_b200 import upside_engine as ue, h5lite
D = int(sys.argv[1]) if len(sys.argv) > 1 else 2
N = 8
T = np.geomspace(0.7, 1.0, N)
sets = [','.join('%d-%d' % (i, i + 1) for i in range(0, N - 1, 2)), ','.join('%d-%d' % (i, i + 1) for i in range(1, N - 1, 2))]
flags = ['--duration', '2.0', '--frame-interval', '0.27', '--temperature', ','.join('%.5f' % t for t in T), '--seed', '11',
         '--replica-interval', '0.135', '--log-level', 'basic']
for s in sets:
    flags += ['--swap-set', s]
res = {}
for label, env in (('one_device', {}), ('host_exchange', {'UPSIDE_B200_HOST_REPLEX': '1'}), ('%d_devices' % D, {'UPSIDE_B200_DEVICES': str(D)}),
                   ('%d_devices_host_exchange' % D, {'UPSIDE_B200_DEVICES': str(D), 'UPSIDE_B200_HOST_REPLEX': '1'})):
    with tempfile.TemporaryDirectory() as tmp:
        files = []
        for i in range(N):
            f = os.path.join(tmp, 'r%d.up' % i)
            shutil.copy(os.path.join(ROOT, 'configs', 'config1_20res.up'), f)
            files.append(f)
        for k in ('UPSIDE_B200_HOST_REPLEX', 'UPSIDE_B200_DEVICES'):
            os.environ.pop(k, None)
        os.environ.update(env)
        ue.in_process_upside(flags + files, verbose=False)
        out = [h5lite.load(f)['output'] for f in files]
        res[label] = dict(replica_index=np.array([np.asarray(o['replica_index'].data) for o in out]),
                          swaps=[np.asarray(o['replica_cumulative_swaps'].data) for o in out],
                          pos=np.array([np.asarray(o['pos'].data) for o in out]), pot=np.array([np.asarray(o['potential'].data) for o in out]))
ref = res['one_device']
report = {}
for label, r in res.items():
    if label == 'one_device':
        continue
    report[label] = dict(replica_index_identical=bool((r['replica_index'] == ref['replica_index']).all()),
                         swap_counts_identical=all((a == b).all() for a, b in zip(r['swaps'], ref['swaps'])),
                         max_pos_diff=float(np.abs(r['pos'] - ref['pos']).max()), max_potential_diff=float(np.abs(r['pot'] - ref['pot']).max()))
for label, r in res.items():
    if label != 'one_device':
        d = (r['replica_index'] != ref['replica_index']).any(axis=(0, 2)) if r['replica_index'].ndim == 3 else (r['replica_index'] != ref['replica_index']).any(axis=0)
        first = int(np.argmax(d)) if d.any() else -1
        report[label]['first_frame_with_other_indices'] = first
        report[label]['potential_diff_per_frame'] = [float(x) for x in np.abs(r['pot'] - ref['pot']).reshape(N, -1).max(axis=0)[:8]]
report['n_swaps_accepted'] = int(sum(int(s[-1, :, 0].sum()) for s in ref['swaps']) // 2)
print(json.dumps(report))
ok = all(v['replica_index_identical'] and v['swap_counts_identical'] for k, v in report.items() if isinstance(v, dict))
sys.exit(0 if ok else 1)
