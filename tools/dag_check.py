"""DAG-captured graph vs linear graph on the bench workload: same forces? (run on the GPU box)"""
import os, sys, subprocess, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import bench
    from upside_md_b200 import upside_engine as ue
    B = int(sys.argv[2]); rounds = int(sys.argv[3])
    pos = bench.workload_positions(B, 0)
    eng = ue.BatchEngine(bench.CONFIG, B)
    eng.set_pos(pos)
    en0, d0 = eng.evaluate()
    eng.md_init(0.8, seed=42)
    eng.md_run(rounds)
    p = eng.get_pos()
    en1, d1 = eng.evaluate()
    np.savez(sys.argv[4], en0=en0, d0=d0, p=p, en1=en1, d1=d1)
    print('ok', np.isfinite(p).all(), en0[:3], en1[:3])
else:
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    outs = []
    for tag, env in (('lin', {'UPSIDE_B200_NO_DAG': '1'}), ('dag', {}), ('dag2', {})):
        f = '/tmp/dagcheck_%s.npz' % tag
        r = subprocess.run([sys.executable, __file__, 'child', str(B), str(rounds), f], env=dict(os.environ, **env), capture_output=True, text=True)
        print(tag, r.stdout.strip()[-200:], r.stderr.strip()[-300:])
        outs.append(np.load(f) if os.path.exists(f) else None)
    a = outs[0]
    for tag, b in zip(('dag', 'dag2'), outs[1:]):
        if a is None or b is None: continue
        print(tag, 'en0 maxdiff', np.abs(a['en0'] - b['en0']).max(), 'd0 maxdiff', np.abs(a['d0'] - b['d0']).max(), 'scale', np.abs(a['d0']).max(),
              'pos maxdiff after md', np.abs(a['p'] - b['p']).max(), 'n bad replicas d0', int((np.abs(a['d0'] - b['d0']).max(axis=(1, 2)) > 1e-2 * np.abs(a['d0']).max(axis=(1,2))).sum()))
