"""throughput of 1 engine x 4096 replicas vs k engines x 4096/k replicas running concurrently (one stream each)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from upside_md_b200 import upside_engine as ue
B = 4096
pos = bench.workload_positions(B, 0)
for k in (1, 2, 4):
    engs = []
    for q in range(k):
        e = ue.BatchEngine(bench.CONFIG, B // k)
        e.set_pos(pos[q * (B // k):(q + 1) * (B // k)])
        e.md_init(0.8, seed=42 + q)
        engs.append(e)
    for e in engs: e.md_run(35, sync=False)
    for e in engs: e.sync()
    t0 = time.perf_counter()
    n = 20
    for i in range(n):
        for e in engs: e.md_run(1, sync=False)
    for e in engs: e.sync()
    dt = time.perf_counter() - t0
    print('engines %d x %d replicas: %.2f ms/round, %.0f replica-timesteps/s' % (k, B // k, 1e3 * dt / n, B * 3 * n / dt))
    for e in engs: e.close()
