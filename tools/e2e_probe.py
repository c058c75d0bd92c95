"""time the pieces of the end-to-end step (host buffers through the C ABI)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, torch, ctypes as ct
from upside_md_b200 import upside_engine as ue
B = 4096
eng = ue.BatchEngine(bench.CONFIG, B)
eng.set_pos(bench.workload_positions(B, 0)); eng.md_init(0.8, seed=42); eng.md_run(35)
host_in = torch.from_numpy(eng.get_pos()).pin_memory(); host_out = torch.empty_like(host_in).pin_memory()
fp = ct.POINTER(ct.c_float); pin, pout = ct.cast(host_in.data_ptr(), fp), ct.cast(host_out.data_ptr(), fp)
L, e = eng.L, eng.e
for rep in range(3):
    t = [time.perf_counter()]
    L.ub_set_pos(e, pin); t.append(time.perf_counter())
    L.ub_md_run(e, 1); L.ub_sync(e); t.append(time.perf_counter())
    L.ub_get_pos(e, pout); t.append(time.perf_counter())
    print('set_pos %.2f ms  md_run+sync %.2f ms  get_pos %.2f ms' % tuple(1e3 * (b - a) for a, b in zip(t, t[1:])))
t0 = time.perf_counter(); eng.md_run(10); print('10 rounds resident: %.2f ms/round' % (1e2 * (time.perf_counter() - t0)))
