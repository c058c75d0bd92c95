#!/usr/bin/env python
"""Generate tests/golden/cli_replex.npz: the reference `upside` binary (oracle/_ref/upside_ref, the unmodified
src/main.cpp) run as a 4-rung temperature ladder with replica exchange on the 20-residue configuration, starting from
relaxed structures.  The GPU test replays the same command line through upside_main of libupside_b200.so.
Run in the build container (needs oracle/_ref)."""
import os, sys, subprocess, tempfile, shutil
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from upside_md_b200 import h5lite

ARGS = ['--duration', '0.27', '--frame-interval', '0.027', '--replica-interval', '0.054', '--swap-set', '0-1,2-3',
        '--swap-set', '1-2', '--temperature', '0.7,0.8,0.9,1.0', '--seed', '7']
KEYS = ('pos', 'kinetic', 'potential', 'time', 'temperature', 'replica_index', 'replica_swap_partner', 'replica_cumulative_swaps')


def write_inputs(dirname, start_pos):
    """4 copies of config1 whose /input/pos are the given relaxed structures"""
    paths = []
    for i, p in enumerate(start_pos):
        t = h5lite.load(os.path.join(ROOT, 'configs', 'config1_20res.up'))
        t['input/pos'].data[:, :, 0] = p
        path = os.path.join(dirname, 's%d.up' % i)
        h5lite.save(t, path)
        paths.append(path)
    return paths


def main():
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'config1.npz'))
    start = np.array([g['pos'][0], g['pos'][1], g['pos'][2], g['pos'][0]], dtype='f4')
    d = tempfile.mkdtemp()
    try:
        paths = write_inputs(d, start)
        exe = os.path.join(ROOT, 'oracle', '_ref', 'upside_ref')
        r = subprocess.run([exe] + ARGS + paths, capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS='1'))
        assert r.returncode == 0, r.stderr
        out = dict(start=start, args=np.array(ARGS))
        for i, p in enumerate(paths):
            o = h5lite.load(p)['output']
            for k in KEYS:
                out['%s_%d' % (k, i)] = np.array(o[k].data)
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'cli_replex.npz'), **out)
        print('replica_index per frame:', np.stack([out['replica_index_%d' % i].ravel() for i in range(4)], 1).tolist())
        print('cumulative swaps of system 1 at the end:', out['replica_cumulative_swaps_1'][-1].tolist())
    finally:
        shutil.rmtree(d)


if __name__ == '__main__':
    main()
