#!/usr/bin/env python
"""Generate the five BASELINE.json benchmark configurations as .up files under configs/.

Needs the reference's parameter libraries (``/root/reference/parameters``), which exist only in the build
container, so the generated files are committed and travel to the GPU box.  Sequences: i.i.d. uniform over
the 20 residue types, ``numpy.random.default_rng(1000+config_id)``; start: ``random_initial_config`` with
seed ``2000+config_id``.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from upside_md_b200 import config, h5lite  # noqa: E402

PARAM = os.environ.get('UPSIDE_REFERENCE_PARAMETERS', '/root/reference/parameters')
CONFIGS = {1: (20, False), 2: (76, False), 3: (100, False), 4: (150, False), 5: (300, True)}


def main():
    sc = h5lite.load(os.path.join(PARAM, 'ff_1/sidechain.h5'))
    env = h5lite.load(os.path.join(PARAM, 'ff_1/environment.h5'))
    hb = float(open(os.path.join(PARAM, 'ff_1/hbond')).read())
    rref = config.load_rama_reference(os.path.join(PARAM, 'common/rama_reference.pkl'))
    os.makedirs(os.path.join(ROOT, 'configs'), exist_ok=True)
    for cid, (n_res, membrane) in CONFIGS.items():
        seq = config.random_sequence(n_res, 1000 + cid)
        pos = config.random_initial_config(n_res, np.random.default_rng(2000 + cid))
        mem = config.synthetic_membrane_library() if membrane else None
        path = os.path.join(ROOT, 'configs', 'config%d_%dres.up' % (cid, n_res))
        config.write_ff1_config(path, seq, pos, sc, env, hb, rref, membrane=mem)
        print(path, os.path.getsize(path))


if __name__ == '__main__':
    main()
