#!/usr/bin/env python
"""configs/config10_two_chains_40res.up: a two-chain system the way upside_config.py --chain-break-from-file sets it up
(/input/chain_break/chain_first_residue, no H-bond donor / acceptor inferred on the residues next to the break:
upside_config.py:1413-1449), the bonded terms and Rama coordinates confined to their chain, plus one rigid-body jump move per chain (src/monte_carlo_sampler.cpp:174-201).  The second chain
starts 12 Angstrom away from the first.  Needs the reference's parameter libraries, so the file is committed."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from upside_md_b200 import config, h5lite  # noqa: E402

PARAM = os.environ.get('UPSIDE_REFERENCE_PARAMETERS', '/root/reference/parameters')


def main():
    sc = h5lite.load(os.path.join(PARAM, 'ff_1/sidechain.h5'))
    env = h5lite.load(os.path.join(PARAM, 'ff_1/environment.h5'))
    hb = float(open(os.path.join(PARAM, 'ff_1/hbond')).read())
    rref = config.load_rama_reference(os.path.join(PARAM, 'common/rama_reference.pkl'))
    n_chain_res = 20
    seq = config.random_sequence(2 * n_chain_res, 1010)
    rng = np.random.default_rng(2010)
    a = config.random_initial_config(n_chain_res, rng)
    b = config.random_initial_config(n_chain_res, rng)
    b = b - b.mean(0) + a.mean(0) + np.array([12., 0., 0.])
    pos = np.concatenate([a, b])
    path = os.path.join(ROOT, 'configs', 'config10_two_chains_40res.up')
    w = config.write_ff1_config(path, seq, pos, sc, env, hb, rref, chain_first_residue=[n_chain_res], split_bonded=True)
    w.write_jump_moves_per_chain(1.0, 0.3)
    w.save(path)
    print(path, os.path.getsize(path))


if __name__ == '__main__':
    main()
