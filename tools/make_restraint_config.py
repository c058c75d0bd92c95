#!/usr/bin/env python
"""configs/config6_restraints_20res.up: BASELINE config 1 plus every restraint / plumbing node of SURVEY.md section 8(f) row 2
(atom_pos_spring, tension, AFM, cavity_radial, z_flat_bottom, contact on side-chain beads, a restraint group, and a
slice -> atom_pos_spring chain), so that the parity tests can compare those nodes with the reference engine.  Derived from
the committed config 1 file, needs no parameter library."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from upside_md_b200 import config  # noqa: E402


def main():
    w = config.ConfigWriter.from_file(os.path.join(ROOT, 'configs', 'config1_20res.up'))
    rng = np.random.default_rng(6)
    n_res = w.n_res
    w.write_cavity_radial(14.0)                                  # small enough that part of the chain is outside
    w.write_z_flat_bottom([2, 7, 11], [0.0, 3.0, -2.0], [1.0, 2.0, 0.5], [1.5, 0.8, 2.0])
    w.write_tension([0, n_res - 1], [[0.1, 0.0, -0.2], [-0.1, 0.0, 0.2]])
    w.write_AFM([1, n_res - 2], [0.05, 0.08], [[-10., 0., 0.], [10., 2., 1.]], [[-0.01, 0., 0.], [0.02, 0., 0.01]], 0.0, 0.009 * 3)
    w.write_pos_spring([4, 16, 31], w.pos[[4, 16, 31]] + rng.normal(0, 0.5, (3, 3)), [0.5, 1.0, 1.0])
    n_bead = w.potential['placement_fixed_point_vector_only/affine_residue'].shape[0]
    pairs = rng.choice(n_bead, (24, 2), replace=False)
    w.write_contact_energies(pairs, rng.uniform(-2., 1., 24), rng.uniform(4., 9., 24), rng.uniform(0.5, 3., 24),
                             argument='placement_fixed_point_vector_only')
    w.make_restraint_group([3, 4, 5, 12, 13], 0.3)
    g = w.group('slice_ends', ['pos'])                          # slice -> spring on the sliced atoms
    w.arr(g, 'id', np.array([1, 58, 1, 30], dtype='i'))          # an atom listed twice exercises the scatter
    g = w.potential.create_group('atom_pos_spring_sliced')
    g.attrs['arguments'] = np.array(['slice_ends'], dtype='S')
    w.arr(g, 'id', np.arange(4))
    w.arr(g, 'x0', w.pos[[1, 58, 1, 30]] + rng.normal(0, 1.0, (4, 3)))
    w.arr(g, 'spring_const', np.array([0.3, 0.6, 0.2, 0.4]))
    path = os.path.join(ROOT, 'configs', 'config6_restraints_20res.up')
    w.save(path)
    print(path, os.path.getsize(path))

    # config 7: constant + concat (the reference's Concat cannot be constructed - it reads its own member before it is
    # initialised, bonds.cpp:633-636 - so these two are checked by finite differences, not against the oracle)
    w = config.ConfigWriter.from_file(os.path.join(ROOT, 'configs', 'config1_20res.up'))
    g = w.group('slice_some', ['pos'])
    w.arr(g, 'id', np.array([4, 22, 40, 57], dtype='i'))
    g = w.group('constant_anchor', [])
    w.arr(g, 'value', w.pos[[10, 47]] + np.array([[3., 0., 0.], [0., -2., 1.]]))
    w.group('concat_points', ['slice_some', 'constant_anchor'])
    g = w.group('dist_spring_tethers', ['concat_points'])          # slice rows 0..3, anchors 4..5
    w.arr(g, 'id', np.array([[0, 4], [1, 4], [2, 5], [3, 5], [0, 3]], dtype='i'))
    w.arr(g, 'equil_dist', np.array([8., 10., 9., 7., 12.]))
    w.arr(g, 'spring_const', np.array([0.5, 0.4, 0.6, 0.3, 0.2]))
    w.arr(g, 'bonded_atoms', np.zeros(5, dtype='i'))
    path = os.path.join(ROOT, 'configs', 'config7_concat_20res.up')
    w.save(path)
    print(path, os.path.getsize(path))

    # config 8: radial + hbond_sc_radial pair nodes (src/sidechain_radial.cpp:16-136) on top of config 1.  Synthetic tables
    # that obey RadialHelper's four conditions (p[0] = 1/dx, p[1] == p[3], p[-3] == p[-1], zero at the cutoff)
    w = config.ConfigWriter.from_file(os.path.join(ROOT, 'configs', 'config1_20res.up'))
    names = list(config.restypes)

    def radial_table(n1, n2, symmetric, seed):
        r = np.random.default_rng(seed)
        p = np.zeros((n1, n2, 17), dtype='f4')
        for a in range(n1):
            for b in range(a if symmetric else 0, n2):
                knots = np.zeros(16)
                depth, width = r.uniform(0.2, 1.2), r.uniform(1.5, 3.0)
                x = np.arange(16) - 1.0
                knots[:] = 3.0 * np.exp(-x / 1.5) - depth * np.exp(-((x - 9.0) / width) ** 2)
                knots[-3:] = 0.0                      # continuity at the cutoff and terminal clamp
                knots[-4] *= 0.3
                knots[0] = knots[2]                   # origin clamp
                inv_dx = r.choice([2.0, 2.5])         # cutoffs of 7 and 5.6 Angstrom: the graph takes the largest
                p[a, b, 0] = inv_dx
                p[a, b, 1:] = knots
                if symmetric:
                    p[b, a] = p[a, b]
        return p
    w.write_sidechain_radial(radial_table(20, 20, True, 81), names)
    w.write_hbond_sc_radial(radial_table(2, 20, False, 82), names)
    path = os.path.join(ROOT, 'configs', 'config8_radial_20res.up')
    w.save(path)
    print(path, os.path.getsize(path))

    # config 9: uniform_transform and both linear_coupling nodes (src/environment.cpp:158-321) on top of config 1
    w = config.ConfigWriter.from_file(os.path.join(ROOT, 'configs', 'config1_20res.up'))
    r = np.random.default_rng(9)
    types = np.array([names.index(x) for x in w.fasta], dtype='i')
    w.write_linear_coupling('environment_coverage', r.uniform(-0.3, 0.3, 20), types)
    w.write_linear_coupling('environment_coverage', r.uniform(-0.3, 0.3, 20), types, inactivation='placement_fixed_point_vector_only_CB',
                            inactivation_dim=3)
    w.write_uniform_transform('environment_coverage', r.uniform(-1., 1., 12), -0.5, 1.5)
    w.write_linear_coupling('uniform_transform', r.uniform(-0.5, 0.5, 20), types, suffix='_transformed')
    path = os.path.join(ROOT, 'configs', 'config9_coupling_20res.up')
    w.save(path)
    print(path, os.path.getsize(path))


if __name__ == '__main__':
    main()
