"""Python-3 restatement of the parts of the reference's ``py/upside_config.py`` (Python 2 + PyTables,
not runnable in this image) that emit the ff_1 force field, written on top of :mod:`h5lite`.

It produces ``.up`` files with exactly the ``/input`` schema the reference engine's node constructors
read, for the invocation documented in the reference ``README.md:143-153``::

    --hbond-energy $(cat ff_1/hbond) --dynamic-rotamer-1body --rotamer-placement ff_1/sidechain.h5
    --rotamer-interaction ff_1/sidechain.h5 --environment ff_1/environment.h5
    --rama-library <synthetic> --reference-state-rama common/rama_reference.pkl  [--membrane-potential <synthetic>]

Each ``write_*`` function cites the reference function it follows.  The real Rama library
(``parameters/common/rama.dat``) is absent from the reference tree, and no membrane library ships at all, so
:func:`synthetic_rama_pot` / :func:`synthetic_membrane_library` provide smooth stand-ins.
This is configuration tooling: nothing here is on the timed path.
"""
import pickle

import numpy as np

from . import h5lite

deg = np.deg2rad(1)
n_bit_rotamer = 4

three_letter_aa = dict(A='ALA', C='CYS', D='ASP', E='GLU', F='PHE', G='GLY', H='HIS', I='ILE', K='LYS', L='LEU',
                       M='MET', N='ASN', P='PRO', Q='GLN', R='ARG', S='SER', T='THR', V='VAL', W='TRP', Y='TYR')
restypes = sorted(three_letter_aa.values())

_BB_REF = np.array([(-1.19280531, -0.83127186, 0.),          # N
                    (0., 0., 0.),                            # CA
                    (1.25222632, -0.87268266, 0.),           # C
                    (0., 0.94375626, 1.2068012)])            # CB


def _s(x):
    return x.decode() if isinstance(x, bytes) else str(x)


def vmag(x):
    return np.sqrt((x ** 2).sum(axis=-1))


# ------------------------------------------------------------------------------------------------
# initial structures (upside_config.py:414-476)
# ------------------------------------------------------------------------------------------------

def make_tab_matrices(phi, theta, bond_length):
    """torsion-angle-bond affine matrices, upside_config.py:414-432"""
    phi, theta, l = np.asarray(phi), np.asarray(theta), np.asarray(bond_length)
    r = np.zeros(phi.shape + (4, 4))
    cp, sp, ct, st = np.cos(phi), np.sin(phi), np.cos(theta), np.sin(theta)
    r[..., 0, 0] = -ct;     r[..., 0, 1] = -st;      r[..., 0, 2] = 0;   r[..., 0, 3] = -l * ct
    r[..., 1, 0] = cp * st; r[..., 1, 1] = -cp * ct; r[..., 1, 2] = -sp; r[..., 1, 3] = l * cp * st
    r[..., 2, 0] = sp * st; r[..., 2, 1] = -sp * ct; r[..., 2, 2] = cp;  r[..., 2, 3] = l * sp * st
    r[..., 3, 3] = 1
    return r


def construct_equilibrium_structure(rama, angles, bond_lengths):
    """upside_config.py:435-458"""
    n_res = rama.shape[0]
    t = np.zeros(3 * n_res)
    t[3::3] = rama[:-1, 1]
    t[4::3] = rama[:-1, 2]
    t[5::3] = rama[1:, 0]
    transforms = make_tab_matrices(t, angles.ravel(), bond_lengths.ravel())
    cur = np.eye(4)
    pos = np.zeros((3 * n_res, 3))
    for i, mat in enumerate(transforms):
        cur = cur @ mat
        pos[i] = cur[:3, 3]
    return pos


def _structure_from_rama(rama):
    angles = np.zeros_like(rama)
    lengths = np.zeros_like(rama)
    angles[:, 0] = 120.0 * deg
    angles[:, 1] = 120.0 * deg
    angles[:, 2] = 109.5 * deg
    lengths[:, 0] = 1.453
    lengths[:, 1] = 1.526
    lengths[:, 2] = 1.300
    return construct_equilibrium_structure(rama, angles, lengths)


def random_initial_config(n_res, rng):
    """upside_config.py:461-476 with an explicit numpy Generator instead of the global RandomState"""
    rama = rng.random((n_res, 3)) * 2 * np.pi - np.pi
    rama[:, 2] = np.pi
    return _structure_from_rama(rama)


def extended_initial_config(n_res, rng, noise=0.05):
    """phi=-120, psi=+130 strand perturbed by N(0, noise) Angstrom (SURVEY.md §8(d) second input set)"""
    rama = np.zeros((n_res, 3))
    rama[:, 0] = -120 * deg
    rama[:, 1] = 130 * deg
    rama[:, 2] = np.pi
    return _structure_from_rama(rama) + rng.normal(0., noise, (3 * n_res, 3))


def random_sequence(n_res, seed):
    rng = np.random.default_rng(seed)
    return np.array([restypes[i] for i in rng.integers(0, 20, n_res)])


# ------------------------------------------------------------------------------------------------
# synthetic libraries (the real ones are not in the reference tree)
# ------------------------------------------------------------------------------------------------

def basin_cond_prob_fcns(a_phi, a_psi):
    """upside_config.py:528-564"""
    def basin_box(phi0, phi1, psi0, psi1):
        if phi0 > phi1: phi1 += 2 * np.pi
        if psi0 > psi1: psi1 += 2 * np.pi
        phi_mid, psi_mid = 0.5 * (phi1 + phi0), 0.5 * (psi1 + psi0)
        phi_switch, psi_switch = np.cos(phi1 - phi_mid), np.cos(psi1 - psi_mid)

        def f(phi, psi):
            dphi, dpsi = np.cos(phi - phi_mid), np.cos(psi - psi_mid)
            return 1. / ((1. + np.exp(-a_phi * (dphi - phi_switch))) * (1. + np.exp(-a_psi * (dpsi - psi_switch))))
        return f
    bb = lambda a, b, c, d: basin_box(a * deg, b * deg, c * deg, d * deg)
    return [bb(-180., 0., -100., 50.), bb(-180., -100., 50., -100.), bb(-100., 0., 50., -100.),
            bb(0., 180., -50., 100.), bb(0., 180., 100., -50.)]


def synthetic_rama_pot(seq, n_bin=36):
    """Per-residue (n_bin x n_bin) Rama potential = -log of a residue-type dependent mixture of the five
    basins of upside_config.py:528-564, normalised like read_rama_maps_and_weights (:627)."""
    phi = np.linspace(-np.pi, np.pi, n_bin, endpoint=False)[:, None]
    psi = np.linspace(-np.pi, np.pi, n_bin, endpoint=False)[None, :]
    basins = np.array([f(phi, psi) for f in basin_cond_prob_fcns(6., 6.)])
    pots = np.zeros((len(seq), n_bin, n_bin), dtype='f4')
    for i, aa in enumerate(seq):
        k = restypes.index(aa)
        w = np.array([0.45, 0.25, 0.20, 0.07, 0.03])
        if aa == 'GLY': w = np.array([0.25, 0.15, 0.15, 0.25, 0.20])
        if aa == 'PRO': w = np.array([0.35, 0.05, 0.58, 0.01, 0.01])
        w = w * (1. + 0.3 * np.cos(np.arange(5) * 1.7 + k))     # mild type dependence
        w /= w.sum()
        p = (w[:, None, None] * basins).sum(axis=0) + 1e-4
        pots[i] = -np.log(p)
    pots -= -np.log(np.exp(-1.0 * pots).sum(axis=(-2, -1), keepdims=True))
    return pots


def synthetic_membrane_library(n_z=61, thickness=30.):
    """Stand-in for the (unshipped) membrane library: schema of upside_config.py:1044-1054."""
    z = np.linspace(-thickness / 2 - 15., thickness / 2 + 15., n_z)
    names = restypes + ['NON']
    hydrophobicity = dict(ALA=0.3, ARG=-1.5, ASN=-0.8, ASP=-1.6, CYS=0.4, GLN=-0.8, GLU=-1.5, GLY=0.0, HIS=-0.6,
                          ILE=1.2, LEU=1.2, LYS=-1.4, MET=0.8, PHE=1.3, PRO=-0.2, SER=-0.3, THR=-0.1, TRP=0.9,
                          TYR=0.4, VAL=1.0, NON=0.0)
    inside = 0.5 * (np.tanh((z + thickness / 2) / 2.5) - np.tanh((z - thickness / 2) / 2.5))
    cb_energy = np.array([-hydrophobicity[n] * inside for n in names])
    uhb_energy = np.array([1.2 * inside, 1.0 * inside])
    return dict(names=names, cb_energy=cb_energy, uhb_energy=uhb_energy, z_min=z[0], z_max=z[-1],
                thickness=thickness, cov_midpoint=np.full(len(names), 5.0), cov_sharpness=np.full(len(names), 0.5))


# ------------------------------------------------------------------------------------------------
# node writers
# ------------------------------------------------------------------------------------------------

class ConfigWriter:
    def __init__(self, fasta, pos, compress=True):
        self.fasta = np.array([_s(x) for x in fasta])
        self.n_res = len(self.fasta)
        self.n_atom = 3 * self.n_res
        self.pos = np.asarray(pos, dtype='f8').reshape(self.n_atom, 3)
        self.root = h5lite.File()
        self.compress = compress
        inp = self.root.create_group('input')
        self.arr(inp, 'sequence', self.fasta)
        p = np.zeros((self.n_atom, 3, 1), dtype='f4')
        p[:, :, 0] = pos
        self.arr(inp, 'pos', p)
        self.potential = inp.create_group('potential')

    def arr(self, grp, name, obj):
        """create_array (upside_config.py:33-34): chunked, zlib-5 + shuffle + fletcher32 like PyTables EArrays"""
        a = np.asarray(obj)
        if a.dtype.kind == 'U':
            a = np.char.encode(a, 'ascii')
        chunks = None
        if self.compress and a.size and a.ndim >= 1:
            rows = max(1, min(a.shape[0], (1 << 18) // max(1, a.dtype.itemsize * int(np.prod(a.shape[1:])))))
            if -(-a.shape[0] // rows) <= 64:
                chunks = (rows,) + a.shape[1:]
        return grp.create_dataset(name, a, chunks=chunks, compress=chunks is not None)

    def group(self, name, arguments):
        g = self.potential.create_group(name)
        g.attrs['arguments'] = np.array(arguments, dtype='S')
        return g

    # -- restraints (upside_config.py:37-160, 383-415, 814-853); tables are passed as arrays instead of text files ------
    def write_cavity_radial(self, cavity_radius, spring_constant=5.):
        g = self.group('cavity_radial', ['pos'])
        self.arr(g, 'id', np.arange(self.n_atom))
        self.arr(g, 'radius', np.ones(self.n_atom) * cavity_radius)
        self.arr(g, 'spring_constant', np.ones(self.n_atom) * spring_constant)

    def _ca(self, residues):
        residues = np.asarray(residues, dtype='i')
        if residues.size and not (0 <= residues.min() and residues.max() < len(self.fasta)):
            raise ValueError('restraint specified for a residue outside the FASTA (zero is first residue)')
        return residues * 3 + 1   # restrain the CA atom in each residue

    def write_z_flat_bottom(self, residues, z0, radius, spring_constant):
        g = self.group('z_flat_bottom', ['pos'])
        self.arr(g, 'atom', self._ca(residues))
        self.arr(g, 'z0', np.asarray(z0, dtype='f8'))
        self.arr(g, 'radius', np.asarray(radius, dtype='f8'))
        self.arr(g, 'spring_constant', np.asarray(spring_constant, dtype='f8'))

    def write_tension(self, residues, tension):
        g = self.group('tension', ['pos'])
        self.arr(g, 'atom', self._ca(residues))
        self.arr(g, 'tension_coeff', np.asarray(tension, dtype='f8').reshape(-1, 3))

    def write_AFM(self, residues, spring_const, starting_tip_pos, pulling_vel, time_initial, time_step):
        g = self.group('AFM', ['pos'])
        self.arr(g, 'atom', self._ca(residues))
        self.arr(g, 'spring_const', np.asarray(spring_const, dtype='f8'))
        self.arr(g, 'starting_tip_pos', np.asarray(starting_tip_pos, dtype='f8').reshape(-1, 3))
        d = self.arr(g, 'pulling_vel', np.asarray(pulling_vel, dtype='f8').reshape(-1, 3))
        d.attrs['time_initial'] = np.float64(time_initial)
        d.attrs['time_step'] = np.float64(time_step)

    def write_pos_spring(self, atoms, x0, spring_const):
        """atom_pos_spring (bonds.cpp:9-50); the reference's Python has no writer for it"""
        g = self.group('atom_pos_spring', ['pos'])
        self.arr(g, 'id', np.asarray(atoms, dtype='i'))
        self.arr(g, 'x0', np.asarray(x0, dtype='f8').reshape(-1, 3))
        self.arr(g, 'spring_const', np.asarray(spring_const, dtype='f8'))

    def write_contact_energies(self, pairs, energy, distance, width, argument='placement_fixed_point_only_CB'):
        if np.min(width) <= 0.:
            raise ValueError('Cannot have negative contact transition_width')
        g = self.group('contact', [argument])
        self.arr(g, 'id', np.asarray(pairs, dtype='i').reshape(-1, 2))
        self.arr(g, 'energy', np.asarray(energy, dtype='f8'))
        self.arr(g, 'distance', np.asarray(distance, dtype='f8'))
        self.arr(g, 'width', np.asarray(width, dtype='f8'))

    def make_restraint_group(self, residues, strength):
        """random intra-group springs appended to dist_spring (upside_config.py:383-415; same pairing rule, numpy Generator
        instead of the legacy global seed)"""
        rng = np.random.default_rng(314159)
        g = self.potential['dist_spring']
        old = {k: np.array(g[k].data) for k in ('id', 'equil_dist', 'spring_const', 'bonded_atoms')}
        r_atoms = np.array([(3 * i, 3 * i + 1, 3 * i + 2) for i in sorted(residues)]).reshape(-1)
        pairs = np.concatenate([np.column_stack((r_atoms, rng.permutation(r_atoms))) for _ in range(2)], axis=0)
        pairs = np.array(sorted(set((min(x, y), max(x, y)) for x, y in pairs if x // 3 != y // 3)))
        dist = np.sqrt(((self.pos[pairs[:, 0]] - self.pos[pairs[:, 1]]) ** 2).sum(axis=-1))
        for k in old:
            del g.children[k]
        self.arr(g, 'id', np.concatenate((old['id'], pairs), axis=0))
        self.arr(g, 'equil_dist', np.concatenate((old['equil_dist'], dist), axis=0))
        self.arr(g, 'spring_const', np.concatenate((old['spring_const'], strength * np.ones(len(pairs))), axis=0))
        self.arr(g, 'bonded_atoms', np.concatenate((old['bonded_atoms'], np.zeros(len(pairs), dtype='int')), axis=0))

    # -- bonded (upside_config.py:480-525) -------------------------------------------------------
    def _within_chain(self, atom_ids):
        """rows of an atom-index table that stay inside one chain (all rows unless write_chain_break was asked to split the
        bonded terms)"""
        ids = np.asarray(atom_ids)
        if not getattr(self, 'split_bonded', False):
            return np.ones(len(ids), dtype=bool)
        chain = np.searchsorted(self.chain_first_residue, ids // 3, side='right')
        return (chain == chain[:, :1]).all(axis=1)

    def write_dist_spring(self, bond_stiffness=48.):
        g = self.group('dist_spring', ['pos'])
        i = np.arange(self.n_atom - 1)
        eq = np.zeros(len(i))
        eq[0::3], eq[1::3], eq[2::3] = 1.453, 1.526, 1.300
        ids = np.column_stack((i, i + 1))
        keep = self._within_chain(ids)
        self.arr(g, 'id', ids[keep])
        self.arr(g, 'equil_dist', eq[keep])
        self.arr(g, 'spring_const', bond_stiffness * np.ones(keep.sum()))
        self.arr(g, 'bonded_atoms', np.ones(keep.sum(), dtype='int'))

    def write_angle_spring(self, angle_stiffness=175.):
        g = self.group('angle_spring', ['pos'])
        i = np.arange(self.n_atom - 2)
        eq = np.zeros(len(i))
        eq[0::3], eq[1::3], eq[2::3] = np.cos(109.5 * deg), np.cos(120.0 * deg), np.cos(120.0 * deg)
        ids = np.column_stack((i, i + 2, i + 1))
        keep = self._within_chain(ids)
        self.arr(g, 'id', ids[keep])
        self.arr(g, 'equil_dist', eq[keep])
        self.arr(g, 'spring_const', angle_stiffness * np.ones(keep.sum()))

    def write_dihedral_spring(self):
        g = self.group('dihedral_spring', ['pos'])
        i = np.arange(1, self.n_atom - 3, 3)
        ids = np.column_stack((i, i + 1, i + 2, i + 3))
        keep = self._within_chain(ids)
        self.arr(g, 'id', ids[keep])
        self.arr(g, 'equil_dist', np.where(self.fasta[1:] == 'CPR', 0. * deg, 180. * deg)[keep])
        self.arr(g, 'spring_const', 30.0 * np.ones(keep.sum()))

    # -- coordinate nodes -------------------------------------------------------------------------
    def write_rama_coord(self):
        """upside_config.py:855-863"""
        g = self.group('rama_coord', ['pos'])
        n = 3 * np.arange(self.n_res)
        idx = np.column_stack((n - 1, n, n + 1, n + 2, n + 3))
        idx[idx >= self.n_atom] = -1
        if getattr(self, 'split_bonded', False):   # phi / psi of a chain end have no partner atom, as at the termini
            res = np.arange(self.n_res)[:, None]
            chain = np.searchsorted(self.chain_first_residue, np.where(idx >= 0, idx // 3, res), side='right')
            idx[chain != np.searchsorted(self.chain_first_residue, res, side='right')] = -1
        self.arr(g, 'id', idx)

    def write_affine_alignment(self):
        """upside_config.py:168-184"""
        g = self.group('affine_alignment', ['pos'])
        ref = np.zeros((self.n_res, 3, 3))
        ref[:] = _BB_REF[:3]
        ref -= ref.mean(axis=1)[:, None]
        n = np.arange(self.n_res) * 3
        self.arr(g, 'atoms', np.column_stack((n, n + 1, n + 2)))
        self.arr(g, 'ref_geom', ref)

    def write_backbone_pair(self):
        """upside_config.py:149-165"""
        g = self.group('backbone_pairs', ['affine_alignment'])
        ref = np.zeros((self.n_res, 4, 3))
        ref[:] = _BB_REF
        ref[self.fasta == 'GLY', 3] = np.nan
        ref -= ref[:, :3].mean(axis=1)[:, None]
        self.arr(g, 'id', np.arange(self.n_res))
        self.arr(g, 'ref_pos', ref)
        self.arr(g, 'n_atom', np.isfinite(ref.sum(axis=-1)).sum(axis=-1))

    def write_infer_H_O(self, excluded=()):
        """upside_config.py:187-212"""
        n_res, fasta = self.n_res, self.fasta
        don = np.array([i for i in range(n_res) if i > 0 and i not in excluded and fasta[i] != 'PRO'], dtype='i8')
        acc = np.array([i for i in range(n_res) if i < n_res - 1 and i not in excluded], dtype='i8')
        g = self.group('infer_H_O', ['pos'])
        d, a = g.create_group('donors'), g.create_group('acceptors')
        self.arr(d, 'residue', don)
        self.arr(a, 'residue', acc)
        self.arr(d, 'bond_length', 0.88 * np.ones(len(don)))
        self.arr(a, 'bond_length', 1.24 * np.ones(len(acc)))
        self.arr(d, 'id', np.array((-1, 0, 1))[None, :] + 3 * don[:, None])
        self.arr(a, 'id', np.array((1, 2, 3))[None, :] + 3 * acc[:, None])
        self.donor_residues, self.acceptor_residues = don, acc

    def write_rotamer_placement(self, lib):
        """upside_config.py:885-1006, fixed placement + --dynamic-rotamer-1body, no --fix-rotamer"""
        restype_num = dict((_s(aa), i) for i, aa in enumerate(lib['restype_order'].data))
        placement_pos = lib['rotamer_center_fixed'].data
        placement_energy = -np.log(lib['rotamer_prob'].data.transpose((2, 0, 1)))[..., None]
        start_stop = lib['rotamer_start_stop_bead'].data
        rama_residue, affine_residue, layer_index, beadtype_seq, id_seq = [], [], [], [], []
        count_by_n_rot = dict()
        for rnum, aa in enumerate(self.fasta):
            start, stop, n_bead = (int(x) for x in start_stop[restype_num[aa]])
            assert (stop - start) % n_bead == 0
            n_rot = (stop - start) // n_bead
            count_by_n_rot.setdefault(n_rot, 0)
            base_id = (count_by_n_rot[n_rot] << n_bit_rotamer) + n_rot
            count_by_n_rot[n_rot] += 1
            rama_residue.extend([rnum] * (stop - start))
            affine_residue.extend([rnum] * (stop - start))
            layer_index.extend(range(start, stop))
            beadtype_seq.extend(['%s_%i' % (aa, i) for i in range(n_bead)] * n_rot)
            id_seq.extend(np.arange(stop - start) // n_bead + (base_id << n_bit_rotamer))
        sc, pl = 'placement_fixed_point_vector_only', 'placement_scalar'
        g = self.group(sc, ['affine_alignment'])
        self.arr(g, 'rama_residue', rama_residue)
        self.arr(g, 'affine_residue', affine_residue)
        self.arr(g, 'layer_index', layer_index)
        self.arr(g, 'placement_data', placement_pos[..., :6])
        self.arr(g, 'beadtype_seq', np.array(beadtype_seq))
        self.arr(g, 'id_seq', np.array(id_seq))
        g = self.group(pl, ['affine_alignment', 'rama_coord'])
        self.arr(g, 'rama_residue', rama_residue)
        self.arr(g, 'affine_residue', affine_residue)
        self.arr(g, 'layer_index', layer_index)
        self.arr(g, 'placement_data', placement_energy)
        self.sc_node, self.pl_node = sc, pl
        self.beadtype_seq = np.array(beadtype_seq)
        self.sc_resnum = np.array(affine_residue)
        self.id_seq = np.array(id_seq)
        return sc, pl

    def write_count_hbond(self, hbond_energy, lib, loose_hbond=False):
        """upside_config.py:295-380"""
        n_res = self.n_res
        nd, na = len(self.donor_residues), len(self.acceptor_residues)
        g = self.group('protein_hbond', ['infer_H_O'])
        self.arr(g, 'index1', np.arange(0, nd))
        self.arr(g, 'type1', np.zeros(nd, dtype='i'))
        self.arr(g, 'id1', self.donor_residues)
        self.arr(g, 'index2', np.arange(nd, nd + na))
        self.arr(g, 'type2', np.zeros(na, dtype='i'))
        self.arr(g, 'id2', self.acceptor_residues)
        self.arr(g, 'interaction_param', np.array([[[(0.5 if loose_hbond else 1.4), 1. / 0.10,
                                                     (3.1 if loose_hbond else 2.5), 1. / 0.125,
                                                     (0.182 if loose_hbond else 0.682), 1. / 0.05, 0., 0.]]]))
        bead_num = dict((_s(k), i) for i, k in enumerate(lib['bead_order'].data))
        c = self.group('hbond_coverage', ['protein_hbond', self.sc_node])
        self.arr(c, 'interaction_param', lib['coverage_interaction'].data)
        self.arr(c, 'index1', np.arange(nd + na))
        self.arr(c, 'type1', 1 * (np.arange(nd + na) >= nd))
        self.arr(c, 'id1', np.concatenate([self.donor_residues, self.acceptor_residues]))
        self.arr(c, 'index2', np.arange(len(self.beadtype_seq)))
        self.arr(c, 'type2', np.array([bead_num[s] for s in self.beadtype_seq]))
        self.arr(c, 'id2', self.sc_resnum)
        g = self.group('placement_fixed_point_vector_scalar', ['affine_alignment'])
        self.arr(g, 'affine_residue', np.arange(3 * n_res) // 3)
        self.arr(g, 'layer_index', np.arange(3 * n_res) % 3)
        self.arr(g, 'placement_data', lib['hydrophobe_placement'].data)
        c = self.group('hbond_coverage_hydrophobe', ['placement_fixed_point_vector_scalar', self.sc_node])
        self.arr(c, 'interaction_param', lib['hydrophobe_interaction'].data)
        self.arr(c, 'index1', np.arange(3 * n_res))
        self.arr(c, 'type1', np.arange(3 * n_res) % 3)
        self.arr(c, 'id1', np.arange(3 * n_res) // 3)
        self.arr(c, 'index2', np.arange(len(self.beadtype_seq)))
        self.arr(c, 'type2', np.array([bead_num[s] for s in self.beadtype_seq]))
        self.arr(c, 'id2', self.sc_resnum)
        g = self.group('hbond_energy', ['protein_hbond'])
        g.attrs['protein_hbond_energy'] = float(hbond_energy)

    def write_environment(self, lib):
        """upside_config.py:215-292"""
        restype_order = dict((_s(x), i) for i, x in enumerate(lib['restype_order'].data))
        coverage_param = lib['coverage_param'].data
        assert coverage_param.shape == (len(restype_order), 1, 4)
        n_res = self.n_res
        p = self.group('placement_fixed_point_vector_only_CB', ['affine_alignment'])
        ref = _BB_REF.copy()
        ref -= ref.mean(axis=0, keepdims=True)      # sic: includes CB in the mean, as the reference does
        pd = np.zeros((1, 6))
        pd[0, 0:3] = ref[3]
        pd[0, 3:6] = (ref[3] - ref[2]) / vmag(ref[3] - ref[2])
        self.arr(p, 'affine_residue', np.arange(n_res))
        self.arr(p, 'layer_index', np.zeros(n_res, dtype='i'))
        self.arr(p, 'placement_data', pd)
        n_sc = len(self.sc_resnum)
        w = self.group('weighted_pos', [self.sc_node, self.pl_node])
        self.arr(w, 'index_pos', np.arange(n_sc))
        self.arr(w, 'index_weight', np.arange(n_sc))
        c = self.group('environment_coverage', ['placement_fixed_point_vector_only_CB', 'weighted_pos'])
        self.arr(c, 'index1', np.arange(n_res))
        self.arr(c, 'type1', np.array([restype_order[s] for s in self.fasta]))
        self.arr(c, 'id1', np.arange(n_res))
        self.arr(c, 'index2', np.arange(n_sc))
        self.arr(c, 'type2', 0 * np.arange(n_sc))
        self.arr(c, 'id2', self.sc_resnum)
        self.arr(c, 'interaction_param', coverage_param)
        e = self.group('nonlinear_coupling_environment', ['environment_coverage'])
        d = self.arr(e, 'coeff', lib['energies'].data)
        d.attrs['spline_offset'] = float(lib['energies'].attrs['offset'])
        d.attrs['spline_inv_dx'] = float(lib['energies'].attrs['inv_dx'])
        self.arr(e, 'coupling_types', np.array([restype_order[s] for s in self.fasta]))

    def write_rama_map_pot(self, rama_pot):
        """upside_config.py:692-734 (library reading replaced by a caller-supplied per-residue map)"""
        g = self.group('rama_map_pot', ['rama_coord'])
        rama_pot = np.array(rama_pot, dtype='f4')
        rama_pot -= (rama_pot * np.exp(-rama_pot)).sum(axis=(-2, -1), keepdims=True)
        self.arr(g, 'residue_id', np.arange(self.n_res))
        self.arr(g, 'rama_map_id', np.arange(rama_pot.shape[0]))
        self.arr(g, 'rama_pot', rama_pot)
        self.rama_pot = rama_pot

    def write_reference_state_rama(self, ref_table):
        """upside_config.py:1446-1457"""
        cor = np.log(ref_table)
        cor -= cor.mean()
        g = self.group('rama_map_pot_ref', ['rama_coord'])
        g.attrs['log_pot'] = 0
        self.arr(g, 'residue_id', np.arange(self.n_res))
        self.arr(g, 'rama_map_id', np.zeros(self.n_res, dtype='i4'))
        self.arr(g, 'rama_pot', cor[None])

    def write_rotamer(self, lib, damping=0.4):
        """upside_config.py:1009-1035"""
        g = self.potential.create_group('rotamer')
        args = [self.sc_node, self.pl_node]
        for nm in ('hbond_coverage', 'hbond_coverage_hydrophobe'):
            if nm in self.potential.children:
                args.append(nm)
        g.attrs['arguments'] = np.array(args, dtype='S')
        g.attrs['max_iter'] = 1000
        g.attrs['tol'] = 1e-3
        g.attrs['damping'] = float(damping)
        g.attrs['iteration_chunk_size'] = 2
        pg = g.create_group('pair_interaction')
        bead_num = dict((_s(k), i) for i, k in enumerate(lib['bead_order'].data))
        self.arr(pg, 'interaction_param', lib['pair_interaction'].data)
        self.arr(pg, 'index', np.arange(len(self.beadtype_seq)))
        self.arr(pg, 'type', np.array([bead_num[s] for s in self.beadtype_seq]))
        self.arr(pg, 'id', self.id_seq)

    def write_CB(self):
        """upside_config.py:795-811"""
        p = self.group('placement_fixed_point_only_CB', ['affine_alignment'])
        ref = _BB_REF.copy()
        ref -= ref[:3].mean(axis=0, keepdims=True)
        self.arr(p, 'affine_residue', np.arange(self.n_res))
        self.arr(p, 'layer_index', np.zeros(self.n_res, dtype='i'))
        self.arr(p, 'placement_data', ref[3][None, :])

    def write_membrane_potential(self, lib, membrane_thickness, exclude=(), hbond_exclude=()):
        """upside_config.py:1038-1149; the scipy InterpolatedUnivariateSpline resampling is replaced by
        numpy linear interpolation of the (already smooth, synthetic) library profiles."""
        g = self.group('membrane_potential', ['placement_fixed_point_only_CB', 'environment_coverage', 'protein_hbond'])
        n_res, fasta = self.n_res, self.fasta
        don = np.array([i for i in range(n_res) if i > 0 and i not in hbond_exclude and fasta[i] != 'PRO'])
        acc = np.array([i for i in range(n_res) if i < n_res - 1 and i not in hbond_exclude])
        default_half, half = lib['thickness'] / 2., membrane_thickness / 2.
        z_ = np.linspace(-half - 15., half + 15., int((membrane_thickness + 30.) / 0.25) + 1)

        def resample(table, z_min, z_max):
            zl = np.linspace(z_min, z_max, table.shape[-1])
            out = np.zeros((len(table), len(z_)))
            for i, y in enumerate(table):
                sp = lambda x, y=y: np.interp(x, zl, y)
                if half < default_half:
                    dt = default_half - half
                    ds = sp(dt) - sp(-dt)
                    out[i] = np.where(z_ < 0, sp(z_ - dt) + 0.5 * ds, sp(z_ + dt) - 0.5 * ds)
                elif half > default_half:
                    dt = half - default_half
                    out[i] = np.select([z_ < -dt, (z_ >= -dt) & (z_ <= dt), z_ > dt], [sp(z_ + dt), sp(0.), sp(z_ - dt)])
                else:
                    out[i] = sp(z_)
            return out
        cb = resample(lib['cb_energy'], lib['z_min'], lib['z_max'])
        uhb = resample(lib['uhb_energy'], lib['z_min'], lib['z_max'])
        seq = list(fasta)
        for num in exclude:
            seq[num] = 'NON'
        r2n = dict((aa, i) for i, aa in enumerate(lib['names']))
        self.arr(g, 'cb_index', np.arange(n_res))
        self.arr(g, 'env_index', np.arange(n_res))
        self.arr(g, 'residue_type', np.array([r2n[aa] for aa in seq]))
        self.arr(g, 'cov_midpoint', lib['cov_midpoint'])
        self.arr(g, 'cov_sharpness', lib['cov_sharpness'])
        d = self.arr(g, 'cb_energy', cb)
        d.attrs['z_min'], d.attrs['z_max'] = float(z_[0]), float(z_[-1])
        d = self.arr(g, 'uhb_energy', uhb)
        d.attrs['z_min'], d.attrs['z_max'] = float(z_[0]), float(z_[-1])
        self.arr(g, 'donor_residue_ids', don)
        self.arr(g, 'acceptor_residue_ids', acc)

    def write_pivot_moves(self):
        """upside_config.py:1655-1666"""
        g = self.root['input'].create_group('pivot_moves')
        pivot_atom = self.potential['rama_coord/id'].data
        nonterm = np.array([(-1 not in tuple(x)) for x in pivot_atom])
        self.arr(g, 'proposal_pot', self.rama_pot)
        self.arr(g, 'pivot_atom', pivot_atom[nonterm])
        self.arr(g, 'pivot_restype', np.arange(self.n_res)[nonterm])
        self.arr(g, 'pivot_range', np.column_stack((pivot_atom[nonterm][:, 4] + 1,
                                                    np.zeros(nonterm.sum(), 'i') + self.n_atom)))

    def write_sidechain_radial(self, interaction_param, restype_names, excluded_residues=(), argument='placement_fixed_point_vector_only_CB',
                               suffix=''):
        """upside_config.py:866-883 (--sidechain-radial): a radial pair potential between one point per residue;
        interaction_param (n_type, n_type, 17) = 1/dx then 16 clamped B-spline knots per residue-type pair"""
        name2type = dict((_s(x), i) for i, x in enumerate(restype_names))
        residues = sorted(set(range(self.n_res)).difference(excluded_residues))
        g = self.group('radial' + suffix, [argument])
        self.arr(g, 'index', np.array(residues, dtype='i'))
        self.arr(g, 'type', np.array([name2type[self.fasta[r]] for r in residues], dtype='i'))
        self.arr(g, 'id', np.array(residues, dtype='i'))
        self.arr(g, 'interaction_param', np.asarray(interaction_param, dtype='f4'))

    def write_hbond_sc_radial(self, interaction_param, restype_names, argument='placement_fixed_point_vector_only_CB'):
        """the two-group radial node of src/sidechain_radial.cpp:108-136 (registered as hbond_sc_radial; the reference's
        config generator has no writer for it): backbone H / O virtual sites of protein_hbond (types 0 / 1) against one point
        per residue; interaction_param (2, n_type, 17)"""
        name2type = dict((_s(x), i) for i, x in enumerate(restype_names))
        ph = self.potential['protein_hbond']
        don, acc = np.asarray(ph['id1'].data), np.asarray(ph['id2'].data)
        g = self.group('hbond_sc_radial', ['protein_hbond', argument])
        self.arr(g, 'index1', np.arange(len(don) + len(acc), dtype='i'))
        self.arr(g, 'type1', np.array([0] * len(don) + [1] * len(acc), dtype='i'))
        self.arr(g, 'id1', np.concatenate([don, acc]).astype('i'))
        self.arr(g, 'index2', np.arange(self.n_res, dtype='i'))
        self.arr(g, 'type2', np.array([name2type[x] for x in self.fasta], dtype='i'))
        self.arr(g, 'id2', np.arange(self.n_res, dtype='i'))
        self.arr(g, 'interaction_param', np.asarray(interaction_param, dtype='f4'))

    def write_uniform_transform(self, argument, bspline_coeff, spline_offset, spline_inv_dx, suffix=''):
        """uniform_transform (src/environment.cpp:158-233): one clamped cubic B-spline applied to every element of a width-1 node"""
        g = self.group('uniform_transform' + suffix, [argument])
        d = self.arr(g, 'bspline_coeff', np.asarray(bspline_coeff, dtype='f4'))
        d.attrs['spline_offset'], d.attrs['spline_inv_dx'] = float(spline_offset), float(spline_inv_dx)

    def write_linear_coupling(self, argument, couplings, coupling_types, inactivation=None, inactivation_dim=0, suffix=''):
        """linear_coupling_uniform / linear_coupling_with_inactivation (src/environment.cpp:235-321)"""
        name = ('linear_coupling_with_inactivation' if inactivation else 'linear_coupling_uniform') + suffix
        g = self.group(name, [argument] + ([inactivation] if inactivation else []))
        self.arr(g, 'couplings', np.asarray(couplings, dtype='f4'))
        self.arr(g, 'coupling_types', np.asarray(coupling_types, dtype='i4'))
        if inactivation:
            g.attrs['inactivation_dim'] = int(inactivation_dim)

    def write_jump_moves(self, atom_ranges, sigma_trans, sigma_rot):
        """/input/jump_moves as JumpSampler reads it (src/monte_carlo_sampler.cpp:174-201): rigid translations (sigma_trans,
        Angstrom) and rotations about the centre of mass (sigma_rot, radians) of the atom ranges [first, next_first)"""
        atom_ranges = np.asarray(atom_ranges, dtype='i4').reshape(-1, 2)
        g = self.root['input'].create_group('jump_moves')
        self.arr(g, 'atom_range', atom_ranges)
        self.arr(g, 'sigma_trans', np.broadcast_to(np.asarray(sigma_trans, dtype='f4'), (len(atom_ranges),)).copy())
        self.arr(g, 'sigma_rot', np.broadcast_to(np.asarray(sigma_rot, dtype='f4'), (len(atom_ranges),)).copy())

    def write_chain_break(self, chain_first_residue, rl_chains=None, split_bonded=False):
        """/input/chain_break as upside_config.py:1413-1449 writes it for --chain-break-from-file: the first residue of every
        chain after the first (what extract_vtf.py / mdtraj_upside.py read back).  Returns the residues next to the breaks,
        which that script adds to --hbond-exclude-residues (no donor / acceptor is inferred across a break).  Like the
        reference generator this leaves the bonded terms alone unless ``split_bonded`` is set: then the bond / angle / dihedral
        springs and the Rama coordinates written AFTER this call skip the atoms of another chain.  Call it first."""
        cfr = np.asarray(chain_first_residue, dtype='i4').reshape(-1)
        if cfr.size and (np.any(np.diff(cfr) <= 0) or cfr[0] <= 0 or cfr[-1] >= self.n_res):
            raise ValueError('chain_first_residue must be ascending and inside (0, n_res)')
        self.chain_first_residue = cfr
        self.split_bonded = bool(split_bonded) and cfr.size > 0
        if cfr.size:
            g = self.root['input'].create_group('chain_break')
            self.arr(g, 'chain_first_residue', cfr)
            if rl_chains is not None:
                self.arr(g, 'rl_chains', np.asarray(rl_chains, dtype='i4'))
        return sorted({int(i) + j for i in cfr for j in (-1, 0)})

    def chain_endpts(self, i):
        """[first residue, first residue of the next chain) of chain i (upside_config.py:1184-1196)"""
        cfr = getattr(self, 'chain_first_residue', np.zeros(0, dtype='i4'))
        bounds = [0] + [int(x) for x in cfr] + [self.n_res]
        return bounds[i], bounds[i + 1]

    def write_jump_moves_per_chain(self, sigma_trans, sigma_rot):
        """one rigid-body jump move per chain of a multi-chain configuration"""
        n_chain = len(getattr(self, 'chain_first_residue', ())) + 1
        self.write_jump_moves([[3 * a, 3 * b] for a, b in (self.chain_endpts(i) for i in range(n_chain))], sigma_trans, sigma_rot)

    def save(self, path):
        h5lite.save(self.root, path)

    @classmethod
    def from_file(cls, path, compress=True):
        """reopen a written configuration to add nodes (restraints) to it"""
        root = h5lite.load(path)
        w = cls.__new__(cls)
        w.fasta = np.array([_s(x) for x in root['input/sequence'].data])
        w.n_res = len(w.fasta)
        w.n_atom = 3 * w.n_res
        w.pos = np.asarray(root['input/pos'].data[:, :, 0], dtype='f8')
        w.root, w.compress = root, compress
        w.potential = root['input/potential']
        return w


def load_rama_reference(path):
    with open(path, 'rb') as fh:
        return pickle.load(fh, encoding='latin1')


def write_ff1_config(path, fasta, pos, sidechain_lib, environment_lib, hbond_energy, rama_reference,
                     rama_pot=None, membrane=None, membrane_thickness=30., compress=True, chain_first_residue=None,
                     hbond_exclude_residues=(), split_bonded=False):
    """Order of calls follows ``main()`` of upside_config.py:1371-1671 for the README invocation; ``chain_first_residue`` is
    the content of its --chain-break-from-file."""
    w = ConfigWriter(fasta, pos, compress=compress)
    excluded = set(int(x) for x in hbond_exclude_residues)
    if chain_first_residue is not None:
        excluded |= set(w.write_chain_break(chain_first_residue, split_bonded=split_bonded))
    w.write_dist_spring()
    w.write_angle_spring()
    w.write_dihedral_spring()
    w.write_rotamer_placement(sidechain_lib)
    w.write_infer_H_O(excluded=excluded)
    w.write_count_hbond(hbond_energy, sidechain_lib)
    w.write_environment(environment_lib)
    w.write_rama_map_pot(synthetic_rama_pot(w.fasta) if rama_pot is None else rama_pot)
    w.write_reference_state_rama(rama_reference)
    w.write_backbone_pair()
    w.write_rotamer(sidechain_lib)
    if membrane is not None:
        w.write_membrane_potential(membrane, membrane_thickness)
        w.write_CB()
    w.write_rama_coord()
    w.write_affine_alignment()
    w.write_pivot_moves()
    w.save(path)
    return w
