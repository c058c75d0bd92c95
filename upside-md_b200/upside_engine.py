"""ctypes binding to the B200 engine, mirroring the reference's py/upside_engine.py.

`Upside` has the same constructor and methods as the reference class (py/upside_engine.py:172-277: energy, deriv,
set_param, get_param, get_param_deriv, get_output, get_sens, get_value_by_name) and goes through the same C entry
points (include/engine_c_library.h).  `BatchEngine` binds the batched ABI (include/upside_b200.h) that replaces the
reference's one-engine-per-OpenMP-thread loop.  The shared library is required: there is no CPU fallback.
"""
import ctypes as ct
import os

import numpy as np

from . import h5lite

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libupside_b200.so')
_lib = None

_fp = ct.POINTER(ct.c_float)
_ip = ct.POINTER(ct.c_int)


def lib():
    """Load libupside_b200.so (built in-tree by __graft_entry__.build / csrc/Makefile); raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError('%s not found: build it with `make -C upside-md_b200/csrc` (no CPU fallback exists)' % LIB_PATH)
    L = ct.CDLL(LIB_PATH)
    # reference ABI (include/engine_c_library.h)
    L.construct_deriv_engine.restype = ct.c_void_p
    L.construct_deriv_engine.argtypes = [ct.c_int, ct.c_char_p, ct.c_bool]
    L.free_deriv_engine.restype = None
    L.free_deriv_engine.argtypes = [ct.c_void_p]
    L.evaluate_energy.argtypes = [_fp, ct.c_void_p, _fp]
    L.evaluate_deriv.argtypes = [_fp, ct.c_void_p, _fp]
    L.set_param.argtypes = [ct.c_int, _fp, ct.c_void_p, ct.c_char_p]
    L.get_param.argtypes = [ct.c_int, _fp, ct.c_void_p, ct.c_char_p]
    L.get_param_deriv.argtypes = [ct.c_int, _fp, ct.c_void_p, ct.c_char_p]
    L.get_output_dims.argtypes = [_ip, _ip, ct.c_void_p, ct.c_char_p]
    L.get_output.argtypes = [ct.c_int, _fp, ct.c_void_p, ct.c_char_p]
    L.get_sens.argtypes = [ct.c_int, _fp, ct.c_void_p, ct.c_char_p]
    L.get_value_by_name.argtypes = [ct.c_int, _fp, ct.c_void_p, ct.c_char_p, ct.c_char_p]
    L.clamped_spline_solve.argtypes = [ct.c_int, _fp, _fp]
    L.clamped_spline_value.argtypes = [ct.c_int, _fp, _fp, ct.c_int, _fp]
    L.get_clamped_value_and_deriv.argtypes = [ct.c_int, _fp, _fp, ct.c_int, _fp]
    L.get_clamped_coeff_deriv.argtypes = [ct.c_int, _fp, _fp, ct.c_float]
    # batched ABI (include/upside_b200.h)
    L.ub_last_error.restype = ct.c_char_p
    L.ub_engine_create.restype = ct.c_void_p
    L.ub_engine_create.argtypes = [ct.c_char_p, ct.c_int, ct.c_int]
    L.ub_engine_destroy.restype = None
    L.ub_engine_destroy.argtypes = [ct.c_void_p]
    L.ub_n_atom.argtypes = [ct.c_void_p]
    L.ub_n_replica.argtypes = [ct.c_void_p]
    for nm in ('ub_initial_pos', 'ub_set_pos', 'ub_get_pos', 'ub_set_mom', 'ub_get_mom', 'ub_kinetic_energy'):
        getattr(L, nm).argtypes = [ct.c_void_p, _fp]
    L.ub_evaluate.argtypes = [ct.c_void_p, _fp, _fp]
    L.ub_n_nodes.argtypes = [ct.c_void_p]
    L.ub_node_name.argtypes = [ct.c_void_p, ct.c_int, ct.c_char_p, ct.c_int, _ip]
    L.ub_get_output_dims.argtypes = [ct.c_void_p, ct.c_char_p, _ip, _ip]
    L.ub_get_output.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_int, ct.c_int, _fp]
    L.ub_get_sens.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_int, ct.c_int, _fp]
    L.ub_get_node_potential.argtypes = [ct.c_void_p, ct.c_char_p, _fp]
    L.ub_get_value_by_name.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_char_p, ct.c_int, ct.c_int, _fp, _ip]
    L.ub_get_param.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_int, _fp, _ip]
    L.ub_set_param.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_int, _fp]
    L.ub_get_param_deriv.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_int, ct.c_int, _fp, _ip]
    L.ub_get_pairlist.argtypes = [ct.c_void_p, ct.c_char_p, ct.c_int, ct.c_int, _ip, _ip, _ip]
    L.ub_md_init.argtypes = [ct.c_void_p, ct.c_uint32, _fp, ct.c_float, ct.c_float, ct.c_int]
    L.ub_md_set_temperature.argtypes = [ct.c_void_p, _fp]
    L.ub_md_run.argtypes = [ct.c_void_p, ct.c_long]
    L.ub_checkpoint_size.restype = ct.c_long
    L.ub_checkpoint_size.argtypes = [ct.c_void_p]
    L.ub_checkpoint_save.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_long, ct.POINTER(ct.c_long)]
    L.ub_checkpoint_load.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_long]
    L.ub_sync.argtypes = [ct.c_void_p]
    L.ub_mc_n_samplers.argtypes = [ct.c_void_p]
    L.ub_mc_sampler_name.argtypes = [ct.c_void_p, ct.c_int, ct.c_char_p, ct.c_int]
    L.ub_mc_execute.argtypes = [ct.c_void_p, ct.c_uint64]
    L.ub_mc_stats.argtypes = [ct.c_void_p, ct.c_int, ct.POINTER(ct.c_uint64), ct.POINTER(ct.c_uint64), ct.c_int]
    L.ub_recenter.argtypes = [ct.c_void_p, ct.c_int]
    L.ub_stream.restype = ct.c_void_p
    L.ub_stream.argtypes = [ct.c_void_p]
    L.ub_launches_per_eval.argtypes = [ct.c_void_p]
    L.ub_profile_eval.argtypes = [ct.c_void_p, ct.c_int, ct.c_char_p, ct.c_int, _fp, _ip]
    L.ub_measure_fp32_peak.argtypes = [ct.c_int, _fp, _fp]
    L.ub_rng_probe.argtypes = [ct.c_uint32, ct.c_uint32, ct.c_uint32, ct.c_uint64, ct.POINTER(ct.c_uint32), _fp]
    L.upside_main.argtypes = [ct.c_int, ct.POINTER(ct.c_char_p), ct.c_int]
    L.ub_md_init_seeds.argtypes = [ct.c_void_p, ct.POINTER(ct.c_uint32), _fp, ct.c_float, ct.c_float, ct.c_int]
    L.ub_set_pos_range.argtypes = [ct.c_void_p, _fp, ct.c_int, ct.c_int]
    L.ub_get_pos_range.argtypes = [ct.c_void_p, _fp, ct.c_int, ct.c_int]
    L.ub_swap_pos.argtypes = [ct.c_void_p, ct.c_int, _ip]
    L.ub_replex_create.restype = ct.c_void_p
    L.ub_replex_create.argtypes = [ct.c_int, ct.c_int, ct.POINTER(ct.c_char_p)]
    L.ub_replex_destroy.restype = None
    L.ub_replex_destroy.argtypes = [ct.c_void_p]
    L.ub_replex_n_sets.argtypes = [ct.c_void_p]
    L.ub_replex_set_size.argtypes = [ct.c_void_p, ct.c_int]
    L.ub_replex_pairs.argtypes = [ct.c_void_p, ct.c_int, _ip]
    L.ub_replex_begin.argtypes = [ct.c_void_p, ct.c_uint32, ct.c_uint64]
    L.ub_replex_decide.argtypes = [ct.c_void_p, ct.c_int, _fp, _fp, _ip]
    L.ub_replex_decide_same_hamiltonian.argtypes = [ct.c_void_p, ct.c_int, _fp, _fp, _ip]
    L.ub_replex_replica_indices.argtypes = [ct.c_void_p, _ip]
    L.ub_replex_counts.argtypes = [ct.c_void_p, ct.c_int, ct.POINTER(ct.c_uint64), ct.POINTER(ct.c_uint64)]
    L.ub_host_rng_uniform.argtypes = [ct.c_uint32, ct.c_uint32, ct.c_uint32, ct.c_uint64, ct.c_int, _fp, ct.POINTER(ct.c_uint32)]
    # device-resident ladder exchange over NCCL (csrc/ladder_nccl.cu)
    _u64p = ct.POINTER(ct.c_uint64)
    L.ub_ladder_last_error.restype = ct.c_char_p
    L.ub_nccl_unique_id.argtypes = [ct.c_char_p, ct.c_int]
    L.ub_nccl_comm_create.restype = ct.c_void_p
    L.ub_nccl_comm_create.argtypes = [ct.c_char_p, ct.c_int, ct.c_int, ct.c_int]
    L.ub_nccl_comm_destroy.restype = None
    L.ub_nccl_comm_destroy.argtypes = [ct.c_void_p]
    L.ub_ladder_create.restype = ct.c_void_p
    L.ub_ladder_create.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_int, ct.c_int, ct.c_int, ct.c_int, ct.POINTER(ct.c_char_p), ct.c_uint32, _fp]
    L.ub_ladder_destroy.restype = None
    L.ub_ladder_destroy.argtypes = [ct.c_void_p]
    L.ub_ladder_attempt.argtypes = [ct.c_void_p, ct.c_uint64]
    L.ub_ladder_set_temperature.argtypes = [ct.c_void_p, _fp]
    L.ub_ladder_n_pairs.argtypes = [ct.c_void_p]
    L.ub_ladder_state.argtypes = [ct.c_void_p, _ip, _ip, _u64p, _u64p, _fp]
    L.ub_ladder_comm_bytes.argtypes = [ct.c_void_p, _u64p, _u64p]
    _lib = L
    return L


def _f(a):
    return a.ctypes.data_as(_fp)


def _err(what):
    return RuntimeError('%s: %s' % (what, (lib().ub_last_error() or b'').decode()))


def read_config(config_file_path):
    """(initial_pos (n_atom,3), sequence) of a .up file - what the reference reads with PyTables (:174-177)"""
    t = h5lite.load(str(config_file_path))
    pos = np.array(t['input/pos'].data[:, :, 0], dtype='f4')
    seq = t['input/sequence'].data if 'input/sequence' in t else None
    return pos, seq


class Upside(object):
    """Single-replica engine; drop-in for the reference's `Upside` (py/upside_engine.py:172-277)."""

    def __init__(self, config_file_path, quiet=True):
        self.config_file_path = str(config_file_path)
        self.initial_pos, self.sequence = read_config(self.config_file_path)
        self.n_atom = self.initial_pos.shape[0]
        self.engine = lib().construct_deriv_engine(self.n_atom, self.config_file_path.encode(), bool(quiet))
        if not self.engine:
            raise RuntimeError('Unable to initialize upside engine for %s' % (config_file_path,))

    def __repr__(self):
        return 'Upside(%r, %r)' % (self.n_atom, self.config_file_path)

    def energy(self, pos):
        pos = np.require(pos, dtype='f4', requirements='C')
        assert pos.shape == (self.n_atom, 3)
        energy = np.zeros(1, dtype='f4')
        if lib().evaluate_energy(_f(energy), self.engine, _f(pos)):
            raise RuntimeError('Unable to evaluate energy')
        return energy[0]

    def deriv(self, pos):
        pos = np.require(pos, dtype='f4', requirements='C')
        assert pos.shape == (self.n_atom, 3)
        deriv = np.zeros_like(pos)
        if lib().evaluate_deriv(_f(deriv), self.engine, _f(pos)):
            raise RuntimeError('Unable to evaluate derivative')
        return deriv

    def set_param(self, param, node_name):
        param_size = param.shape
        param = np.require(param.ravel(), dtype='f4', requirements='C')
        if lib().set_param(int(param.shape[0]), _f(param), self.engine, node_name.encode()):
            raise RuntimeError('Unable to set param with size %s for node %s' % (param_size, node_name))

    def get_param(self, param_shape, node_name):
        param = np.zeros(param_shape, dtype='f4')
        if lib().get_param(int(np.prod(param_shape)), _f(param), self.engine, node_name.encode()):
            raise RuntimeError('Unable to get param')
        return param

    def get_param_deriv(self, param_shape, node_name):
        deriv = np.zeros(param_shape, dtype='f4')
        if lib().get_param_deriv(int(np.prod(param_shape)), _f(deriv), self.engine, node_name.encode()):
            raise RuntimeError('Unable to get param deriv')
        return deriv

    def _dims(self, node_name):
        n_elem, width = ct.c_int(), ct.c_int()
        if lib().get_output_dims(ct.byref(n_elem), ct.byref(width), self.engine, node_name.encode()):
            raise RuntimeError('Unable to get output dims')
        return n_elem.value, width.value

    def get_output(self, node_name):
        output = np.zeros(self._dims(node_name), dtype='f4')
        if lib().get_output(int(output.size), _f(output), self.engine, node_name.encode()):
            raise RuntimeError('Unable to get output')
        return output

    def get_sens(self, node_name):
        output = np.zeros(self._dims(node_name), dtype='f4')
        if lib().get_sens(int(output.size), _f(output), self.engine, node_name.encode()):
            raise RuntimeError('Unable to get sens')
        return output

    def get_value_by_name(self, output_shape, node_name, log_name):
        output = np.zeros(output_shape, dtype='f4')
        if lib().get_value_by_name(int(output.size), _f(output), self.engine, node_name.encode(), log_name.encode()):
            raise RuntimeError('Unable to get value by name')
        return output

    def __del__(self):
        if getattr(self, 'engine', None):
            lib().free_deriv_engine(self.engine)
            self.engine = None


class BatchEngine(object):
    """n_replica copies of one configuration on one GPU (include/upside_b200.h)."""

    def __init__(self, config_file_path, n_replica, device=0):
        self.config_file_path = str(config_file_path)
        self.L = lib()
        self.e = self.L.ub_engine_create(self.config_file_path.encode(), int(n_replica), int(device))
        if not self.e:
            raise _err('Unable to initialize batched upside engine for %s' % config_file_path)
        self.n_atom = self.L.ub_n_atom(self.e)
        self.n_replica = self.L.ub_n_replica(self.e)
        self.initial_pos = np.zeros((self.n_atom, 3), dtype='f4')
        if self.L.ub_initial_pos(self.e, _f(self.initial_pos)):
            raise _err('initial_pos')

    def close(self):
        if getattr(self, 'e', None):
            self.L.ub_engine_destroy(self.e)
            self.e = None

    __del__ = close

    def _arr(self, a):
        a = np.require(a, dtype='f4', requirements='C')
        assert a.shape == (self.n_replica, self.n_atom, 3), a.shape
        return a

    def set_pos(self, pos):
        if self.L.ub_set_pos(self.e, _f(self._arr(pos))): raise _err('set_pos')

    def get_pos(self):
        a = np.zeros((self.n_replica, self.n_atom, 3), dtype='f4')
        if self.L.ub_get_pos(self.e, _f(a)): raise _err('get_pos')
        return a

    def set_mom(self, mom):
        if self.L.ub_set_mom(self.e, _f(self._arr(mom))): raise _err('set_mom')

    def get_mom(self):
        a = np.zeros((self.n_replica, self.n_atom, 3), dtype='f4')
        if self.L.ub_get_mom(self.e, _f(a)): raise _err('get_mom')
        return a

    def evaluate(self, pos=None, want_deriv=True):
        """energies (n_replica,) and dV/dx (n_replica,n_atom,3) of a PotentialAndDerivMode evaluation"""
        if pos is not None:
            self.set_pos(pos)
        en = np.zeros(self.n_replica, dtype='f4')
        d = np.zeros((self.n_replica, self.n_atom, 3), dtype='f4') if want_deriv else None
        if self.L.ub_evaluate(self.e, _f(en), _f(d) if want_deriv else None): raise _err('evaluate')
        return (en, d) if want_deriv else en

    def node_names(self):
        out = []
        buf = ct.create_string_buffer(256)
        isp = ct.c_int()
        for i in range(self.L.ub_n_nodes(self.e)):
            self.L.ub_node_name(self.e, i, buf, 256, ct.byref(isp))
            out.append((buf.value.decode(), bool(isp.value)))
        return out

    def output_dims(self, node):
        n, w = ct.c_int(), ct.c_int()
        if self.L.ub_get_output_dims(self.e, node.encode(), ct.byref(n), ct.byref(w)): raise _err('get_output_dims')
        return n.value, w.value

    def get_output(self, node, replica=0):
        a = np.zeros(self.output_dims(node), dtype='f4')
        if self.L.ub_get_output(self.e, node.encode(), replica, a.size, _f(a)): raise _err('get_output')
        return a

    def get_sens(self, node, replica=0):
        a = np.zeros(self.output_dims(node), dtype='f4')
        if self.L.ub_get_sens(self.e, node.encode(), replica, a.size, _f(a)): raise _err('get_sens')
        return a

    def node_potential(self, node):
        a = np.zeros(self.n_replica, dtype='f4')
        if self.L.ub_get_node_potential(self.e, node.encode(), _f(a)): raise _err('get_node_potential')
        return a

    def get_value_by_name(self, node, log_name, replica=0):
        n = ct.c_int()
        if self.L.ub_get_value_by_name(self.e, node.encode(), log_name.encode(), replica, 0, None, ct.byref(n)):
            raise _err('get_value_by_name')
        a = np.zeros(n.value, dtype='f4')
        if self.L.ub_get_value_by_name(self.e, node.encode(), log_name.encode(), replica, a.size, _f(a), ct.byref(n)):
            raise _err('get_value_by_name')
        return a

    def get_param(self, node):
        n = ct.c_int()
        if self.L.ub_get_param(self.e, node.encode(), 0, None, ct.byref(n)): raise _err('get_param')
        a = np.zeros(n.value, dtype='f4')
        if self.L.ub_get_param(self.e, node.encode(), a.size, _f(a), ct.byref(n)): raise _err('get_param')
        return a

    def get_param_deriv(self, node, replica=0):
        """dV/d(parameter) of `node` for one replica (replica < 0: summed over the batch) from the state the last
        evaluate(want_deriv=True) left behind; empty if the node has no parameter derivative"""
        n = ct.c_int()
        if self.L.ub_get_param_deriv(self.e, node.encode(), replica, 0, None, ct.byref(n)): raise _err('get_param_deriv')
        a = np.zeros(n.value, dtype='f4')
        if n.value and self.L.ub_get_param_deriv(self.e, node.encode(), replica, a.size, _f(a), ct.byref(n)): raise _err('get_param_deriv')
        return a

    def set_param(self, node, param):
        p = np.require(np.ravel(param), dtype='f4', requirements='C')
        if self.L.ub_set_param(self.e, node.encode(), p.size, _f(p)): raise _err('set_param')

    def pairlist(self, node, replica=0):
        """(n_edge,2) pair list of the last evaluation, in the reference's emission order"""
        n = ct.c_int()
        cap = 1 << 16
        while True:
            i1 = np.zeros(cap, dtype='i4')
            i2 = np.zeros(cap, dtype='i4')
            if self.L.ub_get_pairlist(self.e, node.encode(), replica, cap, i1.ctypes.data_as(_ip),
                                      i2.ctypes.data_as(_ip), ct.byref(n)):
                raise _err('get_pairlist')
            if n.value <= cap:
                return np.stack([i1[:n.value], i2[:n.value]], axis=1)
            cap = n.value

    def md_init(self, temperature, seed=42, dt=0.009, timescale=5.0, thermostat_interval=1):
        T = np.ascontiguousarray(np.broadcast_to(np.asarray(temperature, dtype='f4'), (self.n_replica,)))
        if self.L.ub_md_init(self.e, int(seed), _f(T), dt, timescale, int(thermostat_interval)): raise _err('md_init')

    def md_init_seeds(self, temperature, seeds, dt=0.009, timescale=5.0, thermostat_interval=1):
        """as md_init with one RNG key per replica (seed of system ns in the reference = base_seed + ns, main.cpp:459)"""
        T = np.ascontiguousarray(np.broadcast_to(np.asarray(temperature, dtype='f4'), (self.n_replica,)))
        sd = np.ascontiguousarray(np.asarray(seeds, dtype=np.uint32))
        assert sd.shape == (self.n_replica,)
        if self.L.ub_md_init_seeds(self.e, sd.ctypes.data_as(ct.POINTER(ct.c_uint32)), _f(T), dt, timescale, int(thermostat_interval)):
            raise _err('md_init_seeds')

    def get_pos_range(self, first, n):
        a = np.zeros((n, self.n_atom, 3), dtype='f4')
        if self.L.ub_get_pos_range(self.e, _f(a), int(first), int(n)): raise _err('get_pos_range')
        return a

    def set_pos_range(self, pos, first):
        a = np.require(pos, dtype='f4', requirements='C')
        assert a.shape[1:] == (self.n_atom, 3)
        if self.L.ub_set_pos_range(self.e, _f(a), int(first), int(a.shape[0])): raise _err('set_pos_range')

    def swap_pos(self, pairs):
        """exchange the coordinates of replica pairs [(a0,b0),(a1,b1),...]"""
        p = np.require(np.asarray(pairs, dtype='i4').reshape(-1, 2), requirements='C')
        if self.L.ub_swap_pos(self.e, int(p.shape[0]), p.ctypes.data_as(_ip)): raise _err('swap_pos')

    def set_temperature(self, temperature):
        T = np.ascontiguousarray(np.broadcast_to(np.asarray(temperature, dtype='f4'), (self.n_replica,)))
        if self.L.ub_md_set_temperature(self.e, _f(T)): raise _err('md_set_temperature')

    def checkpoint(self):
        """bytes holding positions, momenta, RNG keys and counters, temperatures and integrator settings of every replica"""
        n = self.L.ub_checkpoint_size(self.e)
        buf = ct.create_string_buffer(n)
        w = ct.c_long()
        if self.L.ub_checkpoint_save(self.e, buf, n, ct.byref(w)):
            raise _err('checkpoint_save')
        return buf.raw[:w.value]

    def restore(self, blob):
        """resume from checkpoint(): the run continues with the thermostat noise of the uninterrupted run"""
        if self.L.ub_checkpoint_load(self.e, blob, len(blob)):
            raise _err('checkpoint_load')

    def md_run(self, n_round, sync=True):
        if self.L.ub_md_run(self.e, int(n_round)): raise _err('md_run')
        if sync:
            self.sync()

    def sync(self):
        if self.L.ub_sync(self.e): raise _err('sync')

    # Monte-Carlo moves (reference monte_carlo_sampler.cpp): the samplers of /input/pivot_moves and /input/jump_moves
    def mc_samplers(self):
        names = []
        for i in range(self.L.ub_mc_n_samplers(self.e)):
            buf = ct.create_string_buffer(64)
            if self.L.ub_mc_sampler_name(self.e, i, buf, 64): raise _err('mc_sampler_name')
            names.append(buf.value.decode())
        return names

    def mc_execute(self, round_num, sync=True):
        """one Metropolis step of every sampler for every replica (MultipleMonteCarloSampler::execute); needs md_init"""
        if self.L.ub_mc_execute(self.e, int(round_num)): raise _err('mc_execute')
        if sync:
            self.sync()

    def mc_stats(self, sampler, reset=False):
        """(n_success, n_attempt) per replica of sampler number `sampler`"""
        ok, tr = np.zeros(self.n_replica, dtype='u8'), np.zeros(self.n_replica, dtype='u8')
        p = ct.POINTER(ct.c_uint64)
        if self.L.ub_mc_stats(self.e, int(sampler), ok.ctypes.data_as(p), tr.ctypes.data_as(p), int(reset)): raise _err('mc_stats')
        return ok, tr

    def recenter(self, xy_only=False):
        if self.L.ub_recenter(self.e, int(xy_only)): raise _err('recenter')

    def kinetic_energy(self):
        a = np.zeros(self.n_replica, dtype='f4')
        if self.L.ub_kinetic_energy(self.e, _f(a)): raise _err('kinetic_energy')
        return a

    def stream(self):
        return self.L.ub_stream(self.e)

    def profile_eval(self):
        """[(label, milliseconds)] of one DerivMode evaluation, kernel group by kernel group (CUDA events, engine stream)"""
        cap, ll = 256, 64
        buf = ct.create_string_buffer(cap * ll)
        ms = np.zeros(cap, dtype='f4')
        n = ct.c_int()
        if self.L.ub_profile_eval(self.e, cap, buf, ll, _f(ms), ct.byref(n)): raise _err('profile_eval')
        raw = buf.raw
        return [(raw[i * ll:(i + 1) * ll].split(b'\0')[0].decode(), float(ms[i])) for i in range(min(n.value, cap))]

    def launches_per_eval(self):
        return self.L.ub_launches_per_eval(self.e)


class Ladder(object):
    """Replica exchange of a temperature ladder sharded over GPUs, on the device (csrc/ladder_nccl.cu; reference
    ReplicaExchange::attempt_swaps, src/main.cpp:227-275).  `engine` holds this rank's contiguous block of rungs;
    `temperature_all` the temperatures of all n_global rungs; `swap_sets` the reference's --swap-set strings
    ("0-1,2-3,...").  With world > 1 pass nccl_comm (see nccl_comm_from_torch)."""

    def __init__(self, engine, swap_sets, temperature_all, seed=42, rank=0, world=1, nccl_comm=None):
        self.L, self.engine = lib(), engine
        t = np.ascontiguousarray(temperature_all, dtype='f4')
        self.n_global = len(t)
        arr = (ct.c_char_p * len(swap_sets))(*[s.encode() for s in swap_sets])
        self.h = self.L.ub_ladder_create(engine.e, nccl_comm, rank, world, self.n_global, len(swap_sets), arr, seed, _f(t))
        if not self.h:
            raise RuntimeError('ub_ladder_create: %s' % (self.L.ub_ladder_last_error() or b'').decode())
        self.n_pairs = self.L.ub_ladder_n_pairs(self.h)

    def attempt(self, round_num):
        """asynchronous: enqueued on the engine's stream"""
        if self.L.ub_ladder_attempt(self.h, int(round_num)):
            raise RuntimeError('ub_ladder_attempt: %s' % (self.L.ub_ladder_last_error() or b'').decode())

    def state(self):
        """(replica_indices, accept of the last attempt, n_attempt, n_success, energies of the last attempt); waits for the stream"""
        ri, acc = np.zeros(self.n_global, dtype='i4'), np.zeros(max(1, self.n_pairs), dtype='i4')
        na, ns = np.zeros(max(1, self.n_pairs), dtype='u8'), np.zeros(max(1, self.n_pairs), dtype='u8')
        en = np.zeros(self.n_global, dtype='f4')
        u64p = ct.POINTER(ct.c_uint64)
        if self.L.ub_ladder_state(self.h, ri.ctypes.data_as(_ip), acc.ctypes.data_as(_ip), na.ctypes.data_as(u64p), ns.ctypes.data_as(u64p), _f(en)):
            raise RuntimeError('ub_ladder_state: %s' % (self.L.ub_ladder_last_error() or b'').decode())
        return ri, acc[:self.n_pairs], na[:self.n_pairs], ns[:self.n_pairs], en

    def comm_bytes(self):
        a, c = ct.c_uint64(), ct.c_uint64()
        self.L.ub_ladder_comm_bytes(self.h, ct.byref(a), ct.byref(c))
        return int(a.value), int(c.value)

    def close(self):
        if self.h:
            self.L.ub_ladder_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def nccl_comm_from_torch(dist, device):
    """an ncclComm_t of our own for the ranks of an initialised torch.distributed job: rank 0 creates the unique id, the
    job's existing backend carries its 128 bytes to the other ranks (plumbing only), every rank joins"""
    import torch
    L = lib()
    buf = ct.create_string_buffer(128)
    if dist.get_rank() == 0 and L.ub_nccl_unique_id(buf, 128):
        raise RuntimeError('ub_nccl_unique_id: %s' % (L.ub_ladder_last_error() or b'').decode())
    dev = torch.device('cuda', device) if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().tolist())
    comm = L.ub_nccl_comm_create(raw, dist.get_world_size(), dist.get_rank(), device)
    if not comm:
        raise RuntimeError('ub_nccl_comm_create: %s' % (L.ub_ladder_last_error() or b'').decode())
    return comm


def measure_fp32_peak(device=0):
    """TFLOP/s of the FP32 FMA pipe measured on `device` (csrc/peaks.cu), and the nominal SM clock in MHz"""
    tf, mhz = np.zeros(1, dtype='f4'), np.zeros(1, dtype='f4')
    if lib().ub_measure_fp32_peak(int(device), _f(tf), _f(mhz)): raise _err('measure_fp32_peak')
    return float(tf[0]), float(mhz[0])


def rng_probe(seed, stream, atom, timestep):
    bits = (ct.c_uint32 * 4)()
    f = np.zeros(4, dtype='f4')
    if lib().ub_rng_probe(seed, stream, atom, timestep, bits, _f(f)): raise _err('rng_probe')
    return np.array(list(bits), dtype=np.uint32), f[:3].copy(), float(f[3])


def host_rng_uniform(seed, stream, atom, timestep, n_draw):
    """n_draw successive RandomGenerator(seed, stream, atom, timestep).uniform_open_closed().x on the HOST generator,
    and the raw Threefry bits of the first draw"""
    out = np.zeros(n_draw, dtype='f4')
    bits = (ct.c_uint32 * 4)()
    lib().ub_host_rng_uniform(seed, stream, atom, timestep, n_draw, _f(out), bits)
    return out, np.array(list(bits), dtype=np.uint32)


def in_process_upside(args, verbose=True):
    """Run the `upside` command line in-process (reference py/upside_engine.py:66-91)."""
    exec_args = [b'python_library', b'--re-raise-signal'] + [a.encode() if isinstance(a, str) else a for a in args]
    arr = (ct.c_char_p * len(exec_args))(*exec_args)
    retcode = lib().upside_main(len(exec_args), arr, int(verbose))
    if retcode:
        raise RuntimeError('In process Upside returned %i' % retcode)


def continue_config(config_file_path):
    """Prepare a finished run for continuation as the reference's continue_sim does (py/run_upside.py:231-257): the last
    frame of /output/pos becomes /input/pos and /output is renamed /output_previous_<i>.  Returns the last logged
    temperature, to be passed to the next run (the momenta are drawn afresh, as in the reference; for an exact
    continuation inside one process use BatchEngine.checkpoint / restore)."""
    t = h5lite.load(str(config_file_path))
    i = 0
    while 'output_previous_%i' % i in t:
        i += 1
    n = t['output'] if 'output' in t else t['output_previous_%i' % (i - 1)]
    pos = np.asarray(n['pos'].data)                    # (n_frame, 1, n_atom, 3)
    t['input/pos'].data[:, :, 0] = pos[-1, 0]
    temperature = float(np.asarray(n['temperature'].data)[-1, 0])
    if 'output' in t:
        t.children['output_previous_%i' % i] = t.children.pop('output')
    h5lite.save(t, str(config_file_path))
    return temperature


def clamped_spline_value(bspline_coeff, x):
    x = np.require(x, dtype='f4', requirements='C')
    c = np.require(bspline_coeff, dtype='f4', requirements='C')
    result = np.zeros(len(x), dtype='f4')
    if lib().clamped_spline_value(len(c), _f(result), _f(c), len(x), _f(x)): raise RuntimeError('spline evaluation error')
    return result


def clamped_spline_solve(values):
    values = np.require(values, dtype='f4', requirements='C')
    c = np.zeros(len(values) + 2, dtype='f4')
    if lib().clamped_spline_solve(len(c), _f(c), _f(values)): raise RuntimeError('spline solve error')
    return c


def clamped_value_and_deriv(bspline_coeff, x):
    x = np.require(x, dtype='f4', requirements='C')
    c = np.require(bspline_coeff, dtype='f4', requirements='C')
    result = np.zeros((len(x), 2), dtype='f4')
    if lib().get_clamped_value_and_deriv(len(c), _f(result), _f(c), len(x), _f(x)): raise RuntimeError('spline evaluation error')
    return result


def clamped_coeff_deriv(bspline_coeff, x):
    x = np.asarray(x, dtype='f4')
    c = np.require(bspline_coeff, dtype='f4', requirements='C')
    result = np.zeros((len(x), len(c)), dtype='f4')
    for i, y in enumerate(x):
        if lib().get_clamped_coeff_deriv(len(c), _f(result[i]), _f(c), float(y)): raise RuntimeError('spline evaluation error')
    return result
