"""TEST INFRASTRUCTURE (oracle/): Python-3 ctypes binding to oracle/_ref/libupside_ref_{pinned,fast}.so, i.e. to the
UNMODIFIED reference engine compiled by oracle/Makefile.  Mirrors the reference's py/upside_engine.py:19-186
(same C entry points) and adds the oracle-only accessors of ref_driver.cpp.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
import ctypes as ct
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def lib_path(flavour='pinned'):
    return os.path.join(HERE, '_ref', 'libupside_ref_%s.so' % flavour)


def available(flavour='pinned'):
    return os.path.exists(lib_path(flavour))


def load(flavour='pinned'):
    if flavour in _libs:
        return _libs[flavour]
    L = ct.CDLL(lib_path(flavour))
    fp, ip = ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)
    L.construct_deriv_engine.restype = ct.c_void_p
    L.construct_deriv_engine.argtypes = [ct.c_int, ct.c_char_p, ct.c_bool]
    L.free_deriv_engine.restype = None
    L.free_deriv_engine.argtypes = [ct.c_void_p]
    L.evaluate_energy.argtypes = [fp, ct.c_void_p, fp]
    L.evaluate_deriv.argtypes = [fp, ct.c_void_p, fp]
    L.get_output_dims.argtypes = [ip, ip, ct.c_void_p, ct.c_char_p]
    L.get_output.argtypes = [ct.c_int, fp, ct.c_void_p, ct.c_char_p]
    L.get_sens.argtypes = [ct.c_int, fp, ct.c_void_p, ct.c_char_p]
    L.get_param.argtypes = [ct.c_int, fp, ct.c_void_p, ct.c_char_p]
    L.set_param.argtypes = [ct.c_int, fp, ct.c_void_p, ct.c_char_p]
    L.get_param_deriv.argtypes = [ct.c_int, fp, ct.c_void_p, ct.c_char_p]
    L.get_value_by_name.argtypes = [ct.c_int, fp, ct.c_void_p, ct.c_char_p, ct.c_char_p]
    L.ref_pairlist.argtypes = [ct.c_void_p, ct.c_char_p, ip, ip, ct.c_int]
    L.ref_get_computation.restype = ct.c_void_p
    L.ref_get_computation.argtypes = [ct.c_void_p, ct.c_char_p]
    L.ref_n_nodes.argtypes = [ct.c_void_p]
    L.ref_node_name.argtypes = [ct.c_void_p, ct.c_int, ct.c_char_p, ct.c_int]
    L.ref_node_potential.argtypes = [ct.c_void_p, ct.c_char_p, fp]
    L.ref_rotamer_bead_marginals.argtypes = [ct.c_void_p, fp, ct.c_int]
    L.ref_rotamer_solve_stats.argtypes = [ct.c_void_p, fp]
    u32p = ct.POINTER(ct.c_uint32)
    L.ref_rng_bits.restype = None
    L.ref_rng_bits.argtypes = [ct.c_uint32, ct.c_uint32, ct.c_uint32, ct.c_uint64, u32p]
    L.ref_rng_raw.restype = None
    L.ref_rng_raw.argtypes = [u32p, u32p, u32p]
    L.ref_rng_normal3.restype = None
    L.ref_rng_normal3.argtypes = [ct.c_uint32, ct.c_uint32, ct.c_uint32, ct.c_uint64, fp]
    L.ref_rng_uniform.restype = None
    L.ref_rng_uniform.argtypes = [ct.c_uint32, ct.c_uint32, ct.c_uint32, ct.c_uint64, ct.c_int, fp]
    L.ref_md_run.argtypes = [ct.c_char_p, ct.c_int, ct.c_int, fp, fp, fp, ct.c_uint32, ct.c_float, ct.c_float,
                             ct.c_long, ct.c_int, ct.c_int, ct.POINTER(ct.c_double), fp]
    _libs[flavour] = L
    return L


def _fp(a):
    return a.ctypes.data_as(ct.POINTER(ct.c_float))


class RefEngine:
    """The reference DerivEngine for one configuration (cf. class Upside, py/upside_engine.py:77-186)."""

    def __init__(self, config_path, n_atom, flavour='pinned'):
        self.L = load(flavour)
        self.n_atom = int(n_atom)
        self.e = self.L.construct_deriv_engine(self.n_atom, config_path.encode(), True)
        if not self.e:
            raise RuntimeError('reference engine construction failed for %s' % config_path)

    def close(self):
        if self.e:
            self.L.free_deriv_engine(self.e)
            self.e = None

    __del__ = close

    def energy(self, pos):
        pos = np.ascontiguousarray(pos, dtype='f4').reshape(self.n_atom, 3)
        out = ct.c_float()
        if self.L.evaluate_energy(ct.byref(out), self.e, _fp(pos)):
            raise RuntimeError('evaluate_energy failed')
        return out.value

    def deriv(self, pos):
        pos = np.ascontiguousarray(pos, dtype='f4').reshape(self.n_atom, 3)
        d = np.zeros((self.n_atom, 3), dtype='f4')
        if self.L.evaluate_deriv(_fp(d), self.e, _fp(pos)):
            raise RuntimeError('evaluate_deriv failed')
        return d

    def node_names(self):
        out = []
        buf = ct.create_string_buffer(256)
        for i in range(self.L.ref_n_nodes(self.e)):
            is_pot = self.L.ref_node_name(self.e, i, buf, 256)
            out.append((buf.value.decode(), bool(is_pot)))
        return out

    def output_dims(self, node):
        n, w = ct.c_int(), ct.c_int()
        if self.L.get_output_dims(ct.byref(n), ct.byref(w), self.e, node.encode()):
            raise RuntimeError('get_output_dims failed for ' + node)
        return n.value, w.value

    def get_output(self, node):
        n, w = self.output_dims(node)
        a = np.zeros((n, w), dtype='f4')
        if self.L.get_output(a.size, _fp(a), self.e, node.encode()):
            raise RuntimeError('get_output failed for ' + node)
        return a

    def get_sens(self, node):
        n, w = self.output_dims(node)
        a = np.zeros((n, w), dtype='f4')
        if self.L.get_sens(a.size, _fp(a), self.e, node.encode()):
            raise RuntimeError('get_sens failed for ' + node)
        return a

    def node_potential(self, node):
        out = ct.c_float()
        if self.L.ref_node_potential(self.e, node.encode(), ct.byref(out)):
            raise RuntimeError(node + ' is not a potential node')
        return out.value

    def get_param_deriv(self, node, n):
        """reference get_param_deriv (PARAM_DERIV build, engine_c_library.cpp:93-108) after a deriv() call"""
        a = np.zeros(n, dtype='f4')
        if self.L.get_param_deriv(a.size, _fp(a), self.e, node.encode()):
            raise RuntimeError('get_param_deriv failed for ' + node)
        return a

    def get_value_by_name(self, node, name, n):
        a = np.zeros(n, dtype='f4')
        if self.L.get_value_by_name(n, _fp(a), self.e, node.encode(), name.encode()):
            raise RuntimeError('get_value_by_name failed')
        return a

    def pairlist(self, node, max_edge=1 << 20):
        i1 = np.zeros(max_edge, dtype='i4')
        i2 = np.zeros(max_edge, dtype='i4')
        n = self.L.ref_pairlist(self.e, node.encode(), i1.ctypes.data_as(ct.POINTER(ct.c_int)),
                                i2.ctypes.data_as(ct.POINTER(ct.c_int)), max_edge)
        if n < 0:
            raise RuntimeError('no pair list for node ' + node)
        return np.stack([i1[:n], i2[:n]], axis=1)

    def rotamer_bead_marginals(self, node='rotamer', n_bead=None):
        n_bead = n_bead or self.output_dims('placement_fixed_point_vector_only')[0]
        a = np.zeros(n_bead, dtype='f4')
        c = self.L.ref_get_computation(self.e, node.encode())
        if self.L.ref_rotamer_bead_marginals(c, _fp(a), n_bead) < 0:
            raise RuntimeError('not a rotamer node')
        return a

    def rotamer_solve_stats(self, node='rotamer'):
        a = np.zeros(8, dtype='f4')
        c = self.L.ref_get_computation(self.e, node.encode())
        if self.L.ref_rotamer_solve_stats(c, _fp(a)) < 0:
            raise RuntimeError('not a rotamer node')
        return dict(n_iter=int(a[0]), max_dev=float(a[1]), n33=int(a[2]), n36=int(a[3]), n66=int(a[4]),
                    n11=int(a[5]), n13=int(a[6]), n16=int(a[7]))


def md_run(config_path, pos, temperature, n_round, seed=42, dt=0.009, timescale=5.0, thermostat_interval=1,
           n_thread=0, flavour='fast'):
    """Reference MD loop over independent replicas (ref_driver.cpp:ref_md_run).  pos: (n_sys, n_atom, 3)."""
    L = load(flavour)
    pos = np.array(pos, dtype='f4', order='C')
    n_sys, n_atom = pos.shape[0], pos.shape[1]
    T = np.ascontiguousarray(np.broadcast_to(np.asarray(temperature, dtype='f4'), (n_sys,)))
    mom = np.zeros_like(pos)
    pot = np.zeros(n_sys, dtype='f4')
    sec = ct.c_double()
    rc = L.ref_md_run(config_path.encode(), n_sys, n_atom, _fp(pos), _fp(mom), _fp(T), seed, dt, timescale,
                      int(n_round), int(thermostat_interval), int(n_thread), ct.byref(sec), _fp(pot))
    if rc:
        raise RuntimeError('ref_md_run failed')
    return dict(pos=pos, mom=mom, potential=pot, seconds=sec.value)


def md_bench(config_path, pos, temperature, warm_rounds, n_round, n_rep=3, seed=42, dt=0.009, timescale=5.0, n_thread=0,
             flavour='fast'):
    """Timed reference MD (ref_driver.cpp:ref_md_bench): engines built once, warm_rounds untimed rounds, then n_rep timed
    repetitions of n_round rounds.  Returns dict(pos, seconds=[n_rep])."""
    L = load(flavour)
    pos = np.array(pos, dtype='f4', order='C')
    n_sys, n_atom = pos.shape[0], pos.shape[1]
    T = np.ascontiguousarray(np.broadcast_to(np.asarray(temperature, dtype='f4'), (n_sys,)))
    sec = (ct.c_double * n_rep)()
    L.ref_md_bench.restype = ct.c_int
    L.ref_md_bench.argtypes = [ct.c_char_p, ct.c_int, ct.c_int, ct.POINTER(ct.c_float), ct.POINTER(ct.c_float), ct.c_uint32, ct.c_float,
                               ct.c_float, ct.c_long, ct.c_long, ct.c_int, ct.c_int, ct.POINTER(ct.c_double)]
    if L.ref_md_bench(config_path.encode(), n_sys, n_atom, _fp(pos), _fp(T), seed, dt, timescale, int(warm_rounds), int(n_round),
                      int(n_rep), int(n_thread), sec):
        raise RuntimeError('ref_md_bench failed')
    return dict(pos=pos, seconds=list(sec))


def eval_latency(config_path, pos, n_warm=20, n_eval=200, flavour='fast'):
    """microseconds per evaluate_deriv call of the reference's C ABI on one system (ref_driver.cpp:ref_eval_latency)"""
    L = load(flavour)
    pos = np.array(pos, dtype='f4', order='C')
    us = ct.c_double()
    L.ref_eval_latency.restype = ct.c_int
    L.ref_eval_latency.argtypes = [ct.c_char_p, ct.c_int, ct.POINTER(ct.c_float), ct.c_int, ct.c_int, ct.POINTER(ct.c_double)]
    if L.ref_eval_latency(config_path.encode(), pos.shape[0], _fp(pos), int(n_warm), int(n_eval), ct.byref(us)):
        raise RuntimeError('ref_eval_latency failed')
    return us.value


def mc_steps(config_path, pos, temperature, base_seed, first_round, n_step, flavour='pinned', max_sampler=4):
    """Reference Monte-Carlo samplers (ref_driver.cpp:ref_mc_steps) on copies of `pos` (n_sys, n_atom, 3), system s seeded
    base_seed + s.  Returns (new positions, stats[n_sys, n_sampler, 2] = n_success, n_attempt)."""
    import shutil
    import tempfile
    L = load(flavour)
    pos = np.array(pos, dtype='f4', order='C')
    n_sys, n_atom = pos.shape[0], pos.shape[1]
    T = np.ascontiguousarray(np.broadcast_to(np.asarray(temperature, dtype='f4'), (n_sys,)))
    stats = np.zeros((n_sys, max_sampler, 2), dtype=np.int64)
    L.ref_mc_steps.restype = ct.c_int
    L.ref_mc_steps.argtypes = [ct.c_char_p, ct.c_int, ct.c_int, ct.POINTER(ct.c_float), ct.POINTER(ct.c_float), ct.c_uint32,
                               ct.c_uint64, ct.c_int, ct.POINTER(ct.c_long), ct.c_int]
    n = 0
    with tempfile.TemporaryDirectory() as tmp:
        for s in range(n_sys):   # a fresh scratch copy per system: the samplers' loggers create /output datasets in it
            scratch = os.path.join(tmp, 'mc%d.up' % s)
            shutil.copy(config_path, scratch)
            n = L.ref_mc_steps(scratch.encode(), 1, n_atom, _fp(pos[s]), _fp(T[s:s + 1]), int(base_seed) + s, int(first_round),
                               int(n_step), stats[s].ctypes.data_as(ct.POINTER(ct.c_long)), max_sampler)
            if n < 0:
                break
    if n < 0:
        raise RuntimeError('ref_mc_steps failed')
    return pos, stats[:, :n]
