"""TEST INFRASTRUCTURE (oracle/): numpy restatement of the integer / predicate parts of the reference's MD inner loop,
kept independent of both the CUDA engine and the compiled reference so that each can be checked against it.

Pinned by: tests/test_oracle.py compares every function here with oracle/_ref (the unmodified reference compiled by
oracle/Makefile) and with the committed fixtures in tests/golden/ (generated from oracle/_ref by tools/make_golden.py);
the Threefry known-answer vectors are those of SURVEY.md Appendix B.  The full floating-point force field is NOT
restated here: for it the oracle is oracle/_ref itself (plus its golden dumps), see DESIGN.md.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

# ---- Random123 threefry4x32-20 (src/include/Random123/threefry.h:110-117,172,183,303-428) -----------------------
_R = [(10, 26), (11, 21), (13, 27), (23, 5), (6, 20), (17, 11), (25, 10), (18, 20)]
_M = 0xFFFFFFFF


def _rotl(x, n):
    return ((x << n) | (x >> (32 - n))) & _M


def threefry4x32(ctr, key):
    ks = [int(k) & _M for k in key] + [0x1BD11BDA]
    for k in key:
        ks[4] ^= int(k) & _M
    X = [(int(c) + ks[i]) & _M for i, c in enumerate(ctr)]
    for r in range(20):
        ra, rb = _R[r % 8]
        if r % 2 == 0:
            X[0] = (X[0] + X[1]) & _M; X[1] = _rotl(X[1], ra) ^ X[0]
            X[2] = (X[2] + X[3]) & _M; X[3] = _rotl(X[3], rb) ^ X[2]
        else:
            X[0] = (X[0] + X[3]) & _M; X[3] = _rotl(X[3], ra) ^ X[0]
            X[2] = (X[2] + X[1]) & _M; X[1] = _rotl(X[1], rb) ^ X[2]
        if r % 4 == 3:
            s = r // 4 + 1
            X = [(X[i] + ks[(s + i) % 5]) & _M for i in range(4)]
            X[3] = (X[3] + s) & _M
    return X


def random_bits(seed, stream, atom, t, call=0):
    """RandomGenerator(seed, stream, atom, t) draw number `call` (src/random.h:32-44)"""
    return threefry4x32([t & _M, (t >> 32) & _M, atom, call], [seed, stream, 0, 0])


def u01(w):      # uniform.hpp:145-154
    return np.float32(np.float32(w) * np.float32(2.0 ** -32) + np.float32(2.0 ** -33))


def uneg11(w):   # uniform.hpp:171-180
    return np.float32(np.float32(np.int32(np.uint32(w))) * np.float32(2.0 ** -31) + np.float32(2.0 ** -32))


def normal3(seed, stream, atom, t):
    """Box-Muller of words (0,1) and (2,3), fourth value dropped (boxmuller.hpp:109-117, random.h:55-66)"""
    b = random_bits(seed, stream, atom, t)
    out = []
    for w0, w1 in ((b[0], b[1]), (b[2], b[3])):
        a = np.float32(np.pi) * uneg11(w0)
        r = np.sqrt(np.float32(-2.0) * np.log(u01(w1)))
        out += [np.sin(a) * r, np.cos(a) * r]
    return np.array(out[:3], dtype='f4')


# ---- pair list (src/interaction_graph.h:122-158 emission order, :223-244 refine predicate) ---------------------------
def acceptable(kind, id1, id2):
    if kind == 'rotamer':
        return (id1.astype(np.uint32) >> 4) != (id2.astype(np.uint32) >> 4)     # bead_interaction.h:195-197
    if kind == 'seq2':
        return np.abs(id1 - id2) > 2                                           # hbond.cpp:254-259
    if kind == 'seq1':
        return np.abs(id1 - id2) > 1                                           # backbone_steric.cpp:32-35
    return np.ones(np.broadcast(id1, id2).shape, dtype=bool)


def pairlist(pos1, id1, pos2, id2, cutoff, kind, symmetric):
    """fp32, no FMA, ((dx*dx+dy*dy)+dz*dz) < cutoff^2 on pos1[i1]-pos2[i2]; order (i1>>2, i2, i1&3)"""
    p1 = np.asarray(pos1, dtype='f4')[:, None, :3]
    p2 = np.asarray(pos2, dtype='f4')[None, :, :3]
    d = p1 - p2
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    c = np.float32(cutoff)
    hit = (d2 < c * c) & acceptable(kind, np.asarray(id1)[:, None], np.asarray(id2)[None, :])
    if symmetric:
        hit &= np.arange(len(id1))[:, None] < np.arange(len(id2))[None, :]
    i1, i2 = np.nonzero(hit)
    order = np.lexsort((i1 & 3, i2, i1 >> 2))
    return np.stack([i1[order], i2[order]], axis=1).astype('i4')


# ---- integrator (src/deriv_engine.cpp:11-35,172-192 with Verlet factors {1,1,1}) -----------------------------------------
def integration_stage(mom, pos, deriv, dt):
    mom = (mom - np.float32(dt) * deriv).astype('f4')
    return mom, (pos + np.float32(dt) * mom).astype('f4')


# ---- compact sigmoid (src/vector_math.h:639-658) -----------------------------------------------------------------------
def compact_sigmoid(x, sharpness):
    y = np.float32(x) * np.float32(sharpness)
    val = np.where(y < -1, 1., np.where(y > 1, 0., 0.25 * (y + 2) * (y - 1) * (y - 1)))
    der = np.where(np.abs(y) > 1, 0., sharpness * 0.75 * (y * y - 1))
    return val.astype('f4'), der.astype('f4')


# ---- replica exchange (src/main.cpp:227-275) ---------------------------------------------------------------------
def attempt_swaps(seed, round_num, swap_sets, temperature, energy_of_slot, slots, stats=None):
    """One ReplicaExchange::attempt_swaps.  `slots[i]` = configuration currently held by system i (any hashable id),
    `energy_of_slot(i, conf)` = potential of configuration `conf` evaluated in system slot i (the reference evaluates the
    energy twice per swap set because the slots may carry different Hamiltonians).  Returns the new slots list; `stats`
    (dict (set,pair) -> [n_success, n_attempt]) is updated if given.  fp32 arithmetic as in the reference."""
    f32 = np.float32
    slots = list(slots)
    n = len(slots)
    beta = [f32(1.) / f32(t) for t in temperature]
    draw = 0
    for si, pairs in enumerate(swap_sets):
        old = [f32(-beta[i] * f32(energy_of_slot(i, slots[i]))) for i in range(n)]
        for a, b in pairs:
            slots[a], slots[b] = slots[b], slots[a]
        new = [f32(-beta[i] * f32(energy_of_slot(i, slots[i]))) for i in range(n)]
        for pi, (a, b) in enumerate(pairs):
            diff = f32(f32(new[a] + new[b]) - f32(old[a] + old[b]))
            reject = False
            if diff < 0:   # the random number is only drawn for uphill exchanges (short-circuit &&, main.cpp:268)
                u = u01(random_bits(seed, 1, 0, round_num, draw)[0])
                draw += 1
                reject = f32(np.exp(diff, dtype=f32)) < u
            if reject:
                slots[a], slots[b] = slots[b], slots[a]
            if stats is not None:
                st = stats.setdefault((si, pi), [0, 0])
                st[1] += 1
                st[0] += 0 if reject else 1
    return slots


def parse_swap_sets(strings):
    return [[tuple(int(x) for x in p.split('-')) for p in s.split(',')] for s in strings]
