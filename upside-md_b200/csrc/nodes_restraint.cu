// Restraint and plumbing nodes that upside_config.py emits beside the ff_1 force field:
// position / tension / AFM springs, radial cavity and flat-bottom z walls, bead contacts (reference src/bonds.cpp,
// src/sidechain_radial.cpp:139-205) and the constant / slice / concat coordinate nodes (bonds.cpp:550-672).
// All are element-wise: one thread per term, blockIdx.y = replica, scatter with float atomics (terms may share atoms).
#include <cmath>

#include "engine.h"

namespace ub {
namespace {

constexpr int TPB = 128;
inline dim3 grid_for(int n_elem, int n_rep) { return dim3((n_elem + TPB - 1) / TPB, n_rep); }

__device__ __forceinline__ void accumulate_potential(float v, float* pot) {
    __shared__ float sc[32];
    v = block_sum(v, sc);
    if (threadIdx.x == 0) atomicAdd(pot + blockIdx.y, v);
}
void check_atoms(const std::vector<int>& a, const CoordNode& n) {
    for (int v : a) if (v < 0 || v >= n.n_elem) throw std::string("atom index out of range");
}

// ---------------------------------------------------------------------------------------------- PosSpring / Tension / AFM
// One kernel for the three "spring to a point" potentials of bonds.cpp:9-168.
//   PosSpring : 0.5 k |x - x0|^2                       (:36-47)
//   Tension   : -x . c, force -c                       (:77-88)
//   AFM       : 0.5 k |x - (tip0 + v t)|^2, t = time_initial + time_step * round_num, round_num advanced by every
//               DerivMode evaluation (:151-166); the counter lives on the device so that graph replays advance it
struct PointTerm { int atom; float k; float x0[3]; float v[3]; };
enum PointKind { POINT_SPRING = 0, POINT_TENSION = 1, POINT_AFM = 2 };
__global__ void k_point_terms(const float* __restrict__ pos, float* __restrict__ sens, float* __restrict__ pot,
                              const PointTerm* __restrict__ prm, int n, int n_atom, int wp, int kind, float time_initial, float time_step,
                              const unsigned long long* __restrict__ round_num, int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        PointTerm p = prm[i];
        const float* x = pos + (size_t(r) * n_atom + p.atom) * wp;
        float* s = sens + (size_t(r) * n_atom + p.atom) * wp;
        f3 xx = ld3(x);
        if (kind == POINT_TENSION) {
            f3 c = mk3(p.x0[0], p.x0[1], p.x0[2]);
            e = -dot(xx, c);
            atomic_add3(s, -c);
        } else {
            f3 target = mk3(p.x0[0], p.x0[1], p.x0[2]);
            if (kind == POINT_AFM) {
                const float t = time_initial + time_step * float((long long)round_num[0]);
                target = target + t * mk3(p.v[0], p.v[1], p.v[2]);
            }
            f3 disp = xx - target;
            e = 0.5f * p.k * mag2(disp);
            atomic_add3(s, p.k * disp);
        }
    }
    if (want_pot) accumulate_potential(e, pot);
}
__global__ void k_bump(unsigned long long* c) { c[0] += 1ull; }

struct PointPotential : PotentialNode {
    CoordNode& pos;
    int n_elem = 0, kind;
    bool always_pot;   // Tension and AFM set `potential` in every mode (bonds.cpp:80-87,158-166)
    float time_initial = 0.f, time_step = 0.f;
    std::vector<PointTerm> h;
    DevBuf<PointTerm> prm;
    DevBuf<unsigned long long> round_num;
    PointPotential(CoordNode& pos_, int kind_) : pos(pos_), kind(kind_), always_pot(kind_ != POINT_SPRING) {}
    void upload() {
        std::vector<int> atoms;
        for (auto& t : h) atoms.push_back(t.atom);
        check_atoms(atoms, pos);
        n_elem = (int)h.size();
        prm.upload(h);
        round_num.upload(std::vector<unsigned long long>(1, 0ull));
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (kind == POINT_AFM && mode == DerivMode) k_bump<<<1, 1, 0, s>>>(round_num.p);
        if (!n_elem) return;
        k_point_terms<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, pos.sens, potential, prm.p, n_elem, pos.n_elem, pos.wp, kind,
                                                                       time_initial, time_step, round_num.p,
                                                                       always_pot || mode == PotentialAndDerivMode);
    }
};
struct PosSpring : PointPotential {
    PosSpring(Engine&, const h5l::Node& g, CoordNode& pos_) : PointPotential(pos_, POINT_SPRING) {
        int n = (int)h5_dims(g, "id", 1)[0];
        h5_check_size(g, "id", {(uint64_t)n});
        h5_check_size(g, "x0", {(uint64_t)n, 3});
        h5_check_size(g, "spring_const", {(uint64_t)n});
        auto id = h5_read<int>(g, "id");
        auto x0 = h5_read<float>(g, "x0");
        auto k = h5_read<float>(g, "spring_const");
        for (int i = 0; i < n; ++i) h.push_back(PointTerm{id[i], k[i], {x0[3 * i], x0[3 * i + 1], x0[3 * i + 2]}, {0.f, 0.f, 0.f}});
        upload();
    }
};
RegisterNodeType<PosSpring, 1> pos_spring_node("atom_pos_spring");
struct TensionPotential : PointPotential {
    TensionPotential(Engine&, const h5l::Node& g, CoordNode& pos_) : PointPotential(pos_, POINT_TENSION) {
        int n = (int)h5_dims(g, "atom", 1)[0];
        h5_check_size(g, "atom", {(uint64_t)n});
        h5_check_size(g, "tension_coeff", {(uint64_t)n, 3});
        auto id = h5_read<int>(g, "atom");
        auto c = h5_read<float>(g, "tension_coeff");
        for (int i = 0; i < n; ++i) h.push_back(PointTerm{id[i], 0.f, {c[3 * i], c[3 * i + 1], c[3 * i + 2]}, {0.f, 0.f, 0.f}});
        upload();
    }
};
RegisterNodeType<TensionPotential, 1> tension_node("tension");
struct AFMPotential : PointPotential {
    AFMPotential(Engine&, const h5l::Node& g, CoordNode& pos_) : PointPotential(pos_, POINT_AFM) {
        time_initial = h5_attr<float>(g, "pulling_vel", "time_initial");
        time_step = h5_attr<float>(g, "pulling_vel", "time_step");
        int n = (int)h5_dims(g, "atom", 1)[0];
        h5_check_size(g, "atom", {(uint64_t)n});
        h5_check_size(g, "spring_const", {(uint64_t)n});
        h5_check_size(g, "starting_tip_pos", {(uint64_t)n, 3});
        h5_check_size(g, "pulling_vel", {(uint64_t)n, 3});
        auto id = h5_read<int>(g, "atom");
        auto k = h5_read<float>(g, "spring_const");
        auto x0 = h5_read<float>(g, "starting_tip_pos");
        auto v = h5_read<float>(g, "pulling_vel");
        for (int i = 0; i < n; ++i)
            h.push_back(PointTerm{id[i], k[i], {x0[3 * i], x0[3 * i + 1], x0[3 * i + 2]}, {v[3 * i], v[3 * i + 1], v[3 * i + 2]}});
        upload();
    }
    std::vector<float> get_value_by_name(int, const char* log_name) override {   // the reference's loggers (bonds.cpp:129-146)
        std::string nm(log_name);
        engine->sync_and_check();
        unsigned long long rn = 0;
        UB_CUDA(cudaMemcpy(&rn, round_num.p, sizeof(rn), cudaMemcpyDeviceToHost));
        const float t = time_initial + time_step * float((long long)rn);
        if (nm == "time_estimate") return {t};
        if (nm == "tip_pos") {
            std::vector<float> out;
            for (auto& p : h) for (int d = 0; d < 3; ++d) out.push_back(p.x0[d] + p.v[d] * t);
            return out;
        }
        throw std::string("Value ") + log_name + " not implemented";
    }
};
RegisterNodeType<AFMPotential, 1> AFM_node("AFM");

// ---------------------------------------------------------------------------------------------- CavityRadial / ZFlatBottom
// bonds.cpp:355-372 (harmonic wall outside `radius` from the origin) and :406-424 (flat-bottomed harmonic well in z)
struct WallTerm { int atom; float z0, radius, k; };
__global__ void k_walls(const float* __restrict__ pos, float* __restrict__ sens, float* __restrict__ pot, const WallTerm* __restrict__ prm,
                        int n, int n_atom, int wp, int z_only, int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        WallTerm p = prm[i];
        const float* x = pos + (size_t(r) * n_atom + p.atom) * wp;
        float* s = sens + (size_t(r) * n_atom + p.atom) * wp;
        if (z_only) {
            const float dz = x[2] - p.z0;
            const float excess = dz > p.radius ? dz - p.radius : (dz < -p.radius ? dz + p.radius : 0.f);
            if (excess != 0.f) atomicAdd(s + 2, p.k * excess);
            e = 0.5f * p.k * excess * excess;
        } else {
            f3 xx = ld3(x);
            const float r2 = mag2(xx);
            if (r2 > p.radius * p.radius) {
                const float inv_r = rsqrtf(r2), rr = r2 * inv_r, excess = rr - p.radius;
                e = 0.5f * p.k * excess * excess;
                atomic_add3(s, (p.k * excess * inv_r) * xx);
            }
        }
    }
    if (want_pot) accumulate_potential(e, pot);
}
struct WallPotential : PotentialNode {
    CoordNode& pos;
    int n_elem = 0, z_only;
    DevBuf<WallTerm> prm;
    WallPotential(CoordNode& pos_, int z_only_) : pos(pos_), z_only(z_only_) {}
    void set(const std::vector<WallTerm>& h) {
        std::vector<int> atoms;
        for (auto& t : h) atoms.push_back(t.atom);
        check_atoms(atoms, pos);
        n_elem = (int)h.size();
        prm.upload(h);
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!n_elem) return;
        k_walls<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, pos.sens, potential, prm.p, n_elem, pos.n_elem, pos.wp, z_only,
                                                                 mode == PotentialAndDerivMode);
    }
};
struct CavityRadial : WallPotential {
    CavityRadial(Engine&, const h5l::Node& g, CoordNode& pos_) : WallPotential(pos_, 0) {
        int n = (int)h5_dims(g, "id", 1)[0];
        h5_check_size(g, "id", {(uint64_t)n});
        h5_check_size(g, "radius", {(uint64_t)n});
        h5_check_size(g, "spring_constant", {(uint64_t)n});
        auto id = h5_read<int>(g, "id");
        auto rad = h5_read<float>(g, "radius");
        auto k = h5_read<float>(g, "spring_constant");
        std::vector<WallTerm> h;
        for (int i = 0; i < n; ++i) h.push_back(WallTerm{id[i], 0.f, rad[i], k[i]});
        set(h);
    }
};
RegisterNodeType<CavityRadial, 1> cavity_radial_node("cavity_radial");
struct ZFlatBottom : WallPotential {
    ZFlatBottom(Engine&, const h5l::Node& g, CoordNode& pos_) : WallPotential(pos_, 1) {
        int n = (int)h5_dims(g, "atom", 1)[0];
        h5_check_size(g, "atom", {(uint64_t)n});
        h5_check_size(g, "z0", {(uint64_t)n});
        h5_check_size(g, "radius", {(uint64_t)n});
        h5_check_size(g, "spring_constant", {(uint64_t)n});
        auto id = h5_read<int>(g, "atom");
        auto z0 = h5_read<float>(g, "z0");
        auto rad = h5_read<float>(g, "radius");
        auto k = h5_read<float>(g, "spring_constant");
        std::vector<WallTerm> h;
        for (int i = 0; i < n; ++i) h.push_back(WallTerm{id[i], z0[i], rad[i], k[i]});
        set(h);
    }
};
RegisterNodeType<ZFlatBottom, 1> z_flat_bottom_node("z_flat_bottom");

// ---------------------------------------------------------------------------------------------- ContactEnergy
// sidechain_radial.cpp:186-204: energy * compact_sigmoid(|x0-x1| - dist, 1/width) for listed bead pairs
struct ContactTerm { int a0, a1; float energy, dist, scale, cutoff; };
__global__ void k_contact(const float* __restrict__ pos, float* __restrict__ sens, float* __restrict__ pot, const ContactTerm* __restrict__ prm,
                          int n, int n_elem, int wp, int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        ContactTerm p = prm[i];
        const float* x = pos + size_t(r) * n_elem * wp;
        float* s = sens + size_t(r) * n_elem * wp;
        f3 disp = ld3(x + size_t(p.a0) * wp) - ld3(x + size_t(p.a1) * wp);
        const float dist = sqrtf(mag2(disp));
        if (dist < p.cutoff) {
            float v, d;
            compact_sigmoid(dist - p.dist, p.scale, v, d);
            e = p.energy * v;
            f3 deriv = (p.energy * d / dist) * disp;
            atomic_add3(s + size_t(p.a0) * wp, deriv);
            atomic_add3(s + size_t(p.a1) * wp, -deriv);
        }
    }
    if (want_pot) accumulate_potential(e, pot);
}
struct ContactEnergy : PotentialNode {
    CoordNode& bead_pos;
    int n_contact;
    DevBuf<ContactTerm> prm;
    ContactEnergy(Engine&, const h5l::Node& g, CoordNode& bp) : bead_pos(bp) {
        n_contact = (int)h5_dims(g, "id", 2)[0];
        h5_check_size(g, "id", {(uint64_t)n_contact, 2});
        h5_check_size(g, "energy", {(uint64_t)n_contact});
        h5_check_size(g, "distance", {(uint64_t)n_contact});
        h5_check_size(g, "width", {(uint64_t)n_contact});
        check_elem_width_lower_bound(bead_pos, 3);
        auto id = h5_read<int>(g, "id");
        auto en = h5_read<float>(g, "energy");
        auto di = h5_read<float>(g, "distance");
        auto wi = h5_read<float>(g, "width");
        check_atoms(id, bead_pos);
        std::vector<ContactTerm> h;
        for (int i = 0; i < n_contact; ++i) {
            float scale = 1.f / wi[i];
            h.push_back(ContactTerm{id[2 * i], id[2 * i + 1], en[i], di[i], scale, di[i] + 1.f / scale});
        }
        prm.upload(h);
    }
    void compute_value(cudaStream_t s, ComputeMode) override {   // `potential` is set in every mode (:190)
        if (!n_contact) return;
        k_contact<<<grid_for(n_contact, engine->n_rep), TPB, 0, s>>>(bead_pos.output, bead_pos.sens, potential, prm.p, n_contact,
                                                                      bead_pos.n_elem, bead_pos.wp, 1);
    }
};
RegisterNodeType<ContactEnergy, 1> contact_node("contact");

// ---------------------------------------------------------------------------------------------- constant / slice / concat
// gather rows of `src` (row index from `map`, -1 = constant table) into `dst`; with `back` the sens rows flow the other way
__global__ void k_rows_forward(float* __restrict__ dst, int n_dst, int wp_dst, const float* __restrict__ src, int n_src, int wp_src,
                               const int* __restrict__ map, int dst_offset, int n, int width, int src_batched) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    const int j = map ? map[i] : i;
    const float* s = src + ((src_batched ? size_t(r) * n_src : 0) + j) * wp_src;
    float* d = dst + (size_t(r) * n_dst + dst_offset + i) * wp_dst;
    for (int k = 0; k < width; ++k) d[k] = s[k];
}
__global__ void k_rows_backward(const float* __restrict__ dst_sens, int n_dst, int wp_dst, float* __restrict__ src_sens, int n_src, int wp_src,
                                const int* __restrict__ map, int dst_offset, int n, int width) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    const int j = map ? map[i] : i;
    const float* d = dst_sens + (size_t(r) * n_dst + dst_offset + i) * wp_dst;
    float* s = src_sens + (size_t(r) * n_src + j) * wp_src;
    for (int k = 0; k < width; ++k) atomicAdd(s + k, d[k]);   // a slice may list an element twice
}

// bonds.cpp:550-587: the same (n_elem, elem_width) table for every replica; get/set_param = the table
struct ConstantCoord : CoordNode {
    std::vector<float> value;   // un-padded rows
    DevBuf<float> d_value;      // padded rows
    ConstantCoord(Engine&, const h5l::Node& g) : CoordNode((int)h5_dims(g, "value", 2)[0], (int)h5_dims(g, "value", 2)[1]) {
        value = h5_read<float>(g, "value");
        upload_value();
    }
    void upload_value() {
        std::vector<float> p(size_t(n_elem) * wp, 0.f);
        for (int e = 0; e < n_elem; ++e) for (int d = 0; d < elem_width; ++d) p[size_t(e) * wp + d] = value[size_t(e) * elem_width + d];
        d_value.upload(p);
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        k_rows_forward<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(output, n_elem, wp, d_value.p, n_elem, wp, nullptr, 0, n_elem, elem_width, 0);
    }
    void propagate_deriv(cudaStream_t) override {}
    std::vector<float> get_param() const override { return value; }
    void set_param(const std::vector<float>& p) override {
        if (p.size() != value.size()) throw std::string("invalid size to set_param");
        value = p;
        upload_value();
    }
};
RegisterNodeType<ConstantCoord, 0> constant_coord_node("constant");

// bonds.cpp:589-621
struct Slice : CoordNode {
    CoordNode& src;
    DevBuf<int> d_id;
    Slice(Engine&, const h5l::Node& g, CoordNode& src_) : CoordNode((int)h5_dims(g, "id", 1)[0], src_.elem_width), src(src_) {
        h5_check_size(g, "id", {(uint64_t)n_elem});
        auto id = h5_read<int>(g, "id");
        check_atoms(id, src);
        d_id.upload(id);
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        k_rows_forward<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(output, n_elem, wp, src.output, src.n_elem, src.wp, d_id.p, 0, n_elem,
                                                                        elem_width, 1);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_rows_backward<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(sens, n_elem, wp, src.sens, src.n_elem, src.wp, d_id.p, 0, n_elem,
                                                                         elem_width);
    }
};
RegisterNodeType<Slice, 1> slice_node("slice");

// bonds.cpp:623-672: rows of all arguments, one argument after the other.  (The reference's constructor insists on equal
// element counts and its propagate_deriv indexes the argument's sens with the running output row; what is implemented
// here is what those loops are for: every argument receives the sens of its own rows.)
struct Concat : CoordNode {
    std::vector<CoordNode*> args;
    static int total(const ArgList& a) { int n = 0; for (auto c : a) n += c->n_elem; return n; }
    Concat(Engine&, const h5l::Node&, const ArgList& a) : CoordNode(total(a), a.empty() ? 0 : a[0]->elem_width), args(a) {
        if (args.empty()) throw std::string("concat needs at least one argument");
        for (auto c : args) if (c->elem_width != elem_width) throw std::string("Coord node elem_width mismatch");
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        int off = 0;
        for (auto c : args) {
            if (c->n_elem)
                k_rows_forward<<<grid_for(c->n_elem, engine->n_rep), TPB, 0, s>>>(output, n_elem, wp, c->output, c->n_elem, c->wp, nullptr, off,
                                                                                   c->n_elem, elem_width, 1);
            off += c->n_elem;
        }
    }
    void propagate_deriv(cudaStream_t s) override {
        int off = 0;
        for (auto c : args) {
            if (c->n_elem)
                k_rows_backward<<<grid_for(c->n_elem, engine->n_rep), TPB, 0, s>>>(sens, n_elem, wp, c->sens, c->n_elem, c->wp, nullptr, off,
                                                                                    c->n_elem, elem_width);
            off += c->n_elem;
        }
    }
};
RegisterNodeType<Concat, -1> concat_node("concat");

}  // namespace
}  // namespace ub
