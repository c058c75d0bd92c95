// Element-wise nodes of the ff_1 graph, batched over replicas: bonded springs, Rama coordinates and maps,
// affine alignment, virtual H/O sites, placements, weighted positions, non-linear coupling, H-bond energy.
// Each kernel cites the reference function it restates.  Grid convention: blockIdx.y = replica,
// blockIdx.x*blockDim.x + threadIdx.x = element.
#include <cmath>

#include "engine.h"
#include "spline_fit.h"

namespace ub {
namespace {

constexpr int TPB = 128;
inline dim3 grid_for(int n_elem, int n_rep) { return dim3((n_elem + TPB - 1) / TPB, n_rep); }

// host copy of replicas [r0,r1) of a node's output or sens rows (padded rows, as on the device); accessors only
static std::vector<float> download_rows(const CoordNode& n, const float* base, int r0, int r1) {
    std::vector<float> h(size_t(r1 - r0) * n.stride());
    if (!h.empty()) UB_CUDA(cudaMemcpy(h.data(), base + size_t(r0) * n.stride(), h.size() * sizeof(float), cudaMemcpyDeviceToHost));
    return h;
}

// block partial sum -> one atomicAdd per block into pot[replica]
__device__ __forceinline__ void accumulate_potential(float v, float* pot) {
    __shared__ float sc[32];
    v = block_sum(v, sc);
    if (threadIdx.x == 0) atomicAdd(pot + blockIdx.y, v);
}

// ---------------------------------------------------------------------------------------------- DistSpring
// reference bonds.cpp:297-318
struct SpringParam2 { int a0, a1; float equil, k; };
__global__ void k_dist_spring(const float* __restrict__ pos, float* __restrict__ sens, float* __restrict__ pot,
                              const SpringParam2* __restrict__ prm, int n, int n_atom, int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        SpringParam2 p = prm[i];
        const float* x = pos + size_t(r) * n_atom * 4;
        float* s = sens + size_t(r) * n_atom * 4;
        f3 disp = ld3v(x + 4 * p.a0) - ld3v(x + 4 * p.a1);
        float d2 = mag2(disp);
        float inv = rsqrtf(d2);
        f3 deriv = (p.k * (1.f - p.equil * inv)) * disp;
        float dm = d2 * inv - p.equil;
        e = 0.5f * p.k * dm * dm;
        atomic_add3v(s + 4 * p.a0, deriv);
        atomic_add3v(s + 4 * p.a1, -deriv);
    }
    if (want_pot) accumulate_potential(e, pot);
}

struct DistSpring : PotentialNode {
    CoordNode& pos;
    int n_elem;
    DevBuf<SpringParam2> prm;
    std::vector<SpringParam2> h_prm;
    std::vector<int> bonded_atoms;
    DistSpring(Engine&, const h5l::Node& g, CoordNode& pos_) : pos(pos_) {
        n_elem = (int)h5_dims(g, "id", 2)[0];
        h5_check_size(g, "id", {(uint64_t)n_elem, 2});
        h5_check_size(g, "equil_dist", {(uint64_t)n_elem});
        h5_check_size(g, "spring_const", {(uint64_t)n_elem});
        h5_check_size(g, "bonded_atoms", {(uint64_t)n_elem});
        auto id = h5_read<int>(g, "id");
        auto eq = h5_read<float>(g, "equil_dist");
        auto k = h5_read<float>(g, "spring_const");
        std::vector<SpringParam2> h(n_elem);
        for (int i = 0; i < n_elem; ++i) {
            h[i] = {id[2 * i], id[2 * i + 1], eq[i], k[i]};
            if (h[i].a0 < 0 || h[i].a0 >= pos.n_elem || h[i].a1 < 0 || h[i].a1 >= pos.n_elem) throw std::string("atom index out of range");
        }
        prm.upload(h);
        h_prm = h;
        bonded_atoms = h5_read<int>(g, "bonded_atoms");
    }
    // bonds.cpp:280-295: energy of the springs that are not covalent bonds
    void add_loggers(int level, std::vector<NodeLogger>& out) override {
        if (level < 1) return;
        out.push_back({"nonbonded_spring_energy", {1}, false, [this](int r) {
            auto x = pos.host_rows(pos.output, r);
            float pot = 0.f;
            for (int nt = 0; nt < n_elem; ++nt) {
                if (bonded_atoms[nt]) continue;
                const SpringParam2& p = h_prm[nt];
                const float* a = &x[size_t(p.a0) * pos.wp];
                const float* b = &x[size_t(p.a1) * pos.wp];
                float d = std::sqrt((a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]));
                pot += 0.5f * p.k * (d - p.equil) * (d - p.equil);
            }
            return std::vector<float>{pot};
        }});
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!n_elem) return;
        k_dist_spring<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, pos.sens, potential, prm.p, n_elem,
                                                                       pos.n_elem, mode == PotentialAndDerivMode);
    }
};
RegisterNodeType<DistSpring, 1> dist_spring_node("dist_spring");

// ---------------------------------------------------------------------------------------------- AngleSpring
// reference bonds.cpp:457-487 (third atom is the vertex; equil is cos(theta0))
struct SpringParam3 { int a0, a1, a2; float equil, k; };
__global__ void k_angle_spring(const float* __restrict__ pos, float* __restrict__ sens, float* __restrict__ pot,
                               const SpringParam3* __restrict__ prm, int n, int n_atom, int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        SpringParam3 p = prm[i];
        const float* x = pos + size_t(r) * n_atom * 4;
        float* s = sens + size_t(r) * n_atom * 4;
        f3 v = ld3v(x + 4 * p.a2);
        f3 x1 = ld3v(x + 4 * p.a0) - v, x2 = ld3v(x + 4 * p.a1) - v;
        float i1 = rsqrtf(mag2(x1)), i2 = rsqrtf(mag2(x2));
        f3 h1 = i1 * x1, h2 = i2 * x2;
        float dp = dot(h1, h2);
        float pre = p.k * (dp - p.equil);
        f3 d1 = (pre * i1) * (h2 - dp * h1);
        f3 d2 = (pre * i2) * (h1 - dp * h2);
        atomic_add3v(s + 4 * p.a0, d1);
        atomic_add3v(s + 4 * p.a1, d2);
        atomic_add3v(s + 4 * p.a2, -(d1 + d2));
        e = 0.5f * p.k * (dp - p.equil) * (dp - p.equil);
    }
    if (want_pot) accumulate_potential(e, pot);
}
struct AngleSpring : PotentialNode {
    CoordNode& pos;
    int n_elem;
    DevBuf<SpringParam3> prm;
    AngleSpring(Engine&, const h5l::Node& g, CoordNode& pos_) : pos(pos_) {
        n_elem = (int)h5_dims(g, "id", 2)[0];
        h5_check_size(g, "id", {(uint64_t)n_elem, 3});
        h5_check_size(g, "equil_dist", {(uint64_t)n_elem});
        h5_check_size(g, "spring_const", {(uint64_t)n_elem});
        auto id = h5_read<int>(g, "id");
        auto eq = h5_read<float>(g, "equil_dist");
        auto k = h5_read<float>(g, "spring_const");
        std::vector<SpringParam3> h(n_elem);
        for (int i = 0; i < n_elem; ++i) {
            h[i] = {id[3 * i], id[3 * i + 1], id[3 * i + 2], eq[i], k[i]};
            for (int a : {h[i].a0, h[i].a1, h[i].a2}) if (a < 0 || a >= pos.n_elem) throw std::string("atom index out of range");
        }
        prm.upload(h);
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!n_elem) return;
        k_angle_spring<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, pos.sens, potential, prm.p, n_elem,
                                                                        pos.n_elem, mode == PotentialAndDerivMode);
    }
};
RegisterNodeType<AngleSpring, 1> angle_spring_node("angle_spring");

// ---------------------------------------------------------------------------------------------- dihedrals
// (dihedral_germ: common.cuh)
// reference bonds.cpp:519-545
struct SpringParam4 { int a[4]; float equil, k; };
__global__ void k_dihedral_spring(const float* __restrict__ pos, float* __restrict__ sens, float* __restrict__ pot,
                                  const SpringParam4* __restrict__ prm, int n, int n_atom, int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        SpringParam4 p = prm[i];
        const float* x = pos + size_t(r) * n_atom * 4;
        float* s = sens + size_t(r) * n_atom * 4;
        f3 d[4];
        float dih = dihedral_germ(ld3v(x + 4 * p.a[0]), ld3v(x + 4 * p.a[1]), ld3v(x + 4 * p.a[2]), ld3v(x + 4 * p.a[3]),
                                  d[0], d[1], d[2], d[3]);
        const float PI = 3.1415926535897932f;
        float disp = dih - p.equil;
        disp = (disp > PI) ? disp - 2.f * PI : disp;
        disp = (disp < -PI) ? disp + 2.f * PI : disp;
        float sc = p.k * disp;
#pragma unroll
        for (int a = 0; a < 4; ++a) atomic_add3v(s + 4 * p.a[a], sc * d[a]);
        e = 0.5f * p.k * disp * disp;
    }
    if (want_pot) accumulate_potential(e, pot);
}
struct DihedralSpring : PotentialNode {
    CoordNode& pos;
    int n_elem;
    DevBuf<SpringParam4> prm;
    DihedralSpring(Engine&, const h5l::Node& g, CoordNode& pos_) : pos(pos_) {
        n_elem = (int)h5_dims(g, "id", 2)[0];
        h5_check_size(g, "id", {(uint64_t)n_elem, 4});
        h5_check_size(g, "equil_dist", {(uint64_t)n_elem});
        h5_check_size(g, "spring_const", {(uint64_t)n_elem});
        auto id = h5_read<int>(g, "id");
        auto eq = h5_read<float>(g, "equil_dist");
        auto k = h5_read<float>(g, "spring_const");
        std::vector<SpringParam4> h(n_elem);
        for (int i = 0; i < n_elem; ++i) {
            for (int a = 0; a < 4; ++a) {
                h[i].a[a] = id[4 * i + a];
                if (h[i].a[a] < 0 || h[i].a[a] >= pos.n_elem) throw std::string("atom index out of range");
            }
            h[i].equil = eq[i];
            h[i].k = k[i];
        }
        prm.upload(h);
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!n_elem) return;
        k_dihedral_spring<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, pos.sens, potential, prm.p, n_elem,
                                                                           pos.n_elem, mode == PotentialAndDerivMode);
    }
};
RegisterNodeType<DihedralSpring, 1> dihedral_spring_node("dihedral_spring");

// ---------------------------------------------------------------------------------------------- RamaCoord
// reference bonds.cpp:171-249: phi,psi per residue (dummy angle -1.3963 at the termini) and its Jacobian
struct RamaParam { int atom[5]; int dummy0, dummy1; };
__global__ void k_rama_coord(const float* __restrict__ pos, float* __restrict__ out, const RamaParam* __restrict__ prm, int n,
                             int n_atom) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    RamaParam p = prm[i];
    const float* x = pos + size_t(r) * n_atom * 4;
    f3 a[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) a[k] = ld3v(x + 4 * p.atom[k]);
    f3 d[4];
    float phi = -1.3963f, psi = -1.3963f;
    if (!p.dummy0) phi = dihedral_germ(a[0], a[1], a[2], a[3], d[0], d[1], d[2], d[3]);
    if (!p.dummy1) psi = dihedral_germ(a[1], a[2], a[3], a[4], d[0], d[1], d[2], d[3]);
    reinterpret_cast<float2*>(out)[size_t(r) * n + i] = make_float2(phi, psi);
}
// the Jacobian (2 x 5 x 3, bonds.cpp:222-247) is recomputed from the positions: ~200 flops cost less than storing and
// re-reading 120 bytes per residue with a 120-byte stride between threads
__global__ void k_rama_coord_deriv(const float* __restrict__ pos, float* __restrict__ pos_sens, const float* __restrict__ sens,
                                   const RamaParam* __restrict__ prm, int n, int n_atom) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    RamaParam p = prm[i];
    const float2 sn = reinterpret_cast<const float2*>(sens)[size_t(r) * n + i];
    const float* x = pos + size_t(r) * n_atom * 4;
    f3 a[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) a[k] = ld3v(x + 4 * p.atom[k]);
    f3 v[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) v[k] = mk3(0.f, 0.f, 0.f);
    if (!p.dummy0) {
        f3 d[4];
        dihedral_germ(a[0], a[1], a[2], a[3], d[0], d[1], d[2], d[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] += sn.x * d[k];
    }
    if (!p.dummy1) {
        f3 d[4];
        dihedral_germ(a[1], a[2], a[3], a[4], d[0], d[1], d[2], d[3]);
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k + 1] += sn.y * d[k];
    }
    float* ps = pos_sens + size_t(r) * n_atom * 4;
#pragma unroll
    for (int k = 0; k < 5; ++k) {
        bool used = (k < 4 && !p.dummy0) || (k > 0 && !p.dummy1);
        if (used) atomic_add3v(ps + 4 * p.atom[k], v[k]);
    }
}
struct RamaCoord : CoordNode {
    CoordNode& pos;
    DevBuf<RamaParam> prm;
    RamaCoord(Engine&, const h5l::Node& g, CoordNode& pos_) : CoordNode((int)h5_dims(g, "id", 2)[0], 2), pos(pos_) {
        h5_check_size(g, "id", {(uint64_t)n_elem, 5});
        auto id = h5_read<int>(g, "id");
        std::vector<RamaParam> h(n_elem);
        for (int i = 0; i < n_elem; ++i) {
            for (int k = 0; k < 5; ++k) h[i].atom[k] = id[5 * i + k];
            h[i].dummy0 = h[i].atom[0] == -1;
            h[i].dummy1 = h[i].atom[4] == -1;
            if (h[i].dummy0) h[i].atom[0] = 0;
            if (h[i].dummy1) h[i].atom[4] = 0;
            for (int k = 0; k < 5; ++k) if (h[i].atom[k] < 0 || h[i].atom[k] >= pos.n_elem) throw std::string("atom index out of range");
        }
        prm.upload(h);
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        k_rama_coord<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, output, prm.p, n_elem, pos.n_elem);
    }
    void add_loggers(int level, std::vector<NodeLogger>& out) override {   // bonds.cpp:199-202
        if (level < 1) return;
        out.push_back({"rama", {(uint64_t)n_elem, 2}, false, [this](int r) {
            auto o = host_rows(output, r);
            std::vector<float> v(size_t(n_elem) * 2);
            for (int i = 0; i < n_elem; ++i) { v[2 * i] = o[size_t(i) * wp]; v[2 * i + 1] = o[size_t(i) * wp + 1]; }
            return v;
        }});
    }
    void propagate_deriv(cudaStream_t s) override {
        k_rama_coord_deriv<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, pos.sens, sens, prm.p, n_elem, pos.n_elem);
    }
};
RegisterNodeType<RamaCoord, 1> rama_coord_node("rama_coord");

// ---------------------------------------------------------------------------------------------- bicubic maps
// polynomial-form periodic bicubic patch; reference spline.h:61-80
__device__ __forceinline__ void bicubic_vd(const float* __restrict__ c, float fx, float fy, float& value, float& dx, float& dy) {
    float c_[16];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float4 v = __ldg(reinterpret_cast<const float4*>(c) + q);
        c_[4 * q] = v.x; c_[4 * q + 1] = v.y; c_[4 * q + 2] = v.z; c_[4 * q + 3] = v.w;
    }
    float fx2 = fx * fx, fx3 = fx * fx2, fy2 = fy * fy;
    float vx0 = c_[0] + fy * (c_[1] + fy * (c_[2] + fy * c_[3]));
    float vx1 = c_[4] + fy * (c_[5] + fy * (c_[6] + fy * c_[7]));
    float vx2 = c_[8] + fy * (c_[9] + fy * (c_[10] + fy * c_[11]));
    float vx3 = c_[12] + fy * (c_[13] + fy * (c_[14] + fy * c_[15]));
    float vy1 = c_[1] + fx * (c_[5] + fx * (c_[9] + fx * c_[13]));
    float vy2 = c_[2] + fx * (c_[6] + fx * (c_[10] + fx * c_[14]));
    float vy3 = c_[3] + fx * (c_[7] + fx * (c_[11] + fx * c_[15]));
    dx = vx1 + 2.f * fx * vx2 + 3.f * fx2 * vx3;
    dy = vy1 + 2.f * fy * vy2 + 3.f * fy2 * vy3;
    value = vx0 + fx * vx1 + fx2 * vx2 + fx3 * vx3;
}
__device__ __forceinline__ void rama_cell(float phi, float psi, int nx, int ny, int& xb, int& yb, float& fx, float& fy,
                                          float& scale_x, float& scale_y) {
    const float PI = 3.1415926535897932f;
    scale_x = nx * (0.5f / PI - 1e-7f);
    scale_y = ny * (0.5f / PI - 1e-7f);
    float x = (phi + PI) * scale_x, y = (psi + PI) * scale_y;
    xb = min(max((int)x, 0), nx - 1);
    yb = min(max((int)y, 0), ny - 1);
    fx = x - xb;
    fy = y - yb;
}

// reference rama_map_pot.cpp:57-82
__global__ void k_rama_map_pot(const float* __restrict__ rama, float* __restrict__ rama_sens, float* __restrict__ pot,
                               float* __restrict__ residue_pot, const int* __restrict__ residue,
                               const int* __restrict__ map_id, const float* __restrict__ coeff, int n, int n_rama, int nx,
                               int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        int res = residue[i];
        const float* rp = rama + (size_t(r) * n_rama + res) * 2;
        int xb, yb;
        float fx, fy, sx, sy;
        rama_cell(rp[0], rp[1], nx, nx, xb, yb, fx, fy, sx, sy);
        float v, dx, dy;
        bicubic_vd(coeff + ((size_t(map_id[i]) * nx + xb) * nx + yb) * 16, fx, fy, v, dx, dy);
        float* rs = rama_sens + (size_t(r) * n_rama + res) * 2;
        atomicAdd(rs + 0, dx * sx);
        atomicAdd(rs + 1, dy * sy);
        e = v;
        if (want_pot) residue_pot[size_t(r) * n + i] = v;
    }
    if (want_pot) accumulate_potential(e, pot);
}
struct RamaMapPot : PotentialNode {
    CoordNode& rama;
    int n_residue, n_layer, nx;
    DevBuf<int> residue, map_id;
    DevBuf<float> coeff, residue_pot;
    bool log_pot = true;   // attribute log_pot = 0: never log the per-residue potential (rama_map_pot.cpp:22,34)
    RamaMapPot(Engine&, const h5l::Node& g, CoordNode& rama_) : rama(rama_) {
        if (g.attrs.count("log_pot")) log_pot = h5_attr<int>(g, ".", "log_pot") != 0;
        check_elem_width(rama, 2);
        n_residue = (int)h5_dims(g, "residue_id", 1)[0];
        auto d = h5_dims(g, "rama_pot", 3);
        n_layer = (int)d[0];
        nx = (int)d[1];
        if (d[1] != d[2]) throw std::string("must have same x and y grid spacing for Rama maps");
        h5_check_size(g, "rama_map_id", {(uint64_t)n_residue});
        auto res = h5_read<int>(g, "residue_id");
        auto mid = h5_read<int>(g, "rama_map_id");
        for (int i = 0; i < n_residue; ++i) {
            if (res[i] < 0 || res[i] >= rama.n_elem) throw std::string("residue index out of range");
            if (mid[i] < 0 || mid[i] >= n_layer) throw std::string("rama_map_id out of range");
        }
        residue.upload(res);
        map_id.upload(mid);
        fit(h5_read<double>(g, "rama_pot"));
    }
    void fit(const std::vector<double>& raw) { coeff.upload(fit_periodic_spline_2d(n_layer, nx, nx, 1, raw.data())); }
    void finalize() override { residue_pot.alloc(size_t(engine->n_rep) * n_residue); }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!n_residue) return;
        k_rama_map_pot<<<grid_for(n_residue, engine->n_rep), TPB, 0, s>>>(rama.output, rama.sens, potential, residue_pot.p,
                                                                           residue.p, map_id.p, coeff.p, n_residue,
                                                                           rama.n_elem, nx, mode == PotentialAndDerivMode);
    }
    void add_loggers(int level, std::vector<NodeLogger>& out) override {   // rama_map_pot.cpp:50-54
        if (level < 1 || !log_pot) return;
        out.push_back({"rama_map_potential", {(uint64_t)n_residue}, false, [this](int r) {
            std::vector<float> v(n_residue);
            if (n_residue) UB_CUDA(cudaMemcpy(v.data(), residue_pot.p + size_t(r) * n_residue, v.size() * sizeof(float), cudaMemcpyDeviceToHost));
            return v;
        }});
    }
    void set_param(const std::vector<float>& p) override {
        if (p.size() != size_t(n_layer) * nx * nx) throw std::string("wrong number of parameters");
        fit(std::vector<double>(p.begin(), p.end()));
    }
};
RegisterNodeType<RamaMapPot, 1> rama_map_pot_node("rama_map_pot");

// ---------------------------------------------------------------------------------------------- AffineAlignment
// reference eig.cpp:277-471.  One thread per residue: centroid, 3x3 correlation with the reference geometry, largest
// eigenvector of the 4x4 symmetric quaternion matrix (cyclic Jacobi here instead of Householder+QR; any accurate
// symmetric eigensolver gives the same rotation), all four eigenpairs kept for the backward pass.
__device__ void jacobi_eig4(float A[4][4], float V[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) V[i][j] = (i == j) ? 1.f : 0.f;
    for (int sweep = 0; sweep < 8; ++sweep) {
        float off = 0.f, diag = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            diag += A[i][i] * A[i][i];
#pragma unroll
            for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-14f * diag) break;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                float apq = A[p][q];
                if (fabsf(apq) < 1e-30f) continue;
                float theta = (A[q][q] - A[p][p]) / (2.f * apq);
                float t = (theta >= 0.f ? 1.f : -1.f) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
                float c = rsqrtf(t * t + 1.f), s = t * c;
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // A <- A J
                    float akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // A <- J^T A
                    float apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // V <- V J   (columns are eigenvectors)
                    float vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
}
__device__ __forceinline__ void fill_F(float F[4][4], const float R[3][3]) {
    F[0][0] = R[0][0] + R[1][1] + R[2][2]; F[0][1] = R[1][2] - R[2][1]; F[0][2] = R[2][0] - R[0][2]; F[0][3] = R[0][1] - R[1][0];
    F[1][1] = R[0][0] - R[1][1] - R[2][2]; F[1][2] = R[0][1] + R[1][0]; F[1][3] = R[0][2] + R[2][0];
    F[2][2] = -R[0][0] + R[1][1] - R[2][2]; F[2][3] = R[1][2] + R[2][1];
    F[3][3] = -R[0][0] - R[1][1] + R[2][2];
    F[1][0] = F[0][1]; F[2][0] = F[0][2]; F[3][0] = F[0][3]; F[2][1] = F[1][2]; F[3][1] = F[1][3]; F[3][2] = F[2][3];
}
struct AffineParam { int atom[3]; float ref[3][3]; };   // ref[atom][xyz]
__global__ void k_affine_alignment(const float* __restrict__ pos, float* __restrict__ out, float* __restrict__ eig,
                                   const AffineParam* __restrict__ prm, int n, int n_atom) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    AffineParam p = prm[i];
    const float* x = pos + size_t(r) * n_atom * 4;
    f3 a[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] = ld3v(x + 4 * p.atom[k]);
    f3 center = (1.f / 3.f) * (a[0] + a[1] + a[2]);
    float R[3][3];
#pragma unroll
    for (int k = 0; k < 3; ++k) a[k] -= center;
#pragma unroll
    for (int ii = 0; ii < 3; ++ii) {   // R(i,j) = sum_atoms atom[j]*ref[i]
        R[ii][0] = a[0].x * p.ref[0][ii] + a[1].x * p.ref[1][ii] + a[2].x * p.ref[2][ii];
        R[ii][1] = a[0].y * p.ref[0][ii] + a[1].y * p.ref[1][ii] + a[2].y * p.ref[2][ii];
        R[ii][2] = a[0].z * p.ref[0][ii] + a[1].z * p.ref[1][ii] + a[2].z * p.ref[2][ii];
    }
    float F[4][4], V[4][4];
    fill_F(F, R);
    jacobi_eig4(F, V);
    // order eigenpairs so that the largest eigenvalue is first
    int best = 0;
#pragma unroll
    for (int k = 1; k < 4; ++k) if (F[k][k] > F[best][best]) best = k;
    float* e = eig + (size_t(r) * n + i) * 20;
    int slot = 1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int dst = (k == best) ? 0 : slot++;
        e[dst] = F[k][k];
#pragma unroll
        for (int d = 0; d < 4; ++d) e[4 + dst * 4 + d] = V[d][k];
    }
    float* o = out + (size_t(r) * n + i) * 8;
    reinterpret_cast<float4*>(o)[0] = make_float4(center.x, center.y, center.z, V[0][best]);
    reinterpret_cast<float4*>(o)[1] = make_float4(V[1][best], V[2][best], V[3][best], 0.f);
}
__global__ void k_affine_alignment_deriv(float* __restrict__ pos_sens, const float* __restrict__ sens,
                                         const float* __restrict__ eig, const AffineParam* __restrict__ prm, int n, int n_atom) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    AffineParam p = prm[i];
    const float* e = eig + (size_t(r) * n + i) * 20;
    float ev[4], v[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        ev[k] = e[k];
#pragma unroll
        for (int d = 0; d < 4; ++d) v[k][d] = e[4 + k * 4 + d];
    }
    const float* sn = sens + (size_t(r) * n + i) * 8;
    float4 s0 = reinterpret_cast<const float4*>(sn)[0], s1 = reinterpret_cast<const float4*>(sn)[1];
    f3 s3 = mk3(s0.x, s0.y, s0.z);
    float t0 = s0.w, t1 = s1.x, t2 = s1.y;   // torque about the frame origin (lab frame)
    const float* q = v[0];
    float qs[4] = {2.f * (-t0 * q[1] - t1 * q[2] - t2 * q[3]), 2.f * (t0 * q[0] + t1 * q[3] - t2 * q[2]),
                   2.f * (t1 * q[0] + t2 * q[1] - t0 * q[3]), 2.f * (t2 * q[0] + t0 * q[2] - t1 * q[1])};
    // first-order perturbation of the leading eigenvector: dq = sum_k v_k (v_k^T dF q)/(l0-lk)
    float w[4];   // w[k] = (qs . v_k)/(l0-lk)
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        float dotp = qs[0] * v[k][0] + qs[1] * v[k][1] + qs[2] * v[k][2] + qs[3] * v[k][3];
        w[k] = dotp / (ev[0] - ev[k]);
    }
    // symmetric 4x4 matrix W = sum_k w_k (v_k q^T + q v_k^T)/2 so that dE = <dF, W> (Frobenius)
    float u[4];
#pragma unroll
    for (int d = 0; d < 4; ++d) u[d] = w[1] * v[1][d] + w[2] * v[2][d] + w[3] * v[3][d];
    float W[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) W[a][b] = u[a] * q[b];   // dE = sum_ab dF[a][b] * u[a] q[b]
    float* ps = pos_sens + size_t(r) * n_atom * 4;
#pragma unroll
    for (int na = 0; na < 3; ++na) {
        float d[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            // dR(i,j) = ref[na][i] for this coordinate j; dF is linear in R
            float dR[3][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
            dR[0][j] = p.ref[na][0]; dR[1][j] = p.ref[na][1]; dR[2][j] = p.ref[na][2];
            float dF[4][4];
            fill_F(dF, dR);
            float acc = 0.f;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc += dF[a][b] * W[a][b];
            d[j] = acc;
        }
        atomic_add3v(ps + 4 * p.atom[na], (1.f / 3.f) * s3 + mk3(d[0], d[1], d[2]));
    }
}
struct AffineAlignment : CoordNode {
    CoordNode& pos;
    DevBuf<AffineParam> prm;
    DevBuf<float> eig;
    AffineAlignment(Engine&, const h5l::Node& g, CoordNode& pos_) : CoordNode((int)h5_dims(g, "atoms", 2)[0], 7), pos(pos_) {
        h5_check_size(g, "atoms", {(uint64_t)n_elem, 3});
        h5_check_size(g, "ref_geom", {(uint64_t)n_elem, 3, 3});
        auto atoms = h5_read<int>(g, "atoms");
        auto ref = h5_read<float>(g, "ref_geom");
        std::vector<AffineParam> h(n_elem);
        for (int i = 0; i < n_elem; ++i) {
            for (int k = 0; k < 3; ++k) {
                h[i].atom[k] = atoms[3 * i + k];
                if (h[i].atom[k] < 0 || h[i].atom[k] >= pos.n_elem) throw std::string("atom index out of range");
                for (int d = 0; d < 3; ++d) h[i].ref[k][d] = ref[(i * 3 + k) * 3 + d];
            }
        }
        prm.upload(h);
    }
    void finalize() override { eig.alloc(size_t(engine->n_rep) * n_elem * 20); }
    void compute_value(cudaStream_t s, ComputeMode) override {
        k_affine_alignment<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, output, eig.p, prm.p, n_elem, pos.n_elem);
    }
    void propagate_deriv(cudaStream_t s) override {
        k_affine_alignment_deriv<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.sens, sens, eig.p, prm.p, n_elem, pos.n_elem);
    }
};
RegisterNodeType<AffineAlignment, 1> affine_alignment_node("affine_alignment");

// ---------------------------------------------------------------------------------------------- Infer_H_O
// reference hbond.cpp:14-121: virtual H / O on the bisector of the two bonded neighbours; output pos + unit dir
struct VirtualParam { int atom[3]; float bond_length; };
__global__ void k_infer_ho(const float* __restrict__ pos, float* __restrict__ out, const VirtualParam* __restrict__ prm,
                           int n, int n_atom) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    VirtualParam p = prm[i];
    const float* x = pos + size_t(r) * n_atom * 4;
    f3 c = ld3v(x + 4 * p.atom[1]);
    f3 prev = ld3v(x + 4 * p.atom[0]) - c, next = ld3v(x + 4 * p.atom[2]) - c;
    prev = rsqrtf(mag2(prev)) * prev;
    next = rsqrtf(mag2(next)) * next;
    f3 disp = prev + next;
    disp = rsqrtf(mag2(disp)) * disp;
    f3 dir = -disp;
    f3 hp = c + p.bond_length * dir;
    float* o = out + (size_t(r) * n + i) * 8;
    reinterpret_cast<float4*>(o)[0] = make_float4(hp.x, hp.y, hp.z, dir.x);
    reinterpret_cast<float4*>(o)[1] = make_float4(dir.y, dir.z, 0.f, 0.f);
}
__global__ void k_infer_ho_deriv(const float* __restrict__ pos, float* __restrict__ pos_sens, const float* __restrict__ sens,
                                 const VirtualParam* __restrict__ prm, int n, int n_atom) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    VirtualParam p = prm[i];
    const float* x = pos + size_t(r) * n_atom * 4;
    // recompute the three normalisations (cheaper than storing them; reference keeps data_for_deriv, hbond.cpp:84-87)
    f3 c = ld3v(x + 4 * p.atom[1]);
    f3 prev = ld3v(x + 4 * p.atom[0]) - c, next = ld3v(x + 4 * p.atom[2]) - c;
    float pi = rsqrtf(mag2(prev)), ni = rsqrtf(mag2(next));
    prev = pi * prev;
    next = ni * next;
    f3 disp = prev + next;
    float di = rsqrtf(mag2(disp));
    disp = di * disp;
    const float* sn = sens + (size_t(r) * n + i) * 8;
    float4 a = reinterpret_cast<const float4*>(sn)[0], b = reinterpret_cast<const float4*>(sn)[1];
    f3 s_pos = mk3(a.x, a.y, a.z), s_dir = mk3(a.w, b.x, b.y);
    f3 s_neg = s_dir + p.bond_length * s_pos;
    // hbond.cpp:106-108 (fmsub(a,b,c) = a*b-c)
    f3 s_disp = di * (dot(disp, s_neg) * disp - s_neg);
    f3 s_prev = (-pi) * (dot(prev, s_disp) * prev - s_disp);
    f3 s_next = (-ni) * (dot(next, s_disp) * next - s_disp);
    float* ps = pos_sens + size_t(r) * n_atom * 4;
    atomic_add3v(ps + 4 * p.atom[0], s_prev);
    atomic_add3v(ps + 4 * p.atom[1], s_pos - s_prev - s_next);
    atomic_add3v(ps + 4 * p.atom[2], s_next);
}
struct InferHO : CoordNode {
    CoordNode& pos;
    int n_donor, n_acceptor;
    DevBuf<VirtualParam> prm;
    InferHO(Engine&, const h5l::Node& g, CoordNode& pos_)
        : CoordNode((int)(h5_dims(g, "donors/id", 2)[0] + h5_dims(g, "acceptors/id", 2)[0]), 6), pos(pos_) {
        n_donor = (int)h5_dims(g, "donors/id", 2)[0];
        n_acceptor = (int)h5_dims(g, "acceptors/id", 2)[0];
        h5_check_size(g, "donors/id", {(uint64_t)n_donor, 3});
        h5_check_size(g, "donors/bond_length", {(uint64_t)n_donor});
        h5_check_size(g, "acceptors/id", {(uint64_t)n_acceptor, 3});
        h5_check_size(g, "acceptors/bond_length", {(uint64_t)n_acceptor});
        std::vector<VirtualParam> h(n_elem);
        auto did = h5_read<int>(g, "donors/id"), aid = h5_read<int>(g, "acceptors/id");
        auto dbl = h5_read<float>(g, "donors/bond_length"), abl = h5_read<float>(g, "acceptors/bond_length");
        for (int i = 0; i < n_donor; ++i) { for (int k = 0; k < 3; ++k) h[i].atom[k] = did[3 * i + k]; h[i].bond_length = dbl[i]; }
        for (int i = 0; i < n_acceptor; ++i) { for (int k = 0; k < 3; ++k) h[n_donor + i].atom[k] = aid[3 * i + k]; h[n_donor + i].bond_length = abl[i]; }
        for (auto& v : h) for (int k = 0; k < 3; ++k) if (v.atom[k] < 0 || v.atom[k] >= pos.n_elem) throw std::string("atom index out of range");
        prm.upload(h);
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        k_infer_ho<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, output, prm.p, n_elem, pos.n_elem);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_infer_ho_deriv<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, pos.sens, sens, prm.p, n_elem, pos.n_elem);
    }
    void add_loggers(int level, std::vector<NodeLogger>& out) override {   // hbond.cpp:48-56 (extensive only): site positions
        if (level < 2) return;
        out.push_back({"virtual", {(uint64_t)n_elem, 3}, false, [this](int r) {
            auto o = host_rows(output, r);
            std::vector<float> v(size_t(n_elem) * 3);
            for (int i = 0; i < n_elem; ++i) for (int d = 0; d < 3; ++d) v[size_t(i) * 3 + d] = o[size_t(i) * wp + d];
            return v;
        }});
    }
};
RegisterNodeType<InferHO, 1> infer_node("infer_H_O");

// ---------------------------------------------------------------------------------------------- placements
// reference placement.cpp:233-314.  signature: sequence of POINT(3)/VECTOR(3)/SCALAR(1) blocks encoded as a small code
// array; data source is either a fixed table row (FixedPlacement) or a periodic bicubic spline in (phi,psi)
// (RamaPlacement).  Forward: out = R(q)*ref (+t for POINT).  Backward: inverse-rotated sens to the data source,
// force on the frame origin and torque cross(x-t,s) / cross(x,s) to the affine sens (placement.cpp:209-230).
enum PlaceT { P_SCALAR = 0, P_VECTOR = 1, P_POINT = 2 };
struct PlaceSig { int n_block; int type[3]; int n_dim; };

template <bool RAMA>
__global__ void k_placement(const float* __restrict__ affine, const float* __restrict__ rama, float* __restrict__ out,
                            float* __restrict__ rama_deriv, const int* __restrict__ affine_residue,
                            const int* __restrict__ rama_residue, const int* __restrict__ layer,
                            const float* __restrict__ data, PlaceSig sig, int n, int wp, int n_aff, int n_rama, int nx, int ny) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    const float* aff = affine + (size_t(r) * n_aff + affine_residue[i]) * 8;
    float4 a0 = reinterpret_cast<const float4*>(aff)[0], a1 = reinterpret_cast<const float4*>(aff)[1];
    f3 t = mk3(a0.x, a0.y, a0.z);
    float q[4] = {a0.w, a1.x, a1.y, a1.z};
    float U[9];
    quat_to_rot(U, q);
    float val[7];
    if (RAMA) {
        const float* rp = rama + (size_t(r) * n_rama + rama_residue[i]) * 2;
        int xb, yb;
        float fx, fy, sx, sy;
        rama_cell(rp[0], rp[1], nx, ny, xb, yb, fx, fy, sx, sy);
        const float* c = data + (((size_t(layer[i]) * nx + xb) * ny + yb) * sig.n_dim) * 16;
        float* rd = rama_deriv + (size_t(r) * n + i) * 2 * sig.n_dim;
        for (int d = 0; d < sig.n_dim; ++d) {
            float dx, dy;
            bicubic_vd(c + d * 16, fx, fy, val[d], dx, dy);
            rd[d] = dx;
            rd[sig.n_dim + d] = dy;
        }
    } else {
        const float* c = data + size_t(layer[i]) * sig.n_dim;
        for (int d = 0; d < sig.n_dim; ++d) val[d] = __ldg(c + d);
    }
    float o[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    int off = 0;
    for (int b = 0; b < sig.n_block; ++b) {
        if (sig.type[b] == P_SCALAR) { o[off] = val[off]; off += 1; }
        else {
            f3 v = rot_apply(U, mk3(val[off], val[off + 1], val[off + 2]));
            if (sig.type[b] == P_POINT) v += t;
            o[off] = v.x; o[off + 1] = v.y; o[off + 2] = v.z;
            off += 3;
        }
    }
    // rows are padded to 1/2/4/8 floats (padding written as zero): 16-byte stores
    float* dst = out + (size_t(r) * n + i) * wp;
    if (wp == 8) {
        reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
        reinterpret_cast<float4*>(dst)[1] = make_float4(o[4], o[5], o[6], o[7]);
    } else if (wp == 4) reinterpret_cast<float4*>(dst)[0] = make_float4(o[0], o[1], o[2], o[3]);
    else if (wp == 2) reinterpret_cast<float2*>(dst)[0] = make_float2(o[0], o[1]);
    else dst[0] = o[0];
}
template <bool RAMA>
__global__ void k_placement_deriv(const float* __restrict__ affine, float* __restrict__ affine_sens,
                                  const float* __restrict__ rama, float* __restrict__ rama_sens,
                                  const float* __restrict__ out, const float* __restrict__ sens,
                                  const float* __restrict__ rama_deriv, float* __restrict__ param_deriv,
                                  const int* __restrict__ affine_residue, const int* __restrict__ rama_residue,
                                  const int* __restrict__ layer, PlaceSig sig, int n, int wp, int n_aff, int n_rama,
                                  int nx, int ny) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    const bool active = i < n;
    const unsigned lane = threadIdx.x & 31u;
    // elements placed on the same residue frame are neighbours in every configuration upside_config.py writes: their
    // force and torque are summed across the lanes of a warp first (segmented suffix sums), one atomic per run of <= 8
    const int ar = active ? affine_residue[i] : -1 - (int)lane;
    bool spatial = false;
    for (int b = 0; b < sig.n_block; ++b) spatial |= sig.type[b] != P_SCALAR;
    f3 com = mk3(0.f, 0.f, 0.f), torque = mk3(0.f, 0.f, 0.f);
    if (active) {
    const float* aff = affine + (size_t(r) * n_aff + ar) * 8;
    float4 a0 = reinterpret_cast<const float4*>(aff)[0], a1 = reinterpret_cast<const float4*>(aff)[1];
    f3 t = mk3(a0.x, a0.y, a0.z);
    float q[4] = {a0.w, a1.x, a1.y, a1.z};
    float U[9];
    quat_to_rot(U, q);
    const float* x = out + (size_t(r) * n + i) * wp;
    const float* sn = sens + (size_t(r) * n + i) * wp;
    float ref_sens[7];
    int off = 0;
    for (int b = 0; b < sig.n_block; ++b) {
        if (sig.type[b] == P_SCALAR) { ref_sens[off] = sn[off]; off += 1; }
        else {
            f3 s = ld3(sn + off), xv = ld3(x + off);
            f3 rs = rot_apply_inv(U, s);
            ref_sens[off] = rs.x; ref_sens[off + 1] = rs.y; ref_sens[off + 2] = rs.z;
            if (sig.type[b] == P_POINT) { com += s; torque += cross(xv - t, s); }
            else torque += cross(xv, s);
            off += 3;
        }
    }
    if (RAMA) {
        const float PI = 3.1415926535897932f;
        float sx = nx * (0.5f / PI - 1e-7f), sy = ny * (0.5f / PI - 1e-7f);
        const float* rd = rama_deriv + (size_t(r) * n + i) * 2 * sig.n_dim;
        float dphi = 0.f, dpsi = 0.f;
        for (int d = 0; d < sig.n_dim; ++d) { dphi += ref_sens[d] * rd[d]; dpsi += ref_sens[d] * rd[sig.n_dim + d]; }
        float* rs = rama_sens + (size_t(r) * n_rama + rama_residue[i]) * 2;
        atomicAdd(rs + 0, sx * dphi);
        atomicAdd(rs + 1, sy * dpsi);
    } else if (param_deriv) {
        float* pd = param_deriv + size_t(layer[i]) * sig.n_dim;
        for (int d = 0; d < sig.n_dim; ++d) atomicAdd(pd + d, ref_sens[d]);
    }
    }
    if (!spatial) return;   // scalar placements move no frame (uniform over the grid)
    const int ar_prev = __shfl_up_sync(UB_FULL_MASK, ar, 1);
    const unsigned heads = __ballot_sync(UB_FULL_MASK, lane == 0 || ar_prev != ar);
    const int seg_start = 31 - __clz(heads & (0xffffffffu >> (31 - lane)));
    float v[6] = {com.x, com.y, com.z, torque.x, torque.y, torque.z};
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const int start_o = __shfl_down_sync(UB_FULL_MASK, seg_start, o);
        const bool take = lane + o < 32 && start_o == seg_start;   // same run, whatever the order of the elements
#pragma unroll
        for (int c = 0; c < 6; ++c) { const float w = __shfl_down_sync(UB_FULL_MASK, v[c], o); if (take) v[c] += w; }
    }
    if (active && ((lane - seg_start) & 7) == 0) {
        float* as = affine_sens + (size_t(r) * n_aff + ar) * 8;
        atomicAdd(reinterpret_cast<float4*>(as), make_float4(v[0], v[1], v[2], v[3]));   // force, torque: two 16-byte reductions
        atomicAdd(reinterpret_cast<float4*>(as) + 1, make_float4(v[4], v[5], 0.f, 0.f));
    }
}

template <bool RAMA> struct PlacementNode : CoordNode {
    PlaceSig sig;
    CoordNode& alignment;
    CoordNode* rama;
    int n_layer, nx = 0, ny = 0;
    DevBuf<int> affine_residue, rama_residue, layer;
    DevBuf<float> data, rama_deriv;
    std::vector<float> h_data;
    static PlaceSig make_sig(std::initializer_list<int> types) {
        PlaceSig s{};
        s.n_block = 0;
        s.n_dim = 0;
        for (int t : types) { s.type[s.n_block++] = t; s.n_dim += (t == P_SCALAR ? 1 : 3); }
        return s;
    }
    PlacementNode(const h5l::Node& g, PlaceSig sig_, CoordNode& alignment_, CoordNode* rama_)
        : CoordNode((int)h5_dims(g, "layer_index", 1)[0], sig_.n_dim), sig(sig_), alignment(alignment_), rama(rama_) {
        check_elem_width(alignment, 7);
        h5_check_size(g, "affine_residue", {(uint64_t)n_elem});
        auto ar = h5_read<int>(g, "affine_residue");
        auto li = h5_read<int>(g, "layer_index");
        for (int a : ar) if (a < 0 || a >= alignment.n_elem) throw std::string("affine_residue out of range");
        if (RAMA) {
            check_elem_width(*rama, 2);
            auto d = h5_dims(g, "placement_data", 4);
            n_layer = (int)d[0]; nx = (int)d[1]; ny = (int)d[2];
            h5_check_size(g, "placement_data", {d[0], d[1], d[2], (uint64_t)sig.n_dim});
            h5_check_size(g, "rama_residue", {(uint64_t)n_elem});
            auto rr = h5_read<int>(g, "rama_residue");
            for (int a : rr) if (a < 0 || a >= rama->n_elem) throw std::string("rama_residue out of range");
            rama_residue.upload(rr);
            auto raw = h5_read<double>(g, "placement_data");
            data.upload(fit_periodic_spline_2d(n_layer, nx, ny, sig.n_dim, raw.data()));
        } else {
            auto d = h5_dims(g, "placement_data", 2);
            n_layer = (int)d[0];
            h5_check_size(g, "placement_data", {d[0], (uint64_t)sig.n_dim});
            h_data = h5_read<float>(g, "placement_data");
            data.upload(h_data);
        }
        for (int l : li) if (l < 0 || l >= n_layer) throw std::string("layer_index out of range");
        affine_residue.upload(ar);
        layer.upload(li);
    }
    void finalize() override { if (RAMA) rama_deriv.alloc(size_t(engine->n_rep) * n_elem * 2 * sig.n_dim); }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        k_placement<RAMA><<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(
            alignment.output, RAMA ? rama->output : nullptr, output, rama_deriv.p, affine_residue.p, rama_residue.p,
            layer.p, data.p, sig, n_elem, wp, alignment.n_elem, RAMA ? rama->n_elem : 0, nx, ny);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_placement_deriv<RAMA><<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(
            alignment.output, alignment.sens, RAMA ? rama->output : nullptr, RAMA ? rama->sens : nullptr, output, sens,
            rama_deriv.p, nullptr, affine_residue.p, rama_residue.p, layer.p, sig, n_elem, wp, alignment.n_elem,
            RAMA ? rama->n_elem : 0, nx, ny);
    }
    // placement.cpp:254-261 (extensive only).  The reference names the dataset "placement_pos" for every placement node, so a
    // configuration with several of them fails there when the second dataset is created; the CLI reports the same clash.
    void add_loggers(int level, std::vector<NodeLogger>& out) override {
        if (level < 2) return;
        out.push_back({"placement_pos", {(uint64_t)n_elem, (uint64_t)sig.n_dim}, false, [this](int r) {
            auto o = host_rows(output, r);
            std::vector<float> v(size_t(n_elem) * sig.n_dim);
            for (int i = 0; i < n_elem; ++i) for (int d = 0; d < sig.n_dim; ++d) v[size_t(i) * sig.n_dim + d] = o[size_t(i) * wp + d];
            return v;
        }});
    }
    std::vector<float> get_param() const override { return h_data; }
    // placement.cpp:138-161: sensitivity of every element rotated back into the reference frame, summed per layer
    // (the Rama-dependent placements return nothing, :95-97)
    std::vector<float> get_param_deriv(int replica) override {
        if (RAMA) return {};
        if (replica >= engine->n_rep) throw std::string("replica out of range");
        engine->sync_and_check();
        const int r0 = replica < 0 ? 0 : replica, r1 = replica < 0 ? engine->n_rep : replica + 1;
        auto sn = download_rows(*this, sens, r0, r1);
        auto af = download_rows(alignment, alignment.output, r0, r1);
        auto ar = affine_residue.download();
        auto li = layer.download();
        std::vector<float> deriv(h_data.size(), 0.f);
        for (int r = 0; r < r1 - r0; ++r)
            for (int i = 0; i < n_elem; ++i) {
                const float* a = &af[(size_t(r) * alignment.n_elem + ar[i]) * 8];
                const float q0 = a[3], q1 = a[4], q2 = a[5], q3 = a[6];
                const float U[9] = {q0 * q0 + q1 * q1 - q2 * q2 - q3 * q3, 2.f * (q1 * q2 - q0 * q3), 2.f * (q1 * q3 + q0 * q2),
                                    2.f * (q1 * q2 + q0 * q3), q0 * q0 - q1 * q1 + q2 * q2 - q3 * q3, 2.f * (q2 * q3 - q0 * q1),
                                    2.f * (q1 * q3 - q0 * q2), 2.f * (q2 * q3 + q0 * q1), q0 * q0 - q1 * q1 - q2 * q2 + q3 * q3};
                const float* s = &sn[(size_t(r) * n_elem + i) * wp];
                float* d = &deriv[size_t(li[i]) * sig.n_dim];
                int off = 0;
                for (int b = 0; b < sig.n_block; ++b) {
                    if (sig.type[b] == P_SCALAR) { d[off] += s[off]; off += 1; }
                    else {   // U^T s
                        for (int c = 0; c < 3; ++c) d[off + c] += U[c] * s[off] + U[3 + c] * s[off + 1] + U[6 + c] * s[off + 2];
                        off += 3;
                    }
                }
            }
        return deriv;
    }
    void set_param(const std::vector<float>& p) override {
        if (RAMA) return;
        if (p.size() != h_data.size()) throw std::string("wrong param size");
        h_data = p;
        data.upload(h_data);
    }
};
#define UB_PLACEMENT(cls, rama_flag, nargs, prefix, ...)                                                          \
    struct cls : PlacementNode<rama_flag> {                                                                       \
        cls(Engine&, const h5l::Node& g, CoordNode& a) : PlacementNode<rama_flag>(g, make_sig({__VA_ARGS__}), a, nullptr) {} \
        cls(Engine&, const h5l::Node& g, CoordNode& a, CoordNode& r) : PlacementNode<rama_flag>(g, make_sig({__VA_ARGS__}), a, &r) {} \
    };                                                                                                            \
    RegisterNodeType<cls, nargs> cls##_node(prefix);
// the seven registered variants of placement.cpp:319-325
UB_PLACEMENT(PlScalar, true, 2, "placement_scalar", P_SCALAR)
UB_PLACEMENT(PlFixedScalar, false, 1, "placement_fixed_scalar", P_SCALAR)
UB_PLACEMENT(PlPointOnly, true, 2, "placement_point_only", P_POINT)
UB_PLACEMENT(PlFixedPointOnly, false, 1, "placement_fixed_point_only", P_POINT)
UB_PLACEMENT(PlPointVectorOnly, true, 2, "placement_point_vector_only", P_POINT, P_VECTOR)
UB_PLACEMENT(PlFixedPointVectorOnly, false, 1, "placement_fixed_point_vector_only", P_POINT, P_VECTOR)
UB_PLACEMENT(PlFixedPointVectorScalar, false, 1, "placement_fixed_point_vector_scalar", P_POINT, P_VECTOR, P_SCALAR)

// ---------------------------------------------------------------------------------------------- WeightedPos
// reference environment.cpp:112-156: (x,y,z, exp(-E))
__global__ void k_weighted_pos(const float* __restrict__ pos, const float* __restrict__ energy, float* __restrict__ out,
                               const int* __restrict__ ipos, const int* __restrict__ iw, int n, int n_pos, int wp_pos,
                               int n_en, int wp_en) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    const float* p = pos + (size_t(r) * n_pos + ipos[i]) * wp_pos;
    float e = energy[(size_t(r) * n_en + iw[i]) * wp_en];
    reinterpret_cast<float4*>(out)[size_t(r) * n + i] = make_float4(p[0], p[1], p[2], __expf(-e));
}
__global__ void k_weighted_pos_deriv(float* __restrict__ pos_sens, float* __restrict__ energy_sens,
                                     const float* __restrict__ out, const float* __restrict__ sens,
                                     const int* __restrict__ ipos, const int* __restrict__ iw, int n, int n_pos, int wp_pos,
                                     int n_en, int wp_en) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    float4 s = reinterpret_cast<const float4*>(sens)[size_t(r) * n + i];
    float w = out[(size_t(r) * n + i) * 4 + 3];
    atomic_add3v(pos_sens + (size_t(r) * n_pos + ipos[i]) * wp_pos, mk3(s.x, s.y, s.z));
    atomicAdd(energy_sens + (size_t(r) * n_en + iw[i]) * wp_en, -w * s.w);
}
struct WeightedPos : CoordNode {
    CoordNode &pos, &energy;
    DevBuf<int> ipos, iw;
    WeightedPos(Engine&, const h5l::Node& g, CoordNode& pos_, CoordNode& energy_)
        : CoordNode((int)h5_dims(g, "index_pos", 1)[0], 4), pos(pos_), energy(energy_) {
        check_elem_width_lower_bound(pos, 3);
        h5_check_size(g, "index_weight", {(uint64_t)n_elem});
        auto a = h5_read<int>(g, "index_pos"), b = h5_read<int>(g, "index_weight");
        for (int v : a) if (v < 0 || v >= pos.n_elem) throw std::string("index_pos out of range");
        for (int v : b) if (v < 0 || v >= energy.n_elem) throw std::string("index_weight out of range");
        ipos.upload(a);
        iw.upload(b);
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        k_weighted_pos<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.output, energy.output, output, ipos.p, iw.p, n_elem,
                                                                        pos.n_elem, pos.wp, energy.n_elem, energy.wp);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_weighted_pos_deriv<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(pos.sens, energy.sens, output, sens, ipos.p, iw.p,
                                                                              n_elem, pos.n_elem, pos.wp, energy.n_elem, energy.wp);
    }
};
RegisterNodeType<WeightedPos, 2> weighted_pos_node("weighted_pos");

// ---------------------------------------------------------------------------------------------- NonlinearCoupling
// reference environment.cpp:324-397: per-element clamped B-spline of a width-1 coordinate (scalar clamping rule
// x<=1 / x>=n-2 of spline.h:268-272)
__global__ void k_nonlinear_coupling(const float* __restrict__ in, float* __restrict__ in_sens, float* __restrict__ pot,
                                     const float* __restrict__ coeff, const int* __restrict__ types, int n, int n_coeff,
                                     float offset, float inv_dx, int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        float x = (in[size_t(r) * n + i] - offset) * inv_dx;
        const float* c = coeff + size_t(types[i]) * n_coeff;
        float v, d;
        if (x <= 1.f) { v = (1.f / 6.f) * c[0] + (2.f / 3.f) * c[1] + (1.f / 6.f) * c[2]; d = 0.f; }
        else if (x >= (float)(n_coeff - 2)) { v = (1.f / 6.f) * c[n_coeff - 3] + (2.f / 3.f) * c[n_coeff - 2] + (1.f / 6.f) * c[n_coeff - 1]; d = 0.f; }
        else { int b = (int)x; deboor_core(c[b - 1], c[b], c[b + 1], c[b + 2], x - b, v, d); }
        in_sens[size_t(r) * n + i] += d * inv_dx;   // sole writer of this element in this kernel
        e = v;
    }
    if (want_pot) accumulate_potential(e, pot);
}
struct NonlinearCoupling : PotentialNode {
    CoordNode& input;
    int n_restype, n_coeff;
    float offset, inv_dx;
    std::vector<float> h_coeff;
    DevBuf<float> coeff;
    DevBuf<int> types;
    NonlinearCoupling(Engine&, const h5l::Node& g, CoordNode& in) : input(in) {
        check_elem_width(input, 1);
        auto d = h5_dims(g, "coeff", 2);
        n_restype = (int)d[0];
        n_coeff = (int)d[1];
        offset = h5_attr<float>(g, "coeff", "spline_offset");
        inv_dx = h5_attr<float>(g, "coeff", "spline_inv_dx");
        h5_check_size(g, "coupling_types", {(uint64_t)input.n_elem});
        h_coeff = h5_read<float>(g, "coeff");
        auto t = h5_read<int>(g, "coupling_types");
        for (int v : t) if (v < 0 || v >= n_restype) throw std::string("invalid coupling type");
        coeff.upload(h_coeff);
        types.upload(t);
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!input.n_elem) return;
        k_nonlinear_coupling<<<grid_for(input.n_elem, engine->n_rep), TPB, 0, s>>>(
            input.output, input.sens, potential, coeff.p, types.p, input.n_elem, n_coeff, offset, inv_dx,
            mode == PotentialAndDerivMode);
    }
    std::vector<float> get_param() const override { return h_coeff; }
    // clamped cubic B-spline: first coefficient and the four basis weights at coordinate c (spline.h:318-336,375-392)
    void basis(float c, int& bin, float* w) const {
        if (c <= 1.f) { bin = 0; w[0] = 1.f / 6.f; w[1] = 2.f / 3.f; w[2] = 1.f / 6.f; w[3] = 0.f; return; }
        if (c >= float(n_coeff - 2)) { bin = n_coeff - 4; w[0] = 0.f; w[1] = 1.f / 6.f; w[2] = 2.f / 3.f; w[3] = 1.f / 6.f; return; }
        const int b = (int)c;
        const float y = c - b, z = 1.f - y;
        bin = b - 1;
        w[0] = z * z * z / 6.f; w[3] = y * y * y / 6.f;
        w[1] = (3.f * y * y * y - 6.f * y * y + 4.f) / 6.f; w[2] = (3.f * z * z * z - 6.f * z * z + 4.f) / 6.f;
    }
    void add_loggers(int level, std::vector<NodeLogger>& out) override {   // environment.cpp:348-355: energy per residue
        if (level < 1) return;
        out.push_back({"nonlinear_coupling", {(uint64_t)input.n_elem}, false, [this](int r) {
            auto x = input.host_rows(input.output, r);
            auto t = types.download();
            std::vector<float> v(input.n_elem);
            for (int ne = 0; ne < input.n_elem; ++ne) {
                int bin;
                float w[4];
                basis((x[size_t(ne) * input.wp] - offset) * inv_dx, bin, w);
                const float* c = &h_coeff[size_t(t[ne]) * n_coeff + bin];
                v[ne] = w[0] * c[0] + w[1] * c[1] + w[2] * c[2] + w[3] * c[3];
            }
            return v;
        }});
    }
    // environment.cpp:375-390: d/d(coeff) = the four basis weights of each residue's knot interval (clamped spline)
    std::vector<float> get_param_deriv(int replica) override {
        if (replica >= engine->n_rep) throw std::string("replica out of range");
        engine->sync_and_check();
        const int r0 = replica < 0 ? 0 : replica, r1 = replica < 0 ? engine->n_rep : replica + 1;
        auto x = download_rows(input, input.output, r0, r1);
        auto t = types.download();
        std::vector<float> deriv(h_coeff.size(), 0.f);
        for (int r = 0; r < r1 - r0; ++r)
            for (int ne = 0; ne < input.n_elem; ++ne) {
                const float c = (x[(size_t(r) * input.n_elem + ne) * input.wp] - offset) * inv_dx;
                int bin;
                float w[4];
                basis(c, bin, w);
                for (int i = 0; i < 4; ++i) deriv[size_t(t[ne]) * n_coeff + bin + i] += w[i];
            }
        return deriv;
    }
    void set_param(const std::vector<float>& p) override {
        if (p.size() != h_coeff.size()) throw std::string("attempting to change size of coeff vector on set_param");
        h_coeff = p;
        coeff.upload(h_coeff);
    }
};
RegisterNodeType<NonlinearCoupling, 1> nonlinear_coupling_node("nonlinear_coupling");

// ---------------------------------------------------------------------------------------------- UniformTransform
// reference environment.cpp:158-233: output = clamped cubic B-spline of a width-1 coordinate, one spline for all
// elements; the backward pass recomputes the slope instead of storing a Jacobian
__device__ __forceinline__ void clamped_scalar_spline(const float* __restrict__ c, int n_coeff, float x, float& v, float& d) {
    if (x <= 1.f) { v = (1.f / 6.f) * c[0] + (2.f / 3.f) * c[1] + (1.f / 6.f) * c[2]; d = 0.f; }
    else if (x >= (float)(n_coeff - 2)) { v = (1.f / 6.f) * c[n_coeff - 3] + (2.f / 3.f) * c[n_coeff - 2] + (1.f / 6.f) * c[n_coeff - 1]; d = 0.f; }
    else { const int b = (int)x; deboor_core(c[b - 1], c[b], c[b + 1], c[b + 2], x - b, v, d); }
}
__global__ void k_uniform_transform(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ coeff, int n, int n_coeff,
                                    float offset, float inv_dx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    float v, d;
    clamped_scalar_spline(coeff, n_coeff, (in[size_t(r) * n + i] - offset) * inv_dx, v, d);
    out[size_t(r) * n + i] = v;
}
__global__ void k_uniform_transform_deriv(const float* __restrict__ in, float* __restrict__ in_sens, const float* __restrict__ sens,
                                          const float* __restrict__ coeff, int n, int n_coeff, float offset, float inv_dx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (i >= n) return;
    float v, d;
    clamped_scalar_spline(coeff, n_coeff, (in[size_t(r) * n + i] - offset) * inv_dx, v, d);
    atomicAdd(&in_sens[size_t(r) * n + i], d * inv_dx * sens[size_t(r) * n + i]);   // other consumers of the input add here too
}
struct UniformTransform : CoordNode {
    CoordNode& input;
    int n_coeff;
    float offset, inv_dx;
    std::vector<float> h_coeff;
    DevBuf<float> coeff;
    UniformTransform(Engine&, const h5l::Node& g, CoordNode& in) : CoordNode(in.n_elem, 1), input(in) {
        check_elem_width(input, 1);
        n_coeff = (int)h5_dims(g, "bspline_coeff", 1)[0];
        offset = h5_attr<float>(g, "bspline_coeff", "spline_offset");
        inv_dx = h5_attr<float>(g, "bspline_coeff", "spline_inv_dx");
        h_coeff = h5_read<float>(g, "bspline_coeff");
        if (n_coeff < 4) throw std::string("too small of size for spline");
        coeff.upload(h_coeff);
    }
    void compute_value(cudaStream_t s, ComputeMode) override {
        if (!n_elem) return;
        k_uniform_transform<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(input.output, output, coeff.p, n_elem, n_coeff, offset, inv_dx);
    }
    void propagate_deriv(cudaStream_t s) override {
        if (!n_elem) return;
        k_uniform_transform_deriv<<<grid_for(n_elem, engine->n_rep), TPB, 0, s>>>(input.output, input.sens, sens, coeff.p, n_elem, n_coeff, offset, inv_dx);
    }
    std::vector<float> get_param() const override {   // (offset, inv_dx, coefficients): environment.cpp:198-204
        std::vector<float> p{offset, inv_dx};
        p.insert(p.end(), h_coeff.begin(), h_coeff.end());
        return p;
    }
    void set_param(const std::vector<float>& p) override {   // environment.cpp:222-232
        if (p.size() < size_t(2 + 4)) throw std::string("too small of size for spline");
        offset = p[0]; inv_dx = p[1];
        h_coeff.assign(p.begin() + 2, p.end());
        n_coeff = (int)h_coeff.size();
        coeff.upload(h_coeff);
    }
};
RegisterNodeType<UniformTransform, 1> uniform_transform_node("uniform_transform");

// ---------------------------------------------------------------------------------------------- LinearCoupling
// reference environment.cpp:235-321: E = sum_i c[type_i] * x_i * (1 - inact_i)^2, with the optional inactivation read
// from component `inactivation_dim` of a second node; logger: c * x per element (:270-278)
__global__ void k_linear_coupling(const float* __restrict__ in, float* __restrict__ in_sens, const float* __restrict__ inact,
                                  float* __restrict__ inact_sens, int inact_wp, int inact_dim, const float* __restrict__ couplings,
                                  const int* __restrict__ types, int n, float* __restrict__ pot, int want_pot) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float e = 0.f;
    if (i < n) {
        const float c = couplings[types[i]], val = in[size_t(r) * n + i];
        float act = 1.f;
        if (inact) { const float u = 1.f - inact[(size_t(r) * n + i) * inact_wp + inact_dim]; act = u * u; }
        e = c * val * act;
        atomicAdd(&in_sens[size_t(r) * n + i], c * act);
        if (inact) atomicAdd(&inact_sens[(size_t(r) * n + i) * inact_wp + inact_dim], -c * val);   // sic: the reference's derivative (:293)
    }
    if (want_pot) accumulate_potential(e, pot);
}
struct LinearCoupling : PotentialNode {
    CoordNode& input;
    CoordNode* inactivation;
    int inactivation_dim = 0;
    std::vector<float> h_couplings;
    std::vector<int> h_types;
    DevBuf<float> couplings;
    DevBuf<int> types;
    LinearCoupling(Engine& e, const h5l::Node& g, CoordNode& in) : LinearCoupling(e, g, in, nullptr) {}
    LinearCoupling(Engine& e, const h5l::Node& g, CoordNode& in, CoordNode& inact) : LinearCoupling(e, g, in, &inact) {}
    LinearCoupling(Engine&, const h5l::Node& g, CoordNode& in, CoordNode* inact) : input(in), inactivation(inact) {
        check_elem_width(input, 1);
        if (inactivation) {
            inactivation_dim = h5_attr<int>(g, ".", "inactivation_dim");
            if (input.n_elem != inactivation->n_elem) throw std::string("Inactivation size must match input size");
            check_elem_width_lower_bound(*inactivation, inactivation_dim + 1);
        }
        h_couplings = h5_read<float>(g, "couplings");
        h5_check_size(g, "coupling_types", {(uint64_t)input.n_elem});
        h_types = h5_read<int>(g, "coupling_types");
        for (int v : h_types) if (v < 0 || v >= (int)h_couplings.size()) throw std::string("invalid coupling type");
        couplings.upload(h_couplings);
        types.upload(h_types);
    }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        if (!input.n_elem) return;
        k_linear_coupling<<<grid_for(input.n_elem, engine->n_rep), TPB, 0, s>>>(
            input.output, input.sens, inactivation ? inactivation->output : nullptr, inactivation ? inactivation->sens : nullptr,
            inactivation ? inactivation->wp : 1, inactivation_dim, couplings.p, types.p, input.n_elem, potential, mode == PotentialAndDerivMode);
    }
    std::vector<float> get_param() const override { return h_couplings; }
    void set_param(const std::vector<float>& p) override {
        if (p.size() != h_couplings.size()) throw std::string("attempting to change size of couplings vector on set_param");
        h_couplings = p;
        couplings.upload(h_couplings);
    }
    std::vector<float> get_param_deriv(int replica) override {   // environment.cpp:305-314 (PARAM_DERIV build)
        if (replica >= engine->n_rep) throw std::string("replica out of range");
        engine->sync_and_check();
        std::vector<float> d(h_couplings.size(), 0.f);
        const int r0 = replica < 0 ? 0 : replica, r1 = replica < 0 ? engine->n_rep : replica + 1;
        for (int r = r0; r < r1; ++r) {
            auto x = input.host_rows(input.output, r);
            std::vector<float> u;
            if (inactivation) u = inactivation->host_rows(inactivation->output, r);
            for (int ne = 0; ne < input.n_elem; ++ne) {
                const float act = inactivation ? 1.f - u[size_t(ne) * inactivation->wp + inactivation_dim] : 1.f;
                d[h_types[ne]] += x[size_t(ne) * input.wp] * act;
            }
        }
        return d;
    }
    void add_loggers(int level, std::vector<NodeLogger>& out) override {
        if (level < 1) return;
        out.push_back({inactivation ? "linear_coupling_with_inactivation" : "linear_coupling_uniform", {(uint64_t)input.n_elem}, false, [this](int r) {
            auto x = input.host_rows(input.output, r);
            std::vector<float> v(input.n_elem);
            for (int ne = 0; ne < input.n_elem; ++ne) v[ne] = h_couplings[h_types[ne]] * x[size_t(ne) * input.wp];
            return v;
        }});
    }
};
RegisterNodeType<LinearCoupling, 1> linear_coupling_node1("linear_coupling_uniform");
RegisterNodeType<LinearCoupling, 2> linear_coupling_node2("linear_coupling_with_inactivation");

// ---------------------------------------------------------------------------------------------- HBondEnergy
// reference hbond.cpp:417-456: E = E_hb * sum_sites hb, sens(6) += E_hb
__global__ void k_hbond_energy(const float* __restrict__ hb, float* __restrict__ hb_sens, float* __restrict__ pot,
                               float* __restrict__ n_hbond, int n, float Ep, int want_pot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    float v = 0.f;
    if (i < n) {
        v = hb[(size_t(r) * n + i) * 8 + 6];
        hb_sens[(size_t(r) * n + i) * 8 + 6] += Ep;
    }
    if (want_pot) {
        __shared__ float sc[32];
        v = block_sum(v, sc);
        if (threadIdx.x == 0) { atomicAdd(pot + r, v * Ep); atomicAdd(n_hbond + r, v); }
    }
}
struct HBondEnergy : PotentialNode {
    CoordNode& protein_hbond;
    float Ep;
    DevBuf<float> n_hbond;
    HBondEnergy(Engine&, const h5l::Node& g, CoordNode& ph) : protein_hbond(ph) {
        check_elem_width(protein_hbond, 7);
        Ep = h5_attr<float>(g, ".", "protein_hbond_energy");
    }
    void finalize() override { n_hbond.alloc(engine->n_rep); }
    void compute_value(cudaStream_t s, ComputeMode mode) override {
        int n = protein_hbond.n_elem;
        if (!n) return;
        if (mode == PotentialAndDerivMode) UB_CUDA(cudaMemsetAsync(n_hbond.p, 0, sizeof(float) * engine->n_rep, s));
        k_hbond_energy<<<grid_for(n, engine->n_rep), TPB, 0, s>>>(protein_hbond.output, protein_hbond.sens, potential, n_hbond.p,
                                                                   n, Ep, mode == PotentialAndDerivMode);
    }
    std::vector<float> get_param() const override { return {Ep}; }
    std::vector<float> get_param_deriv(int replica) override {   // hbond.cpp:447-449: dE/dE_hb = number of H-bonds
        if (replica >= engine->n_rep) throw std::string("replica out of range");
        engine->sync_and_check();
        auto v = n_hbond.download();
        float tot = 0.f;
        for (int r = 0; r < engine->n_rep; ++r) if (replica < 0 || r == replica) tot += v[r];
        return {tot};
    }
    void set_param(const std::vector<float>& p) override {
        if (p.size() != 1u) throw "expected 1 param to hbond_energy but got " + std::to_string(p.size());
        Ep = p[0];
        // the captured graphs hold Ep by value: force re-capture
        for (auto& g : engine->graph_eval) if (g) { cudaGraphExecDestroy(g); g = nullptr; }
        if (engine->graph_round) { cudaGraphExecDestroy(engine->graph_round); engine->graph_round = nullptr; }
    }
    std::vector<float> get_value_by_name(int replica, const char* nm) override {
        if (std::string(nm) == "n_hbond") { auto v = n_hbond.download(); return {v.at(replica)}; }
        throw std::string("Value ") + nm + " not implemented";
    }
};
RegisterNodeType<HBondEnergy, 1> hbond_energy_node("hbond_energy");

}  // namespace
}  // namespace ub
