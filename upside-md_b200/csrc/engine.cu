// Engine core: node graph, device arenas, batched evaluation, Verlet integrator, Ornstein-Uhlenbeck thermostat.
// Reference behaviour restated: src/deriv_engine.cpp:11-35 (integration_stage), :37-48 (recenter), :124-169
// (compute), :172-192 (integration_cycle), :195-270 (initialize_engine_from_hdf5); src/thermostat.cpp:9-18;
// src/random.h:19-66 + Random123 threefry4x32-20, uniform.hpp u01/uneg11, boxmuller.hpp.
#include "engine.h"
#include "rng.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace ub {

void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
}

// ================================================================================================ h5 helpers
const h5l::Node& h5_child(const h5l::Node& g, const std::string& name) {
    const h5l::Node* n = h5l::find(&g, name);
    if (!n) throw std::string("'") + name + "' not found in configuration";
    return *n;
}
bool h5_has(const h5l::Node& g, const std::string& name) { return h5l::find(&g, name) != nullptr; }
std::vector<uint64_t> h5_dims(const h5l::Node& g, const std::string& name, int ndim) {
    const h5l::Node& d = h5_child(g, name);
    if (d.is_group) throw std::string("'") + name + "' is a group, expected a dataset";
    if ((int)d.data.dims.size() != ndim)
        throw std::string("while getting size of '") + name + "', wrong number of dimensions (expected " +
            std::to_string(ndim) + ", but got " + std::to_string(d.data.dims.size()) + ")";
    return d.data.dims;
}
void h5_check_size(const h5l::Node& g, const std::string& name, std::vector<uint64_t> sz) {
    auto dims = h5_dims(g, name, (int)sz.size());
    if (dims != sz) {
        std::string msg = "dimensions of '" + name + "', expected (";
        for (size_t i = 0; i < sz.size(); ++i) msg += std::to_string(sz[i]) + (i + 1 < sz.size() ? ", " : "");
        msg += ") but got (";
        for (size_t i = 0; i < dims.size(); ++i) msg += std::to_string(dims[i]) + (i + 1 < dims.size() ? ", " : "");
        throw msg + ")";
    }
}
template <typename T> std::vector<T> h5_read(const h5l::Node& g, const std::string& name) {
    const h5l::Node& d = h5_child(g, name);
    if (d.is_group) throw std::string("'") + name + "' is a group, expected a dataset";
    return h5l::as<T>(d.data);
}
template std::vector<float> h5_read<float>(const h5l::Node&, const std::string&);
template std::vector<double> h5_read<double>(const h5l::Node&, const std::string&);
template std::vector<int> h5_read<int>(const h5l::Node&, const std::string&);

template <typename T> static bool attr_try(const h5l::Node& g, const std::string& path, const std::string& attr, T& out) {
    const h5l::Node& o = h5_child(g, path);
    auto it = o.attrs.find(attr);
    if (it == o.attrs.end()) return false;
    auto v = h5l::as<T>(it->second);
    if (v.empty()) throw std::string("attribute ") + attr + " is empty";
    out = v[0];
    return true;
}
template <typename T> T h5_attr(const h5l::Node& g, const std::string& path, const std::string& attr) {
    T v;
    if (!attr_try(g, path, attr, v)) throw "attribute " + attr + " not present";
    return v;
}
template <typename T> T h5_attr(const h5l::Node& g, const std::string& path, const std::string& attr, T dflt) {
    T v;
    return attr_try(g, path, attr, v) ? v : dflt;
}
template float h5_attr<float>(const h5l::Node&, const std::string&, const std::string&);
template int h5_attr<int>(const h5l::Node&, const std::string&, const std::string&);
template float h5_attr<float>(const h5l::Node&, const std::string&, const std::string&, float);
template int h5_attr<int>(const h5l::Node&, const std::string&, const std::string&, int);

std::vector<std::string> h5_attr_strings(const h5l::Node& g, const std::string& attr) {
    auto it = g.attrs.find(attr);
    if (it == g.attrs.end()) throw "attribute " + attr + " not present";
    return h5l::as_strings(it->second);
}

// ================================================================================================ registry
std::map<std::string, NodeCreationFunction>& node_creation_map() {
    static std::map<std::string, NodeCreationFunction> m;
    return m;
}
static bool is_prefix(const std::string& a, const std::string& b) { return a == b.substr(0, a.size()); }
void add_node_creation_function(std::string name_prefix, NodeCreationFunction fcn) {
    auto& m = node_creation_map();
    // no registered name may be a prefix of another (reference deriv_engine.cpp:54-67)
    for (auto& kv : m)
        if (is_prefix(kv.first, name_prefix) || is_prefix(name_prefix, kv.first)) {
            fprintf(stderr, "Internal error.  Type name %s conflicts with %s.\n", name_prefix.c_str(), kv.first.c_str());
            throw std::string("node type name prefix conflict");
        }
    m[name_prefix] = fcn;
}
void check_elem_width(const CoordNode& n, int w) {
    if (n.elem_width != w)
        throw "expected argument with width " + std::to_string(w) + " but received argument with width " +
            std::to_string(n.elem_width);
}
void check_elem_width_lower_bound(const CoordNode& n, int w) {
    if (n.elem_width < w)
        throw "expected argument with width at least " + std::to_string(w) + " but received argument with width " +
            std::to_string(n.elem_width);
}
void check_arguments_length(const ArgList& a, int n) {
    if ((int)a.size() != n) throw "expected " + std::to_string(n) + " arguments but got " + std::to_string(a.size());
}

void sort_reference_order(std::vector<int>& i1, std::vector<int>& i2) {
    std::vector<size_t> idx(i1.size());
    for (size_t i = 0; i < idx.size(); ++i) idx[i] = i;
    std::sort(idx.begin(), idx.end(), [&](size_t a, size_t b) {
        int ba = i1[a] >> 2, bb = i1[b] >> 2;
        if (ba != bb) return ba < bb;
        if (i2[a] != i2[b]) return i2[a] < i2[b];
        return (i1[a] & 3) < (i1[b] & 3);
    });
    std::vector<int> a(i1.size()), b(i1.size());
    for (size_t i = 0; i < idx.size(); ++i) { a[i] = i1[idx[i]]; b[i] = i2[idx[i]]; }
    i1.swap(a);
    i2.swap(b);
}

// ================================================================================================ engine
Engine::Engine(int n_atom_, int n_rep_, int device_) : n_rep(n_rep_), n_atom(n_atom_), device(device_) {
    if (n_rep < 1) throw std::string("need at least one replica");
    int n_dev = 0;
    cudaError_t e = cudaGetDeviceCount(&n_dev);
    if (e != cudaSuccess || n_dev == 0)
        throw std::string("no CUDA device available: this engine has no CPU fallback (") + cudaGetErrorString(e) + ")";
    UB_CUDA(cudaSetDevice(device));
    UB_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    if (const char* e = getenv("UPSIDE_B200_NO_DAG")) use_dag = atoi(e) == 0;
    nodes.emplace_back();
    nodes[0].name = "pos";
    nodes[0].computation.reset(new Pos(n_atom));
    nodes[0].computation->engine = this;
    nodes[0].computation->name = "pos";
    pos = static_cast<Pos*>(nodes[0].computation.get());
}

Engine::~Engine() {
    cudaSetDevice(device);
    if (mc) mc_destroy(mc);
    for (auto& g : graph_eval) if (g) cudaGraphExecDestroy(g);
    if (graph_round) cudaGraphExecDestroy(graph_round);
    if (graph_host_eval) cudaGraphExecDestroy(graph_host_eval);
    if (pinned_io) cudaFreeHost(pinned_io);
    nodes.clear();
    for (auto st : node_stream) cudaStreamDestroy(st);
    for (auto e : ev_fwd) cudaEventDestroy(e);
    for (auto e : ev_bwd) cudaEventDestroy(e);
    if (ev_start) cudaEventDestroy(ev_start);
    if (ev_zeroed) cudaEventDestroy(ev_zeroed);
    if (zero_stream) cudaStreamDestroy(zero_stream);
    if (stream) cudaStreamDestroy(stream);
}

int Engine::get_idx(const std::string& name, bool must_exist) const {
    for (size_t i = 0; i < nodes.size(); ++i) if (nodes[i].name == name) return (int)i;
    if (must_exist) throw std::string("name not found");
    return -1;
}
DerivComputation& Engine::get(const std::string& name) { return *nodes[get_idx(name)].computation; }

void Engine::add_node(const std::string& name, std::unique_ptr<DerivComputation> c, const std::vector<std::string>& args) {
    if (get_idx(name, false) != -1) throw std::string("name conflict in DerivEngine");
    nodes.emplace_back();
    auto& n = nodes.back();
    n.name = name;
    c->engine = this;
    c->name = name;
    n.computation = std::move(c);
    for (auto& a : args) {
        int p = get_idx(a);
        n.parents.push_back(p);
        nodes[p].children.push_back(nodes.size() - 1);
    }
}

void Engine::allocate() {
    UB_CUDA(cudaSetDevice(device));
    size_t total = 0;
    n_pot_nodes = 0;
    for (auto& n : nodes) {
        if (n.computation->potential_term) { ++n_pot_nodes; continue; }
        auto* c = static_cast<CoordNode*>(n.computation.get());
        total += (c->stride() * n_rep + 3) & ~size_t(3);
    }
    out_arena.alloc(total);
    sens_arena.alloc(total);
    pot_arena.alloc(size_t(std::max(n_pot_nodes, 1)) * n_rep);
    potential.alloc(n_rep);
    error_flag.alloc(1);
    size_t off = 0;
    int ip = 0;
    std::vector<float*> pp;
    for (auto& n : nodes) {
        if (n.computation->potential_term) {
            auto* p = static_cast<PotentialNode*>(n.computation.get());
            p->potential = pot_arena.p + size_t(ip++) * n_rep;
            pp.push_back(p->potential);
        } else {
            auto* c = static_cast<CoordNode*>(n.computation.get());
            c->output = out_arena.p + off;
            c->sens = sens_arena.p + off;
            off += (c->stride() * n_rep + 3) & ~size_t(3);
        }
    }
    if (pp.empty()) pp.push_back(nullptr);
    pot_ptrs.upload(pp);
    mom.alloc(size_t(n_rep) * n_atom * 4);
    for (auto& n : nodes) n.computation->finalize();
}

__global__ void k_sum_potentials(float* __restrict__ total, float* const* __restrict__ ptrs, int n_pot, int n_rep) {
    int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_rep) return;
    float s = 0.f;
    for (int k = 0; k < n_pot; ++k) s += ptrs[k][r];
    total[r] = s;
}

void Engine::enqueue_compute(cudaStream_t s, ComputeMode mode) {
    // The reference zeroes each CoordNode's sens right after its forward pass and lets PotentialNodes add into
    // their parents' sens during compute_value (deriv_engine.cpp:143-151).  Zeroing everything first and running
    // forward in construction (topological) order, then backward in reverse order, is equivalent.
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    UB_CUDA(cudaStreamIsCapturing(s, &cap));
    if (use_dag && cap == cudaStreamCaptureStatusActive) {
        enqueue_compute_dag(s, mode);   // (zeroes the arenas on a branch of its own)
    } else {
        UB_CUDA(cudaMemsetAsync(sens_arena.p, 0, sens_arena.n * sizeof(float), s));
        // every mode: nodes that report their potential in DerivMode too (tension, AFM) add into a zeroed slot, so the slot
        // holds the last evaluation's value as in the reference, not a running sum
        UB_CUDA(cudaMemsetAsync(pot_arena.p, 0, pot_arena.n * sizeof(float), s));
        for (auto& n : nodes) { n.computation->compute_value(s, mode); mark(s, n.name + ":fwd"); }
        for (size_t i = nodes.size(); i-- > 0;)
            if (!nodes[i].computation->potential_term) { nodes[i].computation->propagate_deriv(s); mark(s, nodes[i].name + ":bwd"); }
    }
    if (mode == PotentialAndDerivMode && n_pot_nodes)
        k_sum_potentials<<<(n_rep + 127) / 128, 128, 0, s>>>(potential.p, pot_ptrs.p, n_pot_nodes, n_rep);
}

// The same work as the loop above, issued on one stream per node with event edges:
//   forward  of node i : after the forward of its parents;
//   backward of node i : after the backward (CoordNode) / forward (PotentialNode) of each child, which is what fills sens_i;
//   a kernel set that adds into a parent's sens (backward of a CoordNode, forward of a PotentialNode) additionally waits for
//   the previous writer of that buffer: the gather-form kernels update sens with plain read-modify-writes, so writers of
//   one buffer stay serialised, in the same order as in the sequential schedule (bit-identical sums).
void Engine::enqueue_compute_dag(cudaStream_t s, ComputeMode mode) {
    const size_t N = nodes.size();
    if (node_stream.size() != N) {
        for (auto st : node_stream) cudaStreamDestroy(st);
        for (auto e : ev_fwd) cudaEventDestroy(e);
        for (auto e : ev_bwd) cudaEventDestroy(e);
        node_stream.assign(N, nullptr); ev_fwd.assign(N, nullptr); ev_bwd.assign(N, nullptr);
        for (size_t i = 0; i < N; ++i) {
            UB_CUDA(cudaStreamCreateWithFlags(&node_stream[i], cudaStreamNonBlocking));
            UB_CUDA(cudaEventCreateWithFlags(&ev_fwd[i], cudaEventDisableTiming));
            UB_CUDA(cudaEventCreateWithFlags(&ev_bwd[i], cudaEventDisableTiming));
        }
        if (!ev_start) UB_CUDA(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
        if (!ev_zeroed) UB_CUDA(cudaEventCreateWithFlags(&ev_zeroed, cudaEventDisableTiming));
        if (!zero_stream) UB_CUDA(cudaStreamCreateWithFlags(&zero_stream, cudaStreamNonBlocking));
    }
    UB_CUDA(cudaEventRecord(ev_start, s));
    // The sens and potential arenas (56 KB per replica in ff_1) are zeroed on a branch of their own: only the kernels that ADD
    // into them wait for it - the forward pass of a PotentialNode and every backward pass - so the forward kernels of the
    // coordinate nodes start at once instead of behind a 230 MB memset (B = 4096).
    UB_CUDA(cudaStreamWaitEvent(zero_stream, ev_start, 0));
    UB_CUDA(cudaMemsetAsync(sens_arena.p, 0, sens_arena.n * sizeof(float), zero_stream));
    UB_CUDA(cudaMemsetAsync(pot_arena.p, 0, pot_arena.n * sizeof(float), zero_stream));
    UB_CUDA(cudaEventRecord(ev_zeroed, zero_stream));
    std::vector<cudaEvent_t> last_writer(N, nullptr);   // last kernel set that added into sens of node i
    std::vector<cudaEvent_t> last_event(N, nullptr);    // last thing recorded on node i's stream (for the final join)
    for (size_t i = 0; i < N; ++i) {
        cudaStream_t st = node_stream[i];
        auto& nd = nodes[i];
        UB_CUDA(cudaStreamWaitEvent(st, ev_start, 0));
        for (size_t p : nd.parents) if (p) UB_CUDA(cudaStreamWaitEvent(st, ev_fwd[p], 0));
        const bool pot = nd.computation->potential_term;
        if (pot) UB_CUDA(cudaStreamWaitEvent(st, ev_zeroed, 0));
        if (pot) for (size_t p : nd.parents) if (last_writer[p]) UB_CUDA(cudaStreamWaitEvent(st, last_writer[p], 0));
        nd.computation->compute_value(st, mode);
        UB_CUDA(cudaEventRecord(ev_fwd[i], st));
        last_event[i] = ev_fwd[i];
        if (pot) for (size_t p : nd.parents) last_writer[p] = ev_fwd[i];
    }
    for (size_t i = N; i-- > 0;) {
        auto& nd = nodes[i];
        if (nd.computation->potential_term) continue;
        cudaStream_t st = node_stream[i];
        UB_CUDA(cudaStreamWaitEvent(st, ev_zeroed, 0));
        for (size_t c : nd.children) UB_CUDA(cudaStreamWaitEvent(st, nodes[c].computation->potential_term ? ev_fwd[c] : ev_bwd[c], 0));
        for (size_t p : nd.parents) if (last_writer[p]) UB_CUDA(cudaStreamWaitEvent(st, last_writer[p], 0));
        nd.computation->propagate_deriv(st);
        UB_CUDA(cudaEventRecord(ev_bwd[i], st));
        last_event[i] = ev_bwd[i];
        for (size_t p : nd.parents) last_writer[p] = ev_bwd[i];
    }
    for (size_t i = 0; i < N; ++i) UB_CUDA(cudaStreamWaitEvent(s, last_event[i], 0));
    UB_CUDA(cudaStreamWaitEvent(s, ev_zeroed, 0));
}

void Engine::compute(ComputeMode mode) {
    UB_CUDA(cudaSetDevice(device));
    if (!use_graphs) {
        enqueue_compute(stream, mode);
        UB_CUDA(cudaGetLastError());
        return;
    }
    if (!graph_eval[mode]) {
        cudaGraph_t g;
        UB_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        enqueue_compute(stream, mode);
        UB_CUDA(cudaStreamEndCapture(stream, &g));
        UB_CUDA(cudaGraphInstantiate(&graph_eval[mode], g, 0));
        UB_CUDA(cudaGraphDestroy(g));
    }
    UB_CUDA(cudaGraphLaunch(graph_eval[mode], stream));
}

__global__ void k_pack3to4(float* __restrict__ dst, const float* __restrict__ src, long n);
__global__ void k_unpack4to3(float* __restrict__ dst, const float* __restrict__ src, long n);
void Engine::evaluate_host(const float* pos3, float* energy, float* deriv3) {
    UB_CUDA(cudaSetDevice(device));
    const size_t n3 = size_t(n_rep) * n_atom * 3;
    const long cnt = long(n_rep) * n_atom;
    if (!pinned_io) UB_CUDA(cudaMallocHost((void**)&pinned_io, sizeof(float) * (2 * n3 + n_rep + 1)));
    float* h_in = pinned_io, *h_out = pinned_io + n3, *h_en = pinned_io + 2 * n3;
    int* h_err = reinterpret_cast<int*>(pinned_io + 2 * n3 + n_rep);
    if (!graph_host_eval) {
        float* tmp = io_staging(2 * n3);
        cudaGraph_t g;
        UB_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        UB_CUDA(cudaMemcpyAsync(tmp, h_in, n3 * sizeof(float), cudaMemcpyHostToDevice, stream));
        k_pack3to4<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(pos->output, tmp, cnt);
        enqueue_compute(stream, PotentialAndDerivMode);
        k_unpack4to3<<<(unsigned)((cnt + 255) / 256), 256, 0, stream>>>(tmp + n3, pos->sens, cnt);
        UB_CUDA(cudaMemcpyAsync(h_out, tmp + n3, n3 * sizeof(float), cudaMemcpyDeviceToHost, stream));
        UB_CUDA(cudaMemcpyAsync(h_en, potential.p, n_rep * sizeof(float), cudaMemcpyDeviceToHost, stream));
        UB_CUDA(cudaMemcpyAsync(h_err, error_flag.p, sizeof(int), cudaMemcpyDeviceToHost, stream));
        UB_CUDA(cudaStreamEndCapture(stream, &g));
        UB_CUDA(cudaGraphInstantiate(&graph_host_eval, g, 0));
        UB_CUDA(cudaGraphDestroy(g));
    }
    memcpy(h_in, pos3, n3 * sizeof(float));
    UB_CUDA(cudaGraphLaunch(graph_host_eval, stream));
    UB_CUDA(cudaStreamSynchronize(stream));
    if (*h_err) sync_and_check();   // reports (and clears) the device-side failure
    if (energy) memcpy(energy, h_en, n_rep * sizeof(float));
    if (deriv3) memcpy(deriv3, h_out, n3 * sizeof(float));
}

std::vector<std::pair<std::string, float>> Engine::profile_eval(ComputeMode mode) {
    UB_CUDA(cudaSetDevice(device));
    UB_CUDA(cudaStreamSynchronize(stream));
    profiling = true;
    marks.clear();
    mark(stream, "start");
    enqueue_compute(stream, mode);   // not capturing: linear order, one mark after every kernel group
    UB_CUDA(cudaStreamSynchronize(stream));
    profiling = false;
    std::vector<std::pair<std::string, float>> out;
    for (size_t i = 1; i < marks.size(); ++i) {
        float ms = 0.f;
        UB_CUDA(cudaEventElapsedTime(&ms, marks[i - 1].second, marks[i].second));
        out.emplace_back(marks[i].first, ms);
    }
    for (auto& m : marks) cudaEventDestroy(m.second);
    marks.clear();
    return out;
}

void Engine::sync_and_check() {
    UB_CUDA(cudaSetDevice(device));
    UB_CUDA(cudaStreamSynchronize(stream));
    UB_CUDA(cudaGetLastError());
    int flag = 0;
    UB_CUDA(cudaMemcpy(&flag, error_flag.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        UB_CUDA(cudaMemset(error_flag.p, 0, sizeof(int)));
        if (flag == 1)
            throw std::string("pair-list capacity exceeded on the device (raise UPSIDE_B200_NEIGHBOR_SCALE)");
        if (flag == 2) throw std::string("rotamer residue-pair capacity exceeded on the device");
        if (flag == 3) throw std::string("rotamer bead-pair capacity exceeded on the device (raise UPSIDE_B200_NEIGHBOR_SCALE)");
        if (flag == 4) throw std::string("rotamer build: candidate / active residue-pair capacity exceeded on the device (raise UPSIDE_B200_NEIGHBOR_SCALE)");
        throw std::string("device-side failure flag ") + std::to_string(flag);
    }
}

// ------------------------------------------------------------------------------------------------ host I/O
__global__ void k_pack3to4(float* __restrict__ dst, const float* __restrict__ src, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = make_float4(src[3 * i], src[3 * i + 1], src[3 * i + 2], 0.f);
    reinterpret_cast<float4*>(dst)[i] = v;
}
__global__ void k_unpack4to3(float* __restrict__ dst, const float* __restrict__ src, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 v = reinterpret_cast<const float4*>(src)[i];
    dst[3 * i] = v.x; dst[3 * i + 1] = v.y; dst[3 * i + 2] = v.z;
}

// device staging area for (n,3) <-> (n,4) conversion of host buffers: allocated once, grown on demand (cudaMalloc/cudaFree per
// call would serialise the device and dominate a per-step upload/download loop)
float* Engine::io_staging(size_t n_float) {
    if (io_stage.n < n_float) io_stage.alloc(n_float);
    return io_stage.p;
}
static void upload3(Engine& e, float* dev4, const float* host3, int first_rep, int n) {
    if (n < 0) n = e.n_rep - first_rep;
    if (first_rep < 0 || first_rep + n > e.n_rep) throw std::string("replica range out of bounds");
    long cnt = long(n) * e.n_atom;
    if (!cnt) return;
    float* tmp = e.io_staging(size_t(cnt) * 3);
    UB_CUDA(cudaMemcpyAsync(tmp, host3, size_t(cnt) * 3 * sizeof(float), cudaMemcpyHostToDevice, e.stream));
    k_pack3to4<<<(unsigned)((cnt + 255) / 256), 256, 0, e.stream>>>(dev4 + size_t(first_rep) * e.n_atom * 4, tmp, cnt);
    UB_CUDA(cudaStreamSynchronize(e.stream));
}
static void download3(Engine& e, const float* dev4, float* host3, int first_rep, int n) {
    if (n < 0) n = e.n_rep - first_rep;
    if (first_rep < 0 || first_rep + n > e.n_rep) throw std::string("replica range out of bounds");
    long cnt = long(n) * e.n_atom;
    if (!cnt) return;
    float* tmp = e.io_staging(size_t(cnt) * 3);
    k_unpack4to3<<<(unsigned)((cnt + 255) / 256), 256, 0, e.stream>>>(tmp, dev4 + size_t(first_rep) * e.n_atom * 4, cnt);
    UB_CUDA(cudaMemcpyAsync(host3, tmp, size_t(cnt) * 3 * sizeof(float), cudaMemcpyDeviceToHost, e.stream));
    UB_CUDA(cudaStreamSynchronize(e.stream));
}
void Engine::set_pos(const float* p, int first_rep, int n) { UB_CUDA(cudaSetDevice(device)); upload3(*this, pos->output, p, first_rep, n); }
void Engine::get_pos(float* p, int first_rep, int n) { UB_CUDA(cudaSetDevice(device)); download3(*this, pos->output, p, first_rep, n); }
void Engine::get_deriv(float* p, int first_rep, int n) { UB_CUDA(cudaSetDevice(device)); download3(*this, pos->sens, p, first_rep, n); }
void Engine::set_mom(const float* p) { UB_CUDA(cudaSetDevice(device)); upload3(*this, mom.p, p, 0, -1); }
void Engine::get_mom(float* p) { UB_CUDA(cudaSetDevice(device)); download3(*this, mom.p, p, 0, -1); }
std::vector<float> Engine::get_potential() {
    UB_CUDA(cudaSetDevice(device));
    UB_CUDA(cudaStreamSynchronize(stream));
    return potential.download();
}

// ------------------------------------------------------------------------------------------------ RNG
// (Threefry-4x32-20, u01/uneg11 and Box-Muller: rng.cuh)
// p <- mom_scale*p + noise_scale[r]*N(0,1); key (seed_r, stream 0, 0, 0), counter (t_lo, t_hi, atom, 0)
__global__ void k_thermostat(float* __restrict__ mom, const uint32_t* __restrict__ seed,
                             const float* __restrict__ noise_scale, const unsigned long long* __restrict__ invocation,
                             float mom_scale, int n_atom, int n_rep) {
    long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (idx >= long(n_rep) * n_atom) return;
    int r = (int)(idx / n_atom), na = (int)(idx % n_atom);
    unsigned long long t = invocation[0];
    uint32_t key[4] = {seed[r], 0u, 0u, 0u};
    uint32_t ctr[4] = {(uint32_t)(t & 0xffffffffull), (uint32_t)(t >> 32), (uint32_t)na, 0u};
    uint32_t bits[4];
    threefry4x32_20(bits, ctr, key);
    float n0, n1, n2, n3;
    boxmuller_f(bits[0], bits[1], n0, n1);
    boxmuller_f(bits[2], bits[3], n2, n3);
    float4* m = reinterpret_cast<float4*>(mom) + idx;
    float4 p = *m;
    float ns = noise_scale[r];
    p.x = mom_scale * p.x + ns * n0;
    p.y = mom_scale * p.y + ns * n1;
    p.z = mom_scale * p.z + ns * n2;
    *m = p;
}
__global__ void k_increment(unsigned long long* c) { c[0] += 1ull; }

// test hook: raw generator output for (seed, stream, atom, t)
__global__ void k_rng_probe(uint32_t* bits_out, float* normal_out, uint32_t seed, uint32_t stream, uint32_t atom,
                            unsigned long long t) {
    uint32_t key[4] = {seed, stream, 0u, 0u};
    uint32_t ctr[4] = {(uint32_t)(t & 0xffffffffull), (uint32_t)(t >> 32), atom, 0u};
    uint32_t bits[4];
    threefry4x32_20(bits, ctr, key);
    for (int i = 0; i < 4; ++i) bits_out[i] = bits[i];
    float a, b, c, d;
    boxmuller_f(bits[0], bits[1], a, b);
    boxmuller_f(bits[2], bits[3], c, d);
    normal_out[0] = a; normal_out[1] = b; normal_out[2] = c;
    normal_out[3] = u01_f(bits[0]);
}
void rng_probe(uint32_t seed, uint32_t stream, uint32_t atom, unsigned long long t, uint32_t* bits4, float* normal3_u01) {
    DevBuf<uint32_t> b(4);
    DevBuf<float> f(4);
    k_rng_probe<<<1, 1>>>(b.p, f.p, seed, stream, atom, t);
    UB_CUDA(cudaDeviceSynchronize());
    auto hb = b.download();
    auto hf = f.download();
    for (int i = 0; i < 4; ++i) { bits4[i] = hb[i]; normal3_u01[i] = hf[i]; }
}

// ------------------------------------------------------------------------------------------------ integrator
// p -= vel_factor*dV/dx ; x += pos_factor*p   (unit mass, no force clipping: main.cpp:663 passes max_force=0)
__global__ void k_integration_stage(float* __restrict__ mom, float* __restrict__ pos, const float* __restrict__ deriv,
                                    float vel_factor, float pos_factor, long n) {
    long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 d = reinterpret_cast<const float4*>(deriv)[i];
    float4 p = reinterpret_cast<float4*>(mom)[i];
    float4 x = reinterpret_cast<float4*>(pos)[i];
    p.x -= vel_factor * d.x; p.y -= vel_factor * d.y; p.z -= vel_factor * d.z;
    x.x += pos_factor * p.x; x.y += pos_factor * p.y; x.z += pos_factor * p.z;
    reinterpret_cast<float4*>(mom)[i] = p;
    reinterpret_cast<float4*>(pos)[i] = x;
}

void Engine::set_temperature(const float* T) {
    UB_CUDA(cudaSetDevice(device));
    h_temperature.assign(T, T + n_rep);
    // OrnsteinUhlenbeckThermostat::update_parameters (thermostat.h:11-14)
    double delta_t = double(thermostat_interval) * 3. * double(dt);
    mom_scale = (float)std::exp(-(float)delta_t / thermostat_timescale);
    std::vector<float> ns(n_rep);
    for (int r = 0; r < n_rep; ++r) ns[r] = sqrtf(h_temperature[r] * (1 - mom_scale * mom_scale));
    noise_scale.upload(ns);
    temperature.upload(h_temperature);
}

void Engine::md_init(uint32_t base_seed, const float* T, float dt_, float timescale, int interval) {
    std::vector<uint32_t> seeds(n_rep);
    for (int r = 0; r < n_rep; ++r) seeds[r] = base_seed + (uint32_t)r;   // main.cpp:459
    md_init_seeds(seeds.data(), T, dt_, timescale, interval);
}

void Engine::md_init_seeds(const uint32_t* seeds, const float* T, float dt_, float timescale, int interval) {
    UB_CUDA(cudaSetDevice(device));
    dt = dt_;
    thermostat_timescale = timescale;
    thermostat_interval = std::max(1, interval);
    h_seed.assign(seeds, seeds + n_rep);
    seed.upload(h_seed);
    d_invocation.alloc(1);
    n_thermostat_invocations = 0;
    round_num = 0;
    UB_CUDA(cudaMemset(mom.p, 0, mom.n * sizeof(float)));
    // initial thermalisation: thermostat with delta_t = 1e8 => mom_scale = 0, noise = sqrt(T)  (main.cpp:515-521)
    h_temperature.assign(T, T + n_rep);
    std::vector<float> ns(n_rep);
    for (int r = 0; r < n_rep; ++r) ns[r] = sqrtf(h_temperature[r] * 1.f);
    noise_scale.upload(ns);
    mom_scale = 0.f;
    enqueue_thermostat(stream);
    UB_CUDA(cudaStreamSynchronize(stream));
    set_temperature(T);   // true thermostat interval (main.cpp:522)
    if (graph_round) { cudaGraphExecDestroy(graph_round); graph_round = nullptr; }
}

void Engine::enqueue_thermostat(cudaStream_t s) {
    long n = long(n_rep) * n_atom;
    k_thermostat<<<(unsigned)((n + 127) / 128), 128, 0, s>>>(mom.p, seed.p, noise_scale.p, d_invocation.p, mom_scale,
                                                              n_atom, n_rep);
    k_increment<<<1, 1, 0, s>>>(d_invocation.p);
    ++n_thermostat_invocations;
}

void Engine::enqueue_integration_cycle(cudaStream_t s) {
    // Verlet: a=1/6, b=1/3 => mom_update = pos_update = {1,1,1} (deriv_engine.cpp:176-180)
    const float a = 1.f / 6.f, b = 1.f / 3.f;
    const float mom_update[3] = {1.5f - 3.f * a, 1.5f - 3.f * a, 6.f * a};
    const float pos_update[3] = {3.f * b, 3.0f - 6.f * b, 3.f * b};
    long n = long(n_rep) * n_atom;
    for (int stage = 0; stage < 3; ++stage) {
        enqueue_compute(s, DerivMode);
        k_integration_stage<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(mom.p, pos->output, pos->sens,
                                                                        dt * mom_update[stage], dt * pos_update[stage], n);
    }
}

void Engine::md_run(long n_round) {
    UB_CUDA(cudaSetDevice(device));
    if (!seed.p) throw std::string("md_init must be called before md_run");
    if (use_graphs && !graph_round) {
        cudaGraph_t g;
        UB_CUDA(cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal));
        enqueue_integration_cycle(stream);
        UB_CUDA(cudaStreamEndCapture(stream, &g));
        UB_CUDA(cudaGraphInstantiate(&graph_round, g, 0));
        UB_CUDA(cudaGraphDestroy(g));
    }
    for (long i = 0; i < n_round; ++i, ++round_num) {
        if (!(round_num % (uint64_t)thermostat_interval)) enqueue_thermostat(stream);
        if (use_graphs) UB_CUDA(cudaGraphLaunch(graph_round, stream));
        else enqueue_integration_cycle(stream);
    }
    UB_CUDA(cudaGetLastError());
}

// ---- checkpoint ------------------------------------------------------------------------------------------------------------
namespace {
struct CheckpointHeader {
    char magic[8];            // "UBCKPT1\0"
    int32_t n_rep, n_atom, thermostat_interval, reserved;
    float dt, thermostat_timescale;
    uint64_t n_thermostat_invocations, round_num;
};
}
std::vector<char> Engine::checkpoint_save() {
    UB_CUDA(cudaSetDevice(device));
    if (!seed.p) throw std::string("md_init must be called before a checkpoint is taken");
    sync_and_check();
    CheckpointHeader h{};
    memcpy(h.magic, "UBCKPT1", 8);
    h.n_rep = n_rep; h.n_atom = n_atom; h.thermostat_interval = thermostat_interval;
    h.dt = dt; h.thermostat_timescale = thermostat_timescale;
    h.n_thermostat_invocations = n_thermostat_invocations; h.round_num = round_num;
    const size_t n3 = size_t(n_rep) * n_atom * 3;
    std::vector<char> blob(sizeof(h) + sizeof(float) * (2 * n3 + n_rep) + sizeof(uint32_t) * n_rep);
    char* p = blob.data();
    memcpy(p, &h, sizeof(h)); p += sizeof(h);
    get_pos(reinterpret_cast<float*>(p)); p += sizeof(float) * n3;
    get_mom(reinterpret_cast<float*>(p)); p += sizeof(float) * n3;
    memcpy(p, h_temperature.data(), sizeof(float) * n_rep); p += sizeof(float) * n_rep;
    memcpy(p, h_seed.data(), sizeof(uint32_t) * n_rep);
    return blob;
}
void Engine::checkpoint_load(const char* data, size_t size) {
    UB_CUDA(cudaSetDevice(device));
    CheckpointHeader h;
    if (size < sizeof(h)) throw std::string("checkpoint too small");
    memcpy(&h, data, sizeof(h));
    if (memcmp(h.magic, "UBCKPT1", 8)) throw std::string("not a checkpoint of this engine");
    if (h.n_rep != n_rep || h.n_atom != n_atom)
        throw "checkpoint holds " + std::to_string(h.n_rep) + " replicas of " + std::to_string(h.n_atom) + " atoms, the engine " +
            std::to_string(n_rep) + " of " + std::to_string(n_atom);
    const size_t n3 = size_t(n_rep) * n_atom * 3;
    if (size != sizeof(h) + sizeof(float) * (2 * n3 + n_rep) + sizeof(uint32_t) * n_rep) throw std::string("checkpoint size mismatch");
    const char* p = data + sizeof(h);
    const float* pos_in = reinterpret_cast<const float*>(p); p += sizeof(float) * n3;
    const float* mom_in = reinterpret_cast<const float*>(p); p += sizeof(float) * n3;
    std::vector<float> T(n_rep);
    memcpy(T.data(), p, sizeof(float) * n_rep); p += sizeof(float) * n_rep;
    std::vector<uint32_t> seeds(n_rep);
    memcpy(seeds.data(), p, sizeof(uint32_t) * n_rep);
    md_init_seeds(seeds.data(), T.data(), h.dt, h.thermostat_timescale, h.thermostat_interval);   // (draws momenta: overwritten below)
    set_pos(pos_in);
    set_mom(mom_in);
    n_thermostat_invocations = h.n_thermostat_invocations;
    round_num = h.round_num;
    const unsigned long long inv = h.n_thermostat_invocations;
    UB_CUDA(cudaMemcpy(d_invocation.p, &inv, sizeof(inv), cudaMemcpyHostToDevice));
}

__global__ void k_recenter(float* __restrict__ pos, int n_atom, int xy_only) {
    __shared__ float sc[32];
    __shared__ float ctr[3];
    int r = blockIdx.x;
    float4* p = reinterpret_cast<float4*>(pos) + size_t(r) * n_atom;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = threadIdx.x; i < n_atom; i += blockDim.x) { float4 v = p[i]; sx += v.x; sy += v.y; sz += v.z; }
    sx = block_sum(sx, sc); if (threadIdx.x == 0) ctr[0] = sx / n_atom;
    sy = block_sum(sy, sc); if (threadIdx.x == 0) ctr[1] = sy / n_atom;
    sz = block_sum(sz, sc); if (threadIdx.x == 0) ctr[2] = xy_only ? 0.f : sz / n_atom;
    __syncthreads();
    for (int i = threadIdx.x; i < n_atom; i += blockDim.x) {
        float4 v = p[i];
        v.x -= ctr[0]; v.y -= ctr[1]; v.z -= ctr[2];
        p[i] = v;
    }
}
// exchange the coordinates of replica pairs (replica exchange swaps engine.pos only: main.cpp:240-243)
__global__ void k_swap_pos(float* __restrict__ pos, const int* __restrict__ pairs, int n_atom) {
    float4* a = reinterpret_cast<float4*>(pos) + size_t(pairs[2 * blockIdx.x]) * n_atom;
    float4* b = reinterpret_cast<float4*>(pos) + size_t(pairs[2 * blockIdx.x + 1]) * n_atom;
    for (int i = threadIdx.x; i < n_atom; i += blockDim.x) { float4 t = a[i]; a[i] = b[i]; b[i] = t; }
}
void Engine::swap_pos(const std::vector<int>& pairs) {
    UB_CUDA(cudaSetDevice(device));
    if (pairs.empty()) return;
    for (int v : pairs) if (v < 0 || v >= n_rep) throw std::string("replica index out of range in swap");
    DevBuf<int> d;
    d.upload(pairs);
    k_swap_pos<<<(unsigned)(pairs.size() / 2), 128, 0, stream>>>(pos->output, d.p, n_atom);
    UB_CUDA(cudaStreamSynchronize(stream));
}
void Engine::recenter(bool xy_only) {
    UB_CUDA(cudaSetDevice(device));
    k_recenter<<<n_rep, 128, 0, stream>>>(pos->output, n_atom, xy_only ? 1 : 0);
}

__global__ void k_kinetic(float* __restrict__ out, const float* __restrict__ mom, int n_atom) {
    __shared__ float sc[32];
    int r = blockIdx.x;
    const float4* p = reinterpret_cast<const float4*>(mom) + size_t(r) * n_atom;
    float s = 0.f;
    for (int i = threadIdx.x; i < n_atom; i += blockDim.x) { float4 v = p[i]; s += v.x * v.x + v.y * v.y + v.z * v.z; }
    s = block_sum(s, sc);
    if (threadIdx.x == 0) out[r] = 0.5f * s / n_atom;   // kinetic energy per atom, as the reference logs (main.cpp:529-535)
}
std::vector<float> Engine::kinetic_energy() {
    UB_CUDA(cudaSetDevice(device));
    DevBuf<float> out(n_rep);
    k_kinetic<<<n_rep, 128, 0, stream>>>(out.p, mom.p, n_atom);
    UB_CUDA(cudaStreamSynchronize(stream));
    return out.download();
}

// ================================================================================================ construction
std::unique_ptr<Engine> initialize_engine_from_hdf5(int n_atom, const h5l::Node& potential_group, int n_rep, int device) {
    std::unique_ptr<Engine> engine(new Engine(n_atom, n_rep, device));
    auto& m = node_creation_map();

    // name-sorted dependency map, then repeated sweeps adding every node whose arguments are all placed
    // (reference deriv_engine.cpp:200-226); this fixes the node order and hence the accumulation order.
    std::map<std::string, std::pair<bool, std::vector<std::string>>> dep;
    dep["pos"] = {true, {}};
    for (auto& kv : potential_group.children) {
        if (!kv.second->is_group) continue;
        dep[kv.first] = {true, h5_attr_strings(*kv.second, "arguments")};
    }
    for (auto& kv : dep)
        for (auto& d : kv.second.second)
            if (!dep.count(d))
                throw "Node " + kv.first + " takes " + d + " as an argument, but no node of that name can be found.";
    std::vector<std::string> topo;
    auto in_topo = [&](const std::string& n) { return std::find(topo.begin(), topo.end(), n) != topo.end(); };
    for (size_t round = 0; round < dep.size(); ++round)
        for (auto& kv : dep) {
            if (!kv.second.first) continue;
            if (std::all_of(kv.second.second.begin(), kv.second.second.end(), in_topo)) {
                topo.push_back(kv.first);
                kv.second.first = false;
            }
        }
    for (auto& kv : dep) if (kv.second.first) throw "Unsatisfiable dependency " + kv.first + " in potential computation";

    for (auto& nm : topo) {
        if (nm == "pos") continue;
        std::string type_name;
        for (auto& kv : m) if (is_prefix(kv.first, nm)) type_name = kv.first;
        if (type_name.empty()) throw "No node type found for name '" + nm + "'";
        auto& arg_names = dep[nm].second;
        ArgList args;
        for (auto& an : arg_names) {
            auto* c = dynamic_cast<CoordNode*>(&engine->get(an));
            if (!c) throw an + " is not an intermediate value, but it is an argument of " + nm;
            args.push_back(c);
        }
        try {
            const h5l::Node& grp = h5_child(potential_group, nm);
            std::unique_ptr<DerivComputation> comp(m[type_name](*engine, grp, args));
            engine->add_node(nm, std::move(comp), arg_names);
        } catch (const std::string& e) {
            throw "while adding '" + nm + "', " + e;
        }
    }
    engine->allocate();
    return engine;
}

}  // namespace ub
