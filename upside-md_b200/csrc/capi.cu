// C ABI: include/engine_c_library.h (the reference's single-replica interface) and include/upside_b200.h (batched).
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/engine_c_library.h"
#include "../../include/upside_b200.h"
#include "engine.h"
#include "replica_exchange.h"
#include "spline_fit.h"

namespace ub {
void rng_probe(uint32_t seed, uint32_t stream, uint32_t atom, unsigned long long t, uint32_t* bits4, float* normal3_u01);
}

struct UbEngine {
    std::unique_ptr<ub::Engine> eng;
    std::vector<float> initial_pos;   // (n_atom,3)
    int launches_per_eval = 0;
};
struct DerivEngine {
    UbEngine u;
};

namespace ub { Engine* engine_of(UbEngine* e) { return e->eng.get(); } }   // for the translation units that extend the ABI

static thread_local std::string g_last_error;

static int fail(const std::string& e) {
    g_last_error = e;
    fprintf(stderr, "\n\nERROR: %s\n", e.c_str());
    return 1;
}
#define UB_TRY try {
#define UB_CATCH                                                   \
    }                                                              \
    catch (const std::string& e) { return fail(e); }               \
    catch (const char* e) { return fail(e); }                      \
    catch (const std::exception& e) { return fail(e.what()); }     \
    catch (...) { return fail("unknown error"); }

static void load_engine(UbEngine& u, const char* path, int n_atom_expected, int n_rep, int device) {
    auto root = h5l::load(path);
    const h5l::Node* pot = h5l::find(root.get(), "/input/potential");
    if (!pot || !pot->is_group) throw std::string("unable to open group /input/potential (does it exist?)");
    int n_atom = n_atom_expected;
    const h5l::Node* pos = h5l::find(root.get(), "/input/pos");
    if (pos && !pos->is_group && pos->data.dims.size() == 3) {
        int na = (int)pos->data.dims[0];
        if (pos->data.dims[1] != 3 || pos->data.dims[2] != 1) throw std::string("invalid dimensions for /input/pos");
        if (n_atom < 0) n_atom = na;
        if (na == n_atom) {
            auto p = h5l::as<float>(pos->data);
            u.initial_pos.assign(p.begin(), p.end());
        }
    }
    if (n_atom < 0) throw std::string("/input/pos not found and n_atom not given");
    u.eng = ub::initialize_engine_from_hdf5(n_atom, *pot, n_rep, device);
    const h5l::Node* input = h5l::find(root.get(), "/input");
    if (input && (ub::h5_has(*input, "pivot_moves") || ub::h5_has(*input, "jump_moves"))) u.eng->mc_init(*input);
}

static ub::DerivComputation& node_of(UbEngine* e, const char* name) { return e->eng->get(name); }

static void copy_node(UbEngine* e, const char* node, int replica, int n, float* out, bool want_sens) {
    auto& eng = *e->eng;
    if (replica < 0 || replica >= eng.n_rep) throw std::string("replica out of range");
    eng.sync_and_check();
    auto& c = node_of(e, node);
    if (c.potential_term) {
        // a PotentialNode reports dims (1,1) and its potential for both queries (engine_c_library.cpp:115-118,139-142)
        if (n != 1) throw std::string("wrong size for potential node");
        auto* p = static_cast<ub::PotentialNode*>(&c);
        UB_CUDA(cudaMemcpy(out, p->potential + replica, sizeof(float), cudaMemcpyDeviceToHost));
        return;
    }
    auto* cn = static_cast<ub::CoordNode*>(&c);
    if (n != cn->n_elem * cn->elem_width) throw std::string("wrong number of elements");
    std::vector<float> tmp(cn->stride());
    const float* src = (want_sens ? cn->sens : cn->output) + size_t(replica) * cn->stride();
    UB_CUDA(cudaMemcpy(tmp.data(), src, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int i = 0; i < cn->n_elem; ++i)
        for (int d = 0; d < cn->elem_width; ++d) out[i * cn->elem_width + d] = tmp[size_t(i) * cn->wp + d];
}

extern "C" {

const char* ub_last_error(void) { return g_last_error.c_str(); }
int ub_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

UbEngine* ub_engine_create(const char* config_path, int n_replica, int device) {
    try {
        std::unique_ptr<UbEngine> u(new UbEngine);
        load_engine(*u, config_path, -1, n_replica, device);
        if (!u->initial_pos.empty()) {
            std::vector<float> all(size_t(n_replica) * u->initial_pos.size());
            for (int r = 0; r < n_replica; ++r) std::copy(u->initial_pos.begin(), u->initial_pos.end(), all.begin() + size_t(r) * u->initial_pos.size());
            u->eng->set_pos(all.data());
        }
        return u.release();
    } catch (const std::string& e) { fail(e); }
    catch (const std::exception& e) { fail(e.what()); }
    catch (...) { fail("unknown error"); }
    return nullptr;
}
void ub_engine_destroy(UbEngine* e) { delete e; }
int ub_n_atom(const UbEngine* e) { return e->eng->n_atom; }
int ub_n_replica(const UbEngine* e) { return e->eng->n_rep; }

int ub_initial_pos(UbEngine* e, float* pos) {
    UB_TRY
    if (e->initial_pos.empty()) throw std::string("configuration has no /input/pos");
    std::copy(e->initial_pos.begin(), e->initial_pos.end(), pos);
    return 0;
    UB_CATCH
}
int ub_set_pos(UbEngine* e, const float* pos) { UB_TRY e->eng->set_pos(pos); return 0; UB_CATCH }
int ub_get_pos(UbEngine* e, float* pos) { UB_TRY e->eng->sync_and_check(); e->eng->get_pos(pos); return 0; UB_CATCH }
int ub_set_mom(UbEngine* e, const float* mom) { UB_TRY e->eng->set_mom(mom); return 0; UB_CATCH }
int ub_get_mom(UbEngine* e, float* mom) { UB_TRY e->eng->sync_and_check(); e->eng->get_mom(mom); return 0; UB_CATCH }

int ub_evaluate(UbEngine* e, float* energy, float* deriv) {
    UB_TRY
    e->eng->compute(ub::PotentialAndDerivMode);
    e->eng->sync_and_check();
    if (energy) { auto p = e->eng->get_potential(); std::copy(p.begin(), p.end(), energy); }
    if (deriv) e->eng->get_deriv(deriv);
    return 0;
    UB_CATCH
}

int ub_n_nodes(UbEngine* e) { return (int)e->eng->nodes.size(); }
int ub_node_name(UbEngine* e, int index, char* buf, int buf_len, int* is_potential) {
    UB_TRY
    if (index < 0 || index >= (int)e->eng->nodes.size()) throw std::string("node index out of range");
    strncpy(buf, e->eng->nodes[index].name.c_str(), buf_len);
    if (buf_len) buf[buf_len - 1] = 0;
    if (is_potential) *is_potential = e->eng->nodes[index].computation->potential_term;
    return 0;
    UB_CATCH
}
int ub_get_output_dims(UbEngine* e, const char* node, int* n_elem, int* elem_width) {
    UB_TRY
    auto& c = node_of(e, node);
    if (c.potential_term) { *n_elem = 1; *elem_width = 1; }
    else { auto* cn = static_cast<ub::CoordNode*>(&c); *n_elem = cn->n_elem; *elem_width = cn->elem_width; }
    return 0;
    UB_CATCH
}
int ub_get_output(UbEngine* e, const char* node, int replica, int n, float* out) { UB_TRY copy_node(e, node, replica, n, out, false); return 0; UB_CATCH }
int ub_get_sens(UbEngine* e, const char* node, int replica, int n, float* out) { UB_TRY copy_node(e, node, replica, n, out, true); return 0; UB_CATCH }
int ub_get_node_potential(UbEngine* e, const char* node, float* out) {
    UB_TRY
    auto& c = node_of(e, node);
    if (!c.potential_term) throw std::string(node) + " is not a potential node";
    e->eng->sync_and_check();
    UB_CUDA(cudaMemcpy(out, static_cast<ub::PotentialNode*>(&c)->potential, sizeof(float) * e->eng->n_rep, cudaMemcpyDeviceToHost));
    return 0;
    UB_CATCH
}
int ub_get_value_by_name(UbEngine* e, const char* node, const char* log_name, int replica, int n, float* out, int* n_written) {
    UB_TRY
    auto v = node_of(e, node).get_value_by_name(replica, log_name);
    if (n_written) *n_written = (int)v.size();
    if (out) {
        if ((int)v.size() > n) throw std::string("buffer too small for get_value_by_name");
        std::copy(v.begin(), v.end(), out);
    }
    return 0;
    UB_CATCH
}
int ub_get_param(UbEngine* e, const char* node, int n, float* out, int* n_param) {
    UB_TRY
    auto v = node_of(e, node).get_param();
    if (n_param) *n_param = (int)v.size();
    if (out) {
        if ((int)v.size() != n) throw std::string("wrong number of parameters");
        std::copy(v.begin(), v.end(), out);
    }
    return 0;
    UB_CATCH
}
int ub_get_param_deriv(UbEngine* e, const char* node, int replica, int n, float* out, int* n_param) {
    UB_TRY
    auto v = node_of(e, node).get_param_deriv(replica);
    if (n_param) *n_param = (int)v.size();
    if (out) {
        if ((int)v.size() != n) throw std::string("wrong number of parameters (parameter derivatives are not implemented for this node)");
        std::copy(v.begin(), v.end(), out);
    }
    return 0;
    UB_CATCH
}
int ub_set_param(UbEngine* e, const char* node, int n, const float* param) {
    UB_TRY
    e->eng->sync_and_check();
    node_of(e, node).set_param(std::vector<float>(param, param + n));
    return 0;
    UB_CATCH
}
int ub_get_pairlist(UbEngine* e, const char* node, int replica, int max_edge, int* i1, int* i2, int* n_edge) {
    UB_TRY
    std::vector<int> a, b;
    if (!node_of(e, node).get_pairlist(replica, a, b)) throw std::string("node ") + node + " has no pair list";
    *n_edge = (int)a.size();
    for (int k = 0; k < (int)a.size() && k < max_edge; ++k) { i1[k] = a[k]; i2[k] = b[k]; }
    return 0;
    UB_CATCH
}

int ub_md_init(UbEngine* e, uint32_t base_seed, const float* temperature, float dt, float timescale, int interval) {
    UB_TRY e->eng->md_init(base_seed, temperature, dt, timescale, interval); return 0; UB_CATCH
}
int ub_md_init_seeds(UbEngine* e, const uint32_t* seeds, const float* temperature, float dt, float timescale, int interval) {
    UB_TRY e->eng->md_init_seeds(seeds, temperature, dt, timescale, interval); return 0; UB_CATCH
}
int ub_set_pos_range(UbEngine* e, const float* pos, int first_replica, int n_replica) {
    UB_TRY e->eng->set_pos(pos, first_replica, n_replica); return 0; UB_CATCH
}
int ub_get_pos_range(UbEngine* e, float* pos, int first_replica, int n_replica) {
    UB_TRY e->eng->sync_and_check(); e->eng->get_pos(pos, first_replica, n_replica); return 0; UB_CATCH
}
int ub_swap_pos(UbEngine* e, int n_pair, const int* pairs) {
    UB_TRY e->eng->sync_and_check(); e->eng->swap_pos(std::vector<int>(pairs, pairs + 2 * n_pair)); return 0; UB_CATCH
}
int ub_md_set_temperature(UbEngine* e, const float* temperature) { UB_TRY e->eng->set_temperature(temperature); return 0; UB_CATCH }
long ub_checkpoint_size(UbEngine* e) {
    const ub::Engine& g = *e->eng;
    return long(64 + sizeof(float) * (2 * size_t(g.n_rep) * g.n_atom * 3 + g.n_rep) + sizeof(uint32_t) * g.n_rep);
}
int ub_checkpoint_save(UbEngine* e, void* buf, long buf_size, long* written) {
    UB_TRY
    auto blob = e->eng->checkpoint_save();
    if (written) *written = (long)blob.size();
    if ((long)blob.size() > buf_size) throw std::string("checkpoint buffer too small");
    memcpy(buf, blob.data(), blob.size());
    return 0;
    UB_CATCH
}
int ub_checkpoint_load(UbEngine* e, const void* buf, long size) {
    UB_TRY
    e->eng->checkpoint_load(static_cast<const char*>(buf), (size_t)size);
    return 0;
    UB_CATCH
}
int ub_md_run(UbEngine* e, long n_round) { UB_TRY e->eng->md_run(n_round); return 0; UB_CATCH }
int ub_sync(UbEngine* e) { UB_TRY e->eng->sync_and_check(); return 0; UB_CATCH }
int ub_mc_n_samplers(UbEngine* e) { return e->eng->mc_n_samplers(); }
int ub_mc_sampler_name(UbEngine* e, int index, char* buf, int buf_len) {
    UB_TRY
    std::string nm = e->eng->mc_sampler_name(index);
    strncpy(buf, nm.c_str(), buf_len);
    if (buf_len > 0) buf[buf_len - 1] = 0;
    return 0;
    UB_CATCH
}
int ub_mc_execute(UbEngine* e, uint64_t round) { UB_TRY e->eng->mc_execute(round); return 0; UB_CATCH }
int ub_mc_stats(UbEngine* e, int index, uint64_t* n_success, uint64_t* n_attempt, int reset) {
    UB_TRY e->eng->mc_stats(index, n_success, n_attempt, reset != 0); return 0; UB_CATCH
}
int ub_recenter(UbEngine* e, int xy_only) { UB_TRY e->eng->recenter(xy_only != 0); return 0; UB_CATCH }
int ub_kinetic_energy(UbEngine* e, float* out) {
    UB_TRY
    auto k = e->eng->kinetic_energy();
    std::copy(k.begin(), k.end(), out);
    return 0;
    UB_CATCH
}
void* ub_stream(UbEngine* e) { return (void*)e->eng->stream; }
int ub_launches_per_eval(UbEngine* e) {
    // count kernel nodes of the captured evaluation graph
    try {
        auto& eng = *e->eng;
        UB_CUDA(cudaSetDevice(eng.device));
        cudaGraph_t g;
        UB_CUDA(cudaStreamBeginCapture(eng.stream, cudaStreamCaptureModeThreadLocal));
        eng.enqueue_compute(eng.stream, ub::DerivMode);
        UB_CUDA(cudaStreamEndCapture(eng.stream, &g));
        size_t n = 0;
        UB_CUDA(cudaGraphGetNodes(g, nullptr, &n));
        std::vector<cudaGraphNode_t> nodes(n);
        UB_CUDA(cudaGraphGetNodes(g, nodes.data(), &n));
        int k = 0;
        for (auto& nd : nodes) {
            cudaGraphNodeType t;
            UB_CUDA(cudaGraphNodeGetType(nd, &t));
            if (t == cudaGraphNodeTypeKernel) ++k;
        }
        UB_CUDA(cudaGraphDestroy(g));
        return k;
    } catch (...) { return -1; }
}
int ub_profile_eval(UbEngine* e, int max_entry, char* labels, int label_len, float* ms, int* n_entry) {
    UB_TRY
    auto v = e->eng->profile_eval(ub::DerivMode);
    e->eng->sync_and_check();
    *n_entry = (int)v.size();
    for (int i = 0; i < (int)v.size() && i < max_entry; ++i) {
        strncpy(labels + size_t(i) * label_len, v[i].first.c_str(), label_len);
        labels[size_t(i) * label_len + label_len - 1] = 0;
        ms[i] = v[i].second;
    }
    return 0;
    UB_CATCH
}
int ub_rng_probe(uint32_t seed, uint32_t stream, uint32_t atom, uint64_t t, uint32_t* bits4, float* normal3_u01) {
    UB_TRY ub::rng_probe(seed, stream, atom, t, bits4, normal3_u01); return 0; UB_CATCH
}

// ---- replica exchange: host-side plan (no device work; usable without a GPU) ------------------------------------------
struct UbReplex {
    ub::ReplicaExchangePlan plan;
    ub::HostRandomGenerator rng;
    UbReplex(int n, const std::vector<std::string>& sets) : plan(n, sets), rng(0u, ub::REPLICA_EXCHANGE_RANDOM_STREAM, 0u, 0ull) {}
};
UbReplex* ub_replex_create(int n_system, int n_set, const char* const* swap_sets) {
    try {
        std::vector<std::string> sets(swap_sets, swap_sets + n_set);
        return new UbReplex(n_system, sets);
    } catch (const std::string& e) { fail(e); }
    catch (const std::exception& e) { fail(e.what()); }
    catch (...) { fail("unknown error"); }
    return nullptr;
}
void ub_replex_destroy(UbReplex* h) { delete h; }
int ub_replex_n_sets(const UbReplex* h) { return (int)h->plan.swap_sets.size(); }
int ub_replex_set_size(const UbReplex* h, int set) { return set >= 0 && set < (int)h->plan.swap_sets.size() ? (int)h->plan.swap_sets[set].size() : -1; }
int ub_replex_pairs(const UbReplex* h, int set, int* pairs) {
    UB_TRY
    const auto& ss = h->plan.swap_sets.at(set);
    for (size_t i = 0; i < ss.size(); ++i) { pairs[2 * i] = ss[i].sys1; pairs[2 * i + 1] = ss[i].sys2; }
    return 0;
    UB_CATCH
}
int ub_replex_begin(UbReplex* h, uint32_t seed, uint64_t round) {
    h->rng = ub::HostRandomGenerator(seed, ub::REPLICA_EXCHANGE_RANDOM_STREAM, 0u, round);
    return 0;
}
int ub_replex_decide(UbReplex* h, int set, const float* old_lboltz, const float* new_lboltz, int* accept) {
    UB_TRY h->plan.decide(set, old_lboltz, new_lboltz, h->rng, accept); return 0; UB_CATCH
}
int ub_replex_decide_same_hamiltonian(UbReplex* h, int set, const float* beta, const float* energy, int* accept) {
    UB_TRY h->plan.decide_same_hamiltonian(set, beta, energy, h->rng, accept); return 0; UB_CATCH
}
int ub_replex_replica_indices(const UbReplex* h, int* out) {
    for (int i = 0; i < h->plan.n_system; ++i) out[i] = h->plan.replica_indices[i];
    return 0;
}
int ub_replex_counts(const UbReplex* h, int set, uint64_t* n_attempt, uint64_t* n_success) {
    UB_TRY
    const auto& ss = h->plan.swap_sets.at(set);
    for (size_t i = 0; i < ss.size(); ++i) { n_attempt[i] = ss[i].n_attempt; n_success[i] = ss[i].n_success; }
    return 0;
    UB_CATCH
}
int ub_host_rng_uniform(uint32_t seed, uint32_t stream, uint32_t atom, uint64_t timestep, int n_draw, float* out, uint32_t* bits_first) {
    ub::HostRandomGenerator g(seed, stream, atom, timestep);
    for (int i = 0; i < n_draw; ++i) {
        if (i == 0 && bits_first) {
            ub::HostRandomGenerator g2 = g;
            g2.random_bits(bits_first);
        }
        out[i] = g.uniform_open_closed_x();
    }
    return 0;
}

// ================================================================================ reference single-replica ABI
DerivEngine* construct_deriv_engine(int n_atom, const char* potential_file, bool quiet) {
    (void)quiet;
    try {
        std::unique_ptr<DerivEngine> d(new DerivEngine);
        load_engine(d->u, potential_file, n_atom, 1, 0);
        return d.release();
    } catch (const std::string& e) { fail(e); }
    catch (const std::exception& e) { fail(e.what()); }
    catch (...) { fail("unknown error"); }
    return nullptr;
}
void free_deriv_engine(DerivEngine* engine) { delete engine; }

int evaluate_energy(float* energy, DerivEngine* engine, const float* pos) {
    UB_TRY
    engine->u.eng->evaluate_host(pos, energy, nullptr);
    return 0;
    UB_CATCH
}
int evaluate_deriv(float* deriv, DerivEngine* engine, const float* pos) {
    UB_TRY
    engine->u.eng->evaluate_host(pos, nullptr, deriv);
    return 0;
    UB_CATCH
}
int set_param(int n_param, const float* param, DerivEngine* engine, const char* node_name) {
    return ub_set_param(&engine->u, node_name, n_param, param);
}
int get_param(int n_param, float* param, DerivEngine* engine, const char* node_name) {
    return ub_get_param(&engine->u, node_name, n_param, param, nullptr);
}
int get_param_deriv(int n_param, float* deriv, DerivEngine* engine, const char* node_name) {
    UB_TRY
    auto v = node_of(&engine->u, node_name).get_param_deriv(0);
    if ((int)v.size() != n_param) throw std::string("wrong number of parameters (parameter derivatives are not implemented for this node)");
    std::copy(v.begin(), v.end(), deriv);
    return 0;
    UB_CATCH
}
int get_output_dims(int* n_elem, int* elem_width, DerivEngine* engine, const char* node_name) {
    return ub_get_output_dims(&engine->u, node_name, n_elem, elem_width);
}
int get_output(int n_output, float* output, DerivEngine* engine, const char* node_name) {
    return ub_get_output(&engine->u, node_name, 0, n_output, output);
}
int get_sens(int n_output, float* output, DerivEngine* engine, const char* node_name) {
    return ub_get_sens(&engine->u, node_name, 0, n_output, output);
}
int get_value_by_name(int n_output, float* output, DerivEngine* engine, const char* node_name, const char* log_name) {
    UB_TRY
    auto v = node_of(&engine->u, node_name).get_value_by_name(0, log_name);
    if ((int)v.size() != n_output) throw std::string("wrong number of elements");
    std::copy(v.begin(), v.end(), output);
    return 0;
    UB_CATCH
}

// ---- clamped B-spline helpers (host; reference engine_c_library.cpp:197-276, spline.h:268-272,362-380) ----------
static void deboor_host(const float* c, float x, float& val, float& der) {
    int b = (int)x;
    float y = x - b;
    float y2 = y * y, y3 = y2 * y;
    float w0 = (1 - y) * (1 - y) * (1 - y) / 6.f, w1 = (3 * y3 - 6 * y2 + 4) / 6.f, w2 = (-3 * y3 + 3 * y2 + 3 * y + 1) / 6.f, w3 = y3 / 6.f;
    float d0 = -(1 - y) * (1 - y) / 2.f, d1 = (9 * y2 - 12 * y) / 6.f, d2 = (-9 * y2 + 6 * y + 3) / 6.f, d3 = y2 / 2.f;
    val = w0 * c[b - 1] + w1 * c[b] + w2 * c[b + 1] + w3 * c[b + 2];
    der = d0 * c[b - 1] + d1 * c[b] + d2 * c[b + 1] + d3 * c[b + 2];
}
static void clamped_host(const float* c, int n, float x, float& val, float& der) {
    if (x <= 1.f) { val = c[0] / 6.f + 2.f * c[1] / 3.f + c[2] / 6.f; der = 0.f; }
    else if (x >= n - 2) { val = c[n - 3] / 6.f + 2.f * c[n - 2] / 3.f + c[n - 1] / 6.f; der = 0.f; }
    else deboor_host(c, x, val, der);
}
int clamped_spline_solve(int N, float* bspline_coeff, const float* values) {
    UB_TRY
    std::vector<double> v(values, values + N - 2);
    auto c = ub::clamped_bspline_coefficients(v);
    for (int i = 0; i < N; ++i) bspline_coeff[i] = (float)c[i];
    return 0;
    UB_CATCH
}
int clamped_spline_value(int N, float* result, const float* bspline_coeff, int nx, float* x) {
    for (int i = 0; i < nx; ++i) { float d; clamped_host(bspline_coeff, N, x[i], result[i], d); }
    return 0;
}
int get_clamped_value_and_deriv(int N, float* result, const float* bspline_coeff, int nx, float* x) {
    for (int i = 0; i < nx; ++i) clamped_host(bspline_coeff, N, x[i], result[2 * i], result[2 * i + 1]);
    return 0;
}
int get_clamped_coeff_deriv(int N, float* result, const float* bspline_coeff, float x) {
    (void)bspline_coeff;
    for (int i = 0; i < N; ++i) result[i] = 0.f;
    int start;
    float w[4];
    if (x <= 1.f) { start = 0; w[0] = 1.f / 6.f; w[1] = 2.f / 3.f; w[2] = 1.f / 6.f; w[3] = 0.f; }
    else if (x >= N - 2) { start = N - 4; w[0] = 0.f; w[1] = 1.f / 6.f; w[2] = 2.f / 3.f; w[3] = 1.f / 6.f; }
    else {
        int b = (int)x;
        float y = x - b, y2 = y * y, y3 = y2 * y;
        start = b - 1;
        w[0] = (1 - y) * (1 - y) * (1 - y) / 6.f; w[1] = (3 * y3 - 6 * y2 + 4) / 6.f; w[2] = (-3 * y3 + 3 * y2 + 3 * y + 1) / 6.f; w[3] = y3 / 6.f;
    }
    for (int i = 0; i < 4; ++i) result[start + i] = w[i];
    return 0;
}

}  // extern "C"
