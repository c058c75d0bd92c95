// Replica exchange of a ladder sharded over GPUs, on the device and on the engine's stream (reference
// ReplicaExchange::attempt_swaps, src/main.cpp:227-275).
//
// The reference evaluates every system twice per swap set, serially, and swaps coordinates on the host.  Here rank g of W
// owns the contiguous rungs [g*n_local, (g+1)*n_local) of an n_global-rung ladder as the replicas of its batched engine,
// and one attempt is, with no host synchronisation anywhere:
//   1. ONE batched energy evaluation (the engine's CUDA graph);
//   2. ncclAllGather of the n_local energies -> all n_global energies on every rank (192 bytes for 48 rungs);
//   3. a decision kernel that every rank runs identically: all swap sets in order, Metropolis rule with the reference's
//      counter-based random stream RandomGenerator(seed, REPLICA_EXCHANGE_RANDOM_STREAM, 0, round), one draw per uphill
//      pair.  All rungs of a ladder share one Hamiltonian (they are replicas of one engine configuration), so the trial
//      energies of a set are a permutation of the current ones and the second evaluation of the reference is not needed;
//   4. per swap set: the coordinates of block-boundary rungs cross NVLink as grouped ncclSend/ncclRecv (3*n_atom floats
//      per boundary pair, exchanged unconditionally so that no decision has to reach the host), then one kernel swaps the
//      accepted intra-GPU pairs and adopts the received coordinates of accepted boundary pairs.
// Only energies and boundary coordinates cross GPUs; decisions never do.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): the engine library has no link-time dependency on it and a
// single-GPU ladder (comm == NULL) needs no NCCL at all.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../include/upside_b200.h"
#include "ladder_nccl.h"
#include "rng.cuh"

namespace ub {

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& nccl() {
    static NcclApi api;
    if (api.lib) return api;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) throw std::string("NCCL is not available (dlopen libnccl.so.2 failed): ") + dlerror();
    auto sym = [&](const char* name) {
        void* p = dlsym(api.lib, name);
        if (!p) throw std::string("NCCL symbol missing: ") + name;
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}
void nccl_check(ncclResult_t r, const char* what) {
    if (r != ncclSuccess) throw std::string("NCCL error in ") + what + ": " + nccl().GetErrorString(r);
}
#define UB_NCCL(x) ::ub::nccl_check((x), #x)

// ---- kernels -----------------------------------------------------------------------------------------------------------
// Every rank runs this with the same inputs and reaches the same decisions (reference main.cpp:249-273, restated for one
// Hamiltonian: new_lboltz[s1] = -beta[s1]*E[s2]).  One thread: the draw counter advances only on uphill pairs, so the
// pairs are inherently sequential - and there are at most n_global/2 of them per set.
__global__ void k_ladder_decide(int n_global, int n_set, const int* __restrict__ set_start, const int* __restrict__ pairs,
                                const float* __restrict__ beta, const float* __restrict__ energy_in, uint32_t seed,
                                unsigned long long round, int* __restrict__ accept, unsigned long long* __restrict__ counts,
                                int* __restrict__ replica_index, float* __restrict__ energy_work) {
    if (threadIdx.x || blockIdx.x) return;
    for (int i = 0; i < n_global; ++i) energy_work[i] = energy_in[i];
    DeviceRandom rng(seed, 1u /* REPLICA_EXCHANGE_RANDOM_STREAM */, 0u, round);
    for (int s = 0; s < n_set; ++s)
        for (int k = set_start[s]; k < set_start[s + 1]; ++k) {
            const int s1 = pairs[2 * k], s2 = pairs[2 * k + 1];
            const float e1 = energy_work[s1], e2 = energy_work[s2];
            const float old1 = -beta[s1] * e1, old2 = -beta[s2] * e2;
            const float new1 = -beta[s1] * e2, new2 = -beta[s2] * e1;
            const float diff = __fadd_rn(new1, new2) - __fadd_rn(old1, old2);
            counts[2 * k] += 1ull;
            bool ok = true;
            if (diff < 0.f) {   // the random number is drawn only when the exchange is uphill (short-circuit in main.cpp:268)
                float u[4];
                rng.uniform_open_closed(u);
                ok = !(expf(diff) < u[0]);
            }
            accept[k] = ok ? 1 : 0;
            if (ok) {
                counts[2 * k + 1] += 1ull;
                energy_work[s1] = e2; energy_work[s2] = e1;   // the configurations, and with them their energies, change slots
                const int t = replica_index[s1]; replica_index[s1] = replica_index[s2]; replica_index[s2] = t;
            }
        }
}

// rows of the local replicas that a boundary pair sends to its partner rank
__global__ void k_ladder_pack(const float* __restrict__ pos, int n_atom, const int* __restrict__ cross_slot, float* __restrict__ sendbuf) {
    const float4* src = reinterpret_cast<const float4*>(pos) + size_t(cross_slot[blockIdx.x]) * n_atom;
    float4* dst = reinterpret_cast<float4*>(sendbuf) + size_t(blockIdx.x) * n_atom;
    for (int i = threadIdx.x; i < n_atom; i += blockDim.x) dst[i] = src[i];
}
// blocks [0,n_local_pair): swap two local replicas if their pair was accepted; blocks after that: adopt the received
// coordinates of an accepted boundary pair
__global__ void k_ladder_apply(float* __restrict__ pos, int n_atom, int n_local_pair, const int* __restrict__ local_pairs /* a,b,k */,
                               const int* __restrict__ cross_slot, const int* __restrict__ cross_k, const float* __restrict__ recvbuf,
                               const int* __restrict__ accept) {
    const int b = blockIdx.x;
    if (b < n_local_pair) {
        if (!accept[local_pairs[3 * b + 2]]) return;
        float4* x = reinterpret_cast<float4*>(pos) + size_t(local_pairs[3 * b]) * n_atom;
        float4* y = reinterpret_cast<float4*>(pos) + size_t(local_pairs[3 * b + 1]) * n_atom;
        for (int i = threadIdx.x; i < n_atom; i += blockDim.x) { const float4 t = x[i]; x[i] = y[i]; y[i] = t; }
    } else {
        const int c = b - n_local_pair;
        if (!accept[cross_k[c]]) return;
        float4* x = reinterpret_cast<float4*>(pos) + size_t(cross_slot[c]) * n_atom;
        const float4* y = reinterpret_cast<const float4*>(recvbuf) + size_t(c) * n_atom;
        for (int i = threadIdx.x; i < n_atom; i += blockDim.x) x[i] = y[i];
    }
}

Ladder::Ladder(Engine* e_, void* comm_, int rank_, int world_, int n_global_, const std::vector<std::string>& swap_sets, uint32_t seed_,
               const float* temperature_all)
    : e(e_), comm(comm_), rank(rank_), world(world_), n_global(n_global_), seed(seed_), plan(n_global_, swap_sets) {
    if (world < 1 || rank < 0 || rank >= world) throw std::string("invalid rank / world size");
    if (n_global % world) throw std::string("the ladder must divide evenly over the ranks (n_global % world != 0)");
    if (world > 1 && !comm) throw std::string("a ladder sharded over several ranks needs an NCCL communicator");
    n_local = n_global / world;
    first = rank * n_local;
    if (e->n_rep != n_local) throw "the engine holds " + std::to_string(e->n_rep) + " replicas but this rank owns " + std::to_string(n_local) + " rungs";
    UB_CUDA(cudaSetDevice(e->device));
    n_set = (int)plan.swap_sets.size();
    std::vector<int> h_pairs, h_cross_slot, h_cross_k, lp;
    h_set_start.push_back(0);
    sets.resize(n_set);
    for (int s = 0; s < n_set; ++s) {
        sets[s].local_offset = (int)lp.size();
        sets[s].cross_offset = (int)h_cross_slot.size();
        for (auto& sp : plan.swap_sets[s]) {
            const int k = (int)h_pairs.size() / 2;
            h_pairs.push_back(sp.sys1);
            h_pairs.push_back(sp.sys2);
            const int r1 = sp.sys1 / n_local, r2 = sp.sys2 / n_local;
            if (r1 == rank && r2 == rank) { lp.push_back(sp.sys1 - first); lp.push_back(sp.sys2 - first); lp.push_back(k); }
            else if (r1 == rank || r2 == rank) {
                h_cross_slot.push_back((r1 == rank ? sp.sys1 : sp.sys2) - first);
                h_cross_k.push_back(k);
                sets[s].h_partner.push_back(r1 == rank ? r2 : r1);
            }
        }
        h_set_start.push_back((int)h_pairs.size() / 2);
        sets[s].n_local_pair = ((int)lp.size() - sets[s].local_offset) / 3;
        sets[s].n_cross = (int)sets[s].h_partner.size();
    }
    if (lp.empty()) lp.push_back(0);
    local_pairs.upload(lp);
    n_pair_total = (int)h_pairs.size() / 2;
    if (h_pairs.empty()) h_pairs.push_back(0);
    set_start.upload(h_set_start);
    pairs.upload(h_pairs);
    accept.alloc(std::max(1, n_pair_total));
    counts.alloc(std::max(1, 2 * n_pair_total));
    std::vector<int> ri(n_global);
    for (int i = 0; i < n_global; ++i) ri[i] = i;
    replica_index.upload(ri);
    std::vector<float> b(n_global);
    for (int i = 0; i < n_global; ++i) b[i] = 1.f / temperature_all[i];
    beta.upload(b);
    energy_all.alloc(n_global);
    energy_work.alloc(n_global);
    const size_t n_cross_all = std::max<size_t>(1, h_cross_slot.size());
    if (h_cross_slot.empty()) { h_cross_slot.push_back(0); h_cross_k.push_back(0); }
    cross_slot.upload(h_cross_slot);
    cross_k.upload(h_cross_k);
    sendbuf.alloc(n_cross_all * e->n_atom * 4);
    recvbuf.alloc(n_cross_all * e->n_atom * 4);
}

void Ladder::set_temperature(const float* temperature_all) {
    UB_CUDA(cudaSetDevice(e->device));
    std::vector<float> b(n_global);
    for (int i = 0; i < n_global; ++i) b[i] = 1.f / temperature_all[i];
    UB_CUDA(cudaMemcpyAsync(beta.p, b.data(), b.size() * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    UB_CUDA(cudaStreamSynchronize(e->stream));
}

void Ladder::comm_bytes(size_t* gather, size_t* coords) const {
    *gather = world > 1 ? size_t(n_global - n_local) * sizeof(float) : 0;
    size_t c = 0;
    for (auto& s : sets) c += size_t(s.n_cross) * e->n_atom * 4 * sizeof(float);
    *coords = c;
}

void Ladder::attempt_all(const std::vector<Ladder*>& ranks, unsigned long long round) {
    if (ranks.empty()) return;
    const bool many = ranks.size() > 1;   // one host thread drives several devices: NCCL calls go inside group brackets
    const bool use_nccl = ranks[0]->world > 1;
    for (Ladder* l : ranks) { UB_CUDA(cudaSetDevice(l->e->device)); l->e->compute(PotentialAndDerivMode); }
    if (use_nccl) {
        if (many) UB_NCCL(nccl().GroupStart());
        for (Ladder* l : ranks)
            UB_NCCL(nccl().AllGather(l->e->potential.p, l->energy_all.p, l->n_local, ncclFloat, (ncclComm_t)l->comm, l->e->stream));
        if (many) UB_NCCL(nccl().GroupEnd());
    } else {
        Ladder* l = ranks[0];
        UB_CUDA(cudaMemcpyAsync(l->energy_all.p, l->e->potential.p, l->n_local * sizeof(float), cudaMemcpyDeviceToDevice, l->e->stream));
    }
    for (Ladder* l : ranks) {
        UB_CUDA(cudaSetDevice(l->e->device));
        k_ladder_decide<<<1, 32, 0, l->e->stream>>>(l->n_global, l->n_set, l->set_start.p, l->pairs.p, l->beta.p, l->energy_all.p, l->seed, round,
                                                    l->accept.p, l->counts.p, l->replica_index.p, l->energy_work.p);
    }
    const int n_set = ranks[0]->n_set;
    for (int s = 0; s < n_set; ++s) {
        bool any_cross = false;
        for (Ladder* l : ranks) {
            SetPlan& sp = l->sets[s];
            if (!sp.n_cross) continue;
            any_cross = true;
            UB_CUDA(cudaSetDevice(l->e->device));
            const size_t row = size_t(l->e->n_atom) * 4;
            k_ladder_pack<<<sp.n_cross, 128, 0, l->e->stream>>>(l->e->pos->output, l->e->n_atom, l->cross_slot.p + sp.cross_offset,
                                                                l->sendbuf.p + sp.cross_offset * row);
        }
        if (any_cross) {
            UB_NCCL(nccl().GroupStart());
            for (Ladder* l : ranks) {
                SetPlan& sp = l->sets[s];
                const size_t row = size_t(l->e->n_atom) * 4;
                for (int c = 0; c < sp.n_cross; ++c) {
                    UB_NCCL(nccl().Send(l->sendbuf.p + (sp.cross_offset + c) * row, row, ncclFloat, sp.h_partner[c], (ncclComm_t)l->comm, l->e->stream));
                    UB_NCCL(nccl().Recv(l->recvbuf.p + (sp.cross_offset + c) * row, row, ncclFloat, sp.h_partner[c], (ncclComm_t)l->comm, l->e->stream));
                }
            }
            UB_NCCL(nccl().GroupEnd());
        }
        for (Ladder* l : ranks) {
            SetPlan& sp = l->sets[s];
            if (!(sp.n_local_pair + sp.n_cross)) continue;
            UB_CUDA(cudaSetDevice(l->e->device));
            const size_t row = size_t(l->e->n_atom) * 4;
            k_ladder_apply<<<sp.n_local_pair + sp.n_cross, 128, 0, l->e->stream>>>(l->e->pos->output, l->e->n_atom, sp.n_local_pair,
                                                                                   l->local_pairs.p + sp.local_offset, l->cross_slot.p + sp.cross_offset,
                                                                                   l->cross_k.p + sp.cross_offset, l->recvbuf.p + sp.cross_offset * row,
                                                                                   l->accept.p);
            UB_CUDA(cudaGetLastError());
        }
    }
}

void Ladder::download(std::vector<int>* replica_indices, std::vector<int>* accept_last, std::vector<unsigned long long>* counts_out,
                      std::vector<float>* energies) {
    UB_CUDA(cudaSetDevice(e->device));
    e->sync_and_check();
    if (replica_indices) *replica_indices = replica_index.download();
    if (accept_last) { *accept_last = accept.download(); accept_last->resize(n_pair_total); }
    if (counts_out) { *counts_out = counts.download(); counts_out->resize(2 * size_t(n_pair_total)); }
    if (energies) *energies = energy_all.download();
}

std::vector<void*> nccl_comm_init_all(const std::vector<int>& devices) {
    std::vector<ncclComm_t> c(devices.size());
    UB_NCCL(nccl().CommInitAll(c.data(), (int)devices.size(), devices.data()));
    return std::vector<void*>(c.begin(), c.end());
}
void nccl_comm_destroy(void* comm) {
    try { if (comm) nccl().CommDestroy((ncclComm_t)comm); } catch (...) {}
}

}  // namespace ub

// ---- C ABI -------------------------------------------------------------------------------------------------------------
struct UbEngine;
namespace ub { Engine* engine_of(UbEngine* e); }
struct UbLadder { std::unique_ptr<ub::Ladder> l; };

static thread_local std::string g_ladder_error;
static int ladder_fail(const std::string& e) {
    g_ladder_error = e;
    fprintf(stderr, "\n\nERROR: %s\n", e.c_str());
    return 1;
}
#define LAD_TRY try {
#define LAD_CATCH                                                          \
    }                                                                      \
    catch (const std::string& e) { return ladder_fail(e); }                \
    catch (const char* e) { return ladder_fail(e); }                       \
    catch (const std::exception& e) { return ladder_fail(e.what()); }      \
    catch (...) { return ladder_fail("unknown error"); }

extern "C" {

const char* ub_ladder_last_error(void) { return g_ladder_error.c_str(); }

int ub_nccl_unique_id(char* out, int len) {
    LAD_TRY
    if (len < (int)sizeof(ncclUniqueId)) throw std::string("buffer too small for an NCCL unique id (128 bytes)");
    ncclUniqueId id;
    UB_NCCL(ub::nccl().GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return 0;
    LAD_CATCH
}

void* ub_nccl_comm_create(const char* unique_id, int world, int rank, int device) {
    try {
        UB_CUDA(cudaSetDevice(device));
        ncclUniqueId id;
        memcpy(&id, unique_id, sizeof(id));
        ncclComm_t c = nullptr;
        UB_NCCL(ub::nccl().CommInitRank(&c, world, id, rank));
        return c;
    } catch (const std::string& e) { ladder_fail(e); return nullptr; }
    catch (...) { ladder_fail("unknown error"); return nullptr; }
}

/* one communicator per device of THIS process (ncclCommInitAll); comms[n_device] */
int ub_nccl_comm_create_all(int n_device, const int* devices, void** comms) {
    LAD_TRY
    auto c = ub::nccl_comm_init_all(std::vector<int>(devices, devices + n_device));
    for (int i = 0; i < n_device; ++i) comms[i] = c[i];
    return 0;
    LAD_CATCH
}

void ub_nccl_comm_destroy(void* comm) { ub::nccl_comm_destroy(comm); }

UbLadder* ub_ladder_create(UbEngine* e, void* nccl_comm, int rank, int world, int n_global, int n_set, const char* const* swap_sets,
                           uint32_t seed, const float* temperature_all) {
    try {
        std::vector<std::string> ss(swap_sets, swap_sets + n_set);
        auto* h = new UbLadder;
        h->l.reset(new ub::Ladder(ub::engine_of(e), nccl_comm, rank, world, n_global, ss, seed, temperature_all));
        return h;
    } catch (const std::string& err) { ladder_fail(err); return nullptr; }
    catch (const std::exception& err) { ladder_fail(err.what()); return nullptr; }
    catch (...) { ladder_fail("unknown error"); return nullptr; }
}
void ub_ladder_destroy(UbLadder* h) { delete h; }

int ub_ladder_attempt(UbLadder* h, uint64_t round) {
    LAD_TRY
    h->l->attempt(round);
    return 0;
    LAD_CATCH
}
int ub_ladder_set_temperature(UbLadder* h, const float* temperature_all) {
    LAD_TRY
    h->l->set_temperature(temperature_all);
    return 0;
    LAD_CATCH
}
int ub_ladder_n_pairs(const UbLadder* h) { return h->l->n_pair_total; }

/* synchronises the engine's stream.  All arrays may be NULL.  replica_indices: n_global (which original replica sits on
 * each rung); accept: decisions of the last attempt, pairs of all sets concatenated; n_attempt/n_success per pair;
 * energies: the n_global energies the last attempt decided on */
int ub_ladder_state(UbLadder* h, int* replica_indices, int* accept, uint64_t* n_attempt, uint64_t* n_success, float* energies) {
    LAD_TRY
    auto& l = *h->l;
    std::vector<int> ri, acc;
    std::vector<unsigned long long> cnt;
    std::vector<float> en;
    l.download(replica_indices ? &ri : nullptr, accept ? &acc : nullptr, (n_attempt || n_success) ? &cnt : nullptr, energies ? &en : nullptr);
    if (replica_indices) memcpy(replica_indices, ri.data(), l.n_global * sizeof(int));
    if (accept && l.n_pair_total) memcpy(accept, acc.data(), l.n_pair_total * sizeof(int));
    for (int k = 0; k < l.n_pair_total && (n_attempt || n_success); ++k) {
        if (n_attempt) n_attempt[k] = cnt[2 * k];
        if (n_success) n_success[k] = cnt[2 * k + 1];
    }
    if (energies) memcpy(energies, en.data(), l.n_global * sizeof(float));
    return 0;
    LAD_CATCH
}
int ub_ladder_comm_bytes(const UbLadder* h, uint64_t* allgather_bytes, uint64_t* coordinate_bytes) {
    size_t a, c;
    h->l->comm_bytes(&a, &c);
    *allgather_bytes = a; *coordinate_bytes = c;
    return 0;
}

}  // extern "C"
