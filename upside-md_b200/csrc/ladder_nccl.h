// Device-resident replica exchange of a ladder sharded over GPUs (see ladder_nccl.cu).
#pragma once
#include <string>
#include <vector>

#include "engine.h"
#include "replica_exchange.h"

namespace ub {

struct Ladder {
    Engine* e;
    void* comm;   // ncclComm_t (NULL on a single rank)
    int rank, world, n_global, n_local, first;
    uint32_t seed;
    ReplicaExchangePlan plan;          // parsing + validation of the swap sets (main.cpp:130-191)
    int n_set, n_pair_total;
    std::vector<int> h_set_start;
    DevBuf<int> set_start, pairs, accept, replica_index;
    DevBuf<unsigned long long> counts;
    DevBuf<float> beta, energy_all, energy_work, sendbuf, recvbuf;
    struct SetPlan {
        int n_local_pair = 0, local_offset = 0, n_cross = 0, cross_offset = 0;
        std::vector<int> h_partner;   // partner rank of each boundary pair of this rank
    };
    std::vector<SetPlan> sets;
    DevBuf<int> local_pairs, cross_slot, cross_k;   // all sets, concatenated

    Ladder(Engine* e, void* comm, int rank, int world, int n_global, const std::vector<std::string>& swap_sets, uint32_t seed,
           const float* temperature_all);
    void set_temperature(const float* temperature_all);
    // bytes that cross GPUs per attempt on this rank: (all-gather receive, boundary coordinates sent)
    void comm_bytes(size_t* gather, size_t* coords) const;
    // one attempt of this rank (one process per GPU)
    void attempt(unsigned long long round) { std::vector<Ladder*> one{this}; attempt_all(one, round); }
    // one attempt of several ranks driven by ONE host thread (a single process with one engine per device): the same
    // sequence, with the NCCL calls of all ranks inside group brackets as NCCL requires for that arrangement
    static void attempt_all(const std::vector<Ladder*>& ranks, unsigned long long round);
    // state on the host (waits for the stream): replica indices, decisions of the last attempt, counters, energies
    void download(std::vector<int>* replica_indices, std::vector<int>* accept_last, std::vector<unsigned long long>* counts_out,
                  std::vector<float>* energies);
};

// one NCCL communicator per listed device of this process (ncclCommInitAll); empty on failure to load NCCL -> throws
std::vector<void*> nccl_comm_init_all(const std::vector<int>& devices);
void nccl_comm_destroy(void* comm);

}  // namespace ub
