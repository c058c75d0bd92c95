// Host side of replica exchange (reference src/main.cpp:120-275): swap-set parsing, the Metropolis decisions and the
// counter-based random stream they draw from.  Plain C++ (no CUDA): every rank of a multi-GPU ladder evaluates the same
// decisions from the all-gathered energies, so no decision ever has to be broadcast.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace ub {

// Threefry-4x32-20 (Random123 threefry.h; rotation constants :110-117, key parity 0x1BD11BDA :172), host restatement
void threefry4x32_20_host(uint32_t out[4], const uint32_t ctr[4], const uint32_t key[4]);

enum RandomStreamType {   // reference src/random.h:12-17
    THERMOSTAT_RANDOM_STREAM = 0,
    REPLICA_EXCHANGE_RANDOM_STREAM = 1,
    PIVOT_MOVE_RANDOM_STREAM = 2,
    JUMP_MOVE_RANDOM_STREAM = 3
};

// reference src/random.h:19-66 (key = (seed, generator_id, 0, 0), counter = (t_lo, t_hi, atom, n_draw))
struct HostRandomGenerator {
    uint32_t k[4], c[4];
    HostRandomGenerator(uint32_t seed, uint32_t generator_id, uint32_t atom_number, uint64_t timestep);
    void random_bits(uint32_t out[4]);
    float uniform_open_closed_x();   // first component of uniform_open_closed()
};

struct SwapPair {
    int sys1, sys2;
    uint64_t n_attempt, n_success;
};

struct ReplicaExchangePlan {
    int n_system = 0;
    std::vector<std::vector<SwapPair>> swap_sets;
    std::vector<int> replica_indices;                       // which original replica sits in each system slot
    std::vector<std::vector<std::pair<int, int>>> participating_swaps;   // per system: (set, index in set)

    // same parsing and the same error strings as ReplicaExchange::ReplicaExchange (main.cpp:130-191); throws std::string
    ReplicaExchangePlan(int n_system, const std::vector<std::string>& swap_set_strings);

    // Metropolis pass over one swap set (main.cpp:262-273).  old/new_lboltz are -beta_i*E_i before / after the trial
    // exchange of every pair of the set.  accept[i] receives 1 if pair i of the set stays exchanged.  Updates the
    // attempt/success counters and replica_indices.  `rng` carries the draw counter across the sets of one attempt.
    void decide(int set, const float* old_lboltz, const float* new_lboltz, HostRandomGenerator& rng, int* accept);

    // Temperature ladder with ONE Hamiltonian: the energy of configuration x_j in slot i is E_j, so the trial energies are
    // a permutation of the gathered ones and no second evaluation is needed.
    // energy: current potential of every system slot; beta: 1/T of every slot.
    void decide_same_hamiltonian(int set, const float* beta, const float* energy, HostRandomGenerator& rng, int* accept);
};

}  // namespace ub
