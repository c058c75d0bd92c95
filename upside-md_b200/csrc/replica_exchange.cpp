// See replica_exchange.h.  Reference: src/main.cpp:120-275 (ReplicaExchange), src/random.h, Random123 threefry.h/uniform.hpp.
#include "replica_exchange.h"

#include <cmath>
#include <set>

namespace ub {

static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

void threefry4x32_20_host(uint32_t out[4], const uint32_t ctr[4], const uint32_t key[4]) {
    static const int R[8][2] = {{10, 26}, {11, 21}, {13, 27}, {23, 5}, {6, 20}, {17, 11}, {25, 10}, {18, 20}};
    uint32_t ks[5], X[4];
    ks[4] = 0x1BD11BDAu;
    for (int i = 0; i < 4; ++i) { ks[i] = key[i]; X[i] = ctr[i]; ks[4] ^= key[i]; }
    for (int i = 0; i < 4; ++i) X[i] += ks[i];
    for (int r = 0; r < 20; ++r) {
        const int ra = R[r & 7][0], rb = R[r & 7][1];
        if ((r & 1) == 0) {
            X[0] += X[1]; X[1] = rotl32(X[1], ra); X[1] ^= X[0];
            X[2] += X[3]; X[3] = rotl32(X[3], rb); X[3] ^= X[2];
        } else {
            X[0] += X[3]; X[3] = rotl32(X[3], ra); X[3] ^= X[0];
            X[2] += X[1]; X[1] = rotl32(X[1], rb); X[1] ^= X[2];
        }
        if ((r & 3) == 3) {
            const int s = r / 4 + 1;
            for (int i = 0; i < 4; ++i) X[i] += ks[(s + i) % 5];
            X[3] += (uint32_t)s;
        }
    }
    for (int i = 0; i < 4; ++i) out[i] = X[i];
}

HostRandomGenerator::HostRandomGenerator(uint32_t seed, uint32_t generator_id, uint32_t atom_number, uint64_t timestep) {
    k[0] = seed; k[1] = generator_id; k[2] = 0u; k[3] = 0u;
    c[0] = (uint32_t)(timestep & 0xffffffffull); c[1] = (uint32_t)(timestep >> 32); c[2] = atom_number; c[3] = 0u;
}
void HostRandomGenerator::random_bits(uint32_t out[4]) {
    threefry4x32_20_host(out, c, k);
    c[3]++;
}
float HostRandomGenerator::uniform_open_closed_x() {
    uint32_t b[4];
    random_bits(b);
    return b[0] * 2.3283064365386963e-10f + 1.1641532182693481e-10f;   // r123::u01<float>, uniform.hpp:145-154
}

static std::vector<std::string> split_string(const std::string& src, const std::string& sep) {
    std::vector<std::string> ret;
    for (size_t pos = 0; pos < src.size();) {
        size_t m = src.find(sep, pos);
        if (m == std::string::npos) m = src.size();
        ret.emplace_back(src.substr(pos, m - pos));
        pos = m + sep.size();
    }
    return ret;
}
static int stoi_strict(const std::string& s) {
    size_t nchar = 0;
    int i;
    try { i = std::stoi(s, &nchar); } catch (...) { throw "invalid integer '" + s + "'"; }
    if (nchar != s.size()) throw "invalid integer '" + s + "'";
    return i;
}

ReplicaExchangePlan::ReplicaExchangePlan(int n_system_, const std::vector<std::string>& swap_set_strings) : n_system(n_system_) {
    for (int ns = 0; ns < n_system; ++ns) { replica_indices.push_back(ns); participating_swaps.emplace_back(); }
    for (const std::string& set_string : swap_set_strings) {
        swap_sets.emplace_back();
        auto& set = swap_sets.back();
        for (const auto& pair_string : split_string(set_string, ",")) {
            auto p = split_string(pair_string, "-");
            if (p.size() != 2u)
                throw "invalid swap pair, because it contains " + std::to_string(p.size()) + "elements but should contain 2";
            SwapPair s{stoi_strict(p[0]), stoi_strict(p[1]), 0u, 0u};
            if (s.sys1 >= n_system || s.sys2 >= n_system || s.sys1 < 0 || s.sys2 < 0) throw std::string("invalid system");
            set.push_back(s);
        }
    }
    for (size_t is = 0; is < swap_sets.size(); ++is) {
        std::set<int> in_set;
        for (size_t ip = 0; ip < swap_sets[is].size(); ++ip) {
            auto& sw = swap_sets[is][ip];
            if (in_set.count(sw.sys1) || in_set.count(sw.sys2) || sw.sys1 == sw.sys2)
                throw std::string("Overlapping indices in swap set.  No replica index can appear more than once in a swap set.  "
                                  "You probably (but maybe not; I didn't look that closely) need more swap sets to get "
                                  "non-overlapping pairs.");
            in_set.insert(sw.sys1);
            in_set.insert(sw.sys2);
            participating_swaps[sw.sys1].emplace_back((int)is, (int)ip);
            participating_swaps[sw.sys2].emplace_back((int)is, (int)ip);
        }
    }
}

void ReplicaExchangePlan::decide(int set, const float* old_lboltz, const float* new_lboltz, HostRandomGenerator& rng, int* accept) {
    auto& ss = swap_sets.at(set);
    for (size_t i = 0; i < ss.size(); ++i) {
        auto& sp = ss[i];
        const int s1 = sp.sys1, s2 = sp.sys2;
        sp.n_attempt++;
        const float diff = (new_lboltz[s1] + new_lboltz[s2]) - (old_lboltz[s1] + old_lboltz[s2]);
        // the random number is drawn only when the exchange is uphill (short-circuit && in main.cpp:268)
        if (diff < 0.f && expf(diff) < rng.uniform_open_closed_x()) {
            accept[i] = 0;
        } else {
            accept[i] = 1;
            sp.n_success++;
            std::swap(replica_indices[s1], replica_indices[s2]);
        }
    }
}

void ReplicaExchangePlan::decide_same_hamiltonian(int set, const float* beta, const float* energy, HostRandomGenerator& rng,
                                                  int* accept) {
    std::vector<float> old_l(n_system), new_l(n_system);
    for (int i = 0; i < n_system; ++i) old_l[i] = new_l[i] = -beta[i] * energy[i];
    for (auto& sp : swap_sets.at(set)) {
        new_l[sp.sys1] = -beta[sp.sys1] * energy[sp.sys2];
        new_l[sp.sys2] = -beta[sp.sys2] * energy[sp.sys1];
    }
    decide(set, old_l.data(), new_l.data(), rng, accept);
}

}  // namespace ub
