// TEST INFRASTRUCTURE (oracle/).  Extra extern "C" entry points on top of the reference's own engine_c_library:
//  * ref_pairlist      - the pair list a node built during the last evaluation (reference emission order)
//  * ref_node_potential- per-PotentialNode energy after the last PotentialAndDerivMode evaluation
//  * ref_md_run        - the reference's MD loop (main.cpp:515-523 thermalisation, :657-663 thermostat +
//                        integration_cycle, :618 one OpenMP thread per system) without HDF5 output, used as the
//                        CPU baseline and for trajectory parity.  Calls only the reference's own functions.
//  * ref_rng_*         - known-answer access to the reference's RandomGenerator (random.h)
//  * ref_mc_steps      - the reference's Monte-Carlo samplers (monte_carlo_sampler.cpp) executed on n_sys copies of one
//                        configuration exactly as main.cpp:545,630-631 does, for Monte-Carlo parity
#include <chrono>
#include <cstring>
#include <memory>
#include <omp.h>
#include <string>
#include <vector>

#include "deriv_engine.h"
#include "engine_c_library.h"
#include "monte_carlo_sampler.h"
#include "random.h"
#include "state_logger.h"
#include "thermostat.h"

extern "C" {
int ref_pairlist_rotamer(DerivComputation*, int*, int*, int);
int ref_pairlist_hbond(DerivComputation*, int*, int*, int);
int ref_pairlist_environment(DerivComputation*, int*, int*, int);
int ref_pairlist_backbone(DerivComputation*, int*, int*, int);

int ref_pairlist(DerivEngine* engine, const char* node, int* i1, int* i2, int max_edge) try {
    DerivComputation* c = engine->get(node).computation.get();
    int n;
    if((n = ref_pairlist_rotamer(c, i1, i2, max_edge)) != -2) return n;
    if((n = ref_pairlist_hbond(c, i1, i2, max_edge)) != -2) return n;
    if((n = ref_pairlist_environment(c, i1, i2, max_edge)) != -2) return n;
    if((n = ref_pairlist_backbone(c, i1, i2, max_edge)) != -2) return n;
    return -1;
} catch(...) { return -1; }

DerivComputation* ref_get_computation(DerivEngine* engine, const char* node) try {
    return engine->get(node).computation.get();
} catch(...) { return nullptr; }

int ref_n_nodes(DerivEngine* engine) { return (int)engine->nodes.size(); }
int ref_node_name(DerivEngine* engine, int i, char* buf, int n) {
    if(i<0 || i>=(int)engine->nodes.size()) return -1;
    strncpy(buf, engine->nodes[i].name.c_str(), n); buf[n-1]=0;
    return engine->nodes[i].computation->potential_term ? 1 : 0;
}
int ref_node_potential(DerivEngine* engine, const char* node, float* out) try {
    auto* c = engine->get(node).computation.get();
    if(!c->potential_term) return 1;
    *out = static_cast<PotentialNode*>(c)->potential;
    return 0;
} catch(...) { return 1; }

void ref_rng_bits(uint32_t seed, uint32_t stream, uint32_t atom, uint64_t t, uint32_t* out4) {
    threefry4x32_key_t k = {{seed, stream, 0u, 0u}};
    threefry4x32_ctr_t c = {{uint32_t(t & 0xffffffffu), uint32_t(t>>32), atom, 0u}};
    auto r = threefry4x32(c,k);
    for(int i=0;i<4;++i) out4[i] = r.v[i];
}
void ref_rng_raw(const uint32_t* key4, const uint32_t* ctr4, uint32_t* out4) {
    threefry4x32_key_t k = {{key4[0],key4[1],key4[2],key4[3]}};
    threefry4x32_ctr_t c = {{ctr4[0],ctr4[1],ctr4[2],ctr4[3]}};
    auto r = threefry4x32(c,k);
    for(int i=0;i<4;++i) out4[i] = r.v[i];
}
void ref_rng_normal3(uint32_t seed, uint32_t stream, uint32_t atom, uint64_t t, float* out3) {
    RandomGenerator g(seed, stream, atom, t);
    auto v = g.normal3();
    out3[0]=v.x(); out3[1]=v.y(); out3[2]=v.z();
}
void ref_rng_uniform(uint32_t seed, uint32_t stream, uint32_t atom, uint64_t t, int n_draw, float* out4n) {
    RandomGenerator g(seed, stream, atom, t);
    for(int i=0;i<n_draw;++i) { auto v = g.uniform_open_closed(); for(int d=0;d<4;++d) out4n[4*i+d]=v[d]; }
}

// Runs n_sys independent replicas of one configuration for n_round rounds.  pos: [n_sys][n_atom][3] in/out,
// mom_out (optional): [n_sys][n_atom][3].  temperature: [n_sys].  Returns wall seconds of the timed loop in *seconds.
// Engine construction is serial (the registry/HDF5 layer is not thread-safe, main.cpp:456).
int ref_md_run(const char* config_path, int n_sys, int n_atom, float* pos, float* mom_out, const float* temperature,
               uint32_t base_seed, float dt, float thermostat_timescale, long n_round, int thermostat_interval,
               int n_thread, double* seconds, float* potential_out) try {
    struct Sys { DerivEngine* e; VecArrayStorage mom; OrnsteinUhlenbeckThermostat th; Sys(int n): e(nullptr), mom(3,round_up(n,4)) {} };
    std::vector<std::unique_ptr<Sys>> sys;
    for(int s=0; s<n_sys; ++s) {
        sys.emplace_back(new Sys(n_atom));
        auto& S = *sys.back();
        S.e = construct_deriv_engine(n_atom, config_path, true);
        if(!S.e) return 1;
        for(int na=0; na<n_atom; ++na) for(int d=0; d<3; ++d) S.e->pos->output(d,na) = pos[(size_t(s)*n_atom+na)*3+d];
        fill(S.mom, 0.f);
        S.th = OrnsteinUhlenbeckThermostat(base_seed + s, thermostat_timescale, 1., 1e8);
        S.th.set_temp(temperature[s]);
        S.th.apply(S.mom, n_atom);                       // initial thermalisation (main.cpp:521)
        S.th.set_delta_t(thermostat_interval*3*dt);      // main.cpp:522
    }
    if(n_thread>0) omp_set_num_threads(n_thread);
    auto t0 = std::chrono::high_resolution_clock::now();
    #pragma omp parallel for schedule(static,1)
    for(int s=0; s<n_sys; ++s) {
        auto& S = *sys[s];
        for(long nr=0; nr<n_round; ++nr) {
            if(!(nr%thermostat_interval)) S.th.apply(S.mom, n_atom);
            S.e->integration_cycle(S.mom, dt, 0.f, DerivEngine::Verlet);
        }
    }
    *seconds = std::chrono::duration<double>(std::chrono::high_resolution_clock::now()-t0).count();
    for(int s=0; s<n_sys; ++s) {
        auto& S = *sys[s];
        for(int na=0; na<n_atom; ++na) for(int d=0; d<3; ++d) {
            pos[(size_t(s)*n_atom+na)*3+d] = S.e->pos->output(d,na);
            if(mom_out) mom_out[(size_t(s)*n_atom+na)*3+d] = S.mom(d,na);
        }
        if(potential_out) { S.e->compute(PotentialAndDerivMode); potential_out[s] = S.e->potential; }
        free_deriv_engine(S.e);
    }
    return 0;
} catch(const std::string& e) {
    fprintf(stderr, "ref_md_run: %s\n", e.c_str());
    return 1;
} catch(...) { return 1; }

// Timing variant of ref_md_run for bench.py's CPU arms: engines are constructed once, `warm_round` untimed rounds run first
// (thread team spawned, pair lists built, caches warm), then `n_rep` repetitions of `n_round` rounds are timed one by one
// into seconds[n_rep] (the caller quotes the median).  pos: [n_sys][n_atom][3] in/out.
int ref_md_bench(const char* config_path, int n_sys, int n_atom, float* pos, const float* temperature, uint32_t base_seed,
                 float dt, float thermostat_timescale, long warm_round, long n_round, int n_rep, int n_thread,
                 double* seconds) try {
    struct Sys { DerivEngine* e; VecArrayStorage mom; OrnsteinUhlenbeckThermostat th; Sys(int n): e(nullptr), mom(3,round_up(n,4)) {} };
    std::vector<std::unique_ptr<Sys>> sys;
    for(int s=0; s<n_sys; ++s) {
        sys.emplace_back(new Sys(n_atom));
        auto& S = *sys.back();
        S.e = construct_deriv_engine(n_atom, config_path, true);
        if(!S.e) return 1;
        for(int na=0; na<n_atom; ++na) for(int d=0; d<3; ++d) S.e->pos->output(d,na) = pos[(size_t(s)*n_atom+na)*3+d];
        fill(S.mom, 0.f);
        S.th = OrnsteinUhlenbeckThermostat(base_seed + s, thermostat_timescale, 1., 1e8);
        S.th.set_temp(temperature[s]);
        S.th.apply(S.mom, n_atom);
        S.th.set_delta_t(3*dt);
    }
    if(n_thread>0) omp_set_num_threads(n_thread);
    auto run = [&](long rounds) {
        #pragma omp parallel for schedule(static,1)
        for(int s=0; s<n_sys; ++s) {
            auto& S = *sys[s];
            for(long nr=0; nr<rounds; ++nr) {
                S.th.apply(S.mom, n_atom);
                S.e->integration_cycle(S.mom, dt, 0.f, DerivEngine::Verlet);
            }
        }
    };
    run(warm_round);
    for(int rep=0; rep<n_rep; ++rep) {
        auto t0 = std::chrono::high_resolution_clock::now();
        run(n_round);
        seconds[rep] = std::chrono::duration<double>(std::chrono::high_resolution_clock::now()-t0).count();
    }
    for(int s=0; s<n_sys; ++s) {
        auto& S = *sys[s];
        for(int na=0; na<n_atom; ++na) for(int d=0; d<3; ++d) pos[(size_t(s)*n_atom+na)*3+d] = S.e->pos->output(d,na);
        free_deriv_engine(S.e);
    }
    return 0;
} catch(const std::string& e) {
    fprintf(stderr, "ref_md_bench: %s\n", e.c_str());
    return 1;
} catch(...) { return 1; }

// Latency of single evaluations through the reference's own C ABI (evaluate_deriv, engine_c_library.cpp:47-64): n_eval
// calls on the same coordinates after n_warm untimed ones; returns microseconds per call in *us.
int ref_eval_latency(const char* config_path, int n_atom, const float* pos, int n_warm, int n_eval, double* us) try {
    DerivEngine* e = construct_deriv_engine(n_atom, config_path, true);
    if(!e) return 1;
    std::vector<float> deriv(size_t(n_atom)*3);
    for(int i=0; i<n_warm; ++i) evaluate_deriv(deriv.data(), e, pos);
    auto t0 = std::chrono::high_resolution_clock::now();
    for(int i=0; i<n_eval; ++i) evaluate_deriv(deriv.data(), e, pos);
    *us = std::chrono::duration<double>(std::chrono::high_resolution_clock::now()-t0).count() * 1e6 / n_eval;
    free_deriv_engine(e);
    return 0;
} catch(...) { return 1; }

// n_step Monte-Carlo executes (rounds first_round, first_round+1, ...) on n_sys copies of the configuration `rw_copy_path`
// (a scratch copy without /output: the samplers register loggers, which need a writable file; use n_sys = 1 per copy).  pos: [n_sys][n_atom][3] in/out;
// stats_out: [n_sys][n_sampler][2] (n_success, n_attempt); returns the number of samplers, negative on error.
int ref_mc_steps(const char* rw_copy_path, int n_sys, int n_atom, float* pos, const float* temperature, uint32_t base_seed,
                 uint64_t first_round, int n_step, long* stats_out, int max_sampler) try {
    int n_sampler = 0;
    for(int s=0; s<n_sys; ++s) {
        auto config = h5::h5_obj(H5Fclose, H5Fopen(rw_copy_path, H5F_ACC_RDWR, H5P_DEFAULT));
        auto logger = std::make_shared<H5Logger>(config, "output", LOG_BASIC);
        DerivEngine* e = construct_deriv_engine(n_atom, rw_copy_path, true);
        if(!e) return -1;
        for(int na=0; na<n_atom; ++na) for(int d=0; d<3; ++d) e->pos->output(d,na) = pos[(size_t(s)*n_atom+na)*3+d];
        {
            MultipleMonteCarloSampler mc{h5::open_group(config.get(), "/input").get(), *logger};
            n_sampler = (int)mc.samplers.size();
            for(int k=0; k<n_step; ++k) mc.execute(base_seed + s, first_round + k, temperature[s], *e);
            for(int m=0; m<n_sampler && m<max_sampler; ++m) {
                stats_out[(size_t(s)*max_sampler+m)*2+0] = (long)mc.samplers[m]->move_stats.n_success;
                stats_out[(size_t(s)*max_sampler+m)*2+1] = (long)mc.samplers[m]->move_stats.n_attempt;
            }
            for(int na=0; na<n_atom; ++na) for(int d=0; d<3; ++d) pos[(size_t(s)*n_atom+na)*3+d] = e->pos->output(d,na);
            logger->collect_samples();   // (an empty logger cannot be flushed)
            logger.reset();              // the loggers hold references into mc: flush before it goes away
        }
        free_deriv_engine(e);
    }
    return n_sampler;
} catch(const std::string& e) {
    fprintf(stderr, "ref_mc_steps: %s\n", e.c_str());
    return -1;
} catch(...) { return -1; }
}
