// upside_main: the `upside` command line (reference src/main.cpp:317-752) on the batched B200 engine.
//
// Same flags, same units, same round structure (frame logging -> thermostat -> integration_cycle -> replica exchange),
// same /output datasets for the core loggers.  What differs is the execution model: the reference gives every
// configuration file its own DerivEngine on its own OpenMP thread; here configuration files whose /input/potential and
// atom count are identical (a temperature ladder, independent copies) become the replicas of ONE batched engine on the
// GPU and advance together, one CUDA graph launch per MD round for the whole group.  Replica exchange needs two batched
// energy evaluations per swap set instead of 2*n_system serial ones (the reference's serial bottleneck, main.cpp:227-275).
//
// Monte-Carlo pivot/jump moves (--monte-carlo-interval) run batched on the device (monte_carlo.cu).  --log-level defaults
// to detailed as in the reference (main.cpp:475): the node loggers hbond, rama, rama_map_potential, nonlinear_coupling,
// nonbonded_spring_energy and the rotamer series are written unless --log-level basic is given.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <csignal>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../../include/engine_c_library.h"
#include "engine.h"
#include "ladder_nccl.h"
#include "replica_exchange.h"

namespace {

using std::string;
using std::vector;

volatile sig_atomic_t received_signal = 0;
void abort_like_handler(int sig) { received_signal = sig; }   // main.cpp:37-61: only set a flag

struct SignalHandlerHandler {   // RAII so that a Python caller gets its handlers back (main.cpp:63-91)
    int signum;
    void (*old_handler)(int);
    SignalHandlerHandler(int s, void (*h)(int)) : signum(s), old_handler(signal(s, h)) {
        if (old_handler == SIG_ERR) fprintf(stderr, "Warning: problem installing signal handler. Does not affect correctness of simulation.\n");
    }
    ~SignalHandlerHandler() {
        if (old_handler != SIG_ERR && signal(signum, old_handler) == SIG_ERR)
            fprintf(stderr, "Warning: problem restoring signal handler. Does not affect correctness of simulation.\n");
    }
};

double stod_strict(const string& s) {
    size_t n = 0;
    double x;
    try { x = std::stod(s, &n); } catch (...) { throw "invalid float '" + s + "'"; }
    if (n != s.size()) throw "invalid float '" + s + "'";
    return x;
}
vector<string> split_string(const string& src, const string& sep) {
    vector<string> ret;
    for (size_t pos = 0; pos < src.size();) {
        size_t m = src.find(sep, pos);
        if (m == string::npos) m = src.size();
        ret.emplace_back(src.substr(pos, m - pos));
        pos = m + sep.size();
    }
    return ret;
}

// ---- command line (the TCLAP declarations of main.cpp:324-375) ---------------------------------------------------------
struct Args {
    double time_step = 0.009, duration = -1., anneal_factor = 1., anneal_duration = -1., frame_interval = -1.;
    double replica_interval = 0., mc_interval = 0., thermostat_interval = -1., thermostat_timescale = 5.;
    unsigned long seed = 42;
    string temperature, log_level, set_param;
    vector<string> swap_sets, configs;
    bool disable_recentering = false, disable_z_recentering = false, re_raise_signal = false, deriv_agreement = false;
    bool have_duration = false, have_frame_interval = false;
};
struct ArgError { string msg, arg; };

Args parse_args(int argc, const char* const* argv) {
    Args a;
    std::map<string, double*> dbl = {{"time-step", &a.time_step}, {"duration", &a.duration}, {"anneal-factor", &a.anneal_factor},
                                     {"anneal-duration", &a.anneal_duration}, {"frame-interval", &a.frame_interval},
                                     {"replica-interval", &a.replica_interval}, {"monte-carlo-interval", &a.mc_interval},
                                     {"thermostat-interval", &a.thermostat_interval}, {"thermostat-timescale", &a.thermostat_timescale}};
    std::map<string, string*> str = {{"temperature", &a.temperature}, {"log-level", &a.log_level}, {"set-param", &a.set_param}};
    std::map<string, bool*> sw = {{"disable-recentering", &a.disable_recentering}, {"disable-z-recentering", &a.disable_z_recentering},
                                  {"re-raise-signal", &a.re_raise_signal}, {"potential-deriv-agreement", &a.deriv_agreement}};
    bool only_positional = false;
    for (int i = 1; i < argc; ++i) {
        string t = argv[i];
        if (only_positional || t.size() < 2 || t.substr(0, 2) != "--") {
            if (!only_positional && t.size() > 1 && t[0] == '-' && !isdigit((unsigned char)t[1])) throw ArgError{"Couldn't find match", t};
            a.configs.push_back(t);
            continue;
        }
        if (t == "--") { only_positional = true; continue; }
        string name = t.substr(2), val;
        bool has_val = false;
        size_t eq = name.find('=');
        if (eq != string::npos) { val = name.substr(eq + 1); name = name.substr(0, eq); has_val = true; }
        if (sw.count(name)) { *sw[name] = true; continue; }
        auto need = [&]() {
            if (has_val) return val;
            if (i + 1 >= argc) throw ArgError{"Missing a value for this argument!", "--" + name};
            return string(argv[++i]);
        };
        if (dbl.count(name)) {
            string v = need();
            size_t n = 0;
            double x = 0;
            try { x = std::stod(v, &n); } catch (...) { n = 0; }
            if (!n || n != v.size()) throw ArgError{"Couldn't read argument value from string '" + v + "'", "--" + name};
            *dbl[name] = x;
            if (name == "duration") a.have_duration = true;
            if (name == "frame-interval") a.have_frame_interval = true;
        } else if (str.count(name)) *str[name] = need();
        else if (name == "seed") {
            string v = need();
            try { a.seed = std::stoul(v); } catch (...) { throw ArgError{"Couldn't read argument value from string '" + v + "'", "--seed"}; }
        } else if (name == "swap-set") a.swap_sets.push_back(need());
        else throw ArgError{"Couldn't find match", "--" + name};
    }
    if (!a.have_duration) throw ArgError{"Required argument missing", "--duration"};
    if (!a.have_frame_interval) throw ArgError{"Required argument missing", "--frame-interval"};
    if (a.configs.empty()) throw ArgError{"Required argument missing", "config_files"};
    return a;
}

// ---- systems ---------------------------------------------------------------------------------------------------------------
bool array_equal(const h5l::Array& a, const h5l::Array& b) {
    return a.dt.kind == b.dt.kind && a.dt.size == b.dt.size && a.scalar == b.scalar && a.dims == b.dims && a.raw == b.raw;
}
bool tree_equal(const h5l::Node& a, const h5l::Node& b) {
    if (a.is_group != b.is_group || a.attrs.size() != b.attrs.size() || a.children.size() != b.children.size()) return false;
    for (auto ia = a.attrs.begin(), ib = b.attrs.begin(); ia != a.attrs.end(); ++ia, ++ib)
        if (ia->first != ib->first || !array_equal(ia->second, ib->second)) return false;
    if (!a.is_group) return array_equal(a.data, b.data);
    for (auto ia = a.children.begin(), ib = b.children.begin(); ia != a.children.end(); ++ia, ++ib)
        if (ia->first != ib->first || !tree_equal(*ia->second, *ib->second)) return false;
    return true;
}

struct Group {   // configuration files that share one batched engine
    vector<ub::NodeLogger> loggers;   // node loggers of the requested --log-level
    std::unique_ptr<ub::Engine> engine;
    vector<int> systems;   // slot r of the engine = system systems[r]
    const h5l::Node* potential = nullptr;
    int n_atom = 0;
    int device = 0;    // GPU of this engine
    int origin = -1;   // index of the potential the group was split from (UPSIDE_B200_DEVICES)
};
struct System {
    string path;
    std::unique_ptr<h5l::Node> root;
    int n_atom = 0, group = -1, slot = -1;
    uint32_t random_seed = 0;
    float initial_temperature = 1.f, temperature = 1.f;
    // frame buffers (written to /output at the end or on a signal)
    vector<float> pos;
    vector<double> kinetic, potential, time, temperature_log;
    vector<int> replica_index, cumulative_swaps;
    std::map<string, vector<int>> mc_stats;   // "<sampler>_stats": (n_success, n_attempt) since the previous frame
    // node loggers of --log-level detailed (state_logger.h:56-66): name -> (per-frame shape, int64?, samples)
    struct Series { vector<uint64_t> dims; bool integer = false; vector<float> data; };
    std::map<string, Series> node_series;
    size_t n_frame = 0;
};

void write_output(System& sys, const string& invocation, bool have_replex, const ub::ReplicaExchangePlan* plan, int ns) {
    h5l::Node* out = h5l::ensure_group(sys.root.get(), "output");
    out->attrs["invocation"] = h5l::make_string_scalar(invocation);
    auto put = [&](const char* name, h5l::Array a) {
        auto n = std::make_unique<h5l::Node>();
        n->is_group = false;
        n->data = std::move(a);
        out->children[name] = std::move(n);
    };
    const uint64_t nf = sys.n_frame;
    put("pos", h5l::make_array(sys.pos, {nf, 1, (uint64_t)sys.n_atom, 3}));
    put("kinetic", h5l::make_array(sys.kinetic, {nf, 1}));
    put("potential", h5l::make_array(sys.potential, {nf, 1}));
    put("time", h5l::make_array(sys.time, {nf}));
    put("temperature", h5l::make_array(sys.temperature_log, {nf, 1}));
    for (auto& kv : sys.mc_stats) put(kv.first.c_str(), h5l::make_array(kv.second, {nf, 2}));   // monte_carlo_sampler.h:31-36
    for (auto& kv : sys.node_series) {
        vector<uint64_t> dims{nf};
        dims.insert(dims.end(), kv.second.dims.begin(), kv.second.dims.end());
        if (kv.second.integer) put(kv.first.c_str(), h5l::make_array(vector<long>(kv.second.data.begin(), kv.second.data.end()), dims));
        else put(kv.first.c_str(), h5l::make_array(kv.second.data, dims));
    }
    if (have_replex) {
        put("replica_index", h5l::make_array(sys.replica_index, {nf, 1}));
        const auto& ps = plan->participating_swaps[ns];
        vector<int> partner;
        for (auto& p : ps) { const auto& sw = plan->swap_sets[p.first][p.second]; partner.push_back(sw.sys1 != ns ? sw.sys1 : sw.sys2); }
        put("replica_swap_partner", h5l::make_array(partner, {(uint64_t)partner.size()}));
        put("replica_cumulative_swaps", h5l::make_array(sys.cumulative_swaps, {nf, (uint64_t)ps.size(), 2}));
    }
    h5l::save(*sys.root, sys.path);
}

int run(int argc, const char* const* argv, int verbose) {
    Args args = parse_args(argc, argv);
    if (verbose) printf("invocation: ");
    string invocation(argv[0]);
    for (int i = 1; i < argc; ++i) invocation += string(" ") + argv[i];
    if (verbose) printf("%s\n", invocation.c_str());

    std::map<string, vector<float>> set_param_map;
    if (!args.set_param.empty()) {
        auto pf = h5l::load(args.set_param);
        for (auto& kv : pf->children) if (!kv.second->is_group) set_param_map[kv.first] = h5l::as<float>(kv.second->data);
    }

    const float dt = (float)args.time_step;
    const double duration = args.duration;
    const uint64_t n_round = (uint64_t)std::llround(duration / (3 * dt));
    const int thermostat_interval = (int)std::max(1., std::round(args.thermostat_interval / (3 * dt)));
    const int frame_interval = (int)std::max(1., std::round(args.frame_interval / (3 * dt)));
    const unsigned long big_prime = 4294967291ul;   // largest prime smaller than 2^32
    const uint32_t base_random_seed = uint32_t(args.seed % big_prime);
    if (verbose) printf("random seed: %lu\n", (unsigned long)base_random_seed);
    const int duration_print_width = (int)std::ceil(std::log(1 + duration) / std::log(10));
    const bool do_recenter = !args.disable_recentering;
    const bool xy_recenter_only = do_recenter && args.disable_z_recentering;

    const int n_system = (int)args.configs.size();
    vector<System> systems(n_system);
    auto temperature_strings = split_string(args.temperature, ",");
    if (temperature_strings.empty()) temperature_strings.push_back("1.0");
    if (temperature_strings.size() != 1u && (int)temperature_strings.size() != n_system)
        throw "Received " + std::to_string(temperature_strings.size()) + " temperatures but have " + std::to_string(n_system) + " systems";
    for (int ns = 0; ns < n_system; ++ns)
        systems[ns].initial_temperature = systems[ns].temperature =
            (float)stod_strict(temperature_strings.size() > 1u ? temperature_strings[ns] : temperature_strings[0]);

    const double anneal_factor = args.anneal_factor;
    const double anneal_duration = args.anneal_duration == -1. ? duration : args.anneal_duration;
    const double anneal_start = duration - anneal_duration;
    auto anneal_temp = [=](double T0, double time) {   // tighter spacing at the low end (main.cpp:425-430)
        double fraction = std::max(0., (time - anneal_start) / anneal_duration);
        double T1 = T0 * anneal_factor;
        double s = std::sqrt(T0) * (1. - fraction) + std::sqrt(T1) * fraction;
        return s * s;
    };
    int replica_interval = 0;
    if (args.replica_interval) replica_interval = (int)std::max(1., args.replica_interval / (3 * dt));
    const int mc_interval = args.mc_interval > 0. ? std::max(1, int(args.mc_interval / (3 * dt))) : 0;   // main.cpp:409-411
    if (!args.log_level.empty() && args.log_level != "basic" && args.log_level != "detailed" && args.log_level != "extensive")
        throw string("Illegal value for --log-level");
    // main.cpp:474-479.  extensive adds placement_pos, virtual and environment_coverage; with several placement nodes the
    // reference cannot create its second placement_pos dataset, and the duplicate check below stops the run the same way.
    // an empty --log-level means detailed, as in the reference (main.cpp:475)
    const int log_level = args.log_level == "basic" ? 0 : (args.log_level == "extensive" ? 2 : 1);

    // ---- load the configurations and group identical ones into batched engines --------------------------------------------
    // (errors while setting the systems up return 2, as the reference does: main.cpp:562-573)
    vector<Group> groups;
    try {
    for (int ns = 0; ns < n_system; ++ns) {
        System& sys = systems[ns];
        sys.path = args.configs[ns];
        sys.random_seed = base_random_seed + ns;
        try { sys.root = h5l::load(sys.path); } catch (const string&) { throw "Unable to open configuration file at " + sys.path; }
        sys.root->children.erase("output");
        const h5l::Node* pos = h5l::find(sys.root.get(), "/input/pos");
        if (!pos || pos->is_group || pos->data.dims.size() != 3) throw string("unable to read /input/pos");
        sys.n_atom = (int)pos->data.dims[0];
        if (pos->data.dims[1] != 3) throw string("invalid dimensions for initial position");
        if (pos->data.dims[2] != 1) throw string("must have n_system 1 from config");
        const h5l::Node* pot = h5l::find(sys.root.get(), "/input/potential");
        if (!pot || !pot->is_group) throw string("unable to open group /input/potential (does it exist?)");
        // the samplers of a group are read from its first system: systems whose move sets differ are not batched together
        // (the reference builds its samplers per system, main.cpp:543-546)
        auto same_moves = [&](const System& other) {
            if (!mc_interval) return true;
            for (const char* nm : {"/input/pivot_moves", "/input/jump_moves"}) {
                const h5l::Node *a = h5l::find(other.root.get(), nm), *b = h5l::find(sys.root.get(), nm);
                if (bool(a) != bool(b) || (a && !tree_equal(*a, *b))) return false;
            }
            return true;
        };
        for (size_t g = 0; g < groups.size() && sys.group < 0; ++g)
            if (groups[g].n_atom == sys.n_atom && tree_equal(*groups[g].potential, *pot) && same_moves(systems[groups[g].systems[0]]))
                sys.group = (int)g;
        if (sys.group < 0) {
            groups.emplace_back();
            groups.back().potential = pot;
            groups.back().n_atom = sys.n_atom;
            sys.group = (int)groups.size() - 1;
        }
        sys.slot = (int)groups[sys.group].systems.size();
        groups[sys.group].systems.push_back(ns);
        if (verbose) printf("%s\nn_atom %i\n\n", sys.path.c_str(), sys.n_atom);
    }
    // UPSIDE_B200_DEVICES=D: the systems of a group are dealt to D GPUs in contiguous blocks, one batched engine per device.
    // All engines advance concurrently (every call below only enqueues on its engine's stream); a ladder over one
    // Hamiltonian then exchanges on the devices over NCCL (csrc/ladder_nccl.cu), everything else goes through the host.
    int n_device = 1;
    if (const char* d = getenv("UPSIDE_B200_DEVICES")) {
        int have = 0;
        cudaGetDeviceCount(&have);
        n_device = std::max(1, std::min(atoi(d), have));
    }
    for (size_t g = 0; g < groups.size(); ++g) groups[g].origin = (int)g;
    if (n_device > 1) {
        vector<Group> split;
        for (auto& g : groups) {
            const int n = (int)g.systems.size(), D = std::min(n_device, n);
            for (int d = 0; d < D; ++d) {
                const int lo = d * (n / D) + std::min(d, n % D), cnt = n / D + (d < n % D ? 1 : 0);
                split.emplace_back();
                Group& s = split.back();
                s.potential = g.potential; s.n_atom = g.n_atom; s.device = d; s.origin = g.origin;
                s.systems.assign(g.systems.begin() + lo, g.systems.begin() + lo + cnt);
            }
        }
        groups = std::move(split);
        for (size_t g = 0; g < groups.size(); ++g)
            for (size_t r = 0; r < groups[g].systems.size(); ++r) { systems[groups[g].systems[r]].group = (int)g; systems[groups[g].systems[r]].slot = (int)r; }
        if (verbose) printf("%i systems on %i devices\n", n_system, n_device);
    }
    for (auto& g : groups) {
        g.engine = ub::initialize_engine_from_hdf5(g.n_atom, *g.potential, (int)g.systems.size(), g.device);
        for (const auto& p : set_param_map) g.engine->get(p.first).set_param(p.second);
        if (log_level > 0)
            for (auto& n : g.engine->nodes) {
                const size_t before = g.loggers.size();
                n.computation->add_loggers(log_level, g.loggers);
                for (size_t a = before; a < g.loggers.size(); ++a)
                    for (size_t b = 0; b < a; ++b)   // the reference's H5Logger cannot create a second dataset of the same name
                        if (g.loggers[a].name == g.loggers[b].name) throw "while adding '" + n.name + "', logger " + g.loggers[a].name + " exists already";
            }
        vector<float> all_pos;
        vector<float> T;
        vector<uint32_t> seeds;
        for (int ns : g.systems) {
            auto p = h5l::as<float>(h5l::find(systems[ns].root.get(), "/input/pos")->data);
            all_pos.insert(all_pos.end(), p.begin(), p.end());
            T.push_back(systems[ns].initial_temperature);
            seeds.push_back(systems[ns].random_seed);
        }
        g.engine->set_pos(all_pos.data());
        // quick hack of a check for z-centering and membrane potential (main.cpp:544-561)
        for (auto& n : g.engine->nodes) {
            auto pre = [&](const char* p) { return n.name.compare(0, strlen(p), p) == 0; };
            if (do_recenter && !xy_recenter_only && (pre("membrane_potential") || pre("z_flat_bottom") || pre("tension") || pre("AFM")))
                throw string("You have z-centering and a z-dependent potential turned on.  This is not what you want.  "
                             "Consider --disable-z-recentering or --disable-recentering.");
            if (do_recenter && pre("cavity_radial"))
                throw string("You have re-centering and a radial potential turned on.  This is not what you want.  "
                             "Consider --disable-recentering.");
        }
        if (args.deriv_agreement) {
            // central differences of the potential against dV/dx for the initial structure of the group's first system
            // (main.cpp:279-315), all 6*n_atom displaced copies evaluated as one batch
            const int n_coord = 3 * g.n_atom;
            auto probe = ub::initialize_engine_from_hdf5(g.n_atom, *g.potential, 2 * n_coord, 0);
            vector<float> p0(all_pos.begin(), all_pos.begin() + n_coord), batch(size_t(2) * n_coord * n_coord);
            const float eps = 1e-3f;
            for (int c = 0; c < n_coord; ++c)
                for (int sgn = 0; sgn < 2; ++sgn) {
                    float* dst = &batch[size_t(2 * c + sgn) * n_coord];
                    std::copy(p0.begin(), p0.end(), dst);
                    dst[c] += sgn ? -eps : eps;
                }
            probe->set_pos(batch.data());
            probe->compute(ub::PotentialAndDerivMode);
            probe->sync_and_check();
            auto e = probe->get_potential();
            g.engine->compute(ub::PotentialAndDerivMode);
            g.engine->sync_and_check();
            vector<float> deriv(size_t(g.engine->n_rep) * n_coord);
            g.engine->get_deriv(deriv.data());
            if (verbose) {
                printf("Initial potential:\n");
                for (auto& n : g.engine->nodes)
                    if (n.computation->potential_term) {
                        float v;
                        cudaMemcpy(&v, static_cast<ub::PotentialNode*>(n.computation.get())->potential, sizeof(float), cudaMemcpyDeviceToHost);
                        printf("%s: % 4.3f\n", n.name.c_str(), v);
                    }
                printf("\n\n");
            }
            double num = 0., den = 0.;
            for (int c = 0; c < n_coord; ++c) {
                double fd = (double(e[2 * c]) - double(e[2 * c + 1])) / (2. * eps);
                num += (fd - deriv[c]) * (fd - deriv[c]);
                den += double(deriv[c]) * deriv[c];
            }
            if (verbose) printf("overall potential relative error:  %.5f\n", std::sqrt(num / den));
        }
        g.engine->md_init_seeds(seeds.data(), T.data(), dt, (float)args.thermostat_timescale, thermostat_interval);
        if (mc_interval) {   // samplers of the group's first system (the systems of a group share /input/potential; main.cpp:543-546)
            const h5l::Node* input = h5l::find(systems[g.systems[0]].root.get(), "/input");
            if (!input) throw string("unable to open group /input");
            g.engine->mc_init(*input);
        }
    }
    } catch (const string& e) {
        fprintf(stderr, "\n\nERROR: %s\n", e.c_str());
        return 2;
    }

    std::unique_ptr<ub::ReplicaExchangePlan> replex;
    if (replica_interval) {
        if (verbose) printf("initializing replica exchange\n");
        replex.reset(new ub::ReplicaExchangePlan(n_system, args.swap_sets));
        if (replex->swap_sets.empty()) throw string("replica exchange requested but no swap sets proposed");
        for (auto& sys : systems)
            if (sys.n_atom != systems[0].n_atom) throw string("Replica exchange requires all systems have the same number of atoms");
    }
    // A ladder whose rungs all share one Hamiltonian and sit in equal contiguous blocks, one per engine, exchanges on the
    // devices: one evaluation per attempt instead of two per swap set, energies all-gathered and boundary coordinates sent
    // over NCCL, no host synchronisation (csrc/ladder_nccl.cu).  Annealing changes the temperatures: host path.
    vector<void*> ladder_comms;
    struct CommGuard { vector<void*>& v; ~CommGuard() { for (void* c : v) ub::nccl_comm_destroy(c); } } comm_guard{ladder_comms};
    vector<std::unique_ptr<ub::Ladder>> ladders;   // (destroyed before the communicators)
    if (replex && !getenv("UPSIDE_B200_HOST_REPLEX") && anneal_factor == 1.) {
        bool ok = true;
        for (size_t g = 0; g < groups.size(); ++g) {
            ok = ok && groups[g].origin == groups[0].origin && groups[g].systems.size() == groups[0].systems.size();
            for (size_t r = 0; r < groups[g].systems.size() && ok; ++r) ok = groups[g].systems[r] == int(g * groups[0].systems.size() + r);
        }
        if (ok) {
            vector<float> T(n_system);
            for (int ns = 0; ns < n_system; ++ns) T[ns] = systems[ns].temperature;
            if (groups.size() > 1) {
                vector<int> devs;
                for (auto& g : groups) devs.push_back(g.device);
                ladder_comms = ub::nccl_comm_init_all(devs);
            }
            for (size_t g = 0; g < groups.size(); ++g)
                ladders.emplace_back(new ub::Ladder(groups[g].engine.get(), ladder_comms.empty() ? nullptr : ladder_comms[g], (int)g,
                                                    (int)groups.size(), n_system, args.swap_sets, base_random_seed, T.data()));
            if (verbose) printf("replica exchange on the device%s\n", groups.size() > 1 ? "s, over NCCL" : "");
        }
    }
    // host copy of the exchange bookkeeping (replica indices, attempt / success counters) for the frame logger
    auto pull_ladder_state = [&]() {
        if (ladders.empty()) return;
        vector<int> ri;
        vector<unsigned long long> cnt;
        ladders[0]->download(&ri, nullptr, &cnt, nullptr);
        replex->replica_indices = ri;
        size_t k = 0;
        for (auto& set : replex->swap_sets)
            for (auto& sp : set) { sp.n_attempt = cnt[2 * k]; sp.n_success = cnt[2 * k + 1]; ++k; }
    };
    if (verbose) {
        printf("\n");
        for (int ns = 0; ns < n_system; ++ns) printf("%i %.2f\n", ns, systems[ns].temperature);
        printf("\n");
    }

    // energies of every system slot; one batched evaluation per engine
    vector<float> energy(n_system);
    auto compute_energies = [&]() {
        for (auto& g : groups) g.engine->compute(ub::PotentialAndDerivMode);
        for (auto& g : groups) {
            g.engine->sync_and_check();
            auto e = g.engine->get_potential();
            for (size_t r = 0; r < g.systems.size(); ++r) energy[g.systems[r]] = e[r];
        }
    };
    compute_energies();
    if (verbose) {
        printf("Initial potential energy:");
        for (int ns = 0; ns < n_system; ++ns) printf(" %.2f", energy[ns]);
        printf("\n");
    }

    // exchange coordinates of two system slots (same group: a device swap; different groups: through the host)
    auto coord_swap = [&](int s1, int s2) {
        System &a = systems[s1], &b = systems[s2];
        if (a.group == b.group) { groups[a.group].engine->swap_pos({a.slot, b.slot}); return; }
        vector<float> pa(size_t(3) * a.n_atom), pb(size_t(3) * b.n_atom);
        groups[a.group].engine->get_pos(pa.data(), a.slot, 1);
        groups[b.group].engine->get_pos(pb.data(), b.slot, 1);
        groups[a.group].engine->set_pos(pb.data(), a.slot, 1);
        groups[b.group].engine->set_pos(pa.data(), b.slot, 1);
    };
    auto attempt_swaps = [&](uint32_t seed, uint64_t round) {   // main.cpp:227-275
        if (!ladders.empty()) {
            vector<ub::Ladder*> all;
            for (auto& l : ladders) all.push_back(l.get());
            ub::Ladder::attempt_all(all, round);
            return;
        }
        vector<float> beta(n_system), old_l(n_system), new_l(n_system);
        for (int i = 0; i < n_system; ++i) beta[i] = 1.f / systems[i].temperature;
        ub::HostRandomGenerator random(seed, ub::REPLICA_EXCHANGE_RANDOM_STREAM, 0u, round);
        for (size_t is = 0; is < replex->swap_sets.size(); ++is) {
            auto& set = replex->swap_sets[is];
            compute_energies();
            for (int i = 0; i < n_system; ++i) old_l[i] = -beta[i] * energy[i];
            for (auto& sp : set) coord_swap(sp.sys1, sp.sys2);
            compute_energies();   // the Hamiltonians of the two slots may differ (Hamiltonian exchange): evaluate again
            for (int i = 0; i < n_system; ++i) new_l[i] = -beta[i] * energy[i];
            vector<int> accept(set.size());
            replex->decide((int)is, old_l.data(), new_l.data(), random, accept.data());
            for (size_t i = 0; i < set.size(); ++i) if (!accept[i]) coord_swap(set[i].sys1, set[i].sys2);   // reverse rejected swaps
        }
    };

    auto log_frame = [&](uint64_t nr) {
        pull_ladder_state();
        for (auto& g : groups) {
            if (do_recenter) g.engine->recenter(xy_recenter_only);
            g.engine->compute(ub::PotentialAndDerivMode);
        }
        for (auto& g : groups) {
            g.engine->sync_and_check();
            const int B = g.engine->n_rep;
            vector<float> pos(size_t(B) * 3 * g.n_atom);
            g.engine->get_pos(pos.data());
            auto pot = g.engine->get_potential();
            auto kin = g.engine->kinetic_energy();
            vector<vector<uint64_t>> mc_ok(g.engine->mc_n_samplers(), vector<uint64_t>(B)), mc_try(mc_ok);
            for (int m = 0; m < g.engine->mc_n_samplers(); ++m) g.engine->mc_stats(m, mc_ok[m].data(), mc_try[m].data(), true);
            for (int r = 0; r < B; ++r) {
                const int ns = g.systems[r];
                System& sys = systems[ns];
                for (int m = 0; m < g.engine->mc_n_samplers(); ++m) {
                    auto& v = sys.mc_stats[g.engine->mc_sampler_name(m) + "_stats"];
                    v.push_back((int)mc_ok[m][r]);
                    v.push_back((int)mc_try[m][r]);
                }
                const float* p = &pos[size_t(r) * 3 * g.n_atom];
                sys.pos.insert(sys.pos.end(), p, p + 3 * g.n_atom);
                sys.kinetic.push_back(kin[r]);
                sys.potential.push_back(pot[r]);
                sys.time.push_back(3 * double(dt) * nr);
                sys.temperature_log.push_back(sys.temperature);
                for (auto& lg : g.loggers) {
                    auto& series = sys.node_series[lg.name];
                    series.dims = lg.dims;
                    series.integer = lg.integer;
                    auto v = lg.sample(r);
                    series.data.insert(series.data.end(), v.begin(), v.end());
                }
                if (replex) {
                    sys.replica_index.push_back(replex->replica_indices[ns]);
                    for (auto& ps : replex->participating_swaps[ns]) {
                        const auto& sw = replex->swap_sets[ps.first][ps.second];
                        sys.cumulative_swaps.push_back((int)sw.n_success);
                        sys.cumulative_swaps.push_back((int)sw.n_attempt);
                    }
                }
                sys.n_frame++;
                if (verbose) {
                    double cx = 0, cy = 0, cz = 0, Rg = 0;
                    for (int na = 0; na < g.n_atom; ++na) { cx += p[3 * na]; cy += p[3 * na + 1]; cz += p[3 * na + 2]; }
                    cx /= g.n_atom; cy /= g.n_atom; cz /= g.n_atom;
                    for (int na = 0; na < g.n_atom; ++na)
                        Rg += (p[3 * na] - cx) * (p[3 * na] - cx) + (p[3 * na + 1] - cy) * (p[3 * na + 1] - cy) + (p[3 * na + 2] - cz) * (p[3 * na + 2] - cz);
                    Rg = std::sqrt(Rg / g.n_atom);
                    double n_hbond = -1.;
                    if (g.engine->get_idx("hbond_energy", false) != -1) n_hbond = g.engine->get("hbond_energy").get_value_by_name(r, "n_hbond")[0];
                    printf("%*.0f / %*.0f elapsed %2i system %.2f temp %5.1f hbonds, Rg %5.1f A, potential % 8.2f\n", duration_print_width,
                           nr * 3 * double(dt), duration_print_width, duration, ns, sys.temperature, n_hbond, Rg, pot[r]);
                }
            }
        }
        if (verbose) fflush(stdout);
    };

    SignalHandlerHandler sigint_handler(SIGINT, abort_like_handler);
    SignalHandlerHandler sigterm_handler(SIGTERM, abort_like_handler);
    received_signal = 0;

    auto tstart = std::chrono::high_resolution_clock::now();
    uint64_t nr = 0, last_start = 0;
    // a failure inside the loop (pair-list capacity, CUDA error) must not lose the frames already sampled: they are
    // written before the error propagates (the reference's H5Logger flushes as it goes)
    auto flush_all = [&]() {
        try { pull_ladder_state(); } catch (...) {}
        for (int ns = 0; ns < n_system; ++ns) write_output(systems[ns], invocation, bool(replex), replex.get(), ns);
    };
    try {
    while (nr < n_round && !received_signal) {
        // no pivot at t=0, so that a partially strained system may relax first (main.cpp:628-631)
        if (nr && mc_interval && !(nr % mc_interval)) for (auto& g : groups) g.engine->mc_execute(nr);
        if (!(nr % frame_interval)) log_frame(nr);
        // rounds that can run on the device without the host: up to the next frame, the next exchange, the next annealing step
        uint64_t n = std::min<uint64_t>(n_round - nr, frame_interval - nr % frame_interval);
        if (mc_interval) n = std::min<uint64_t>(n, mc_interval - nr % mc_interval);
        uint64_t swap_round = 0;
        if (replica_interval) {   // the reference leaves its inner loop at the first nr > last_start with (nr+1) % interval == 0
            swap_round = ((last_start + 2 + replica_interval - 1) / replica_interval) * replica_interval;
            n = std::min<uint64_t>(n, swap_round - nr);
        }
        if (anneal_factor != 1.) {
            if (!(nr % thermostat_interval)) {
                for (auto& g : groups) {
                    vector<float> T;
                    for (int ns : g.systems) {
                        systems[ns].temperature = (float)anneal_temp(systems[ns].initial_temperature, 3 * dt * double(nr + 1));
                        T.push_back(systems[ns].temperature);
                    }
                    g.engine->set_temperature(T.data());
                }
            }
            n = std::min<uint64_t>(n, thermostat_interval - nr % thermostat_interval);
        }
        for (auto& g : groups) g.engine->md_run((long)n);
        for (auto& g : groups) g.engine->sync_and_check();
        nr += n;
        if (replica_interval && nr == swap_round) {
            last_start = nr;
            attempt_swaps(base_random_seed, nr);
        } else if (replica_interval && nr == n_round && !(nr % replica_interval)) {
            attempt_swaps(base_random_seed, nr);
        }
    }
    } catch (...) {
        try { flush_all(); } catch (...) {}
        throw;
    }
    if (received_signal) fprintf(stderr, "Received early termination signal\n");
    double elapsed = std::chrono::duration<double>(std::chrono::high_resolution_clock::now() - tstart).count();
    flush_all();
    if (verbose) {
        printf("\n\nfinished in %.1f seconds (%.2f us/systems/step, %.1e simulation_time_unit/hour)\n", elapsed,
               elapsed * 1e6 / n_system / std::max<uint64_t>(nr, 1) / 3, nr * 3 * dt / elapsed * 3600.);
        if (mc_interval) {   // main.cpp:697-724
            for (const char* nm : {"pivot", "jump"}) {
                bool any = false;
                for (auto& sys : systems) any = any || sys.mc_stats.count(string(nm) + "_stats");
                if (!any) continue;
                printf("\n%s_success:\n", nm);
                for (auto& sys : systems) {
                    long ok = 0, tr = 0;
                    auto it = sys.mc_stats.find(string(nm) + "_stats");
                    if (it != sys.mc_stats.end()) for (size_t i = 0; i < it->second.size(); i += 2) { ok += it->second[i]; tr += it->second[i + 1]; }
                    printf(" % .4f", double(ok) / double(tr));
                }
                printf("\n");
            }
        }
        printf("\navg_kinetic_energy/1.5kT");
        for (auto& sys : systems) {
            double sum = 0.;
            long cnt = 0;
            for (size_t nf = 0; nf < sys.kinetic.size(); ++nf) if (nf > sys.kinetic.size() / 2) { sum += sys.kinetic[nf]; cnt++; }
            printf(" % .3f", sum / cnt / (1.5 * sys.temperature));
        }
        printf("\n");
    }
    if (args.re_raise_signal && received_signal) raise(received_signal);
    return 0;
}

}  // namespace

extern "C" int upside_main(int argc, const char* const* argv, int verbose) {
    try {
        return run(argc, argv, verbose);
    } catch (const ArgError& e) {
        fprintf(stderr, "\n\nERROR: %s for argument %s\n", e.msg.c_str(), e.arg.c_str());
        return 1;
    } catch (const string& e) {
        fprintf(stderr, "\n\nERROR: %s\n", e.c_str());
        return 1;
    } catch (const char* e) {
        fprintf(stderr, "\n\nERROR: %s\n", e);
        return 1;
    } catch (const std::exception& e) {
        fprintf(stderr, "\n\nERROR: %s\n", e.what());
        return 1;
    } catch (...) {
        fprintf(stderr, "\n\nERROR: unknown error\n");
        return 1;
    }
}
