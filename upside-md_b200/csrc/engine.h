// Host-side framework of the B200 engine: a batched re-design of the reference's DerivEngine
// (src/deriv_engine.h:48-335).  The node-registry "plugin" API is kept - a node type derives CoordNode or
// PotentialNode, is constructed from its HDF5 group plus its argument nodes, and is registered by a file-scope
// RegisterNodeType<T,n_args> under a name PREFIX - but instead of computing on the CPU a node enqueues CUDA
// kernels that process ALL replicas of the batch at once, and the engine captures a whole evaluation / MD round
// into a CUDA graph.
//
// Device layout: a CoordNode owns output[B][n_elem][wp] and sens[B][n_elem][wp] (wp = padded row width, rows are
// float4-aligned); all sens buffers live in one arena so an evaluation zeroes them with a single memset.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "common.cuh"
#include "h5lite.h"

namespace ub {

void cuda_check(cudaError_t e, const char* what);
#define UB_CUDA(x) ::ub::cuda_check((x), #x)

enum ComputeMode { DerivMode = 0, PotentialAndDerivMode = 1 };   // reference deriv_engine.h:41-44

struct Engine;

// ---- config access helpers (mirror src/h5_support.h semantics on the in-memory HDF5 tree) ----------------
const h5l::Node& h5_child(const h5l::Node& g, const std::string& name);
bool h5_has(const h5l::Node& g, const std::string& name);
std::vector<uint64_t> h5_dims(const h5l::Node& g, const std::string& name, int ndim);
void h5_check_size(const h5l::Node& g, const std::string& name, std::vector<uint64_t> dims);
template <typename T> std::vector<T> h5_read(const h5l::Node& g, const std::string& name);
template <typename T> T h5_attr(const h5l::Node& g, const std::string& path, const std::string& attr);
template <typename T> T h5_attr(const h5l::Node& g, const std::string& path, const std::string& attr, T dflt);
std::vector<std::string> h5_attr_strings(const h5l::Node& g, const std::string& attr);

// ---- device memory ------------------------------------------------------------------------------------
template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    DevBuf() {}
    explicit DevBuf(size_t n_) { alloc(n_); }
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { if (p) cudaFree(p); }
    void alloc(size_t n_) {
        if (p) cudaFree(p);
        p = nullptr;
        n = n_;
        if (n) { UB_CUDA(cudaMalloc((void**)&p, n * sizeof(T))); UB_CUDA(cudaMemset(p, 0, n * sizeof(T))); }
    }
    void upload(const std::vector<T>& h) {
        if (h.size() != n) alloc(h.size());
        if (n) UB_CUDA(cudaMemcpy(p, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
    }
    std::vector<T> download() const {
        std::vector<T> h(n);
        if (n) UB_CUDA(cudaMemcpy(h.data(), p, n * sizeof(T), cudaMemcpyDeviceToHost));
        return h;
    }
};

// ---- nodes ----------------------------------------------------------------------------------------------
// A named per-frame series a node offers to the /output writer of upside_main (reference H5Logger::add_logger,
// state_logger.h:107-141; registered in the node constructors when logging(level) holds).
struct NodeLogger {
    std::string name;
    std::vector<uint64_t> dims;                              // shape of one frame
    bool integer = false;                                    // written as int64 (rotamer_bad_solves_cumulative)
    std::function<std::vector<float>(int replica)> sample;   // called after the frame's evaluation
};

struct DerivComputation {
    const bool potential_term;
    Engine* engine = nullptr;
    std::string name;
    explicit DerivComputation(bool pt) : potential_term(pt) {}
    virtual ~DerivComputation() {}
    // enqueue the kernels of compute_value / propagate_deriv for every replica on `s`
    virtual void compute_value(cudaStream_t s, ComputeMode mode) = 0;
    virtual void propagate_deriv(cudaStream_t s) = 0;
    // called once after all nodes exist and device buffers are assigned
    virtual void finalize() {}
    virtual std::vector<float> get_param() const { return {}; }
    virtual void set_param(const std::vector<float>&) {}
    // dV/d(parameter) from the state the last evaluation with derivatives left behind (reference PARAM_DERIV builds);
    // replica < 0 sums over the batch.  Empty = the node has no parameter derivative.
    virtual std::vector<float> get_param_deriv(int replica = 0) { return {}; }
    virtual std::vector<float> get_value_by_name(int replica, const char* log_name) {
        throw std::string("No values implemented");
    }
    // loggers of this node at `level` (1 = detailed, 2 = extensive; state_logger.h:56-66)
    virtual void add_loggers(int level, std::vector<NodeLogger>& out) {}
    // pair list of `replica` built by the last evaluation, in the reference's emission order; false if none
    virtual bool get_pairlist(int replica, std::vector<int>& i1, std::vector<int>& i2) { return false; }
};

struct CoordNode : DerivComputation {
    int n_elem, elem_width, wp;
    float* output = nullptr;   // [B][n_elem][wp]  (assigned by the engine)
    float* sens = nullptr;     // [B][n_elem][wp]
    size_t stride() const { return size_t(n_elem) * wp; }
    // host copy of one replica's padded rows of `base` (= output or sens); accessors and loggers only
    std::vector<float> host_rows(const float* base, int replica) const {
        std::vector<float> h(stride());
        if (!h.empty()) UB_CUDA(cudaMemcpy(h.data(), base + size_t(replica) * stride(), h.size() * sizeof(float), cudaMemcpyDeviceToHost));
        return h;
    }
    CoordNode(int n_elem_, int elem_width_)
        : DerivComputation(false), n_elem(n_elem_), elem_width(elem_width_), wp(ub_padded_width(elem_width_)) {}
};

struct PotentialNode : DerivComputation {
    float* potential = nullptr;   // [B] (assigned by the engine)
    PotentialNode() : DerivComputation(true) {}
    void propagate_deriv(cudaStream_t) override {}
};

struct Pos : CoordNode {
    explicit Pos(int n_atom) : CoordNode(n_atom, 3) {}
    void compute_value(cudaStream_t, ComputeMode) override {}
    void propagate_deriv(cudaStream_t) override {}
};

typedef std::vector<CoordNode*> ArgList;
typedef std::function<DerivComputation*(Engine&, const h5l::Node&, const ArgList&)> NodeCreationFunction;
std::map<std::string, NodeCreationFunction>& node_creation_map();
void add_node_creation_function(std::string name_prefix, NodeCreationFunction fcn);
void check_elem_width(const CoordNode& n, int w);
void check_elem_width_lower_bound(const CoordNode& n, int w);
void check_arguments_length(const ArgList& a, int n);

template <typename NodeClass, int n_args> struct RegisterNodeType;
template <typename NodeClass> struct RegisterNodeType<NodeClass, -1> {
    explicit RegisterNodeType(std::string p) {
        add_node_creation_function(p, [](Engine& e, const h5l::Node& g, const ArgList& a) -> DerivComputation* {
            if (a.empty()) throw std::string("Expected at least 1 arg");
            return new NodeClass(e, g, a);
        });
    }
};
template <typename NodeClass> struct RegisterNodeType<NodeClass, 0> {
    explicit RegisterNodeType(std::string p) {
        add_node_creation_function(p, [](Engine& e, const h5l::Node& g, const ArgList& a) -> DerivComputation* {
            check_arguments_length(a, 0);
            return new NodeClass(e, g);
        });
    }
};
template <typename NodeClass> struct RegisterNodeType<NodeClass, 1> {
    explicit RegisterNodeType(std::string p) {
        add_node_creation_function(p, [](Engine& e, const h5l::Node& g, const ArgList& a) -> DerivComputation* {
            check_arguments_length(a, 1);
            return new NodeClass(e, g, *a[0]);
        });
    }
};
template <typename NodeClass> struct RegisterNodeType<NodeClass, 2> {
    explicit RegisterNodeType(std::string p) {
        add_node_creation_function(p, [](Engine& e, const h5l::Node& g, const ArgList& a) -> DerivComputation* {
            check_arguments_length(a, 2);
            return new NodeClass(e, g, *a[0], *a[1]);
        });
    }
};
template <typename NodeClass> struct RegisterNodeType<NodeClass, 3> {
    explicit RegisterNodeType(std::string p) {
        add_node_creation_function(p, [](Engine& e, const h5l::Node& g, const ArgList& a) -> DerivComputation* {
            check_arguments_length(a, 3);
            return new NodeClass(e, g, *a[0], *a[1], *a[2]);
        });
    }
};

// ---- engine -----------------------------------------------------------------------------------------------
struct MonteCarlo;                 // pivot / jump samplers of the batch (monte_carlo.cu)
void mc_destroy(MonteCarlo* m);

struct Engine {
    struct Node {
        std::string name;
        std::unique_ptr<DerivComputation> computation;
        std::vector<size_t> parents, children;
    };

    int n_rep = 1;      // B: replicas in the batch
    int n_atom = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    std::vector<Node> nodes;   // construction order = reference topological order (deriv_engine.cpp:200-226)
    Pos* pos = nullptr;

    // device arenas
    DevBuf<float> out_arena, sens_arena, pot_arena;
    DevBuf<float> potential;         // [B] total potential of the last PotentialAndDerivMode evaluation
    DevBuf<float*> pot_ptrs;         // device table of per-node potential pointers
    int n_pot_nodes = 0;
    DevBuf<int> error_flag;          // device-side failure flag (pair-list overflow etc.), checked at sync points
    DevBuf<float> io_stage;          // staging for host I/O in (n,3) layout
    float* io_staging(size_t n_float);

    // MD state
    DevBuf<float> mom;               // [B][n_atom][4]
    DevBuf<float> temperature;       // [B]
    DevBuf<float> noise_scale;       // [B] OU noise scale per replica
    std::vector<float> h_temperature;
    std::vector<uint32_t> h_seed;    // per replica RNG key word 0
    DevBuf<uint32_t> seed;           // [B]
    float dt = 0.009f, thermostat_timescale = 5.f, mom_scale = 0.f;
    int thermostat_interval = 1;
    uint64_t n_thermostat_invocations = 0;
    uint64_t round_num = 0;
    DevBuf<unsigned long long> d_invocation;   // device copy of the thermostat counter (graph-updatable)

    // CUDA graphs (one per compute mode, one for an MD round)
    cudaGraphExec_t graph_eval[2] = {nullptr, nullptr};
    cudaGraphExec_t graph_round = nullptr;
    bool use_graphs = true;
    // While a graph is being captured the evaluation is issued as the node DAG, not as a chain: every node records its
    // kernels on its own stream, ordered after its inputs only, so independent branches of the force field (springs, Rama
    // terms, environment coverage, H-bond geometry ...) become parallel branches of the CUDA graph and overlap on the GPU.
    bool use_dag = true;
    std::vector<cudaStream_t> node_stream;
    std::vector<cudaEvent_t> ev_fwd, ev_bwd;
    cudaEvent_t ev_start = nullptr, ev_zeroed = nullptr;
    cudaStream_t zero_stream = nullptr;   // branch of the evaluation graph that zeroes the sens / potential arenas

    // Live timing of one evaluation kernel group by kernel group (CUDA events on the engine's stream, linear order, no graph):
    // while `profiling` is set, mark(s, label) records an event; the engine marks every node's forward / backward, nodes
    // with several kernels add finer marks (e.g. "rotamer/bp").  profile_eval returns (label, milliseconds) pairs.
    bool profiling = false;
    std::vector<std::pair<std::string, cudaEvent_t>> marks;
    void mark(cudaStream_t s, const std::string& label) {
        if (!profiling) return;
        cudaEvent_t e;
        UB_CUDA(cudaEventCreate(&e));
        UB_CUDA(cudaEventRecord(e, s));
        marks.emplace_back(label, e);
    }
    std::vector<std::pair<std::string, float>> profile_eval(ComputeMode mode);

    Engine(int n_atom, int n_rep, int device);
    ~Engine();
    Engine(const Engine&) = delete;

    void add_node(const std::string& name, std::unique_ptr<DerivComputation> c, const std::vector<std::string>& args);
    int get_idx(const std::string& name, bool must_exist = true) const;
    DerivComputation& get(const std::string& name);
    template <typename T> T& get_computation(const std::string& name) { return dynamic_cast<T&>(get(name)); }

    void allocate();                                // assign device buffers after all nodes are added
    void enqueue_compute(cudaStream_t s, ComputeMode mode);   // raw kernel sequence (used for capture)
    void enqueue_compute_dag(cudaStream_t s, ComputeMode mode);
    void compute(ComputeMode mode);                 // one evaluation of all replicas (graph replay), asynchronous
    void sync_and_check();                          // cudaStreamSynchronize + device error flag -> throws
    // One evaluation with HOST buffers in one CUDA graph: positions (n_rep,n_atom,3) from a pinned staging buffer, layout
    // conversion, the evaluation, derivatives and energies back into pinned memory - one launch, one synchronisation (the
    // reference ABI's evaluate_energy / evaluate_deriv pay for four synchronous copies otherwise).  energy / deriv may be NULL.
    void evaluate_host(const float* pos3, float* energy, float* deriv3);
    float* pinned_io = nullptr;                     // [n_rep*n_atom*3 (in) | n_rep*n_atom*3 (out) | n_rep (energy) | 1 (error flag)]
    cudaGraphExec_t graph_host_eval = nullptr;

    // I/O in caller layout: pos/deriv/mom [B][n_atom][3]
    void set_pos(const float* p, int first_rep = 0, int n = -1);
    void get_pos(float* p, int first_rep = 0, int n = -1);
    void get_deriv(float* p, int first_rep = 0, int n = -1);
    void set_mom(const float* p);
    void get_mom(float* p);
    std::vector<float> get_potential();

    // MD (reference main.cpp:515-523,657-663; deriv_engine.cpp:11-35,172-192; thermostat.cpp:9-18)
    void md_init(uint32_t base_seed, const float* temperature, float dt, float timescale, int thermostat_interval);
    // explicit RNG key per replica (replica r of a sharded or regrouped set keeps the seed of its global index)
    void md_init_seeds(const uint32_t* seeds, const float* temperature, float dt, float timescale, int thermostat_interval);
    void set_temperature(const float* temperature);
    void enqueue_thermostat(cudaStream_t s);
    void enqueue_integration_cycle(cudaStream_t s);
    void md_run(long n_round);
    void recenter(bool xy_only);
    void swap_pos(const std::vector<int>& pairs);   // pairs = (a0,b0,a1,b1,...): exchange coordinates of replicas a_k and b_k
    std::vector<float> kinetic_energy();

    // Checkpoint of the MD state: positions, momenta, per-replica RNG keys and temperatures, the thermostat invocation
    // counter and the round number (the counter-based random stream resumes exactly where it stopped), integrator settings.
    // The reference has no such thing: its continue_sim (py/run_upside.py:231-257) restarts from the last logged frame with
    // freshly drawn momenta.  A blob restores into an engine of the same configuration and batch size.
    std::vector<char> checkpoint_save();
    void checkpoint_load(const char* data, size_t size);

    // Monte-Carlo moves (reference monte_carlo_sampler.cpp; main.cpp:545,630-631): samplers read from the /input group of a
    // configuration (pivot_moves, jump_moves); one execute = one Metropolis step of every sampler for every replica, with
    // the replica's md_init seed and current temperature
    MonteCarlo* mc = nullptr;
    void mc_init(const h5l::Node& input_group);
    void mc_execute(uint64_t round);
    int mc_n_samplers() const;
    std::string mc_sampler_name(int i) const;
    void mc_stats(int i, uint64_t* n_success, uint64_t* n_attempt, bool reset);   // per replica
};

std::unique_ptr<Engine> initialize_engine_from_hdf5(int n_atom, const h5l::Node& potential_group, int n_rep,
                                                    int device);

// reference-order pair list: sort edges by (i1>>2, i2, i1&3)   (interaction_graph.h:122-158 emission order)
void sort_reference_order(std::vector<int>& i1, std::vector<int>& i2);

}  // namespace ub
