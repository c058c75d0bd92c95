/* Batched C ABI of the B200 engine (plain pointers and sizes, no C++/torch types).
 *
 * The reference runs one DerivEngine per OpenMP thread (src/main.cpp:618-667).  Here one engine holds a BATCH of
 * n_replica copies of one configuration on one GPU and advances all of them per kernel launch.  These entry points
 * are what a reference-side driver binds instead of looping over `System`s; each cites the reference code it
 * replaces.  All functions return 0 on success, 1 on failure (message retrievable with ub_last_error, also printed to
 * stderr), like the reference's engine_c_library.  Buffers are caller-owned, C-contiguous.
 */
#ifndef UPSIDE_B200_H
#define UPSIDE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct UbEngine UbEngine;

const char* ub_last_error(void);
int ub_device_count(void);

/* initialize_engine_from_hdf5 for n_replica copies (src/deriv_engine.cpp:195-270, src/main.cpp:486-502).
 * config_path: .up file; the potential is read from /input/potential, n_atom from /input/pos.  NULL on failure. */
UbEngine* ub_engine_create(const char* config_path, int n_replica, int device);
void ub_engine_destroy(UbEngine* e);
int ub_n_atom(const UbEngine* e);
int ub_n_replica(const UbEngine* e);

/* /input/pos of the configuration, (n_atom,3) */
int ub_initial_pos(UbEngine* e, float* pos);

/* positions / momenta of all replicas, (n_replica,n_atom,3) */
int ub_set_pos(UbEngine* e, const float* pos);
int ub_get_pos(UbEngine* e, float* pos);
int ub_set_mom(UbEngine* e, const float* mom);
int ub_get_mom(UbEngine* e, float* mom);

/* DerivEngine::compute(PotentialAndDerivMode) for every replica (src/deriv_engine.cpp:124-169):
 * energy (n_replica) and/or deriv (n_replica,n_atom,3) may be NULL */
int ub_evaluate(UbEngine* e, float* energy, float* deriv);

/* per-node access for one replica (src/engine_c_library.cpp:112-194) */
int ub_n_nodes(UbEngine* e);
int ub_node_name(UbEngine* e, int index, char* buf, int buf_len, int* is_potential);
int ub_get_output_dims(UbEngine* e, const char* node, int* n_elem, int* elem_width);
int ub_get_output(UbEngine* e, const char* node, int replica, int n, float* out);
int ub_get_sens(UbEngine* e, const char* node, int replica, int n, float* out);
int ub_get_node_potential(UbEngine* e, const char* node, float* out /* n_replica */);
int ub_get_value_by_name(UbEngine* e, const char* node, const char* log_name, int replica, int n, float* out, int* n_written);
int ub_get_param(UbEngine* e, const char* node, int n, float* out, int* n_param);
int ub_set_param(UbEngine* e, const char* node, int n, const float* param);
/* dV/d(parameter) of `node` from the state of the last evaluation with derivatives (reference get_param_deriv of a
 * PARAM_DERIV build, src/engine_c_library.cpp:93-108, src/interaction_graph.h:404-415); replica < 0 sums over the
 * batch; out == NULL only reports the size (0 = the node has no parameter derivative) */
int ub_get_param_deriv(UbEngine* e, const char* node, int replica, int n, float* out, int* n_param);

/* pair list a node built in the last evaluation, in the reference's emission order
 * (PairlistComputation::find_edges, src/interaction_graph.h:201-257); *n_edge receives the full count */
int ub_get_pairlist(UbEngine* e, const char* node, int replica, int max_edge, int* i1, int* i2, int* n_edge);

/* MD: System setup of src/main.cpp:515-523 (seed_r = base_seed + r, OU thermostat, initial thermalisation),
 * then n_round x { thermostat every `thermostat_interval` rounds ; integration_cycle } (src/main.cpp:657-663,
 * src/deriv_engine.cpp:172-192, src/thermostat.cpp:9-18).  temperature: (n_replica). */
int ub_md_init(UbEngine* e, uint32_t base_seed, const float* temperature, float dt, float thermostat_timescale,
               int thermostat_interval);
int ub_md_set_temperature(UbEngine* e, const float* temperature);
/* as ub_md_init with an explicit RNG key per replica: a replica of a sharded or regrouped set keeps the seed of its global
 * index (reference: sys->random_seed = base_random_seed + ns, src/main.cpp:459) */
int ub_md_init_seeds(UbEngine* e, const uint32_t* seeds, const float* temperature, float dt, float thermostat_timescale,
                     int thermostat_interval);

/* coordinates of a range of replicas, (n_replica,n_atom,3); exchange of coordinates between replica pairs
 * (pairs = a0,b0,a1,b1,...), the coord_swap of ReplicaExchange::attempt_swaps (src/main.cpp:240-243) */
int ub_set_pos_range(UbEngine* e, const float* pos, int first_replica, int n_replica);
int ub_get_pos_range(UbEngine* e, float* pos, int first_replica, int n_replica);
int ub_swap_pos(UbEngine* e, int n_pair, const int* pairs);

/* Replica exchange, host side (src/main.cpp:120-275): swap-set parsing/validation, the Metropolis pass and its
 * counter-based random stream (src/random.h:19-66, stream REPLICA_EXCHANGE_RANDOM_STREAM).  No device work: every rank
 * of a sharded ladder calls this with the all-gathered energies and reaches the same decisions.
 *   ub_replex_begin   once per attempt_swaps call (seed, round): restarts the draw counter
 *   ub_replex_decide  one swap set: old/new_lboltz[i] = -beta_i*E_i before/after the trial exchange; accept[k] = 1 if
 *                     pair k stays exchanged (a rejected pair must be swapped back by the caller)
 *   ub_replex_decide_same_hamiltonian  temperature ladder over ONE Hamiltonian: trial energies are a permutation of the
 *                     current ones, no second evaluation */
typedef struct UbReplex UbReplex;
UbReplex* ub_replex_create(int n_system, int n_set, const char* const* swap_sets);
void ub_replex_destroy(UbReplex* h);
int ub_replex_n_sets(const UbReplex* h);
int ub_replex_set_size(const UbReplex* h, int set);
int ub_replex_pairs(const UbReplex* h, int set, int* pairs /* 2*set_size */);
int ub_replex_begin(UbReplex* h, uint32_t seed, uint64_t round);
int ub_replex_decide(UbReplex* h, int set, const float* old_lboltz, const float* new_lboltz, int* accept);
int ub_replex_decide_same_hamiltonian(UbReplex* h, int set, const float* beta, const float* energy, int* accept);
int ub_replex_replica_indices(const UbReplex* h, int* out /* n_system */);
int ub_replex_counts(const UbReplex* h, int set, uint64_t* n_attempt, uint64_t* n_success);
/* Replica exchange of a ladder SHARDED OVER GPUs, on the device and on the engine's stream (src/main.cpp:227-275; the 8-GPU
 * ladder of BASELINE.json config 4).  Rank g of `world` owns the contiguous rungs [g*n_local, (g+1)*n_local) as the replicas
 * of its engine (n_local = n_global/world = ub_n_replica(e)).  One attempt = one batched energy evaluation, an
 * ncclAllGather of the energies, the Metropolis pass of ALL swap sets on the device (same counter-based random stream as
 * the reference; identical on every rank, nothing is broadcast), and per swap set a grouped ncclSend/ncclRecv of the
 * coordinates of block-boundary rungs.  Asynchronous like ub_md_run; nothing synchronises with the host.
 *   nccl_comm: an ncclComm_t whose rank/size match (NULL for world == 1).  ub_nccl_unique_id + ub_nccl_comm_create make
 *   one (rank 0 creates the 128-byte id and hands it to the other ranks by any means); ub_nccl_comm_create_all makes one
 *   per device for a single process driving several GPUs.  NCCL is bound at run time (dlopen), not linked. */
typedef struct UbLadder UbLadder;
const char* ub_ladder_last_error(void);
int ub_nccl_unique_id(char* out, int len /* >= 128 */);
void* ub_nccl_comm_create(const char* unique_id, int world, int rank, int device);
int ub_nccl_comm_create_all(int n_device, const int* devices, void** comms /* n_device */);
void ub_nccl_comm_destroy(void* comm);
UbLadder* ub_ladder_create(UbEngine* e, void* nccl_comm, int rank, int world, int n_global, int n_set, const char* const* swap_sets,
                           uint32_t seed, const float* temperature_all /* n_global */);
void ub_ladder_destroy(UbLadder* h);
int ub_ladder_attempt(UbLadder* h, uint64_t round);
int ub_ladder_set_temperature(UbLadder* h, const float* temperature_all);
int ub_ladder_n_pairs(const UbLadder* h);
/* waits for the stream; any pointer may be NULL.  replica_indices: n_global; accept / n_attempt / n_success: one entry per
 * swap pair, sets concatenated; energies: the n_global energies the last attempt decided on */
int ub_ladder_state(UbLadder* h, int* replica_indices, int* accept, uint64_t* n_attempt, uint64_t* n_success, float* energies);
/* bytes this rank receives in the all-gather and sends as boundary coordinates per attempt */
int ub_ladder_comm_bytes(const UbLadder* h, uint64_t* allgather_bytes, uint64_t* coordinate_bytes);

/* known-answer access to the host generator: n_draw successive uniform_open_closed().x values (+ raw bits of the first) */
int ub_host_rng_uniform(uint32_t seed, uint32_t stream, uint32_t atom, uint64_t timestep, int n_draw, float* out,
                        uint32_t* bits_first /* 4 or NULL */);
/* Checkpoint of the MD state of all replicas: positions, momenta, RNG keys, temperatures, thermostat invocation counter and
 * round number, integrator settings - a run resumed from it draws the same thermostat noise as the uninterrupted run
 * (the reference's continue_sim, py/run_upside.py:231-257, restarts from the last frame with
 * fresh momenta).  ub_checkpoint_size: upper bound of the blob; save writes it into buf (*written = its size); load
 * restores into an engine of the same configuration and batch size (no md_init needed before). */
long ub_checkpoint_size(UbEngine* e);
int ub_checkpoint_save(UbEngine* e, void* buf, long buf_size, long* written);
int ub_checkpoint_load(UbEngine* e, const void* buf, long size);
int ub_md_run(UbEngine* e, long n_round);      /* asynchronous; ub_sync waits and reports device-side failures */
int ub_sync(UbEngine* e);
int ub_recenter(UbEngine* e, int xy_only);     /* src/deriv_engine.cpp:37-48 */
int ub_kinetic_energy(UbEngine* e, float* out /* n_replica, per atom */);

/* CUDA stream the engine launches on (cudaStream_t), for timing with CUDA events on that stream */
void* ub_stream(UbEngine* e);
/* number of kernels the engine launches per evaluation / per MD round (for the bench's gpu_launches claim) */
int ub_launches_per_eval(UbEngine* e);

/* One DerivMode evaluation timed kernel group by kernel group with CUDA events on the engine's stream (linear order, no
 * graph): label i (NUL-terminated, label_len bytes each) took ms[i].  Labels are "<node>:fwd", "<node>:bwd" and finer
 * marks inside multi-kernel nodes ("rotamer/pairlist", "rotamer/prep", "rotamer/energy", "rotamer/bp", "rotamer/deriv"). */
int ub_profile_eval(UbEngine* e, int max_entry, char* labels, int label_len, float* ms, int* n_entry);

/* known-answer access to the device RNG (src/random.h:19-66): threefry bits, normal3 and u01(word0) */
int ub_rng_probe(uint32_t seed, uint32_t stream, uint32_t atom, uint64_t timestep, uint32_t* bits4, float* normal3_u01);

/* Monte-Carlo pivot / jump moves (reference src/monte_carlo_sampler.cpp:255-308, src/main.cpp:545,630-631).  The samplers
 * are read from /input/pivot_moves and /input/jump_moves of the configuration when the engine is created.  One execute =
 * one Metropolis step of every sampler for EVERY replica: two batched energy evaluations per sampler, proposal and
 * acceptance on the device with the reference's random stream RandomGenerator(seed_r, stream, 0, round); seeds and
 * temperatures are those of ub_md_init*.  Asynchronous like ub_md_run.  Stats are per replica (n_success, n_attempt). */
int ub_mc_n_samplers(UbEngine* e);
int ub_mc_sampler_name(UbEngine* e, int index, char* buf, int buf_len);   /* "pivot", "jump" */
int ub_mc_execute(UbEngine* e, uint64_t round);
int ub_mc_stats(UbEngine* e, int index, uint64_t* n_success /* n_replica */, uint64_t* n_attempt, int reset);

/* FP32 FMA throughput this GPU sustains (independent FFMA chains on every SM, CUDA-event timed, best of 4 after warm-up):
 * the measured denominator of the FP32 roofline bench.py reports (MEASURED_PEAKS.json holds only HBM and bf16 numbers). */
int ub_measure_fp32_peak(int device, float* tflops, float* sm_mhz_nominal);

#ifdef __cplusplus
}
#endif
#endif
