/* Drop-in C ABI of the B200 engine: the SAME entry points, argument meaning and error convention as the reference's
 * src/engine_c_library.h:8-33 (+ src/main.h:1-3), which py/upside_engine.py:19-64 binds through ctypes.
 * Every function returns 0 on success and 1 on failure (message on stderr); constructors return NULL on failure
 * (reference src/engine_c_library.cpp:9-20,29-64).  Buffers are caller-owned C-contiguous float32.
 *
 * A DerivEngine created here holds ONE replica on the current CUDA device; the batched interface the MD driver and
 * the benchmarks use is in upside_b200.h.  There is no CPU fallback: construction fails if no CUDA device exists.
 */
#ifndef UPSIDE_B200_ENGINE_C_LIBRARY_H
#define UPSIDE_B200_ENGINE_C_LIBRARY_H
#include <stdbool.h>

#ifdef __cplusplus
extern "C" {
#endif

struct DerivEngine; /* opaque */

/* reference src/main.h:1-3, src/main.cpp:317 - the `upside` command line (same flags) */
int upside_main(int argc, const char* const* argv, int verbose);

/* reference src/engine_c_library.cpp:9-27 */
struct DerivEngine* construct_deriv_engine(int n_atom, const char* potential_file, bool quiet);
void free_deriv_engine(struct DerivEngine* engine);

/* reference :29-64  pos, deriv: (n_atom,3); deriv is dV/dx of a PotentialAndDerivMode evaluation */
int evaluate_energy(float* energy, struct DerivEngine* engine, const float* pos);
int evaluate_deriv(float* deriv, struct DerivEngine* engine, const float* pos);

/* reference :67-110 */
int set_param(int n_param, const float* param, struct DerivEngine* engine, const char* node_name);
int get_param_deriv(int n_param, float* deriv, struct DerivEngine* engine, const char* node_name);
int get_param(int n_param, float* param, struct DerivEngine* engine, const char* node_name);

/* reference :112-194  node outputs are (n_elem, elem_width) un-padded; a potential node reports (1,1) */
int get_output_dims(int* n_elem, int* elem_width, struct DerivEngine* engine, const char* node_name);
int get_output(int n_output, float* output, struct DerivEngine* engine, const char* node_name);
int get_sens(int n_output, float* output, struct DerivEngine* engine, const char* node_name);
int get_value_by_name(int n_output, float* output, struct DerivEngine* engine, const char* node_name, const char* log_name);

/* reference :197-276  clamped B-spline helpers used by the parameter-training scripts (host arithmetic) */
int clamped_spline_solve(int N, float* bspline_coeff, const float* values);
int clamped_spline_value(int N, float* result, const float* bspline_coeff, int nx, float* x);
int get_clamped_value_and_deriv(int N, float* result, const float* bspline_coeff, int nx, float* x);
int get_clamped_coeff_deriv(int N, float* result, const float* bspline_coeff, float x);

#ifdef __cplusplus
}
#endif
#endif
