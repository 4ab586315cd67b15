/* pqc_b200.h -- C ABI of the B200-native PQC statevector + capacity-measure engine.
 *
 * The reference (rmdocherty/pyramaterised) has no FFI: its "operator API" is Python
 * duck typing on top of QuTiP.  Each entry point below therefore names
 * the reference *Python call site* it replaces (paths relative to /root/reference).
 * The reference-facing Python shim (pyramaterised_b200/) binds these with ctypes;
 * INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; pqc_last_error() gives the
 *     message for the calling thread.  No exception crosses this boundary.
 *   - the caller owns every buffer.  Pointers named d_* are DEVICE pointers (e.g.
 *     torch-allocated); h_* are HOST pointers.  The library owns only pqc_program.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it and the
 *     functions do not synchronise unless stated.
 *   - states are complex128, interleaved (re,im), one row of D = 2^n amplitudes per
 *     sample.  Qubit 0 is the MOST significant bit of the basis index (qt.tensor order,
 *     circuit.py:22, gates.py:39-42); for two-qubit ops q0 is the first listed qubit
 *     (control / q1 in the reference).
 *   - angles are float64, row-major [S, ld_angles], column = parameter slot in the
 *     reference's set_params order (circuit.py:86-116).
 */
#ifndef PQC_B200_H
#define PQC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PQC_ABI_VERSION 1
#if defined(__GNUC__)
#define PQC_API __attribute__((visibility("default")))
#else
#define PQC_API
#endif
#define PQC_MAX_QUBITS 30

typedef struct pqc_program pqc_program;
typedef struct pqc_c128 { double re, im; } pqc_c128;

/* Primitive opcodes a reference Gate lowers to (gates.py). */
enum pqc_opcode {
  PQC_OP_RX = 0,        /* rx(a): [[c,-is],[-is,c]]           gates.py:159-162 (R_x)            */
  PQC_OP_RY = 1,        /* ry(a): [[c,-s],[s,c]]              gates.py:165-168 (R_y, fixed_R_y) */
  PQC_OP_RZ = 2,        /* rz(a): diag(e^-ia/2, e^ia/2)       gates.py:171-205,269-283          */
  PQC_OP_H = 3,         /* x_gate*ry(pi/2) = (1/sqrt2)[[1,1],[1,-1]]   gates.py:211-232          */
  PQC_OP_X = 4,         /* gates.py:245-249 */
  PQC_OP_S = 5,         /* phasegate(pi/2) = diag(1,i)        gates.py:286-291 */
  PQC_OP_T = 6,         /* diag(1,e^{i pi/4})                 gates.py:294-300 */
  PQC_OP_CNOT = 7,      /* q0 = control, q1 = target          gates.py:327-330 */
  PQC_OP_CZ = 8,        /* also CPHASE (quirk Q11)            gates.py:333-337,346-349 */
  PQC_OP_SQRTISWAP = 9, /* gates.py:340-343 */
  PQC_OP_RXX = 10,      /* cos(a/2) - i sin(a/2) XX           gates.py:492-543 */
  PQC_OP_RYY = 11,      /* gates.py:546-551 */
  PQC_OP_RZZ = 12,      /* gates.py:530-535 */
  PQC_OP_FSIM = 13,     /* two angles (param, param2)         gates.py:588-606,651-673 */
  PQC_OP_FIXED_FSIM = 14, /* gates.py:700-716,740-751 */
  PQC_OP_IDENT = 15,    /* `I` gate: identity that consumes a parameter (quirk Q5) gates.py:150-156 */
  PQC_OP__COUNT = 16
};

/* One primitive operation.  angle = scale * angles[s, param] + offset when param >= 0,
 * angle = offset when param < 0 (fixed_R_y/fixed_R_z).  negative_R_z: scale = -1
 * (gates.py:180-187); offset_R_z: offset (gates.py:190-205).  fSim family (full angles,
 * no half): theta = offset + angles[s, param]; phi = angles[s, param2], or the `scale`
 * field when param2 < 0 (a frozen fSim).
 * `group` is the index of the reference Gate object in PQC.gates (circuit.py:53-61)
 * this primitive came from: CHAIN / ALLTOALL / shared_parameter / RR_block expand to
 * several primitives sharing one group. */
typedef struct pqc_op {
  int32_t kind;
  int32_t q0;
  int32_t q1;      /* -1 for one-qubit ops */
  int32_t param;   /* -1: fixed */
  int32_t param2;  /* -1 unless FSIM */
  int32_t group;
  double scale;
  double offset;
} pqc_op;

/* A Pauli-string term  coef * prod_b P_b  with X on bits (xmask & ~zmask), Z on
 * (zmask & ~xmask), Y on (xmask & zmask).  Masks are in BASIS-INDEX bit positions
 * (bit b  <->  qubit n-1-b). */
typedef struct pqc_pauli_term {
  uint32_t xmask;
  uint32_t zmask;
  double re, im;
} pqc_pauli_term;

PQC_API const char* pqc_last_error(void);
PQC_API int pqc_abi_version(void);
/* Number of CUDA kernels this library has launched in the calling process. */
PQC_API long long pqc_launch_count(void);
/* Bracket a region in which every launch of the gate-apply tile kernel is timed with a
 * CUDA event pair on its own stream.  pqc_profile_end synchronises those events and
 * returns out[0] = summed kernel ms, out[1] = launches, out[2] = algorithmic bytes
 * (vectors x 2 x 16 x 2^n per launch), out[3] = reserved. */
PQC_API int pqc_profile_begin(void);
PQC_API int pqc_profile_end(double* out4);
/* Per-kernel split of the region closed by the last pqc_profile_end: out[3 k + {0, 1, 2}] =
 * summed ms, launches, algorithmic bytes of pass-kernel kind k = 0 k_apply_pass,
 * 1 k_sweep_pass, 2 k_layer_pass, 3 k_layer_seq, 4 k_tile_pipe. */
PQC_API int pqc_profile_kinds(double* out, int n_kinds);
/* Fails (<0) unless the current CUDA device is compute capability 10.x. */
PQC_API int pqc_device_check(int* cc_major, int* cc_minor, int* n_sms);

/* ---- gate programs: replaces PQC.set_gates / the per-gate 2^n x 2^n operator rebuild
 * (circuit.py:53-72, gates.py:122-131,469-477) ---------------------------------------- */
PQC_API int pqc_program_create(int n_qubits, int n_params, int n_ops, const pqc_op* h_ops,
                       pqc_program** out);   /* planning only: works without a GPU */
PQC_API int pqc_program_destroy(pqc_program* prog);
/* out[0]=n_qubits out[1]=n_params out[2]=n_ops out[3]=n_passes (forward plan)
 * out[4]=tile_bits out[5]=1 if derivative states are supported out[6]=passes of the
 * derivative / QFIM plan out[7]=reserved */
PQC_API int pqc_program_stats(const pqc_program* prog, int64_t* out8);

/* Text description of the execution plan (one line per pass / gather / dots stage); needs no
 * GPU.  Writes at most `cap` bytes including the terminator. */
PQC_API int pqc_program_describe(const pqc_program* prog, char* out, int64_t cap);

/* ---- PQC.run for a batch (circuit.py:118-125).  d_init: NULL = |0..0>; otherwise
 * init_stride = 0 broadcasts one [D] vector, init_stride = D gives one per sample
 * (the latter is `Gate * state`, gates.py:63-67).  d_out [S, D]. */
PQC_API int pqc_run_batch(const pqc_program* prog, const double* d_angles, int64_t ld_angles,
                  int64_t n_samples, const pqc_c128* d_init, int64_t init_stride,
                  pqc_c128* d_out, void* stream);

/* ---- PQC.get_gradients (circuit.py:149-192): d_out [S, P+1, D]; slot 0 = the final
 * state, slot 1+p = derivative state for parameter slot p, i.e. U_G..(D_p U_p)..U_1|init>
 * with D_p the sum of the Pauli generators of the gate's members (gates.py:133-138,
 * 454-457,519-522).  init_stride: 0 (one shared initial state) or 2^n (one per sample; used
 * when a circuit is run in segments around a dense ARBGATE). */
PQC_API int pqc_gradients_batch(const pqc_program* prog, const double* d_angles, int64_t ld_angles,
                        int64_t n_samples, const pqc_c128* d_init, int64_t init_stride,
                        pqc_c128* d_out, void* stream);

/* ---- Measurements.get_QFI (measure.py:33-71) from explicit states:
 * F[s,p,q] = 4 Re(<d_p|d_q> - conj<psi|d_p> <psi|d_q>).  d_states [S,D],
 * d_grads [S,P,D], d_qfim [S,P,P] float64. */
PQC_API int pqc_qfim_from_grads(const pqc_c128* d_states, const pqc_c128* d_grads, int n_qubits,
                        int n_params, int64_t n_samples, double* d_qfim, void* stream);

/* ---- fused batch path for update_state + get_QFI (circuit.py:127-130, measure.py:33-71)
 * that never materialises the P derivative states of a sample at once outside the
 * caller-provided workspace.  d_states_out may be NULL. */
PQC_API int pqc_qfim_workspace_bytes(const pqc_program* prog, int64_t n_samples, int64_t* bytes);
PQC_API int pqc_qfim_batch(const pqc_program* prog, const double* d_angles, int64_t ld_angles,
                   int64_t n_samples, const pqc_c128* d_init, void* d_work,
                   int64_t work_bytes, double* d_qfim, pqc_c128* d_states_out, void* stream);

/* ---- scipy.linalg.eigh eigenvalues (measure.py:73-75,84) for S symmetric PxP
 * matrices, ascending; and the `eigvals > cutoff` count (measure.py:85-86). */
PQC_API int pqc_eigvalsh_batch(const double* d_mats, int64_t n_mats, int dim, double* d_eig,
                       void* stream);
/* scipy.linalg.eigh with eigenvectors (measure.py:73-75): d_vecs [S,P,P] row-major,
 * column k = unit eigenvector of the k-th (ascending) eigenvalue; may be NULL. */
PQC_API int pqc_eigh_batch(const double* d_mats, int64_t n_mats, int dim, double* d_eig,
                   double* d_vecs, void* stream);
PQC_API int pqc_count_greater(const double* d_vals, int64_t rows, int cols, double cutoff,
                      int32_t* d_counts, void* stream);

/* ---- Measurements.single_Q / entanglement (measure.py:226-249): Q[s]. */
PQC_API int pqc_meyer_wallach(const pqc_c128* d_states, int64_t n_samples, int n_qubits, double* d_Q,
                      void* stream);
/* Qobj.ptrace(k) for one qubit (measure.py:233): d_rho = 2x2 row-major. */
PQC_API int pqc_ptrace_1q(const pqc_c128* d_state, int n_qubits, int qubit, pqc_c128* d_rho,
                  void* stream);
/* Qobj.overlap (measure.py:52,58,135; circuit.py:141): out[i] = <a_i|b_i>. */
PQC_API int pqc_overlap_batch(const pqc_c128* d_a, int64_t stride_a, const pqc_c128* d_b,
                      int64_t stride_b, int64_t dim, int64_t count, pqc_c128* d_out,
                      void* stream);

/* ---- _gen_f_samples + np.histogram (measure.py:123-159).  F = |<A_i|B_j>|^2.
 * triangular != 0: A == B block, only i < j (itertools.combinations).  Counts are ADDED
 * into d_hist[bins] (int64) with np.histogram(range=(0,1)) semantics.  d_F (optional):
 * triangular -> packed combinations order [SA(SA-1)/2]; else row-major [SA,SB]. */
PQC_API int pqc_fidelity_hist(const pqc_c128* d_A, int64_t n_a, const pqc_c128* d_B, int64_t n_b,
                      int n_qubits, int triangular, int64_t bins, long long* d_hist,
                      double* d_F, void* stream);
/* np.histogram(F, bins, range=(0,1)) for caller-supplied samples (expr(F_samples, N)). */
PQC_API int pqc_hist_f64(const double* d_F, int64_t count, int64_t bins, long long* d_hist,
                 void* stream);
/* Measurements.expr on histogram counts (measure.py:161-180): KL(P_pqc || P_haar(N)).
 * d_scratch: 4 doubles. */
PQC_API int pqc_kl_haar(const long long* d_hist, int64_t bins, double hilbert_dim, double* d_out,
                double* d_scratch, void* stream);

/* ---- renyi_entropy_fast / gkp_fast (measure.py:318-368) via Walsh-Hadamard over
 * Z-masks for every X-mask.  d_out [n_alpha, S]: 1/(1-a) ln(sum |2^{-n/2} W|^{2a}) - n ln 2. */
PQC_API int pqc_magic_batch(const pqc_c128* d_states, int64_t n_samples, int n_qubits, int n_alpha,
                    const double* h_alphas, double* d_out, void* stream);

/* ---- qt.expect(H, psi) with H a Pauli sum (circuit.py:28-31,132-137): out[s] = <psi_s|H|psi_s>
 * and H|psi> itself (measure.py:468). */
PQC_API int pqc_pauli_expect_batch(const pqc_c128* d_states, int64_t n_samples, int n_qubits,
                           int n_terms, const pqc_pauli_term* h_terms, pqc_c128* d_out,
                           void* stream);
PQC_API int pqc_pauli_apply_batch(const pqc_c128* d_states, int64_t n_samples, int n_qubits,
                          int n_terms, const pqc_pauli_term* h_terms, pqc_c128* d_out,
                          void* stream);

/* ---- ARBGATE (gates.py:407-435): exp(-i theta H) of a dense Hermitian H = V diag(lambda) V^dagger,
 * as products with the eigenvector matrix instead of a matrix exponential per angle:
 *   out[s][r] = sum_c M[r][c] * w(s, c) * in[s][c],   M row-major [2^n][2^n],
 *   w = 1 when d_lambda is NULL; else exp(-i theta_s lambda_c), times (-i lambda_c / 2) when
 *   `deriv` != 0 (the reference's derivative -i H / 2, gates.py:426-428); theta_s =
 *   d_theta[s * theta_stride].  The gate is the call with M = V^dagger followed by the call with
 *   M = V and the eigenvalues.  Out of place; n <= 13. */
PQC_API int pqc_dense_apply_batch(const pqc_c128* d_in, int64_t n_samples, int n_qubits,
                          const pqc_c128* d_M, const double* d_lambda, const double* d_theta,
                          int64_t theta_stride, int deriv, pqc_c128* d_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PQC_B200_H */
