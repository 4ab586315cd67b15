// pqc_common.cuh -- shared declarations of the B200 PQC engine (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/pqc_b200.h"

typedef double2 c128;

// ------------------------------------------------------------------ error plumbing
void pqc_set_error(const std::string& msg);
#define PQC_FAIL(code, msg)      \
  do {                           \
    pqc_set_error(msg);          \
    return (code);               \
  } while (0)
#define PQC_CUDA(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      pqc_set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));            \
      return -2;                                                                    \
    }                                                                               \
  } while (0)
extern long long g_pqc_launches;      // kernels launched by this library (pqc_launch_count)
#define PQC_LAUNCH_CHECK()                                                          \
  do {                                                                              \
    ++g_pqc_launches;                                                               \
    cudaError_t _e = cudaGetLastError();                                            \
    if (_e != cudaSuccess) {                                                        \
      pqc_set_error(std::string("kernel launch: ") + cudaGetErrorString(_e));       \
      return -2;                                                                    \
    }                                                                               \
  } while (0)

// ------------------------------------------------------------------ device op / pass
// Bit positions are BASIS-INDEX bits: bit b <-> qubit n-1-b.
struct DOp {
  int kind;            // pqc_opcode
  int b0, b1;          // global bit positions (b1 = -1 for 1q)
  int l0, l1;          // local (tile) bit positions, -1 if the bit is outside the tile
  int param, param2;   // angle slots
  int trig;            // slot in the per-item trig table, -1 if none
  double scale, offset;
};

#define PQC_MAX_TILE_BITS 13
struct PassArgs {
  c128* buf;                 // destination / in-place buffer
  const c128* init;          // source for init_mode 2/3
  long long init_stride;
  int init_mode;             // 0 in place, 1 |0..0>, 2 broadcast init, 3 per-sample init
  const double* angles;
  long long ld;
  const DOp* ops;
  int nops;
  int ntrig;
  int n;                     // qubits
  int T;                     // local index bits (amplitude bits + packed item bits)
  int tb;                    // amplitude bits in the tile = min(n, T)
  int items_log2;            // T - tb
  int low_run;               // lbit[j] == j for j < low_run
  int lbit[PQC_MAX_TILE_BITS];   // global bit of local amplitude bit j
  int obit[PQC_MAX_QUBITS];      // global bits outside the tile (n - tb of them)
  long long n_items;
  int slots_active, slots_total, slot_base;
};

struct Pass {
  int op_begin, op_end;      // range in the program's primitive op list
  int tb;                    // amplitude bits in tile
  int low_run;
  int lbit[PQC_MAX_TILE_BITS];
  int obit[PQC_MAX_QUBITS];
  int dev_off;               // offset of this pass' DOps in the device op array
  int nops, ntrig;
};

struct GenTerm {             // one Pauli term of a parameter's generator sum
  uint32_t xmask, zmask;
  double re, im;             // coefficient (already includes -i/2 * scale)
};

struct pqc_program {
  int n = 0, P = 0;
  std::vector<pqc_op> ops;
  int tile_bits = 12;
  // forward plan over the whole op list (PQC.run)
  std::vector<Pass> run_passes;
  // QFIM / gradient plan: one segment of passes per parameter, plus the trailing ops
  std::vector<std::vector<Pass>> seg_passes;   // size P + 1 (last = trailing ops)
  std::vector<int> gen_off;                    // P + 1 offsets into gens
  std::vector<GenTerm> gens;
  bool grad_supported = true;
  std::string grad_reason;
  DOp* d_ops = nullptr;                        // all passes' ops
  GenTerm* d_gens = nullptr;
};

// ------------------------------------------------------------------ small device helpers
__device__ __forceinline__ c128 cmul(c128 a, c128 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ c128 cmul_conj_a(c128 a, c128 b) {  // conj(a) * b
  return make_double2(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ c128 cadd(c128 a, c128 b) { return make_double2(a.x + b.x, a.y + b.y); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum of `v`; result valid in thread 0.  `red` needs >= 32 doubles.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < NT / 32) ? red[l] : 0.0;
    r = warp_sum(r);
  }
  return r;
}

// ------------------------------------------------------------------ cross-file host API
int pqc_plan_program(pqc_program* prog);
int pqc_pauli_apply_slots(const c128* src, c128* dst, int n, long long S, int slots_total,
                          int src_slot, int dst_slot, const GenTerm* d_terms, int nterms,
                          cudaStream_t st);
int pqc_qfim_finalize(const c128* d_G, long long S, int P, double* d_F, cudaStream_t st);
int pqc_launch_pass(const pqc_program* prog, const Pass& ps, c128* buf, const c128* init,
                    long long init_stride, int init_mode, const double* d_angles, long long ld,
                    long long n_items, int slots_active, int slots_total, int slot_base,
                    cudaStream_t st);
