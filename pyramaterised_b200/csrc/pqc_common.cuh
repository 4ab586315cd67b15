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

// ------------------------------------------------------------------ v1 (sweep engine)
// internal opcodes produced by the planner's fusion step
#define PQC_K_ZZSUM 32   // product of same-angle R_zz over a pair set: one table-lookup phase
#define PQC_K_RXY 33     // R_yy * R_xx on one pair, same angle: rotation of the odd-parity pair
#define PQC_K_GEN 34     // multiply by a diagonal generator sum (spawn items only)
#define PQC_K_LAYER_RX4 35    // rx-type gate on each of the 4 register bits (identity = c 1, s 0)
#define PQC_K_LAYER_REAL4 36  // real 2x2 on each register bit: RY, H or identity

#ifndef V1_LOCAL_BITS
#define V1_LOCAL_BITS 12          // 4096 amplitudes per CTA = 256 threads x 16 registers
#endif
#define V1_TBITS (V1_LOCAL_BITS - 4)   // thread bits: every thread holds 16 amplitudes
#define V1_NT (1 << V1_TBITS)          // threads per CTA of the sweep kernel
#define V1_MAX_SPAWN 32
#define V1_MAX_PART 64

struct MOp {               // one micro-op inside a sweep
  int kind;
  int k0, k1;              // register-bit index (0..3) if the bit is a register bit, else -1
  int l0, l1;              // local tile position of the bit, -1 if outside the tile
  int b0, b1;              // global bit positions
  int trig;                // first trig slot
  int aux0, aux1;          // ZZSUM / GEN: aux1 = offset of the op's linear-form table in `wtab`;
                           // GEN: aux0 = in-pass spawn index
  int npairs;              // ZZSUM: number of pairs;  GEN: number of generator terms
  int subk;                // LAYER_*4: gate on register bit K in byte K: 0 none, else opcode+1
  int subt[4];             // LAYER_*4: trig slot of the gate on register bit K
};
struct SweepD {
  int rb[4];               // local positions held in registers, ascending
  int mop_begin, mop_end;
  int io;                  // bit0: load straight from global, bit1: store straight to global
  int pad;                 // npre | npost << 16
  unsigned char tl[8];     // local position fed by thread bit t (the 8 non-register positions)
  unsigned char tg[8];     // global amplitude bit of that position, 255 = item bit (n < 12)
  unsigned short sz[4];    // swizzled shared-memory mask of register bit k
  unsigned gm[4];          // global amplitude-index mask of register bit k
  unsigned char gb[4];     // global bit number of register bit k
  // thread -> tile position / amplitude bits, split by nibble of tid so the kernel needs two
  // table reads instead of an 8-step bit deposit:
  unsigned short tpos[2][16];   // base(tid) = tpos[0][tid & 15] | tpos[1][tid >> 4]
  unsigned tamp[2][16];         // amp bits  = tamp[0][tid & 15] | tamp[1][tid >> 4]
};
struct TrigJob {           // per-item trig table entry (or run of entries) to fill
  int kind, param, param2, slot, npairs, pad;
  double scale, offset;
};
// Linear forms over GF(2) for ZZSUM / GEN ops: a table of V1_WTAB words per op; word b holds,
// in bit k, whether global index bit b takes part in term k (ZZSUM: pair k = {b0, b1}; GEN:
// z-mask of generator term k).  w(x) = XOR of the words of x's set bits, and the number of
// anti-aligned pairs (odd-parity terms) of basis state x is popc(w(x)).  Word 32 is zero.
#define V1_WTAB 33
#define V1_MAX_TERMS 32
#define V1_MAX_WT 8          // linear-form tables per pass (staged in shared memory)

// "Layer pass" fast path (k_layer_pass): a pass made of the aligned nibble sweeps
// (registers = tile positions 8-11, then 0-3, then 4-7; or 8-11 then 4-7), loading and storing
// global memory directly, whose ops are 1-qubit layer macro-ops, ZZSUM phases and in-pass
// generator multiplies.  Its whole description travels in the kernel arguments (constant
// bank): no staging of a micro-program, compile-time shared-memory addressing.
#define FAST_MAX_OPS 4
#define FAST_MAX_SPAWN 4
struct FastOp {
  int kind, subk;
  int t[4];                // LAYER_*4: trig slot per register bit;  ZZSUM: t[0] = first table slot
  int wt, nterms, spawn;   // ZZSUM / GEN: linear-form table index, term count; GEN: spawn index
};
#define FAST_MAX_WT 4      // linear-form tables per fast pass
struct FastPlan {
  int ns;                  // 3: sweeps A, B, C;  2: sweeps A, C;  1: sweep A only
  int nops[3];
  FastOp ops[3][FAST_MAX_OPS];
  // linear forms folded per tile nibble (positions 0-3, 4-7, 8-11) and per out-of-tile bit, so
  // the kernel's prologue copies them from the constant bank instead of chasing global tables
  uint32_t wn[FAST_MAX_WT][3][16];
  uint32_t wo[FAST_MAX_WT][PQC_MAX_QUBITS - 12];
};

// "Layer sequence" path (k_layer_seq): the same idea for passes that visit the aligned nibbles
// in ANY order and up to SEQ_MAX_SWEEPS times (the XXZ template's XY bonds alternate between
// nibbles), and whose ops may also be XY pair rotations (PQC_K_RXY) on two register bits.
#define SEQ_MAX_SWEEPS 32
#define SEQ_MAX_OPS 400              // per pass, all sweeps together
#define SEQ_MAX_SPAWN 8
struct SeqPlan {
  int nsw;
  int has_diag;                      // any R_z / CZ op: the kernel instantiation that folds them
  int geom[SEQ_MAX_SWEEPS];          // registers hold tile positions 0: 8-11, 1: 0-3, 2: 4-7
  int off[SEQ_MAX_SWEEPS], nops[SEQ_MAX_SWEEPS];   // the sweep's ops: ops[off .. off + nops)
  // RXY: subk = ka * 4 + kb (ka < kb), t[0] = trig slot
  // RZ:  t[0] = trig slot, t[1] = register bit or -1, t[2] = tile position or -1, t[3] = global bit
  // CZ:  t[0], t[1] = register bits or -1, t[2], t[3] = tile positions or -1, wt, nterms = global bits
  FastOp ops[SEQ_MAX_OPS];
  uint32_t wn[FAST_MAX_WT][3][16];
  uint32_t wo[FAST_MAX_WT][PQC_MAX_QUBITS - 12];
};

// "Tile pipe" (k_tile_pipe, pqc_pipe.cu): the persistent, async-copy-fed form of the pass kernels.
// One CTA per SM walks over (vector, tile) work items; tiles arrive in a ring of TP_NBUF
// shared-memory buffers by per-thread 16-byte async copies (cp.async, completion counted on one
// mbarrier per slot) issued ahead of the two consumer groups, so HBM loads never wait for a free
// register file.  A pass is a list of
// sweeps with RUNTIME geometry: any 4 of the 12 tile positions in registers, thread bits on the
// other 8, shared-memory slots an affine function of the logical tile index.  X / CNOT are pure
// relabelings of that index (they change the planner's slot tables, not the kernel).
#define TP_MAX_SWEEPS 24
#define TP_MAX_OPS 320
#define TP_MAX_TRIG 256
#define TP_MAX_SPAWN 8
#define TP_MAX_WT 24                 // linear-form tables (ZZSUM subsets, GEN) per pass
#define TP_NBUF 3
#define PQC_K_LAYER_RZ4 37           // rz on each of the 4 register bits, tangent form
#define PQC_K_LAYER_RY4 38           // ry on each register bit; a slot may absorb a fixed ry and a
                                     // CZ with a bit that is constant for the thread (see below)
#define PQC_K_CZF 39                 // CZ(register bit, thread-constant bit): a pending Z "frame"
#define PQC_K_ZFLUSH 40              // apply the pending Z frame now
#define PQC_K_RZZ1 41                // one R_zz on two register bits: a = ka * 4 + kb, t[0] = (cos, sin) of
                                     // the half angle; amplitude x (c - i s) where the bits agree, (c + i s) else
#define PQC_K_RZZ2 42                // two same-angle R_zz on disjoint register-bit pairs (all four bits):
                                     // a = 0 (01|23), 1 (02|13), 2 (03|12), t[0] = (cos, sin) of the FULL
                                     // angle; both pairs aligned: x (c - i s), both anti-aligned: x (c + i s)
// Z frame (k_tile_pipe): CZ between a register bit k and a bit that is constant for the thread is
// Z^b on bit k.  Z anticommutes with the X / Y rotations and commutes with everything diagonal,
// so instead of touching 16 amplitudes it flips a per-thread flag; later rx / ry / xy rotations
// on k take the opposite angle sign and the sign itself is applied once, when the sweep ends.
// A partner byte: 0xff none, else tile position (bit 7 clear) or 0x80 | amplitude bit.
struct TPOp {              // 16 bytes
  uint8_t kind;            // pqc_opcode or PQC_K_*
  uint8_t sub;             // LAYER_*4: 2 bits per register bit: 0 none, 1 rotation, 2 Hadamard
  uint8_t a, b;            // LAYER_RY4: partner bytes of slots 0, 1 (slots 2, 3: wt, nterms); a slot with
                           // a partner holds two trig entries, t[k] + (partner bit): ry(fixed) CZ ry(theta)
                           // CZF: a = register bit, b = partner byte
                           // RXY: a = ka * 4 + kb;  RZ: a = register bit | 0xff, b = tile position | 0xff
                           // CZ: a, b = register bits | 0xff
                           // CNOT (index permutation): a = control register bit | 0xff, b = target
                           // register bit;  X: b = target register bit
  uint16_t t[4];           // trig slots (LAYER_*4, RXY, RZ: t[0]; ZZSUM: t[0] = first table slot)
                           // RZ: t[1] = global bit;  CZ: t[0], t[1] = tile positions | 0xffff, t[2], t[3] = global bits
                           // CNOT: t[0] = control tile position | 0xffff, t[1] = control global bit
  uint8_t wt, nterms, spawn, pad;   // ZZSUM / GEN: linear-form table, term count; GEN: spawn index
                                    // LAYER_RY4: pad = 1 when some slot carries a partner byte
};
#define TP_MAX_INJ 6
// X / CNOT relabeling folded into a sweep's load or store, worked out by the planner: register
// index j lives at slot  base ^ XOR_{k in j} l[k] ^ XOR_i (control bit of injection i ? inj[i].lm : 0).
// An injection is a CNOT whose control bit is constant for the thread (a thread bit of the tile or
// a bit of the tile's offset); more than TP_MAX_INJ distinct controls: the planner gives the pass up.
struct TPAff {
  uint16_t l[4];
  uint16_t base;
  uint8_t ninj, pad;
  struct { uint8_t src, pad; uint16_t lm; } inj[TP_MAX_INJ];   // src: a partner byte
};
struct TPSweep {
  // thread part, per nibble of the thread index: byte offset of the slot (low 16 bits) and
  // logical tile index (high 16 bits); word = tt[0][tid & 15] ^ tt[1][tid >> 4]
  uint32_t tt[2][16];
  uint16_t rs[4];          // slot byte-offset masks of the 4 register bits
  uint8_t rpos[4];         // tile positions of the register bits
  uint16_t op_begin, op_end;
  // X / CNOT index permutations folded into the load (the first npre ops) and into the store
  // (the last npost ops) of the sweep: targets are register bits, so a thread only permutes its
  // own 16 slots
  uint8_t npre, npost, pad[2];
  TPAff pre, post;
};
struct PipePlan {
  int nsw, nops, ntrig, nwt;
  int lbit[12];            // amplitude bit of tile position p
  // tile load (cp.async): thread bits 0-3 sit on the tile positions of amplitude bits 0-3
  uint32_t ld_amp[2][16];  // amplitude offset of the thread part: ld_amp[0][tid & 15] | ld_amp[1][tid >> 4]
  uint32_t ld_r[4];        //   and of the 4 bits walked by the copy index j
  uint16_t ld_slot[2][16]; // slot byte offset of the thread part
  uint16_t ld_sr[4];
  uint32_t st_t[2][16];    // last sweep, direct store: amplitude offset of the thread part
  uint32_t st_r[4];        //   and of the register bits
  uint32_t st_q[4], st_base, st_gm[TP_MAX_INJ];   // the last sweep's store relabeling (TPAff) as amplitude masks
  uint32_t wn[TP_MAX_WT][3][16];
  uint32_t wo[TP_MAX_WT][PQC_MAX_QUBITS - 12];
  TPSweep sw[TP_MAX_SWEEPS];
  TPOp ops[TP_MAX_OPS];
};
struct PipeArgs {
  const c128* src;
  c128* dst;
  const double2* gtrig;
  int tstride, toff;
  int n;
  int obit[PQC_MAX_QUBITS];
  int slots_total, active, nspawn;
  int spawn_slot[TP_MAX_SPAWN];
  double spawn_cr[TP_MAX_SPAWN], spawn_ci[TP_MAX_SPAWN];
  long long n_items;       // vectors (samples x (active + nspawn))
  const PipePlan* plan;    // device copy
};

struct V1Pass {
  bool fast_ok = false;
  FastPlan fast;
  bool seq_ok = false;
  SeqPlan seq;
  int pipe_idx = -1;             // index into pqc_program::h_pipe, -1: not convertible
  bool front = false;            // made by the front planner: runs on k_tile_pipe only

  int tb, low_run;
  int lbit[V1_LOCAL_BITS];
  int obit[PQC_MAX_QUBITS];
  int sweep_off, nsweeps;  // into the program's device arrays
  int mop_off, nmops;
  int tj_off, ntjobs, ntrig;
  int trig_goff;           // first slot of this pass in its plan's per-sample trig table
  int wt_off, nwt;         // linear-form tables of this pass (V1_WTAB words each) in `wtab`
  // in-pass diagonal spawns, in order of their K_GEN micro-ops
  std::vector<int> spawn_param;
  std::vector<int> op_ids;       // program ops this pass executes
  bool direct_ok;
  int io_first, io_last;
};
struct V1Stage {
  int type;                      // 0 = pass, 1 = gather spawn
  int pass = -1;                 // index into v1_passes
  std::vector<int> gather_params;
  std::vector<int> partners;     // parameters whose Gram column is taken on this pass' load
};

// how the derivative vector of a parameter is created from psi
struct ParamSpawn {
  int type = 0;            // 0: sum of Pauli generators (gens); 1: fSim-family pair matrix
  int b0 = -1, b1 = -1;    // pair bits (type 1)
  int p_theta = -1, p_phi = -1;
  int which = 0;           // 1 d/dtheta of fSim, 2 d/dphi of fSim, 3 d/dtheta of fixed_fSim
  double offset = 0.0, phi_fixed = 0.0;
};

struct pqc_program {
  int n = 0, P = 0;
  int n_nodiff = 0;              // the last n_nodiff parameters only supply angles (no derivative)
  std::vector<ParamSpawn> pspawn;
  std::vector<pqc_op> ops;
  int tile_bits = 12;
  // ---- v1 plans
  bool v1_ok = false;            // run plan usable
  bool v1_grad_ok = false;       // derivative / QFIM plan usable
  std::vector<V1Pass> v1_passes;
  std::vector<int> v1_run;                   // pass indices of the plain run plan
  std::vector<V1Stage> v1_grad;              // stages of the derivative / QFIM plan
  std::vector<int> v1_gen_diag_off;          // per parameter: offset/count of diagonal terms
  // trig jobs / per-sample trig-table slots of the run plan and of the derivative plan
  int v1_run_tj0 = 0, v1_run_ntj = 0, v1_run_slots = 0;
  int v1_grad_tj0 = 0, v1_grad_ntj = 0, v1_grad_slots = 0;
  double2* d_trig = nullptr;                 // library-owned scratch: trig table [S][slots]
  size_t trig_cap = 0;                       // (one stream at a time per program)
  // host copies of the device arrays (uploaded lazily by pqc_program_upload)
  std::vector<MOp> h_mops;
  std::vector<SweepD> h_sweeps;
  std::vector<TrigJob> h_tjobs;
  std::vector<uint32_t> h_zz;                 // linear-form tables (V1_WTAB words each)
  std::vector<DOp> h_dops;
  std::vector<PipePlan> h_pipe;              // tile-pipe form of the passes that have one
  // run plan of the front planner (pqc_front.cu): pass indices, its trig jobs / table slots
  bool front_ok = false;
  std::vector<int> front_run;
  int front_tj0 = 0, front_ntj = 0, front_slots = 0;
  PipePlan* d_pipe = nullptr;
  bool uploaded = false;
  int device = -1;               // CUDA device holding the uploaded plan (and the trig scratch)
  MOp* d_mops = nullptr;
  SweepD* d_sweeps = nullptr;
  TrigJob* d_tjobs = nullptr;
  uint32_t* d_zz = nullptr;
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
  // ---- meet-in-the-middle QFIM plan (pqc_api.cu): the op list is cut at `bi_cut`; the
  // parameters before the cut are differentiated by the forward sub-program bi_F, those after
  // it by bi_B = the inverted tail run backwards from psi(T) (made by bi_M); every vector ends
  // at the cut, where the Gram matrix is taken.
  pqc_program* bi_F = nullptr;
  pqc_program* bi_M = nullptr;
  pqc_program* bi_B = nullptr;
  int bi_cut = -1, bi_PF = 0, bi_PB = 0;
  // F stops after its last spawn; the ops of its pure-propagation tail (bi_ntrail of them) are
  // run by M and undone at the end of B, whose bi_extra extra parameters are angle-only
  // copies of the F parameters those ops use
  int bi_ntrail = 0, bi_extra = 0;
  std::vector<int> bi_cols;                    // Gram column v (F params then B params) -> parameter
  std::vector<int> bi_inv;                     // parameter -> v | (sign bit: backward vector)
  long long bi_cost = 0, fwd_cost = 0;         // vector-passes of either plan
  int* d_bi_cols = nullptr;
  int* d_bi_inv = nullptr;
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
int pqc_plan_v1(pqc_program* prog);
int pqc_program_upload(const pqc_program* prog);   // idempotent; needs a CUDA device
int pqc_v1_run(const pqc_program* prog, const double* d_angles, long long ld, long long S,
               const c128* d_init, long long init_stride, c128* d_out, cudaStream_t st);
int pqc_v1_derivatives(const pqc_program* prog, const double* d_angles, long long ld, long long S,
                       const c128* d_init, long long init_stride, c128* buf_a, c128* buf_b,
                       c128* d_gpart, bool want_dots, bool need_final, c128** final_buf,
                       cudaStream_t st);
long long pqc_v1_plan_cost(const pqc_program* prog);   // vector-passes of the QFIM plan
std::vector<int> pqc_v1_trailing_ops(const pqc_program* prog);
int pqc_v1_n_passes(const pqc_program* prog, bool need_final);
long long pqc_v1_gpart_elems(const pqc_program* prog, long long S);
int pqc_v1_qfim_reduce(const pqc_program* prog, const c128* d_gpart, long long S, double* d_F,
                       cudaStream_t st);
bool pqc_use_v0();
bool pqc_v1_gram_ok(const pqc_program* prog);
int pqc_v1_gram_qfim(const pqc_program* prog, const c128* buf, long long S, c128* d_gpart,
                     double* d_F, cudaStream_t st);
int pqc_v1_gram_qfim2(const pqc_program* prog, const c128* buf, int slots1, int M1,
                      const c128* buf2, int slots2, const int* d_inv, long long S, c128* d_gpart,
                      double* d_F, cudaStream_t st);
bool pqc_pipe_build(const V1Pass& ps, int n, PipePlan& out);
void pqc_pipe_fill_load_tables(PipePlan& pp);
bool pqc_plan_front(pqc_program* prog, std::vector<TrigJob>& tjobs);
bool pqc_use_front(const pqc_program* prog);
int pqc_pipe_launch(const PipeArgs& a, const PipePlan& hplan, cudaStream_t st);
bool pqc_pipe_enabled();
// Per-device one-time setup (the shared-memory opt-in of a kernel is a per-device attribute):
// first() is true the first time it is asked on the current device.
struct PqcDeviceOnce {
  bool done[64] = {};
  bool first() {
    int d = 0;
    if (cudaGetDevice(&d) != cudaSuccess || d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};
// pass-kernel kinds reported separately by pqc_profile_end_kinds
enum { PQC_PROF_APPLY_PASS = 0, PQC_PROF_SWEEP_PASS = 1, PQC_PROF_LAYER_PASS = 2,
       PQC_PROF_LAYER_SEQ = 3, PQC_PROF_TILE_PIPE = 4, PQC_PROF_KINDS = 5 };
int pqc_prof_launch_begin(double bytes, cudaStream_t st, int kind);
void pqc_prof_launch_end(int h, cudaStream_t st);
int pqc_pauli_apply_slots(const c128* src, c128* dst, int n, long long S, int slots_total,
                          int src_slot, int dst_slot, const GenTerm* d_terms, int nterms,
                          cudaStream_t st);
int pqc_qfim_finalize(const c128* d_G, long long S, int P, double* d_F, cudaStream_t st);
int pqc_pair_spawn(c128* buf, int n, long long S, int slots_total, int dst_slot,
                   const ParamSpawn& ps, const double* d_angles, long long ld, cudaStream_t st);
int pqc_launch_pass(const pqc_program* prog, const Pass& ps, c128* buf, const c128* init,
                    long long init_stride, int init_mode, const double* d_angles, long long ld,
                    long long n_items, int slots_active, int slots_total, int slot_base,
                    cudaStream_t st);
