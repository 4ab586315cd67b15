// pqc_api.cu -- C-ABI entry points for programs, state generation, derivative states and
// the fused QFIM pipeline (include/pqc_b200.h).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pqc_common.cuh"

static thread_local std::string g_last_error;
void pqc_set_error(const std::string& msg) { g_last_error = msg; }

long long g_pqc_launches = 0;

extern "C" const char* pqc_last_error(void) { return g_last_error.c_str(); }
extern "C" long long pqc_launch_count(void) { return g_pqc_launches; }
extern "C" int pqc_abi_version(void) { return PQC_ABI_VERSION; }

extern "C" int pqc_device_check(int* cc_major, int* cc_minor, int* n_sms) {
  int dev = 0;
  PQC_CUDA(cudaGetDevice(&dev));
  int maj = 0, min = 0, sms = 0;
  PQC_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  PQC_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  PQC_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (n_sms) *n_sms = sms;
  if (maj != 10) PQC_FAIL(-3, "this library is built for sm_100a (B200) only");
  return 0;
}

// ---------------------------------------------------------------------------------
// programs
// ---------------------------------------------------------------------------------
static int plan_bidir(pqc_program* p);

static int program_create(int n_qubits, int n_params, int n_ops, const pqc_op* h_ops,
                          bool with_bidir, pqc_program** out, int n_nodiff = 0) {
  if (!out) PQC_FAIL(-1, "null output handle");
  *out = nullptr;
  if (n_qubits < 1 || n_qubits > PQC_MAX_QUBITS) PQC_FAIL(-1, "n_qubits must be in [1, 30]");
  if (n_params < 0 || n_ops < 0) PQC_FAIL(-1, "negative counts");
  for (int i = 0; i < n_ops; ++i) {
    const pqc_op& op = h_ops[i];
    if (op.kind < 0 || op.kind >= PQC_OP__COUNT) PQC_FAIL(-1, "unknown opcode");
    const bool two = op.kind == PQC_OP_CNOT || op.kind == PQC_OP_CZ ||
                     op.kind == PQC_OP_SQRTISWAP || op.kind == PQC_OP_RXX ||
                     op.kind == PQC_OP_RYY || op.kind == PQC_OP_RZZ || op.kind == PQC_OP_FSIM ||
                     op.kind == PQC_OP_FIXED_FSIM;
    if (op.q0 < 0 || op.q0 >= n_qubits) PQC_FAIL(-1, "qubit index out of range");
    if (two && (op.q1 < 0 || op.q1 >= n_qubits || op.q1 == op.q0))
      PQC_FAIL(-1, "two-qubit op needs two distinct qubits in range");
    if (op.param >= n_params || op.param2 >= n_params) PQC_FAIL(-1, "parameter slot out of range");
  }
  pqc_program* p = new pqc_program();
  p->n = n_qubits;
  p->P = n_params;
  p->n_nodiff = n_nodiff;
  p->ops.assign(h_ops, h_ops + n_ops);
  for (auto& op : p->ops) {
    const bool two = op.kind == PQC_OP_CNOT || op.kind == PQC_OP_CZ ||
                     op.kind == PQC_OP_SQRTISWAP || op.kind == PQC_OP_RXX ||
                     op.kind == PQC_OP_RYY || op.kind == PQC_OP_RZZ || op.kind == PQC_OP_FSIM ||
                     op.kind == PQC_OP_FIXED_FSIM;
    if (!two) op.q1 = -1;
  }
  int rc = pqc_plan_program(p);
  if (rc == 0) rc = pqc_plan_v1(p);
  if (rc == 0 && with_bidir) rc = plan_bidir(p);
  if (rc) {
    pqc_program_destroy(p);
    return rc;
  }
  *out = p;
  return 0;
}

extern "C" int pqc_program_create(int n_qubits, int n_params, int n_ops, const pqc_op* h_ops,
                                  pqc_program** out) {
  return program_create(n_qubits, n_params, n_ops, h_ops, true, out);
}

// ---------------------------------------------------------------------------------
// Meet-in-the-middle QFIM plan.  Gram entries <d_j|d_p> are invariant under a unitary applied
// to both vectors, so they can be taken at ANY common time.  Cutting the circuit at op `c`:
//   F: ops[0,c) forward from |init>, spawning the derivative vectors of its parameters;
//   M: ops[c,T) forward on psi alone -> psi(T);
//   B: ops[c,T) inverted and reversed, from psi(T) back to time c, spawning -d_p on the way
//      (the inverse of exp(-i a G/2) is the same gate at -a, so its generator flips sign).
// A vector then lives for |t_p - c| passes instead of T - t_p: about half the vector-passes of
// the forward-only pipeline.  Only for circuits whose every op is inverted by negating its
// angle, and whose parameters / reference gates do not straddle the cut.
// ---------------------------------------------------------------------------------
static bool bidir_enabled() {
  const char* e = getenv("PQC_BIDIR");
  return !(e && strcmp(e, "0") == 0);
}

static bool trailing_enabled() {             // PQC_BIDIR_TRAIL=0: F runs on to the cut
  const char* e = getenv("PQC_BIDIR_TRAIL");
  return !(e && strcmp(e, "0") == 0);
}

static bool op_negatable(int kind) {
  switch (kind) {
    case PQC_OP_RX: case PQC_OP_RY: case PQC_OP_RZ: case PQC_OP_RXX: case PQC_OP_RYY:
    case PQC_OP_RZZ: case PQC_OP_H: case PQC_OP_X: case PQC_OP_CNOT: case PQC_OP_CZ:
    case PQC_OP_IDENT:
      return true;
    default:
      return false;
  }
}

static void bidir_free(pqc_program* p) {
  if (p->bi_F) pqc_program_destroy(p->bi_F);
  if (p->bi_M) pqc_program_destroy(p->bi_M);
  if (p->bi_B) pqc_program_destroy(p->bi_B);
  p->bi_F = p->bi_M = p->bi_B = nullptr;
  p->bi_cut = -1;
}

// build the three sub-programs for a cut; returns false if one of them is unusable
static bool bidir_build(pqc_program* p, int cut, const std::vector<int>& first,
                        const std::vector<int>& last) {
  const int nops = (int)p->ops.size(), P = p->P;
  std::vector<int> loc(P, -1);
  std::vector<int> cols;
  std::vector<pqc_op> f_ops(p->ops.begin(), p->ops.begin() + cut), b_ops;
  int PF = 0, PB = 0;
  // parameters that never appear in an op go to the forward side (their vectors are zero)
  for (int q = 0; q < P; ++q)
    if (first[q] < 0) { loc[q] = -2; }
  for (auto& op : f_ops)
    if (op.param >= 0) {
      if (loc[op.param] < 0) { loc[op.param] = PF++; cols.push_back(op.param); }
      op.param = loc[op.param];
    }
  for (int q = 0; q < P; ++q)
    if (loc[q] == -2) { loc[q] = PF++; cols.push_back(q); }
  for (int i = nops - 1; i >= cut; --i) {
    pqc_op op = p->ops[i];
    op.scale = -op.scale;
    op.offset = -op.offset;
    if (op.param >= 0) {
      if (loc[op.param] < 0) { loc[op.param] = PB++; cols.push_back(op.param); }
      op.param = loc[op.param];
    }
    b_ops.push_back(op);
  }
  if (PF + PB != P || PB == 0 || PF == 0) return false;
  bidir_free(p);
  if (program_create(p->n, PF, (int)f_ops.size(), f_ops.data(), false, &p->bi_F)) return false;
  if (!p->bi_F->v1_grad_ok || !p->bi_F->grad_supported || p->bi_F->v1_grad.empty() ||
      p->bi_F->v1_grad[0].type != 0)
    return false;
  // F stops after its last spawn.  The ops of its pure-propagation tail (the closing pass that
  // would finish e.g. the last R_x layer on every live vector) go to the front of M and, inverted,
  // to the end of B, where they merge into passes that exist anyway.  B reads their angles from
  // angle-only copies of the F parameters (appended columns).
  std::vector<int> trail = trailing_enabled() ? pqc_v1_trailing_ops(p->bi_F) : std::vector<int>();
  std::vector<pqc_op> m_ops;
  for (int i : trail) m_ops.push_back(p->ops[i]);
  m_ops.insert(m_ops.end(), p->ops.begin() + cut, p->ops.end());
  std::vector<int> extra_loc(P, -1);
  int n_extra = 0;
  for (int k = (int)trail.size() - 1; k >= 0; --k) {
    pqc_op op = p->ops[trail[k]];
    op.scale = -op.scale;
    op.offset = -op.offset;
    if (op.param2 >= 0) return false;
    if (op.param >= 0) {
      if (extra_loc[op.param] < 0) { extra_loc[op.param] = PB + n_extra++; cols.push_back(op.param); }
      op.param = extra_loc[op.param];
    }
    b_ops.push_back(op);
  }
  if (program_create(p->n, P, (int)m_ops.size(), m_ops.data(), false, &p->bi_M)) return false;
  if (program_create(p->n, PB + n_extra, (int)b_ops.size(), b_ops.data(), false, &p->bi_B, n_extra))
    return false;
  if (!p->bi_B->v1_grad_ok || !p->bi_B->grad_supported || p->bi_B->v1_grad.empty() ||
      p->bi_B->v1_grad[0].type != 0)
    return false;
  if (!p->bi_M->v1_ok) return false;
  p->bi_cut = cut;
  p->bi_PF = PF;
  p->bi_PB = PB;
  p->bi_ntrail = (int)trail.size();
  p->bi_extra = n_extra;
  p->bi_cols = cols;
  p->bi_inv.assign(P, 0);
  for (int v = 0; v < P; ++v) p->bi_inv[cols[v]] = v | (v >= PF ? (int)0x80000000 : 0);
  // B runs to its end (need_final): count all of its passes
  long long bcost = pqc_v1_plan_cost(p->bi_B);
  p->bi_cost = pqc_v1_plan_cost(p->bi_F) + bcost + (long long)p->bi_M->v1_run.size() + 2;
  (void)last;
  return true;
}

static int plan_bidir(pqc_program* p) {
  const int nops = (int)p->ops.size(), P = p->P;
  if (!bidir_enabled() || !p->v1_grad_ok || !p->grad_supported || P < 4 || p->n < 10) return 0;
  for (const pqc_op& op : p->ops)
    if (!op_negatable(op.kind)) return 0;
  p->fwd_cost = pqc_v1_plan_cost(p);
  std::vector<int> first(P, -1), last(P, -1);
  for (int i = 0; i < nops; ++i) {
    const int q = p->ops[i].param;
    if (q < 0) continue;
    if (first[q] < 0) first[q] = i;
    last[q] = i;
  }
  // valid cuts: between two reference gates, no parameter on both sides
  std::vector<std::pair<long long, int>> cand;
  for (int c = 1; c < nops; ++c) {
    if (p->ops[c].group == p->ops[c - 1].group) continue;
    bool ok = true;
    long long cost = nops - c;
    for (int q = 0; q < P && ok; ++q) {
      if (first[q] < 0) continue;
      if (first[q] < c && last[q] >= c) ok = false;
      else if (last[q] < c) cost += c - first[q];
      else cost += last[q] + 1 - c;
    }
    if (ok) cand.push_back({cost, c});
  }
  for (int q = 0; q < P; ++q)
    if (first[q] < 0) return 0;   // a parameter no op uses: keep the plain pipeline
  std::sort(cand.begin(), cand.end());
  // plan the few best cuts of the op-count proxy and keep the cheapest real plan
  int best_cut = -1;
  long long best = p->fwd_cost;
  for (size_t k = 0; k < cand.size() && k < 6; ++k) {
    if (!bidir_build(p, cand[k].second, first, last)) continue;
    if (p->bi_cost < best) { best = p->bi_cost; best_cut = cand[k].second; }
  }
  if (best_cut < 0 || best * 10 > p->fwd_cost * 9) {   // < 10 % gain: not worth the extra launches
    bidir_free(p);
    return 0;
  }
  if (p->bi_cut != best_cut && !bidir_build(p, best_cut, first, last)) bidir_free(p);
  return 0;
}

extern "C" int pqc_program_destroy(pqc_program* prog) {
  if (!prog) return 0;
  bidir_free(prog);
  if (prog->d_bi_cols) cudaFree(prog->d_bi_cols);
  if (prog->d_bi_inv) cudaFree(prog->d_bi_inv);
  if (prog->d_ops) cudaFree(prog->d_ops);
  if (prog->d_gens) cudaFree(prog->d_gens);
  if (prog->d_mops) cudaFree(prog->d_mops);
  if (prog->d_sweeps) cudaFree(prog->d_sweeps);
  if (prog->d_tjobs) cudaFree(prog->d_tjobs);
  if (prog->d_zz) cudaFree(prog->d_zz);
  if (prog->d_pipe) cudaFree(prog->d_pipe);
  if (prog->d_trig) cudaFree(prog->d_trig);
  delete prog;
  return 0;
}

extern "C" int pqc_program_stats(const pqc_program* prog, int64_t* out8) {
  if (!prog || !out8) PQC_FAIL(-1, "null argument");
  out8[0] = prog->n;
  out8[1] = prog->P;
  out8[2] = (int64_t)prog->ops.size();
  const bool v1r = prog->v1_ok && !pqc_use_v0();
  const bool v1g = prog->v1_grad_ok && !pqc_use_v0();
  out8[3] = v1r ? (int64_t)(pqc_use_front(prog) ? prog->front_run.size() : prog->v1_run.size())
                : (int64_t)prog->run_passes.size();
  out8[4] = v1r ? V1_LOCAL_BITS : prog->tile_bits;
  out8[5] = prog->grad_supported ? 1 : 0;
  int64_t q = 0;
  if (v1g) q = pqc_v1_n_passes(prog, false);
  else for (auto& v : prog->seg_passes) q += (int64_t)v.size();
  out8[6] = q;
  out8[7] = (v1r ? 1 : 0) | (v1g ? 2 : 0);
  return 0;
}

// ---------------------------------------------------------------------------------
// PQC.run (circuit.py:118-125)
// ---------------------------------------------------------------------------------
static int init_mode_of(const pqc_c128* d_init, int64_t init_stride) {
  if (!d_init) return 1;
  return init_stride == 0 ? 2 : 3;
}

extern "C" int pqc_run_batch(const pqc_program* prog, const double* d_angles, int64_t ld,
                             int64_t S, const pqc_c128* d_init, int64_t init_stride,
                             pqc_c128* d_out, void* stream) {
  if (!prog || !d_out) PQC_FAIL(-1, "null argument");
  if (S <= 0) return 0;
  if (prog->P > 0 && (!d_angles || ld < prog->P)) PQC_FAIL(-1, "No parameters supplied!");
  cudaStream_t st = (cudaStream_t)stream;
  if (pqc_program_upload(prog)) return -2;
  if (prog->v1_ok && !pqc_use_v0())
    return pqc_v1_run(prog, d_angles, ld, S, (const c128*)d_init, init_stride, (c128*)d_out, st);
  int mode = init_mode_of(d_init, init_stride);
  for (const Pass& ps : prog->run_passes) {
    const int rc = pqc_launch_pass(prog, ps, (c128*)d_out, (const c128*)d_init, init_stride, mode,
                                   d_angles, ld, S, 1, 1, 0, st);
    if (rc) return rc;
    mode = 0;
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// PQC.get_gradients (circuit.py:149-192), forward form: the state and every already
// spawned derivative vector advance together; derivative p is spawned as
// (sum of generators) * psi right after parameter p's gate (gates.py:133-138,454-457).
// d_out [S][P+1][D]: slot 0 = final state, slot 1+p = derivative state p.
// ---------------------------------------------------------------------------------
static int forward_derivatives(const pqc_program* prog, const double* d_angles, int64_t ld,
                               int64_t S, const pqc_c128* d_init, int64_t init_stride, c128* buf,
                               c128* d_G, bool trailing_all, bool trailing_state_only,
                               cudaStream_t st);

extern "C" int pqc_gradients_batch(const pqc_program* prog, const double* d_angles, int64_t ld,
                                   int64_t S, const pqc_c128* d_init, int64_t init_stride,
                                   pqc_c128* d_out, void* stream) {
  if (!prog || !d_out) PQC_FAIL(-1, "null argument");
  if (S <= 0) return 0;
  if (!prog->grad_supported) PQC_FAIL(-4, "derivative states unsupported: " + prog->grad_reason);
  if (prog->P > 0 && (!d_angles || ld < prog->P)) PQC_FAIL(-1, "No parameters supplied!");
  if (init_stride != 0 && init_stride != ((int64_t)1 << prog->n))
    PQC_FAIL(-1, "initial states: one shared state (stride 0) or one per sample (stride 2^n)");
  cudaStream_t st = (cudaStream_t)stream;
  if (pqc_program_upload(prog)) return -2;
  if (prog->v1_grad_ok && !pqc_use_v0()) {
    // second ping-pong copy; ordered so that the last pass lands in the caller's buffer
    const size_t bytes = sizeof(c128) * (size_t)S * (prog->P + 1) * ((size_t)1 << prog->n);
    c128* scratch = nullptr;
    PQC_CUDA(cudaMallocAsync(&scratch, bytes, st));
    const bool even = (pqc_v1_n_passes(prog, true) % 2) == 0;
    c128* fin = nullptr;
    int rc = pqc_v1_derivatives(prog, d_angles, ld, S, (const c128*)d_init, init_stride,
                                even ? (c128*)d_out : scratch, even ? scratch : (c128*)d_out,
                                nullptr, false, true, &fin, st);
    if (rc == 0 && fin != (c128*)d_out)
      if (cudaMemcpyAsync(d_out, fin, bytes, cudaMemcpyDeviceToDevice, st) != cudaSuccess) rc = -2;
    cudaFreeAsync(scratch, st);
    return rc;
  }
  return forward_derivatives(prog, d_angles, ld, S, d_init, init_stride, (c128*)d_out, nullptr,
                             true, false, st);
}

// <slot j | slot k> for j = 0..k-1... written into G[(s*(P+1)+j)*P + p]; one CTA per (s, j)
__global__ void __launch_bounds__(256) k_slot_dots(const c128* __restrict__ buf, int n,
                                                   int slots_total, int new_slot, int P, int p,
                                                   c128* __restrict__ G) {
  __shared__ double red[32];
  const long long D = 1ll << n;
  const int rows = new_slot + 1;
  const long long s = blockIdx.x / rows;
  const int j = blockIdx.x % rows;
  const c128* a = buf + ((s * slots_total + j) << n);
  const c128* b = buf + ((s * slots_total + new_slot) << n);
  double re = 0.0, im = 0.0;
  for (long long i = threadIdx.x; i < D; i += 256) {
    const c128 x = a[i], y = b[i];
    re += x.x * y.x + x.y * y.y;
    im += x.x * y.y - x.y * y.x;
  }
  re = block_sum<256>(re, red);
  im = block_sum<256>(im, red);
  if (threadIdx.x == 0) G[(s * (P + 1) + j) * P + p] = make_double2(re, im);
}

static int forward_derivatives(const pqc_program* prog, const double* d_angles, int64_t ld,
                               int64_t S, const pqc_c128* d_init, int64_t init_stride, c128* buf,
                               c128* d_G, bool trailing_all, bool trailing_state_only,
                               cudaStream_t st) {
  const int P = prog->P, n = prog->n;
  const int slots_total = P + 1;
  int mode = init_mode_of(d_init, init_stride);
  if (P == 0) {
    for (const Pass& ps : prog->run_passes) {
      const int rc = pqc_launch_pass(prog, ps, buf, (const c128*)d_init, init_stride, mode,
                                     d_angles, ld, S, 1, 1, 0, st);
      if (rc) return rc;
      mode = 0;
    }
    return 0;
  }
  for (int p = 0; p < P; ++p) {
    for (const Pass& ps : prog->seg_passes[p]) {
      const int active = mode != 0 ? 1 : p + 1;
      int rc = pqc_launch_pass(prog, ps, buf, (const c128*)d_init, init_stride, mode, d_angles,
                               ld, S * active, active, slots_total, 0, st);
      if (rc) return rc;
      mode = 0;
    }
    const int g0 = prog->gen_off[p], g1 = prog->gen_off[p + 1];
    int rc = prog->pspawn[p].type == 1
                 ? pqc_pair_spawn(buf, n, S, slots_total, p + 1, prog->pspawn[p], d_angles, ld, st)
                 : pqc_pauli_apply_slots(buf, buf, n, S, slots_total, 0, p + 1,
                                         prog->d_gens + g0, g1 - g0, st);
    if (rc) return rc;
    if (d_G) {
      const long long grid = S * (p + 2);
      if (grid > 0x7fffffffLL) PQC_FAIL(-1, "dot grid too large");
      k_slot_dots<<<(unsigned)grid, 256, 0, st>>>(buf, n, slots_total, p + 1, P, p, d_G);
      PQC_LAUNCH_CHECK();
    }
  }
  if (trailing_all || trailing_state_only) {
    const int active = trailing_all ? P + 1 : 1;
    for (const Pass& ps : prog->seg_passes[P]) {
      const int rc = pqc_launch_pass(prog, ps, buf, nullptr, 0, 0, d_angles, ld, S * active,
                                     active, slots_total, 0, st);
      if (rc) return rc;
    }
  }
  return 0;
}

// ---------------------------------------------------------------------------------
// fused update_state + get_QFI over a batch (circuit.py:127-130, measure.py:33-71).
// Overlaps are taken at spawn time: every later gate is a unitary applied to both
// vectors, so <d_p|d_q> and <psi|d_p> do not change afterwards.
// ---------------------------------------------------------------------------------
static bool use_bidir(const pqc_program* prog) {
  return prog->bi_cut >= 0 && prog->v1_grad_ok && !pqc_use_v0() && pqc_v1_gram_ok(prog) &&
         bidir_enabled();
}

static int64_t qfim_bytes_per_sample(const pqc_program* prog) {
  const int64_t D = 1ll << prog->n;
  if (use_bidir(prog))   // ping-pong pairs of both pipelines, psi(T), permuted angles, Gram
    return (2 * (int64_t)(prog->P + prog->bi_extra + 2) + 1) * D * (int64_t)sizeof(c128) +
           pqc_v1_gpart_elems(prog, 1) * (int64_t)sizeof(c128) +
           (int64_t)(((prog->P + prog->bi_extra) * sizeof(double) + 255) & ~(size_t)255);
  if (prog->v1_grad_ok && !pqc_use_v0())   // two ping-pong copies + per-tile Gram partials
    return 2 * (int64_t)(prog->P + 1) * D * (int64_t)sizeof(c128) +
           pqc_v1_gpart_elems(prog, 1) * (int64_t)sizeof(c128);
  return (int64_t)(prog->P + 1) * D * (int64_t)sizeof(c128) +
         (int64_t)(prog->P + 1) * std::max(1, prog->P) * (int64_t)sizeof(c128);
}

// samples processed together: small enough that the freshly spawned vectors (the Gram
// partners every CTA re-reads) stay L2 resident, large enough to fill 148 SMs
static int64_t qfim_chunk_target() {
  static int64_t v = -1;
  if (v < 0) {
    const char* e = getenv("PQC_QFIM_CHUNK");
    v = e ? atoll(e) : 256;
    if (v < 1) v = 256;
  }
  return v;
}

// out[s][v] = angles[s][cols[v]]: the sub-programs number their parameters in spawn order
__global__ void k_permute_cols(const double* __restrict__ angles, long long ld, long long S, int P,
                               const int* __restrict__ cols, double* __restrict__ out) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S * P) return;
  out[e] = angles[(e / P) * ld + cols[e % P]];
}

static int qfim_bidir(const pqc_program* prog, const double* d_angles, int64_t ld, int64_t S,
                      const c128* d_init, c128* work, int64_t C, double* d_qfim,
                      c128* d_states_out, cudaStream_t st) {
  const int P = prog->P, n = prog->n, PF = prog->bi_PF, PB = prog->bi_PB;
  const int64_t D = 1ll << n;
  if (!prog->d_bi_cols) {
    pqc_program* mp = const_cast<pqc_program*>(prog);
    PQC_CUDA(cudaMalloc(&mp->d_bi_cols, sizeof(int) * prog->bi_cols.size()));
    PQC_CUDA(cudaMalloc(&mp->d_bi_inv, sizeof(int) * P));
    PQC_CUDA(cudaMemcpy(mp->d_bi_cols, prog->bi_cols.data(), sizeof(int) * prog->bi_cols.size(),
                        cudaMemcpyHostToDevice));
    PQC_CUDA(cudaMemcpy(mp->d_bi_inv, prog->bi_inv.data(), sizeof(int) * P, cudaMemcpyHostToDevice));
  }
  const int PA = P + prog->bi_extra;           // columns of the permuted angle array
  const int PBx = PB + prog->bi_extra;         // B's parameters incl. the angle-only ones
  for (int64_t c0 = 0; c0 < S; c0 += C) {
    const int64_t c = std::min<int64_t>(C, S - c0);
    c128* fa = work;
    c128* fb = fa + c * (int64_t)(PF + 1) * D;
    c128* ba = fb + c * (int64_t)(PF + 1) * D;
    c128* bb = ba + c * (int64_t)(PBx + 1) * D;
    c128* psiT = bb + c * (int64_t)(PBx + 1) * D;
    c128* G = psiT + c * D;
    double* ang = (double*)(G + pqc_v1_gpart_elems(prog, c));
    const double* a0 = d_angles + c0 * ld;
    {
      const long long tot = c * PA;
      k_permute_cols<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(a0, ld, c, PA, prog->d_bi_cols, ang);
      PQC_LAUNCH_CHECK();
    }
    c128 *ffin = nullptr, *bfin = nullptr;
    // F stops after its last spawn when its tail was handed to M / B (bi_ntrail > 0); otherwise
    // psi must reach the cut even if the last spawn is earlier
    int rc = pqc_v1_derivatives(prog->bi_F, ang, PA, c, d_init, 0, fa, fb, nullptr, false,
                                prog->bi_ntrail == 0, &ffin, st);
    if (rc) return rc;
    rc = pqc_v1_run(prog->bi_M, a0, ld, c, ffin, (int64_t)(PF + 1) * D, psiT, st);
    if (rc) return rc;
    rc = pqc_v1_derivatives(prog->bi_B, ang + PF, PA, c, psiT, D, ba, bb, nullptr, false, true,
                            &bfin, st);
    if (rc) return rc;
    rc = pqc_v1_gram_qfim2(prog, ffin, PF + 1, PF + 1, bfin, PBx + 1, prog->d_bi_inv, c, G,
                           d_qfim + c0 * (int64_t)P * P, st);
    if (rc) return rc;
    if (d_states_out)
      PQC_CUDA(cudaMemcpyAsync(d_states_out + c0 * D, psiT, sizeof(c128) * c * D,
                               cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

extern "C" int pqc_qfim_workspace_bytes(const pqc_program* prog, int64_t S, int64_t* bytes) {
  if (!prog || !bytes) PQC_FAIL(-1, "null argument");
  *bytes = qfim_bytes_per_sample(prog) * std::max<int64_t>(1, S) + 256;
  return 0;
}

extern "C" int pqc_qfim_batch(const pqc_program* prog, const double* d_angles, int64_t ld,
                              int64_t S, const pqc_c128* d_init, void* d_work, int64_t work_bytes,
                              double* d_qfim, pqc_c128* d_states_out, void* stream) {
  if (!prog || !d_work || !d_qfim) PQC_FAIL(-1, "null argument");
  if (S <= 0) return 0;
  if (!prog->grad_supported) PQC_FAIL(-4, "QFIM unsupported: " + prog->grad_reason);
  if (prog->P > 0 && (!d_angles || ld < prog->P)) PQC_FAIL(-1, "No parameters supplied!");
  cudaStream_t st = (cudaStream_t)stream;
  if (pqc_program_upload(prog)) return -2;
  const int P = prog->P, n = prog->n;
  const int64_t D = 1ll << n;
  const int64_t per = qfim_bytes_per_sample(prog);
  // 256 B aligned start
  uintptr_t w0 = ((uintptr_t)d_work + 255) & ~(uintptr_t)255;
  const int64_t usable = work_bytes - (int64_t)(w0 - (uintptr_t)d_work);
  int64_t C = usable / per;
  if (C < 1) PQC_FAIL(-1, "QFIM workspace too small for one sample");
  C = std::min<int64_t>(C, S);
  if (use_bidir(prog)) {
    C = std::min<int64_t>(C, qfim_chunk_target());
    return qfim_bidir(prog, d_angles, ld, S, (const c128*)d_init, (c128*)w0, C, d_qfim,
                      (c128*)d_states_out, st);
  }
  if (prog->v1_grad_ok && !pqc_use_v0()) {
    C = std::min<int64_t>(C, qfim_chunk_target());
    for (int64_t c0 = 0; c0 < S; c0 += C) {
      const int64_t c = std::min<int64_t>(C, S - c0);
      c128* buf_a = (c128*)w0;
      c128* buf_b = buf_a + c * (int64_t)(P + 1) * D;
      c128* G = buf_b + c * (int64_t)(P + 1) * D;
      c128* fin = nullptr;
      // Gram mode: run the pipeline without in-pass Gram partials, then ONE batched V^H V
      // on the FP64 tensor cores at the common time all vectors have reached
      const bool gram = pqc_v1_gram_ok(prog);
      int rc = pqc_v1_derivatives(prog, d_angles + c0 * ld, ld, c, (const c128*)d_init, 0, buf_a,
                                  buf_b, G, !gram, d_states_out != nullptr, &fin, st);
      if (rc) return rc;
      rc = gram ? pqc_v1_gram_qfim(prog, fin, c, G, d_qfim + c0 * (int64_t)P * P, st)
                : pqc_v1_qfim_reduce(prog, G, c, d_qfim + c0 * (int64_t)P * P, st);
      if (rc) return rc;
      if (d_states_out) {
        PQC_CUDA(cudaMemcpy2DAsync((c128*)d_states_out + c0 * D, D * sizeof(c128), fin,
                                   (size_t)(P + 1) * D * sizeof(c128), D * sizeof(c128), c,
                                   cudaMemcpyDeviceToDevice, st));
      }
    }
    return 0;
  }
  for (int64_t c0 = 0; c0 < S; c0 += C) {
    const int64_t c = std::min<int64_t>(C, S - c0);
    c128* buf = (c128*)w0;
    c128* G = buf + c * (int64_t)(P + 1) * D;
    int rc = forward_derivatives(prog, d_angles + c0 * ld, ld, c, d_init, 0, buf, G, false,
                                 d_states_out != nullptr, st);
    if (rc) return rc;
    rc = pqc_qfim_finalize(G, c, P, d_qfim + c0 * (int64_t)P * P, st);
    if (rc) return rc;
    if (d_states_out) {
      PQC_CUDA(cudaMemcpy2DAsync((c128*)d_states_out + c0 * D, D * sizeof(c128), buf,
                                 (size_t)(P + 1) * D * sizeof(c128), D * sizeof(c128), c,
                                 cudaMemcpyDeviceToDevice, st));
    }
  }
  return 0;
}
