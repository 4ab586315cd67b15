// pqc_apply.cu -- pass planner and the shared-memory tile gate-application kernel.
//
// Replaces `for g in self.gates: circuit_state = g * circuit_state`
// (/root/reference/pyramaterised/circuit.py:123-124) and the per-gate 2^n x 2^n operator
// construction behind it (gates.py:39-46,122-131).  A *pass* stages a tile of 2^T
// amplitudes in shared memory (T tile bits chosen by the planner, always containing
// the low index bits so global accesses are 256 B-contiguous), applies every
// primitive op whose pair-mixing bits are inside the tile, and writes the tile back:
// one HBM read + one write of the state per pass instead of one per gate.
#include <algorithm>
#include <cstring>

#include "pqc_common.cuh"

// ---------------------------------------------------------------------------------
// planner (host)
// ---------------------------------------------------------------------------------
static inline bool op_has_trig(int kind) {
  switch (kind) {
    case PQC_OP_RX: case PQC_OP_RY: case PQC_OP_RZ: case PQC_OP_RXX: case PQC_OP_RYY:
    case PQC_OP_RZZ: case PQC_OP_FSIM: case PQC_OP_FIXED_FSIM:
      return true;
    default:
      return false;
  }
}

// bits whose amplitude pairs mix (must be inside the tile); controls / diagonal bits may
// live anywhere because their value is constant over a tile.
static inline uint32_t op_mix_mask(const pqc_op& op, int n) {
  const int b0 = n - 1 - op.q0;
  const int b1 = op.q1 >= 0 ? n - 1 - op.q1 : -1;
  switch (op.kind) {
    case PQC_OP_RX: case PQC_OP_RY: case PQC_OP_H: case PQC_OP_X:
      return 1u << b0;
    case PQC_OP_CNOT:
      return 1u << b1;
    case PQC_OP_SQRTISWAP: case PQC_OP_RXX: case PQC_OP_RYY: case PQC_OP_FSIM:
    case PQC_OP_FIXED_FSIM:
      return (1u << b0) | (1u << b1);
    default:
      return 0u;
  }
}

static void close_pass(const pqc_program* prog, uint32_t mask, int begin, int end,
                       std::vector<Pass>& out, std::vector<DOp>& dops) {
  const int n = prog->n;
  const int tb = std::min(n, prog->tile_bits);
  for (int b = 0; b < n && __builtin_popcount(mask) < tb; ++b) mask |= 1u << b;
  Pass ps;
  memset(&ps, 0, sizeof(ps));
  ps.op_begin = begin;
  ps.op_end = end;
  ps.tb = tb;
  int j = 0, o = 0;
  int local_of[32];
  for (int b = 0; b < n; ++b) {
    if (mask >> b & 1u) {
      local_of[b] = j;
      ps.lbit[j++] = b;
    } else {
      local_of[b] = -1;
      ps.obit[o++] = b;
    }
  }
  ps.low_run = 0;
  while (ps.low_run < tb && ps.lbit[ps.low_run] == ps.low_run) ps.low_run++;
  ps.dev_off = (int)dops.size();
  int ntrig = 0;
  for (int i = begin; i < end; ++i) {
    const pqc_op& op = prog->ops[i];
    DOp d;
    d.kind = op.kind;
    d.b0 = n - 1 - op.q0;
    d.b1 = op.q1 >= 0 ? n - 1 - op.q1 : -1;
    d.l0 = local_of[d.b0];
    d.l1 = d.b1 >= 0 ? local_of[d.b1] : -1;
    d.param = op.param;
    d.param2 = op.param2;
    d.scale = op.scale;
    d.offset = op.offset;
    d.trig = op_has_trig(op.kind) ? ntrig++ : -1;
    dops.push_back(d);
  }
  ps.nops = end - begin;
  ps.ntrig = ntrig;
  out.push_back(ps);
}

static void plan_range(const pqc_program* prog, int begin, int end, std::vector<Pass>& out,
                       std::vector<DOp>& dops) {
  if (begin >= end) return;
  const int n = prog->n, T = prog->tile_bits;
  const int tb = std::min(n, T);
  const int items = 1 << (T - tb);
  const int max_trig = std::max(16, 1024 / items);
  const uint32_t low = n <= T ? ((n >= 32 ? 0xffffffffu : (1u << n) - 1u))
                              : ((1u << 4) - 1u);   // keep >= 256 B contiguous runs
  uint32_t mask = low;
  int start = begin, ntrig = 0;
  for (int i = begin; i < end; ++i) {
    const uint32_t need = op_mix_mask(prog->ops[i], n);
    const int t = op_has_trig(prog->ops[i].kind) ? 1 : 0;
    if (__builtin_popcount(mask | need) > tb || ntrig + t > max_trig) {
      close_pass(prog, mask, start, i, out, dops);
      start = i;
      mask = low;
      ntrig = 0;
    }
    mask |= need;
    ntrig += t;
  }
  close_pass(prog, mask, start, end, out, dops);
}

static bool generator_terms(const pqc_op& op, int n, std::vector<GenTerm>& out) {
  const uint32_t m0 = 1u << (n - 1 - op.q0);
  const uint32_t m1 = op.q1 >= 0 ? 1u << (n - 1 - op.q1) : 0u;
  GenTerm t;
  t.re = 0.0;
  t.im = -0.5 * op.scale;      // d/dtheta of exp(-i (scale*theta+offset) P / 2) = -i scale/2 P U
  switch (op.kind) {
    case PQC_OP_RX: t.xmask = m0; t.zmask = 0; break;
    case PQC_OP_RY: t.xmask = m0; t.zmask = m0; break;
    case PQC_OP_RZ: t.xmask = 0; t.zmask = m0; break;
    case PQC_OP_RXX: t.xmask = m0 | m1; t.zmask = 0; break;
    case PQC_OP_RYY: t.xmask = m0 | m1; t.zmask = m0 | m1; break;
    case PQC_OP_RZZ: t.xmask = 0; t.zmask = m0 | m1; break;
    case PQC_OP_IDENT: t.xmask = 0; t.zmask = 0; break;   // -i/2 * identity (quirk Q5)
    default: return false;
  }
  out.push_back(t);
  return true;
}

int pqc_plan_program(pqc_program* prog) {
  const int n = prog->n;
  prog->tile_bits = n >= 12 ? 12 : (n >= 8 ? n : 8);
  std::vector<DOp> dops;
  const int nops = (int)prog->ops.size();
  if (nops == 0)   // identity circuit: one empty pass still copies the initial state out
    close_pass(prog, 0u, 0, 0, prog->run_passes, dops);
  else
    plan_range(prog, 0, nops, prog->run_passes, dops);

  // parameter segments: segment p = ops after the end of parameter p-1's gate up to and
  // including the last op carrying parameter slot p; the generator of slot p is the sum
  // of its members' Pauli generators (gates.py:133-138,454-457,519-522).
  prog->seg_passes.assign(prog->P + 1, std::vector<Pass>());
  prog->gen_off.assign(prog->P + 1, 0);
  std::vector<int> last(prog->P, -1);
  for (int i = 0; i < nops; ++i) {
    const pqc_op& op = prog->ops[i];
    if (op.param >= 0) {
      if (op.param >= prog->P) PQC_FAIL(-1, "op parameter slot out of range");
      last[op.param] = i;
    }
    if (op.param2 >= 0) {
      if (op.param2 >= prog->P) PQC_FAIL(-1, "op parameter slot out of range");
      last[op.param2] = i;
    }
  }
  int prev = 0;
  prog->pspawn.assign(prog->P, ParamSpawn());
  for (int p = 0; p < prog->P; ++p) {
    prog->gen_off[p] = (int)prog->gens.size();
    if (last[p] < 0 || last[p] + 1 < prev) {
      prog->grad_supported = false;
      prog->grad_reason = "parameter slots are not used in gate order";
      break;
    }
    // every op carrying slot p (both slots of an fSim sit on the same op, so the segment of
    // its second slot is empty and the op is found behind `prev`)
    for (int i = std::min(prev, last[p]); i <= last[p]; ++i) {
      const pqc_op& op = prog->ops[i];
      if (op.param != p && op.param2 != p) continue;
      if (op.kind == PQC_OP_FSIM || op.kind == PQC_OP_FIXED_FSIM) {
        // the reference's "derivative" matrices of gates.py:609-648,719-737 (quirk Q3)
        ParamSpawn& ps = prog->pspawn[p];
        ps.type = 1;
        ps.b0 = n - 1 - op.q0;
        ps.b1 = n - 1 - op.q1;
        ps.p_theta = op.param;
        ps.p_phi = op.param2;
        ps.offset = op.offset;
        ps.phi_fixed = op.scale;
        ps.which = op.kind == PQC_OP_FIXED_FSIM ? 3 : (op.param == p ? 1 : 2);
      } else if (!generator_terms(op, n, prog->gens)) {
        prog->grad_supported = false;
        prog->grad_reason = "no derivative rule for this gate kind";
      }
    }
    plan_range(prog, prev, last[p] + 1, prog->seg_passes[p], dops);
    prev = std::max(prev, last[p] + 1);
  }
  prog->gen_off[prog->P] = (int)prog->gens.size();
  if (prog->grad_supported) plan_range(prog, prev, nops, prog->seg_passes[prog->P], dops);

  prog->h_dops = dops;
  return 0;
}

// Copy the plan to the device once (planning itself needs no GPU, so the planner can be
// exercised by the CPU test-suite).
int pqc_program_upload(const pqc_program* cprog) {
  pqc_program* prog = const_cast<pqc_program*>(cprog);
  int dev = 0;
  PQC_CUDA(cudaGetDevice(&dev));
  if (prog->uploaded) {
    // the plan's device arrays live on the GPU that was current at the first use
    if (dev != prog->device)
      PQC_FAIL(-1, "this program was uploaded on another CUDA device; create one program per device");
    return 0;
  }
  prog->device = dev;
  auto up = [&](auto& vec, auto** dptr) -> int {
    if (vec.empty()) return 0;
    PQC_CUDA(cudaMalloc(dptr, vec.size() * sizeof(vec[0])));
    PQC_CUDA(cudaMemcpy(*dptr, vec.data(), vec.size() * sizeof(vec[0]), cudaMemcpyHostToDevice));
    return 0;
  };
  if (up(prog->h_dops, &prog->d_ops) || up(prog->gens, &prog->d_gens) ||
      up(prog->h_mops, &prog->d_mops) || up(prog->h_sweeps, &prog->d_sweeps) ||
      up(prog->h_tjobs, &prog->d_tjobs) || up(prog->h_zz, &prog->d_zz) ||
      up(prog->h_pipe, &prog->d_pipe))
    return -2;
  prog->uploaded = true;
  return 0;
}

// ---------------------------------------------------------------------------------
// tile kernel
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t local_to_amp(uint32_t i, const PassArgs& a) {
  uint32_t r = i & ((1u << a.low_run) - 1u);
  for (int j = a.low_run; j < a.tb; ++j) r |= ((i >> j) & 1u) << a.lbit[j];
  return r;
}

__device__ __forceinline__ uint32_t insert_zero(uint32_t p, int l) {
  return ((p >> l) << (l + 1)) | (p & ((1u << l) - 1u));
}

template <int NT>
__global__ void __launch_bounds__(NT) k_apply_pass(const PassArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  c128* sm = reinterpret_cast<c128*>(smraw);
  double2* trig = reinterpret_cast<double2*>(sm + (1u << a.T));
  const int tid = threadIdx.x;
  const int tiles_log2 = a.n - a.tb;
  const long long blk = blockIdx.x;
  const long long group = blk >> tiles_log2;
  const uint32_t tile = (uint32_t)(blk & ((1ll << tiles_log2) - 1));
  uint32_t base = 0;
  for (int j = 0; j < tiles_log2; ++j) base |= ((tile >> j) & 1u) << a.obit[j];
  const int items = 1 << a.items_log2;
  const long long item0 = group << a.items_log2;
  const uint32_t tsize = 1u << a.T;
  const uint32_t amask = (1u << a.tb) - 1u;

  // ---- per-item trig table: (cos, sin) of the half angle (full angle for fSim) -------
  for (int e = tid; e < items * a.nops; e += NT) {
    const int li = e / a.nops, j = e - li * a.nops;
    const DOp op = a.ops[j];
    if (op.trig < 0) continue;
    const long long item = item0 + li;
    if (item >= a.n_items) continue;
    const long long sample = item / a.slots_active;
    double th = op.offset;
    double s, c;
    double2* t = trig + ((size_t)li * a.ntrig + op.trig) * 2;
    if (op.kind == PQC_OP_FSIM || op.kind == PQC_OP_FIXED_FSIM) {
      // fSim family: theta = offset + angles[param] (full angle, no scale);
      // phi = angles[param2], or the `scale` field when frozen (param2 < 0)
      if (op.param >= 0) th += a.angles[sample * a.ld + op.param];
      sincos(th, &s, &c);
      t[0] = make_double2(c, s);
      const double ph = op.param2 >= 0 ? a.angles[sample * a.ld + op.param2] : op.scale;
      sincos(ph, &s, &c);
      t[1] = make_double2(c, s);
    } else {
      if (op.param >= 0) th += op.scale * a.angles[sample * a.ld + op.param];
      sincos(0.5 * th, &s, &c);
      t[0] = make_double2(c, s);
    }
  }

  // ---- load tile -----------------------------------------------------------------------
  for (uint32_t i = tid; i < tsize; i += NT) {
    const long long item = item0 + (i >> a.tb);
    c128 v = make_double2(0.0, 0.0);
    if (item < a.n_items) {
      const uint32_t amp = base | local_to_amp(i & amask, a);
      const long long sample = item / a.slots_active;
      const long long slot = a.slot_base + item % a.slots_active;
      if (a.init_mode == 0) {
        v = a.buf[((sample * a.slots_total + slot) << a.n) + amp];
      } else if (a.init_mode == 1) {
        v.x = amp == 0 ? 1.0 : 0.0;
      } else if (a.init_mode == 2) {
        v = a.init[amp];
      } else {
        v = a.init[sample * a.init_stride + amp];
      }
    }
    sm[i] = v;
  }
  __syncthreads();

  // ---- ops -------------------------------------------------------------------------------
  const uint32_t npairs = tsize >> 1;
  for (int j = 0; j < a.nops; ++j) {
    const DOp op = a.ops[j];
    const int kind = op.kind;
    if (kind == PQC_OP_IDENT) continue;
    const int tslot = op.trig;
    if (kind == PQC_OP_RX || kind == PQC_OP_RY || kind == PQC_OP_H || kind == PQC_OP_X) {
      const int l = op.l0;
      for (uint32_t p = tid; p < npairs; p += NT) {
        const uint32_t i0 = insert_zero(p, l), i1 = i0 | (1u << l);
        const c128 x = sm[i0], y = sm[i1];
        c128 nx, ny;
        if (kind == PQC_OP_RX) {
          const double2 cs = trig[((size_t)(i0 >> a.tb) * a.ntrig + tslot) * 2];
          nx = make_double2(cs.x * x.x + cs.y * y.y, cs.x * x.y - cs.y * y.x);
          ny = make_double2(cs.x * y.x + cs.y * x.y, cs.x * y.y - cs.y * x.x);
        } else if (kind == PQC_OP_RY) {
          const double2 cs = trig[((size_t)(i0 >> a.tb) * a.ntrig + tslot) * 2];
          nx = make_double2(cs.x * x.x - cs.y * y.x, cs.x * x.y - cs.y * y.y);
          ny = make_double2(cs.y * x.x + cs.x * y.x, cs.y * x.y + cs.x * y.y);
        } else if (kind == PQC_OP_H) {
          const double r = 0.70710678118654752440;
          nx = make_double2(r * (x.x + y.x), r * (x.y + y.y));
          ny = make_double2(r * (x.x - y.x), r * (x.y - y.y));
        } else {
          nx = y;
          ny = x;
        }
        sm[i0] = nx;
        sm[i1] = ny;
      }
    } else if (kind == PQC_OP_RZ || kind == PQC_OP_S || kind == PQC_OP_T) {
      const int l = op.l0;
      const int ext = l < 0 ? (int)((base >> op.b0) & 1u) : 0;
      for (uint32_t i = tid; i < tsize; i += NT) {
        const int bit = l >= 0 ? (int)((i >> l) & 1u) : ext;
        c128 v = sm[i];
        if (kind == PQC_OP_RZ) {
          const double2 cs = trig[((size_t)(i >> a.tb) * a.ntrig + tslot) * 2];
          const double s = bit ? cs.y : -cs.y;            // e^{-+ i a/2}
          v = make_double2(v.x * cs.x - v.y * s, v.y * cs.x + v.x * s);
        } else if (bit) {
          if (kind == PQC_OP_S) {
            v = make_double2(-v.y, v.x);
          } else {
            const double r = 0.70710678118654752440;
            v = make_double2(r * (v.x - v.y), r * (v.x + v.y));
          }
        }
        sm[i] = v;
      }
    } else if (kind == PQC_OP_CNOT) {
      const int lc = op.l0, lt = op.l1;
      const int ext = lc < 0 ? (int)((base >> op.b0) & 1u) : 1;
      if (ext) {
        for (uint32_t p = tid; p < npairs; p += NT) {
          const uint32_t i0 = insert_zero(p, lt), i1 = i0 | (1u << lt);
          if (lc >= 0 && !((i0 >> lc) & 1u)) continue;
          const c128 x = sm[i0];
          sm[i0] = sm[i1];
          sm[i1] = x;
        }
      }
    } else if (kind == PQC_OP_CZ || kind == PQC_OP_RZZ) {
      const int la = op.l0, lb = op.l1;
      const int ea = la < 0 ? (int)((base >> op.b0) & 1u) : 0;
      const int eb = lb < 0 ? (int)((base >> op.b1) & 1u) : 0;
      for (uint32_t i = tid; i < tsize; i += NT) {
        const int ba = la >= 0 ? (int)((i >> la) & 1u) : ea;
        const int bb = lb >= 0 ? (int)((i >> lb) & 1u) : eb;
        c128 v = sm[i];
        if (kind == PQC_OP_CZ) {
          if (ba & bb) v = make_double2(-v.x, -v.y);
        } else {
          const double2 cs = trig[((size_t)(i >> a.tb) * a.ntrig + tslot) * 2];
          const double s = (ba ^ bb) ? cs.y : -cs.y;      // e^{-i a/2 z0 z1}
          v = make_double2(v.x * cs.x - v.y * s, v.y * cs.x + v.x * s);
        }
        sm[i] = v;
      }
    } else {
      // two-qubit mixing ops: both bits are local (planner guarantee)
      const int la = op.l0, lb = op.l1;      // la <-> q0 (first listed qubit), lb <-> q1
      const int lo = la < lb ? la : lb, hi = la < lb ? lb : la;
      const uint32_t nquads = tsize >> 2;
      for (uint32_t p = tid; p < nquads; p += NT) {
        const uint32_t i00 = insert_zero(insert_zero(p, lo), hi);
        const uint32_t ia = 1u << la, ib = 1u << lb;
        const size_t tb_ = ((size_t)(i00 >> a.tb) * a.ntrig + tslot) * 2;
        if (kind == PQC_OP_RXX || kind == PQC_OP_RYY) {
          const double2 cs = trig[tb_];
          // (P P psi)[x] = sgn(x) psi[x ^ m]; XX: sgn = +1; YY: -1 when the two bits are equal
          const double se = kind == PQC_OP_RYY ? -cs.y : cs.y;   // equal bits (00 <-> 11)
          const double sd = cs.y;                                 // differing bits (01 <-> 10)
          c128 x = sm[i00], y = sm[i00 | ia | ib];
          sm[i00] = make_double2(cs.x * x.x + se * y.y, cs.x * x.y - se * y.x);
          sm[i00 | ia | ib] = make_double2(cs.x * y.x + se * x.y, cs.x * y.y - se * x.x);
          x = sm[i00 | ia];
          y = sm[i00 | ib];
          sm[i00 | ia] = make_double2(cs.x * x.x + sd * y.y, cs.x * x.y - sd * y.x);
          sm[i00 | ib] = make_double2(cs.x * y.x + sd * x.y, cs.x * y.y - sd * x.x);
        } else if (kind == PQC_OP_SQRTISWAP) {
          const double r = 0.70710678118654752440;
          const c128 x = sm[i00 | ia], y = sm[i00 | ib];      // |10>, |01> in (q0,q1) order
          // [[r, i r],[i r, r]] on the odd-parity subspace (symmetric in the two states)
          sm[i00 | ia] = make_double2(r * (x.x - y.y), r * (x.y + y.x));
          sm[i00 | ib] = make_double2(r * (y.x - x.y), r * (y.y + x.x));
        } else {   // FSIM / FIXED_FSIM
          const double2 cs = trig[tb_];
          const c128 x = sm[i00 | ia], y = sm[i00 | ib];
          sm[i00 | ia] = make_double2(cs.x * x.x + cs.y * y.y, cs.x * x.y - cs.y * y.x);
          sm[i00 | ib] = make_double2(cs.x * y.x + cs.y * x.y, cs.x * y.y - cs.y * x.x);
          if (kind == PQC_OP_FSIM) {
            const double2 ph = trig[tb_ + 1];                 // e^{-i phi} on |11>
            const c128 z = sm[i00 | ia | ib];
            sm[i00 | ia | ib] = make_double2(z.x * ph.x + z.y * ph.y, z.y * ph.x - z.x * ph.y);
          }
        }
      }
    }
    __syncthreads();
  }

  // ---- store tile ----------------------------------------------------------------------
  for (uint32_t i = tid; i < tsize; i += NT) {
    const long long item = item0 + (i >> a.tb);
    if (item < a.n_items) {
      const uint32_t amp = base | local_to_amp(i & amask, a);
      const long long sample = item / a.slots_active;
      const long long slot = a.slot_base + item % a.slots_active;
      a.buf[((sample * a.slots_total + slot) << a.n) + amp] = sm[i];
    }
  }
}

// ---- optional per-launch timing of the gate-apply kernel (bench.py roofline) ------------
struct ProfRec { cudaEvent_t e0, e1; double bytes; int kind; };
static double g_prof_kinds[PQC_PROF_KINDS][3];   // of the last pqc_profile_end: ms, launches, bytes
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;

static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

extern "C" int pqc_profile_begin(void) {
  for (auto& r : g_prof) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
  g_prof.clear();
  g_prof_on = true;
  return 0;
}

extern "C" int pqc_profile_end(double* out4) {
  g_prof_on = false;
  double ms = 0.0, bytes = 0.0;
  memset(g_prof_kinds, 0, sizeof(g_prof_kinds));
  for (auto& r : g_prof) {
    PQC_CUDA(cudaEventSynchronize(r.e1));
    float t = 0.f;
    PQC_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    ms += t;
    bytes += r.bytes;
    const int k = r.kind >= 0 && r.kind < PQC_PROF_KINDS ? r.kind : 0;
    g_prof_kinds[k][0] += t;
    g_prof_kinds[k][1] += 1.0;
    g_prof_kinds[k][2] += r.bytes;
  }
  if (out4) {
    out4[0] = ms;
    out4[1] = (double)g_prof.size();
    out4[2] = bytes;
    out4[3] = 0.0;
  }
  for (auto& r : g_prof) { g_prof_pool.push_back(r.e0); g_prof_pool.push_back(r.e1); }
  g_prof.clear();
  return 0;
}

// the per-kernel split of the region closed by the last pqc_profile_end:
// out[3 * kind + {0, 1, 2}] = ms, launches, algorithmic bytes
extern "C" int pqc_profile_kinds(double* out, int n_kinds) {
  if (!out) return -1;
  for (int k = 0; k < n_kinds; ++k)
    for (int c = 0; c < 3; ++c) out[3 * k + c] = k < PQC_PROF_KINDS ? g_prof_kinds[k][c] : 0.0;
  return 0;
}

int pqc_prof_launch_begin(double bytes, cudaStream_t st, int kind) {
  if (!g_prof_on) return -1;
  ProfRec rec;
  rec.kind = kind;
  rec.e0 = prof_event();
  rec.e1 = prof_event();
  rec.bytes = bytes;
  cudaEventRecord(rec.e0, st);
  g_prof.push_back(rec);
  return (int)g_prof.size() - 1;
}

void pqc_prof_launch_end(int h, cudaStream_t st) {
  if (h >= 0) cudaEventRecord(g_prof[h].e1, st);
}

template <int NT>
static int launch_nt(const PassArgs& a, long long grid, size_t smem, cudaStream_t st) {
  static PqcDeviceOnce attr_once;
  if (attr_once.first()) {
    PQC_CUDA(cudaFuncSetAttribute(k_apply_pass<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  200 * 1024));
  }
  // algorithmic traffic of one pass: every vector read once and written once
  const int h = pqc_prof_launch_begin(
      (double)a.n_items * 2.0 * sizeof(c128) * (double)(1ll << a.n), st, PQC_PROF_APPLY_PASS);
  k_apply_pass<NT><<<(unsigned)grid, NT, smem, st>>>(a);
  pqc_prof_launch_end(h, st);
  PQC_LAUNCH_CHECK();
  return 0;
}

int pqc_launch_pass(const pqc_program* prog, const Pass& ps, c128* buf, const c128* init,
                    long long init_stride, int init_mode, const double* d_angles, long long ld,
                    long long n_items, int slots_active, int slots_total, int slot_base,
                    cudaStream_t st) {
  if (n_items <= 0) return 0;
  PassArgs a;
  memset(&a, 0, sizeof(a));
  a.buf = buf;
  a.init = init;
  a.init_stride = init_stride;
  a.init_mode = init_mode;
  a.angles = d_angles;
  a.ld = ld;
  a.ops = prog->d_ops + ps.dev_off;
  a.nops = ps.nops;
  a.ntrig = ps.ntrig;
  a.n = prog->n;
  a.T = prog->tile_bits;
  a.tb = ps.tb;
  a.items_log2 = a.T - a.tb;
  a.low_run = ps.low_run;
  memcpy(a.lbit, ps.lbit, sizeof(a.lbit));
  memcpy(a.obit, ps.obit, sizeof(a.obit));
  a.n_items = n_items;
  a.slots_active = slots_active;
  a.slots_total = slots_total;
  a.slot_base = slot_base;
  const int items = 1 << a.items_log2;
  const long long groups = (n_items + items - 1) / items;
  const long long grid = groups << (a.n - a.tb);
  if (grid > 0x7fffffffLL) PQC_FAIL(-1, "pass grid too large; split the batch");
  const size_t smem = ((size_t)1 << a.T) * sizeof(c128) +
                      (size_t)items * std::max(1, a.ntrig) * 2 * sizeof(double2);
  const int half = 1 << (a.T - 1);
  if (half >= 256) return launch_nt<256>(a, grid, smem, st);
  if (half >= 128) return launch_nt<128>(a, grid, smem, st);
  return launch_nt<64>(a, grid, smem, st);
}
