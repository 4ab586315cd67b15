// pqc_pipe.cu -- k_tile_pipe: the persistent, async-copy-fed pass kernel (gate application of
// circuit.py:118-125 and the derivative passes of circuit.py:149-192 for n >= 12 qubits).
//
// Why it exists.  k_layer_pass / k_layer_seq load a 4096-amplitude tile from HBM straight into
// registers, so the bytes a CTA can have in flight are bounded by its register file and an SM only
// has loads outstanding while one of its two CTAs sits in its load phase: the 3-sweep TFIM pass
// stalls at 68 % of the measured copy bandwidth although DRAM, FP64 and shared memory each have
// slack.  Here ONE CTA per SM stays resident and walks over (vector, tile) work items:
//   * tiles land in a ring of three 64 KB shared-memory buffers through asynchronous 16-byte
//     copies (cp.async / LDGSTS, completion counted on an mbarrier per ring slot) issued by the
//     group that has just freed the slot -- one whole tile per SM is in flight at all times and
//     registers hold only the sweep being computed.  (The TMA bulk form, one cp.async.bulk per
//     contiguous run, was measured first: a tile is 16 - 256 runs, every UBLKCP is issued by one
//     elected lane at a time, and the issuing warp became the critical path -- 2.1 TB/s,
//     profiles/r2_pipe_tma_vs_ldgsts.md);
//   * two consumer groups of 256 threads work on different tiles, so one group's FP64 phase
//     overlaps the other's shared-memory transposition;
//   * the per-pass tables (sweep geometry, ops, linear forms) are staged ONCE per CTA;
//   * the finished tile leaves from registers with 256-byte-run stores, as before.
// A sweep's geometry is data, not code: any 4 of the 12 tile positions are the register bits, the
// other 8 are thread bits, and the shared-memory slot of a logical tile index is an affine map
// chosen by the planner (the XOR swizzle of the per-tile kernels).  X / CNOT are relabelings of the logical index and only change those tables.
// For the plans of the per-tile kernels (PQC_PIPE=1) arithmetic and op order per amplitude are
// those of k_layer_pass / k_layer_seq: results are bitwise identical (tests/test_gpu_parity.py
// compares the paths).  The front planner's own plans (pqc_front.cu) additionally use the 4-slot
// R_y / R_z layer ops, the merged ry-CZ-ry rotation and the pending Z frame defined below.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pqc_ops.cuh"

#define TP_THREADS 512
// Op-set specialisation: k_tile_pipe is an interpreter at the 128-register cap, and its speed is set
// by how ptxas lays out the op loop (profiles/r2_relabel_tables.md) -- so the kernel is compiled once
// per FAMILY of op kinds, each instance carrying only the branches its passes can reach (and the Z
// frame / the relabeling code only where they can occur).  The launcher picks the smallest set
// that covers a pass' ops; TP_ALL is the catch-all (and the only one with in-pass spawns).
#define TP_O_RY4 1
#define TP_O_RY4PAD 2      // merged ry-CZ-ry slots (partner bytes): needs the Z frame
#define TP_O_RZ4 4
#define TP_O_RX4 8
#define TP_O_CZF 16        // needs the Z frame
#define TP_O_RXY 32
#define TP_O_ZZSUM 64      // ZZSUM / GEN table phases
#define TP_O_DIAG 128      // runs of R_z / CZ
#define TP_O_REAL4 256
#define TP_O_RZZ 512       // RZZ1 / RZZ2
#define TP_O_PERM 1024     // X / CNOT relabelings at the sweep ends
#define TP_ALL 2047
#define TP_SET_XXZ (TP_O_RXY | TP_O_ZZSUM | TP_O_RZZ)
#define TP_SET_HE (TP_O_RY4 | TP_O_RZ4 | TP_O_PERM)
#define TP_SET_NPQC (TP_O_RY4 | TP_O_RY4PAD | TP_O_RZ4 | TP_O_CZF | TP_O_DIAG | TP_O_REAL4)
#define TP_HAS_FZ(OPS) (((OPS) & (TP_O_RY4PAD | TP_O_CZF)) != 0)
#define TP_TILE_BYTES 65536
#define TP_TRIG_BYTES (TP_MAX_TRIG * 16)
#define TP_SMEM_TILES (TP_NBUF * TP_TILE_BYTES)
#define TP_SMEM_TRIG (4 * TP_TRIG_BYTES)          // [group][parity] trig areas
#define TP_SMEM_TOTAL (TP_SMEM_TILES + TP_SMEM_TRIG + sizeof(PipePlan))

// PQC_PIPE=1 routes the plans of the per-tile kernels (k_layer_pass / k_layer_seq) through
// k_tile_pipe as well; by default only the plans made for it (front planner) run here, because
// the compile-time geometry of k_layer_pass is faster on the TFIM passes (profiles/r2_pipe_ab.md).
// Read per launch so tests can compare both paths.
bool pqc_pipe_enabled() {
  const char* e = getenv("PQC_PIPE");
  return e && strcmp(e, "1") == 0;
}

__device__ __forceinline__ void tp_group_bar(int grp) {
  asm volatile("bar.sync %0, 256;" ::"r"(grp + 1) : "memory");
}
__device__ __forceinline__ void tp_cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tp_cp_async16_cg(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void tp_cp_async_commit() {
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tp_cp_async_wait_1() {      // all but the newest group are done
  asm volatile("cp.async.wait_group 1;" ::: "memory");
}
// the mbarrier gets one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void tp_cp_async_arrive(uint64_t* bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// rz on register bit K in tangent form: a *= (1 - i t) where the bit is 0, (1 + i t) where it is
// 1; the cos factor goes to the pass' scalar like those of op_rx_t / op_ry_t
template <int K>
__device__ __forceinline__ void op_rz_t(c128 (&a)[16], double t) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const double s = (j & (1 << K)) ? t : -t;
    const c128 v = a[j];
    a[j] = make_double2(fma(-s, v.y, v.x), fma(s, v.x, v.y));
  }
}

// R_zz on the register bits KA < KB: exp(-i theta/2 z_a z_b), (c, s) = (cos, sin)(theta / 2)
template <int KA, int KB>
__device__ __forceinline__ void op_rzz1(c128 (&a)[16], double c, double s) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const double sj = (((j >> KA) ^ (j >> KB)) & 1) ? s : -s;
    const c128 v = a[j];
    a[j] = make_double2(fma(-sj, v.y, c * v.x), fma(sj, v.x, c * v.y));
  }
}
// two same-angle R_zz on the disjoint register-bit pairs (KA, KB) and (the other two bits): the
// phase is exp(-i theta) where both pairs are aligned, exp(+i theta) where both are anti-aligned
// and 1 where they differ; (c, s) = (cos, sin)(theta)
template <int KA, int KB>
__device__ __forceinline__ void op_rzz2(c128 (&a)[16], double c, double s) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int p0 = ((j >> KA) ^ (j >> KB)) & 1;
    const int rest = 15 & ~((1 << KA) | (1 << KB));
    int p1 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if ((rest >> k) & 1) p1 ^= (j >> k) & 1;
    if (p0 != p1) continue;
    const double sj = p0 ? s : -s;
    const c128 v = a[j];
    a[j] = make_double2(fma(-sj, v.y, c * v.x), fma(sj, v.x, c * v.y));
  }
}

__device__ __forceinline__ uint32_t tp_partner_bit(uint32_t pb, uint32_t lidx, uint32_t tbase) {
  return (pb & 0x80u) ? ((tbase >> (pb & 0x7fu)) & 1u) : ((lidx >> pb) & 1u);
}
__device__ __forceinline__ double tp_flip(double v, uint32_t s31) {
  return __hiloint2double(__double2hiint(v) ^ (int)s31, __double2loint(v));
}
// apply the pending Z frame: amplitudes whose register bit k is 1 change sign where fz bit k is set
__device__ __forceinline__ void tp_zflush(c128 (&a)[16], uint32_t& fz) {
  const uint32_t m0 = (fz & 1u) << 31, m1 = ((fz >> 1) & 1u) << 31, m2 = ((fz >> 2) & 1u) << 31,
                 m3 = ((fz >> 3) & 1u) << 31;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const uint32_t m = ((j & 1) ? m0 : 0u) ^ ((j & 2) ? m1 : 0u) ^ ((j & 4) ? m2 : 0u) ^ ((j & 8) ? m3 : 0u);
    a[j] = make_double2(tp_flip(a[j].x, m), tp_flip(a[j].y, m));
  }
  fz = 0u;
}

// the ops of one sweep on the 16 register amplitudes.  lidx = the thread's logical tile index
// (register bits zero), tbase = the tile's amplitude offset, tile = its index.
// The 4-slot layer ops run all four slots unconditionally (identity parameters for absent gates):
// straight-line code with an even number of register-set hand-overs, so the compiler needs no
// register moves at the interpreter's merge point (a lone conditional rotation costs 32 moves).
template <bool GEN, int OPS>
__device__ __forceinline__ void tp_ops(c128 (&a)[16], const PipePlan* P, const TPSweep& sw, int ob,
                                       int oe, const double2* trig, uint32_t lidx, uint32_t tbase,
                                       uint32_t tile, int tiles_log2, int gen, const PipeArgs& A,
                                       double& fscale) {
  uint32_t fz = 0u;                               // pending Z frame, one bit per register bit
  // cos factors that depend on a partner bit differ from thread to thread, and the thread ->
  // amplitude map changes with the sweep: they are applied before this sweep ends (fscale, which
  // is the same for every thread of a sample, rides on to the last sweep)
  double lscale = 1.0;
  // tangent-form amplitudes grow by 1 / |cos| per rotation (1.6e16 at an angle of exactly pi):
  // a deep pass at Clifford angles would overflow, so the pending scale is applied early once
  // it is tiny.  Never taken at generic angles.
#define TP_RESCALE()                                                         \
  {                                                                          \
    if (fabs(fscale) < 1e-100) { op_scale(a, fscale); fscale = 1.0; }        \
    if (fabs(lscale) < 1e-100) { op_scale(a, lscale); lscale = 1.0; }        \
  }
  const double2 ident = make_double2(0.0, 1.0);   // (tan, cos) of angle 0
  for (int oi = ob; oi < oe; ++oi) {
    const TPOp op = P->ops[oi];
    const int kind = op.kind;
    if ((OPS & TP_O_RY4) && kind == PQC_K_LAYER_RY4) {
      const int sk = op.sub;
      double2 c0 = ident, c1 = ident, c2 = ident, c3 = ident;
#define TP_RY_SLOT(K, PB, C)                                                       \
  if (sk & (3 << (2 * K))) {                                                       \
    uint32_t e = op.t[K];                                                          \
    if ((PB) != 0xff) {                                                            \
      const uint32_t bb = tp_partner_bit(PB, lidx, tbase);                         \
      e += bb;                                                                     \
      fz ^= bb << K;                                                               \
      C = trig[e];                                                                 \
      lscale *= C.y;                                                               \
      C.y = 1.0;                                                                   \
    } else {                                                                       \
      C = trig[e];                                                                 \
    }                                                                              \
    if ((fz >> K) & 1u) C.x = -C.x;                                                \
  }
      if ((OPS & TP_O_RY4PAD) && op.pad) {        // some slot is a merged ry-CZ-ry rotation
        TP_RY_SLOT(0, op.a, c0) TP_RY_SLOT(1, op.b, c1) TP_RY_SLOT(2, op.wt, c2) TP_RY_SLOT(3, op.nterms, c3)
      } else {                                    // plain layer: four predicated table reads
        if (sk & 0x03) c0 = trig[op.t[0]];
        if (sk & 0x0c) c1 = trig[op.t[1]];
        if (sk & 0x30) c2 = trig[op.t[2]];
        if (sk & 0xc0) c3 = trig[op.t[3]];
        if (TP_HAS_FZ(OPS) && fz) {
          if (fz & 1u) c0.x = -c0.x;
          if (fz & 2u) c1.x = -c1.x;
          if (fz & 4u) c2.x = -c2.x;
          if (fz & 8u) c3.x = -c3.x;
        }
      }
#undef TP_RY_SLOT
      op_ry_t<0>(a, c0.x);
      op_ry_t<1>(a, c1.x);
      op_ry_t<2>(a, c2.x);
      op_ry_t<3>(a, c3.x);
      fscale *= (c0.y * c1.y) * (c2.y * c3.y);
      TP_RESCALE()
    } else if ((OPS & TP_O_RZ4) && kind == PQC_K_LAYER_RZ4) {
      const int sk = op.sub;
      double2 c0 = ident, c1 = ident, c2 = ident, c3 = ident;
      if (sk & 0x03) c0 = trig[op.t[0]];
      if (sk & 0x0c) c1 = trig[op.t[1]];
      if (sk & 0x30) c2 = trig[op.t[2]];
      if (sk & 0xc0) c3 = trig[op.t[3]];
      op_rz_t<0>(a, c0.x);
      op_rz_t<1>(a, c1.x);
      op_rz_t<2>(a, c2.x);
      op_rz_t<3>(a, c3.x);
      fscale *= (c0.y * c1.y) * (c2.y * c3.y);
      TP_RESCALE()
    } else if ((OPS & TP_O_RX4) && kind == PQC_K_LAYER_RX4) {
      const int sk = op.sub;
      double2 c0 = ident, c1 = ident, c2 = ident, c3 = ident;
      if (sk & 0x03) c0 = trig[op.t[0]];
      if (sk & 0x0c) c1 = trig[op.t[1]];
      if (sk & 0x30) c2 = trig[op.t[2]];
      if (sk & 0xc0) c3 = trig[op.t[3]];
      if (TP_HAS_FZ(OPS)) {
        if (fz & 1u) c0.x = -c0.x;
        if (fz & 2u) c1.x = -c1.x;
        if (fz & 4u) c2.x = -c2.x;
        if (fz & 8u) c3.x = -c3.x;
      }
      op_rx_t<0>(a, c0.x);
      op_rx_t<1>(a, c1.x);
      op_rx_t<2>(a, c2.x);
      op_rx_t<3>(a, c3.x);
      fscale *= (c0.y * c1.y) * (c2.y * c3.y);
      TP_RESCALE()
    } else if ((OPS & TP_O_CZF) && kind == PQC_K_CZF) {
      fz ^= tp_partner_bit(op.b, lidx, tbase) << op.a;
    } else if ((OPS & TP_O_RXY) && kind == PQC_K_RXY) {
      double2 cs = trig[op.t[0]];
      if (TP_HAS_FZ(OPS)) {
        const int ka = op.a >> 2, kb = op.a & 3;
        if (((fz >> ka) ^ (fz >> kb)) & 1u) cs.y = -cs.y;
      }
      switch (op.a) {
        case 1: op_xy<0, 1>(a, cs.x, cs.y); break;
        case 2: op_xy<0, 2>(a, cs.x, cs.y); break;
        case 3: op_xy<0, 3>(a, cs.x, cs.y); break;
        case 6: op_xy<1, 2>(a, cs.x, cs.y); break;
        case 7: op_xy<1, 3>(a, cs.x, cs.y); break;
        default: op_xy<2, 3>(a, cs.x, cs.y); break;
      }
    } else if ((OPS & TP_O_RZZ) && kind == PQC_K_RZZ1) {
      const double2 cs = trig[op.t[0]];
      switch (op.a) {
        case 1: op_rzz1<0, 1>(a, cs.x, cs.y); break;
        case 2: op_rzz1<0, 2>(a, cs.x, cs.y); break;
        case 3: op_rzz1<0, 3>(a, cs.x, cs.y); break;
        case 6: op_rzz1<1, 2>(a, cs.x, cs.y); break;
        case 7: op_rzz1<1, 3>(a, cs.x, cs.y); break;
        default: op_rzz1<2, 3>(a, cs.x, cs.y); break;
      }
    } else if ((OPS & TP_O_RZZ) && kind == PQC_K_RZZ2) {
      const double2 cs = trig[op.t[0]];
      if (op.a == 0) op_rzz2<0, 1>(a, cs.x, cs.y);
      else if (op.a == 1) op_rzz2<0, 2>(a, cs.x, cs.y);
      else op_rzz2<0, 3>(a, cs.x, cs.y);
    } else if ((OPS & TP_O_ZZSUM) && (kind == PQC_K_ZZSUM || kind == PQC_K_GEN)) {
      // w(x) = w(tile) ^ w(thread part) ^ w(register value j); wn holds w per tile nibble value
      const uint32_t(*wn)[16] = P->wn[op.wt];
      uint32_t w0 = wn[0][lidx & 15u] ^ wn[1][(lidx >> 4) & 15u] ^ wn[2][lidx >> 8];
      for (int jb = 0; jb < tiles_log2; ++jb) w0 ^= ((tile >> jb) & 1u) ? P->wo[op.wt][jb] : 0u;
      const uint32_t w1 = wn[sw.rpos[0] >> 2][1u << (sw.rpos[0] & 3)];
      const uint32_t w2 = wn[sw.rpos[1] >> 2][1u << (sw.rpos[1] & 3)];
      const uint32_t w3 = wn[sw.rpos[2] >> 2][1u << (sw.rpos[2] & 3)];
      const uint32_t w4 = wn[sw.rpos[3] >> 2][1u << (sw.rpos[3] & 3)];
      if (kind == PQC_K_ZZSUM) {
        const double2* tz = trig + op.t[0];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double2 ph = tz[__popc(w0 ^ XSEL4R(j, w1, w2, w3, w4))];
          const c128 v = a[j];
          a[j] = make_double2(v.x * ph.x - v.y * ph.y, v.y * ph.x + v.x * ph.y);
        }
      } else if (GEN && gen == op.spawn) {
        const double cr = A.spawn_cr[gen], ci = A.spawn_ci[gen];
        const int nt = op.nterms;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const double f = (double)(nt - 2 * __popc(w0 ^ XSEL4R(j, w1, w2, w3, w4)));
          const double fr = f * cr, fi = f * ci;
          const c128 v = a[j];
          a[j] = make_double2(v.x * fr - v.y * fi, v.y * fr + v.x * fi);
        }
      }
    } else if ((OPS & TP_O_DIAG) && (kind == PQC_OP_RZ || kind == PQC_OP_CZ)) {
      // ---- a run of diagonal ops: accumulate, then apply once (as k_layer_seq)
      double tc = 1.0, ts = 0.0;                  // thread-level phase (tc + i ts)
      double pc[4] = {1.0, 1.0, 1.0, 1.0}, pn[4] = {0.0, 0.0, 0.0, 0.0};   // bit k = 1: (pc + i pn)
      uint32_t sg = 0u, touched = 0u;             // sign mask over j; bit 4 of touched: T set
      for (; oi < oe; ++oi) {
        const TPOp o2 = P->ops[oi];
        if (o2.kind == PQC_OP_RZ) {
          const double2 cs = trig[o2.t[0]];
          const int k0 = o2.a == 0xff ? -1 : (int)o2.a;
          if (k0 >= 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              if (k == k0) {                      // (pc + i pn) *= (c + i s)
                const double c = pc[k], sn = pn[k];
                pc[k] = c * cs.x - sn * cs.y;
                pn[k] = sn * cs.x + c * cs.y;
              }
            touched |= 1u << k0;
          } else {
            const uint32_t bit = o2.b != 0xff ? ((lidx >> o2.b) & 1u) : ((tbase >> o2.t[1]) & 1u);
            const double sn = bit ? cs.y : -cs.y, c = tc, s0 = ts;
            tc = c * cs.x - s0 * sn;
            ts = s0 * cs.x + c * sn;
            touched |= 16u;
          }
        } else if (o2.kind == PQC_OP_CZ) {
          auto mask_of = [&](int k, int l, int b) -> uint32_t {
            if (k != 0xff) return k == 0 ? 0xAAAAu : (k == 1 ? 0xCCCCu : (k == 2 ? 0xF0F0u : 0xFF00u));
            const uint32_t bit = l != 0xffff ? ((lidx >> l) & 1u) : ((tbase >> b) & 1u);
            return bit ? 0xFFFFu : 0u;
          };
          sg ^= mask_of(o2.a, o2.t[0], o2.t[2]) & mask_of(o2.b, o2.t[1], o2.t[3]);
        } else {
          break;
        }
      }
      --oi;                                       // the outer loop steps past the run's last op
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (touched & (1u << k)) {                // warp-uniform
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const double sn = (j & (1 << k)) ? pn[k] : -pn[k];
            const c128 v = a[j];
            a[j] = make_double2(v.x * pc[k] - v.y * sn, v.y * pc[k] + v.x * sn);
          }
        }
      }
      if (touched & 16u) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const c128 v = a[j];
          a[j] = make_double2(v.x * tc - v.y * ts, v.y * tc + v.x * ts);
        }
      }
      if (sg) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const uint32_t m = ((sg >> j) & 1u) << 31;
          a[j] = make_double2(tp_flip(a[j].x, m), tp_flip(a[j].y, m));
        }
      }
    } else if ((OPS & TP_O_REAL4) && kind == PQC_K_LAYER_REAL4) {
      // ry / Hadamard mix (initial layers): H does not commute with a pending Z frame
      if (TP_HAS_FZ(OPS) && fz) tp_zflush(a, fz);
      const int sk = op.sub;
      double f = 1.0;
#define TP_REAL_SLOT(K)                                                    \
  {                                                                        \
    const int kd = (sk >> (2 * K)) & 3;                                    \
    if (kd == 1) { const double2 tc = trig[op.t[K]]; op_ry_t<K>(a, tc.x); f *= tc.y; } \
    else if (kd == 2) { op_h_u<K>(a); f *= 0.70710678118654752440; }       \
  }
      TP_REAL_SLOT(0) TP_REAL_SLOT(1) TP_REAL_SLOT(2) TP_REAL_SLOT(3)
#undef TP_REAL_SLOT
      fscale *= f;
      TP_RESCALE()
    } else if (TP_HAS_FZ(OPS) && kind == PQC_K_ZFLUSH) {
      if (fz) tp_zflush(a, fz);
    }
  }
  if (TP_HAS_FZ(OPS) && fz) tp_zflush(a, fz);
  if ((OPS & TP_O_RY4PAD) && lscale != 1.0) op_scale(a, lscale);
#undef TP_RESCALE
}

// X / CNOT index permutations of a sweep are affine maps of the 4-bit register index, label(j) =
// XOR_{k in j} col[k] ^ v (k_sweep_pass' `affine`); the front planner works them out once per sweep
// (TPAff, pqc_front.cu) and the kernel only XORs slot masks.

// base slot (or amplitude) mask of a relabeled load / store: the constant part plus one XOR per CNOT
// whose control bit is fixed for this thread.  (While the kernel was one monolithic interpreter this
// had to be a non-inlined call: inlined, it changed the register allocation and slowed the XXZ passes,
// which never come here, by 3 %.  With one kernel instance per op family the families no longer share
// an allocation, and inlined it is worth 7 % on the CNOT-chain instance: profiles/r2_relabel_tables.md.)
__device__ __forceinline__ uint32_t tp_aff_base(const TPAff* af, uint32_t lidx, uint32_t tbase) {
  uint32_t x = af->base;
  for (int i = 0; i < af->ninj; ++i)
    if (tp_partner_bit(af->inj[i].src, lidx, tbase)) x ^= af->inj[i].lm;
  return x;
}
template <bool GEN, int OPS>
__global__ void __launch_bounds__(TP_THREADS, 1) k_tile_pipe(const PipeArgs A) {
  extern __shared__ __align__(128) unsigned char tp_sm[];
  double2* trigs = reinterpret_cast<double2*>(tp_sm + TP_SMEM_TILES);
  PipePlan* P = reinterpret_cast<PipePlan*>(tp_sm + TP_SMEM_TILES + TP_SMEM_TRIG);
  __shared__ __align__(8) uint64_t full[TP_NBUF];
  const int tid = threadIdx.x, grp = tid >> 8, t = tid & 255, lo = t & 15, hi = t >> 4;
  {
    const int* g = reinterpret_cast<const int*>(A.plan);
    int* d = reinterpret_cast<int*>(P);
    for (int e = tid; e < (int)(sizeof(PipePlan) / 4); e += TP_THREADS) d[e] = g[e];
  }
  if (tid == 0) {
    for (int b = 0; b < TP_NBUF; ++b) gr_mbar_init(&full[b], 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int tiles_log2 = A.n - 12;
  const uint32_t tmask = (1u << tiles_log2) - 1u;
  const uint32_t total = (uint32_t)(A.n_items << tiles_log2);   // < 2^31 (checked by the host)
  const uint32_t stride = gridDim.x;
  const int nk = (int)((total - blockIdx.x + stride - 1) / stride);
  const uint32_t ips = (uint32_t)(A.active + A.nspawn);
  const int ntrig = P->ntrig;

  auto item_of = [&](int k, uint32_t& sample, int& r, uint32_t& tile) {
    const uint32_t item = blockIdx.x + (uint32_t)k * stride;
    const uint32_t vec = item >> tiles_log2;
    tile = item & tmask;
    sample = vec / ips;
    r = (int)(vec - sample * ips);
  };
  auto tile_base = [&](uint32_t tile) -> uint32_t {
    uint32_t tb = 0;
    for (int j = 0; j < tiles_log2; ++j) tb |= ((tile >> j) & 1u) << A.obit[j];
    return tb;
  };
  // the 256 threads of one group: asynchronous copy of the tile of work item k into ring slot
  // k % TP_NBUF (16 x 16 bytes per thread; lanes cover the low amplitude bits, so a warp reads
  // 256-byte runs and writes conflict-free swizzled slots)
  auto load_tile = [&](int k) {
    const int b = k % TP_NBUF;
    uint32_t sample, tile;
    int r;
    item_of(k, sample, r, tile);
    const int src_slot = (GEN && r >= A.active) ? 0 : r;
    const c128* src = A.src + (((long long)sample * A.slots_total + src_slot) << A.n) +
                      (tile_base(tile) | P->ld_amp[0][lo] | P->ld_amp[1][hi]);
    unsigned char* dst = tp_sm + (size_t)b * TP_TILE_BYTES;
    const uint32_t slot0 = (uint32_t)P->ld_slot[0][lo] ^ (uint32_t)P->ld_slot[1][hi];
    const uint32_t g0 = P->ld_r[0], g1 = P->ld_r[1], g2 = P->ld_r[2], g3 = P->ld_r[3];
    const uint32_t s0 = P->ld_sr[0], s1 = P->ld_sr[1], s2 = P->ld_sr[2], s3 = P->ld_sr[3];
#pragma unroll
    for (int j = 0; j < 16; ++j)
      tp_cp_async16_cg(dst + (slot0 ^ XSEL4R(j, s0, s1, s2, s3)), src + XSEL4R(j, g0, g1, g2, g3));
    tp_cp_async_arrive(&full[b]);
  };
  // the group's trig entries of work item k -> area [grp][(k >> 1) & 1] (cp.async, 16 B per thread)
  auto trig_fetch = [&](int k) {
    uint32_t sample, tile;
    int r;
    item_of(k, sample, r, tile);
    double2* dst = trigs + (size_t)(grp * 2 + ((k >> 1) & 1)) * TP_MAX_TRIG;
    const double2* srcp = A.gtrig + (long long)sample * A.tstride + A.toff;
    for (int e = t; e < ntrig; e += 256) tp_cp_async16(dst + e, srcp + e);
  };

  // prologue: every group fetches the trig entries and the tile of its first item; group 0
  // also the third ring slot.  cp.async groups per thread, oldest first: [trig k] [tile ..],
  // then per item [trig k + 2] [tile k + 3] -- so "all but the newest group" at the top of an
  // item always covers that item's trig entries
  if (grp < nk) trig_fetch(grp);
  tp_cp_async_commit();
  if (grp < nk) load_tile(grp);
  if (grp == 0 && 2 < nk) load_tile(2);
  tp_cp_async_commit();

  for (int k = grp; k < nk; k += 2) {
    const int b = k % TP_NBUF;
    unsigned char* buf = tp_sm + (size_t)b * TP_TILE_BYTES;
    const double2* trig = trigs + (size_t)(grp * 2 + ((k >> 1) & 1)) * TP_MAX_TRIG;
    uint32_t sample, tile;
    int r;
    item_of(k, sample, r, tile);
    const uint32_t tbase = tile_base(tile);
    int dst_slot = r, gen = -1;
    if (GEN && r >= (int)A.active) {
      gen = r - A.active;
      dst_slot = A.spawn_slot[gen];
    }
    // this item's trig entries have landed (prefetched one item ahead); fetch the next item's
    tp_cp_async_wait_1();
    tp_group_bar(grp);
    if (k + 2 < nk) trig_fetch(k + 2);
    tp_cp_async_commit();
    gr_mbar_wait(&full[b], (unsigned)((k / TP_NBUF) & 1));

    double fscale = 1.0;
    c128 a[16];
    const int nsw = P->nsw;
    for (int s = 0; s < nsw; ++s) {
      const TPSweep& sw = P->sw[s];
      const uint32_t tw = sw.tt[0][lo] ^ sw.tt[1][hi];
      const uint32_t sb = tw & 0xffffu, lidx = tw >> 16;
      const uint32_t r0 = sw.rs[0], r1 = sw.rs[1], r2 = sw.rs[2], r3 = sw.rs[3];
      const int npre = sw.npre, npost = sw.npost, ob = sw.op_begin, oe = sw.op_end;
      const bool last = s + 1 == nsw;
      if ((OPS & TP_O_PERM) && npre) {
        // planner-made relabeling: four slot masks, a base and one conditional XOR per CNOT whose
        // control is fixed for the thread
        const uint32_t lb = sb ^ tp_aff_base(&sw.pre, lidx, tbase);
        const uint32_t l0 = sw.pre.l[0], l1 = sw.pre.l[1], l2 = sw.pre.l[2], l3 = sw.pre.l[3];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          a[j] = *reinterpret_cast<const c128*>(buf + (lb ^ XSEL4R(j, l0, l1, l2, l3)));
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          a[j] = *reinterpret_cast<const c128*>(buf + (sb ^ XSEL4R(j, r0, r1, r2, r3)));
      }
      if (last) {
        // every thread of the group has read its last amplitudes from the ring slot: refill it
        // with the item three steps ahead before doing this sweep's arithmetic
        tp_group_bar(grp);
        if (k + TP_NBUF < nk) load_tile(k + TP_NBUF);
        tp_cp_async_commit();
      }
      tp_ops<GEN, OPS>(a, P, sw, ob + npre, oe - npost, trig, lidx, tbase, tile, tiles_log2, gen, A, fscale);
      if (last) {
        if (fscale != 1.0) op_scale(a, fscale);
        const uint32_t amp = tbase | P->st_t[0][lo] | P->st_t[1][hi];
        uint32_t g0 = P->st_r[0], g1 = P->st_r[1], g2 = P->st_r[2], g3 = P->st_r[3];
        uint32_t gb = amp;
        if ((OPS & TP_O_PERM) && npost) {
          gb = amp ^ P->st_base;
          const uint32_t lidx2 = (sw.tt[0][lo] ^ sw.tt[1][hi]) >> 16;   // re-read, not kept live
          for (int i = 0; i < sw.post.ninj; ++i)
            if (tp_partner_bit(sw.post.inj[i].src, lidx2, tbase)) gb ^= P->st_gm[i];
          g0 = P->st_q[0]; g1 = P->st_q[1]; g2 = P->st_q[2]; g3 = P->st_q[3];
        }
        c128* dp = A.dst + (((long long)sample * A.slots_total + dst_slot) << A.n);
#pragma unroll
        for (int j = 0; j < 16; ++j) dp[gb ^ XSEL4R(j, g0, g1, g2, g3)] = a[j];
        break;
      }
      // the sweep's slot masks are read again instead of being kept live across the ops: five
      // registers fewer under the 128-register cap, fewer spills (hardware-efficient 16q x 16:
      // 36.0 -> 34.0 ms, XXZ 16q x 16 on this kernel: 44.0 -> 40.8 ms).  Also measured and dropped
      // (profiles/r2_final_summary.md): reading the next op ahead of the arithmetic (+25 - 60 %:
      // spills), a sentinel op instead of the loop's end index (+4 - 9 %)
      const uint32_t tw2 = sw.tt[0][lo] ^ sw.tt[1][hi];
      const uint32_t sb2 = tw2 & 0xffffu;
      const uint32_t u0 = sw.rs[0], u1 = sw.rs[1], u2 = sw.rs[2], u3 = sw.rs[3];
      if ((OPS & TP_O_PERM) && npost) {
        const uint32_t lb = sb2 ^ tp_aff_base(&sw.post, tw2 >> 16, tbase);
        const uint32_t l0 = sw.post.l[0], l1 = sw.post.l[1], l2 = sw.post.l[2], l3 = sw.post.l[3];
#pragma unroll
        for (int j = 0; j < 16; ++j)
          *reinterpret_cast<c128*>(buf + (lb ^ XSEL4R(j, l0, l1, l2, l3))) = a[j];
      } else {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          *reinterpret_cast<c128*>(buf + (sb2 ^ XSEL4R(j, u0, u1, u2, u3))) = a[j];
      }
      tp_group_bar(grp);
    }
  }
}

// =====================================================================================
// host: V1Pass (FastPlan / SeqPlan) -> PipePlan
// =====================================================================================
static uint32_t h_swz(uint32_t i) { return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u); }

namespace {
struct HSweep {
  int geom;                      // registers hold tile positions 0: 8-11, 1: 0-3, 2: 4-7
  std::vector<FastOp> ops;
};
}  // namespace

bool pqc_pipe_build(const V1Pass& ps, int n, PipePlan& pp) {
  if (!(ps.fast_ok || ps.seq_ok) || n < 12 || ps.tb != 12 || ps.low_run < 4) return false;
  std::vector<HSweep> hs;
  const uint32_t(*wn)[3][16] = nullptr;
  const uint32_t(*wo)[PQC_MAX_QUBITS - 12] = nullptr;
  if (ps.fast_ok) {
    const FastPlan& f = ps.fast;
    const int slots3[3] = {0, 1, 2}, slots2[2] = {0, 2};
    for (int i = 0; i < f.ns; ++i) {
      const int slot = f.ns == 3 ? slots3[i] : (f.ns == 2 ? slots2[i] : 0);
      HSweep h;
      h.geom = slot;
      for (int o = 0; o < f.nops[slot]; ++o) h.ops.push_back(f.ops[slot][o]);
      hs.push_back(h);
    }
    wn = f.wn;
    wo = f.wo;
  } else {
    const SeqPlan& q = ps.seq;
    for (int i = 0; i < q.nsw; ++i) {
      HSweep h;
      h.geom = q.geom[i];
      for (int o = 0; o < q.nops[i]; ++o) h.ops.push_back(q.ops[q.off[i] + o]);
      hs.push_back(h);
    }
    wn = q.wn;
    wo = q.wo;
  }
  if (hs.empty()) return false;
  // the last sweep stores to global memory from registers: lanes must cover the low amplitude
  // bits, so a last sweep on tile positions 0-3 gets an op-less transposition sweep after it
  if (hs.back().geom == 1) hs.push_back(HSweep{2, {}});
  if ((int)hs.size() > TP_MAX_SWEEPS) return false;
  memset(&pp, 0, sizeof(pp));
  pp.nsw = (int)hs.size();
  pp.ntrig = ps.ntrig;
  pp.nwt = ps.nwt;
  if (pp.ntrig > TP_MAX_TRIG || pp.nwt > FAST_MAX_WT) return false;
  for (int p = 0; p < 12; ++p) pp.lbit[p] = ps.lbit[p];
  for (int w = 0; w < FAST_MAX_WT; ++w) {
    memcpy(pp.wn[w], wn[w], sizeof(pp.wn[w]));
    memcpy(pp.wo[w], wo[w], sizeof(pp.wo[w]));
  }
  {
    // tile load: thread bits 0-3 on the tile positions of amplitude bits 0-3 (256-byte runs per
    // half warp), thread bits 4-7 and the copy index j on the other positions in ascending order
    int tpos[12], nt = 0;
    for (int b = 0; b < 4; ++b)
      for (int p = 0; p < 12; ++p)
        if (ps.lbit[p] == b) tpos[nt++] = p;
    if (nt != 4) return false;
    for (int p = 0; p < 12; ++p)
      if (ps.lbit[p] >= 4) tpos[nt++] = p;
    for (int h = 0; h < 2; ++h)
      for (int v = 0; v < 16; ++v) {
        uint32_t idx = 0, amp = 0;
        for (int i = 0; i < 4; ++i)
          if ((v >> i) & 1) {
            idx |= 1u << tpos[4 * h + i];
            amp |= 1u << ps.lbit[tpos[4 * h + i]];
          }
        pp.ld_slot[h][v] = (uint16_t)(h_swz(idx) << 4);
        pp.ld_amp[h][v] = amp;
      }
    for (int k = 0; k < 4; ++k) {
      pp.ld_sr[k] = (uint16_t)(h_swz(1u << tpos[8 + k]) << 4);
      pp.ld_r[k] = 1u << ps.lbit[tpos[8 + k]];
    }
  }
  int nops = 0;
  for (int s = 0; s < pp.nsw; ++s) {
    TPSweep& sw = pp.sw[s];
    const int g = hs[s].geom;
    const int r0 = g == 0 ? 8 : (g == 1 ? 0 : 4);
    int tp[8], nt = 0;
    for (int p = 0; p < 12; ++p)
      if (p < r0 || p >= r0 + 4) tp[nt++] = p;
    for (int k = 0; k < 4; ++k) {
      sw.rpos[k] = (uint8_t)(r0 + k);
      sw.rs[k] = (uint16_t)(h_swz(1u << (r0 + k)) << 4);
    }
    for (int h = 0; h < 2; ++h)
      for (int v = 0; v < 16; ++v) {
        uint32_t idx = 0;
        for (int i = 0; i < 4; ++i)
          if ((v >> i) & 1) idx |= 1u << tp[4 * h + i];
        sw.tt[h][v] = (h_swz(idx) << 4) | (idx << 16);
        if (s + 1 == pp.nsw) {
          uint32_t amp = 0;
          for (int p = 0; p < 12; ++p)
            if ((idx >> p) & 1u) amp |= 1u << ps.lbit[p];
          pp.st_t[h][v] = amp;
        }
      }
    if (s + 1 == pp.nsw)
      for (int k = 0; k < 4; ++k) pp.st_r[k] = 1u << ps.lbit[r0 + k];
    sw.op_begin = (uint16_t)nops;
    for (const FastOp& f : hs[s].ops) {
      if (nops >= TP_MAX_OPS) return false;
      TPOp o;
      memset(&o, 0, sizeof(o));
      o.kind = (uint8_t)f.kind;
      o.a = o.b = 0xff;
      if (f.kind == PQC_K_LAYER_RX4 || f.kind == PQC_K_LAYER_REAL4) {
        // REAL4 stays REAL4 (not LAYER_RY4): its cos factors multiply in slot order, which is
        // what keeps these plans bit-identical to k_layer_pass / k_layer_seq
        for (int k = 0; k < 4; ++k) {
          const int kd = ((f.subk >> (8 * k)) & 0xff) - 1;
          int code = 0;
          if (kd == PQC_OP_RX || kd == PQC_OP_RY) code = 1;
          else if (kd == PQC_OP_H) code = 2;
          else if (kd >= 0) return false;
          if (f.kind == PQC_K_LAYER_RX4 && kd >= 0 && kd != PQC_OP_RX) return false;
          if (f.kind == PQC_K_LAYER_REAL4 && kd == PQC_OP_RX) return false;
          o.sub |= (uint8_t)(code << (2 * k));
          o.t[k] = (uint16_t)f.t[k];
        }
      } else if (f.kind == PQC_K_ZZSUM) {
        o.t[0] = (uint16_t)f.t[0];
        o.wt = (uint8_t)f.wt;
        o.nterms = (uint8_t)f.nterms;
      } else if (f.kind == PQC_K_GEN) {
        o.wt = (uint8_t)f.wt;
        o.nterms = (uint8_t)f.nterms;
        o.spawn = (uint8_t)f.spawn;
      } else if (f.kind == PQC_K_RXY) {
        o.a = (uint8_t)f.subk;
        o.t[0] = (uint16_t)f.t[0];
      } else if (f.kind == PQC_OP_RZ) {
        o.t[0] = (uint16_t)f.t[0];
        o.a = f.t[1] >= 0 ? (uint8_t)f.t[1] : 0xff;
        o.b = f.t[2] >= 0 ? (uint8_t)f.t[2] : 0xff;
        o.t[1] = (uint16_t)f.t[3];
      } else if (f.kind == PQC_OP_CZ) {
        o.a = f.t[0] >= 0 ? (uint8_t)f.t[0] : 0xff;
        o.b = f.t[1] >= 0 ? (uint8_t)f.t[1] : 0xff;
        o.t[0] = f.t[2] >= 0 ? (uint16_t)f.t[2] : 0xffff;
        o.t[1] = f.t[3] >= 0 ? (uint16_t)f.t[3] : 0xffff;
        o.t[2] = (uint16_t)f.wt;
        o.t[3] = (uint16_t)f.nterms;
      } else {
        return false;
      }
      pp.ops[nops++] = o;
    }
    sw.op_end = (uint16_t)nops;
  }
  pp.nops = nops;
  return true;
}

// the op kinds (and sweep-end relabelings) a plan uses, as a TP_O_* mask
static int tp_plan_ops(const PipePlan& pp) {
  int m = 0;
  for (int i = 0; i < pp.nops; ++i) {
    const TPOp& o = pp.ops[i];
    switch (o.kind) {
      case PQC_K_LAYER_RY4: m |= TP_O_RY4 | (o.pad ? TP_O_RY4PAD : 0); break;
      case PQC_K_LAYER_RZ4: m |= TP_O_RZ4; break;
      case PQC_K_LAYER_RX4: m |= TP_O_RX4; break;
      case PQC_K_CZF: case PQC_K_ZFLUSH: m |= TP_O_CZF; break;
      case PQC_K_RXY: m |= TP_O_RXY; break;
      case PQC_K_ZZSUM: case PQC_K_GEN: m |= TP_O_ZZSUM; break;
      case PQC_OP_RZ: case PQC_OP_CZ: m |= TP_O_DIAG; break;
      case PQC_K_LAYER_REAL4: m |= TP_O_REAL4; break;
      case PQC_K_RZZ1: case PQC_K_RZZ2: m |= TP_O_RZZ; break;
      case PQC_OP_X: case PQC_OP_CNOT: m |= TP_O_PERM; break;
      default: m |= TP_ALL; break;
    }
  }
  for (int s2 = 0; s2 < pp.nsw; ++s2)
    if (pp.sw[s2].npre || pp.sw[s2].npost) m |= TP_O_PERM;
  return m;
}

int pqc_pipe_launch(const PipeArgs& a, const PipePlan& hplan, cudaStream_t st) {
  int dev = 0;
  PQC_CUDA(cudaGetDevice(&dev));
  static int sms[64] = {0};
  static bool attr[64] = {false};
  if (dev < 0 || dev >= 64) PQC_FAIL(-1, "device index out of range");
  if (!attr[dev]) {
    PQC_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
#define TP_ATTR(K) PQC_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TP_SMEM_TOTAL))
    TP_ATTR((k_tile_pipe<false, TP_ALL>));
    TP_ATTR((k_tile_pipe<true, TP_ALL>));
    TP_ATTR((k_tile_pipe<false, TP_SET_XXZ>));
    TP_ATTR((k_tile_pipe<false, TP_SET_HE>));
    TP_ATTR((k_tile_pipe<false, TP_SET_NPQC>));
#undef TP_ATTR
    attr[dev] = true;
  }
  const long long total = a.n_items << (a.n - 12);
  if (total <= 0) return 0;
  if (total > 0x7fffffffLL) PQC_FAIL(-1, "pass grid too large; split the batch");
  const unsigned grid = (unsigned)std::min<long long>(total, sms[dev]);
  // the smallest compiled op set that covers this plan (PQC_PIPE_OPSET=all: always the catch-all)
  static const bool all_only = getenv("PQC_PIPE_OPSET") && strcmp(getenv("PQC_PIPE_OPSET"), "all") == 0;
  const int need = tp_plan_ops(hplan);
  const int h = pqc_prof_launch_begin((double)a.n_items * 2.0 * sizeof(c128) * (double)(1ll << a.n), st, PQC_PROF_TILE_PIPE);
  if (a.nspawn > 0) k_tile_pipe<true, TP_ALL><<<grid, TP_THREADS, TP_SMEM_TOTAL, st>>>(a);
  else if (!all_only && !(need & ~TP_SET_XXZ)) k_tile_pipe<false, TP_SET_XXZ><<<grid, TP_THREADS, TP_SMEM_TOTAL, st>>>(a);
  else if (!all_only && !(need & ~TP_SET_HE)) k_tile_pipe<false, TP_SET_HE><<<grid, TP_THREADS, TP_SMEM_TOTAL, st>>>(a);
  else if (!all_only && !(need & ~TP_SET_NPQC)) k_tile_pipe<false, TP_SET_NPQC><<<grid, TP_THREADS, TP_SMEM_TOTAL, st>>>(a);
  else k_tile_pipe<false, TP_ALL><<<grid, TP_THREADS, TP_SMEM_TOTAL, st>>>(a);
  pqc_prof_launch_end(h, st);
  PQC_LAUNCH_CHECK();
  return 0;
}
