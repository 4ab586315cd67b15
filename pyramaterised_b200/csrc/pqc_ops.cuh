// pqc_ops.cuh -- device helpers shared by the pass kernels (pqc_v1.cu, pqc_pipe.cu):
// the XOR swizzle, the register-resident gate micro-ops and the mbarrier / TMA bulk-copy
// wrappers.
#pragma once
#include "pqc_common.cuh"

__device__ __forceinline__ uint32_t swz(uint32_t i) {
  return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u);
}

// ---- pair-mixing micro-ops on the 16 register amplitudes ---------------------------------
// Rotations are applied in tangent form: rx = c [[1, -i t], [-i t, 1]], ry = c [[1, -t], [t, 1]]
// with t = tan(angle/2): two FMAs per amplitude instead of four.  The scalar c factors of a
// layer op are multiplied in once afterwards (op_scale).  t is finite for every double angle
// (cos never rounds to exactly 0 at pi/2: |c| >= 6e-17) and the result carries the usual
// relative rounding error of c x + s y because x + t y is computed with one rounding and the
// final multiplication by c is exact to one more.
template <int K>
__device__ __forceinline__ void op_rx_t(c128 (&a)[16], double t) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j & (1 << K)) continue;
    const c128 x = a[j], y = a[j | (1 << K)];
    a[j] = make_double2(fma(t, y.y, x.x), fma(-t, y.x, x.y));
    a[j | (1 << K)] = make_double2(fma(t, x.y, y.x), fma(-t, x.x, y.y));
  }
}
template <int K>
__device__ __forceinline__ void op_ry_t(c128 (&a)[16], double t) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j & (1 << K)) continue;
    const c128 x = a[j], y = a[j | (1 << K)];
    a[j] = make_double2(fma(-t, y.x, x.x), fma(-t, y.y, x.y));
    a[j | (1 << K)] = make_double2(fma(t, x.x, y.x), fma(t, x.y, y.y));
  }
}
// Hadamard without its 1/sqrt(2): (x + y, x - y)
template <int K>
__device__ __forceinline__ void op_h_u(c128 (&a)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j & (1 << K)) continue;
    const c128 x = a[j], y = a[j | (1 << K)];
    a[j] = make_double2(x.x + y.x, x.y + y.y);
    a[j | (1 << K)] = make_double2(x.x - y.x, x.y - y.y);
  }
}
__device__ __forceinline__ void op_scale(c128 (&a)[16], double f) {
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = make_double2(a[j].x * f, a[j].y * f);
}

// generic symmetric two-bit rotation: even-parity pair (00,11) by (ce, se), odd-parity pair
// (01,10) by (co, so), each as [[c, -i s], [-i s, c]]; optional phase (pc - i ps) on |11>.
template <int KA, int KB>
__device__ __forceinline__ void op_pair(c128 (&a)[16], double ce, double se, double co, double so,
                                        bool even, bool ph, double pc, double psn) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j & ((1 << KA) | (1 << KB))) continue;
    const int j01 = j | (1 << KA), j10 = j | (1 << KB), j11 = j01 | j10;
    {
      const c128 x = a[j01], y = a[j10];
      a[j01] = make_double2(co * x.x + so * y.y, co * x.y - so * y.x);
      a[j10] = make_double2(co * y.x + so * x.y, co * y.y - so * x.x);
    }
    if (even) {
      const c128 x = a[j], y = a[j11];
      a[j] = make_double2(ce * x.x + se * y.y, ce * x.y - se * y.x);
      a[j11] = make_double2(ce * y.x + se * x.y, ce * y.y - se * x.x);
    }
    if (ph) {
      const c128 z = a[j11];
      a[j11] = make_double2(z.x * pc + z.y * psn, z.y * pc - z.x * psn);
    }
  }
}

#define SEL4R(j, a0, a1, a2, a3) \
  ((((j)&1) ? (a0) : 0u) | (((j)&2) ? (a1) : 0u) | (((j)&4) ? (a2) : 0u) | (((j)&8) ? (a3) : 0u))
#define XSEL4R(j, a0, a1, a2, a3) \
  ((((j)&1) ? (a0) : 0u) ^ (((j)&2) ? (a1) : 0u) ^ (((j)&4) ? (a2) : 0u) ^ (((j)&8) ? (a3) : 0u))

// rotation of the odd-parity pair (01, 10) of register bits KA < KB by [[c, -i s], [-i s, c]]
template <int KA, int KB>
__device__ __forceinline__ void op_xy(c128 (&a)[16], double c, double s) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j & ((1 << KA) | (1 << KB))) continue;
    const int j01 = j | (1 << KA), j10 = j | (1 << KB);
    const c128 x = a[j01], y = a[j10];
    a[j01] = make_double2(c * x.x + s * y.y, c * x.y - s * y.x);
    a[j10] = make_double2(c * y.x + s * x.y, c * y.y - s * x.x);
  }
}

__device__ __forceinline__ void gr_mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void gr_mbar_expect(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void gr_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void gr_mbar_wait(uint64_t* bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "GR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra GR_DONE;\n"
      "bra GR_WAIT;\n"
      "GR_DONE:\n"
      "}\n" ::"r"(a), "r"(parity) : "memory");
}
__device__ __forceinline__ void gr_bulk_load(void* dst, const void* src, unsigned bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
      ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes),
      "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
