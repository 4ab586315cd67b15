// pqc_dense.cu -- ARBGATE (gates.py:407-435): exp(-i theta H) for an arbitrary dense Hermitian
// H on the whole register, and the reference's "derivative" (-i H / 2) exp(-i theta H).
//
// The reference rebuilds a dense matrix exponential (Qobj.expm) on every set_theta.  Here H is
// diagonalised ONCE on the host (H = V diag(lambda) V^dagger) and a batch of states is updated by
//   y = V^dagger x,   z_k = w(theta_s, lambda_k) y_k,   out = V z
// i.e. two dense complex128 matrix products with the batch (M [D x D] times X^T [D x S]) and a
// diagonal phase folded into the second product's operand load:
//   out[s][r] = sum_c M[r][c] * w(s, c) * in[s][c],
//   w = 1 (no lambda)  |  exp(-i theta_s lambda_c)  |  (-i lambda_c / 2) exp(-i theta_s lambda_c).
// 64 x 64 output tile per CTA, 16-deep K slabs in shared memory, 4 x 4 complex outputs per thread
// (FP64 FMA; D <= 2^13, a side path next to the gate-program kernels).
#include "pqc_common.cuh"

#define DN_T 64
#define DN_K 16

__global__ void __launch_bounds__(256) k_dense_apply(const c128* __restrict__ in, long long S, int n,
                                                     const c128* __restrict__ M,
                                                     const double* __restrict__ lambda,
                                                     const double* __restrict__ theta,
                                                     long long theta_stride, int deriv,
                                                     c128* __restrict__ out) {
  __shared__ c128 Xs[DN_K][DN_T + 1];      // [k][s]
  __shared__ c128 Ms[DN_K][DN_T + 1];      // [k][r]
  const long long D = 1ll << n;
  const long long s0 = (long long)blockIdx.y * DN_T, r0 = (long long)blockIdx.x * DN_T;
  const int tid = threadIdx.x, ts = tid >> 4, tr = tid & 15;
  double ar[4][4], ai[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) ar[i][j] = ai[i][j] = 0.0;
  for (long long k0 = 0; k0 < D; k0 += DN_K) {
    // stage: 64 x 16 elements of each operand, 4 per thread; consecutive threads walk k
    for (int e = tid; e < DN_T * DN_K; e += 256) {
      const int row = e / DN_K, k = e % DN_K;
      const long long c = k0 + k;
      c128 x = make_double2(0.0, 0.0), m = make_double2(0.0, 0.0);
      if (c < D) {
        if (s0 + row < S) {
          x = in[(s0 + row) * D + c];
          if (lambda) {
            double sn, cs;
            const double lam = lambda[c];
            sincos(-theta[(s0 + row) * theta_stride] * lam, &sn, &cs);
            double wr = cs, wi = sn;
            if (deriv) {                      // times (-i lambda / 2)
              const double h = 0.5 * lam;
              const double t = wr;
              wr = wi * h;
              wi = -t * h;
            }
            x = make_double2(x.x * wr - x.y * wi, x.x * wi + x.y * wr);
          }
        }
        if (r0 + row < D) m = M[(r0 + row) * D + c];
      }
      Xs[k][row] = x;
      Ms[k][row] = m;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < DN_K; ++k) {
      c128 xv[4], mv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) xv[i] = Xs[k][ts * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) mv[j] = Ms[k][tr * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ar[i][j] = fma(mv[j].x, xv[i].x, fma(-mv[j].y, xv[i].y, ar[i][j]));
          ai[i][j] = fma(mv[j].x, xv[i].y, fma(mv[j].y, xv[i].x, ai[i][j]));
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long s = s0 + ts * 4 + i;
    if (s >= S) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long r = r0 + tr * 4 + j;
      if (r < D) out[s * D + r] = make_double2(ar[i][j], ai[i][j]);
    }
  }
}

extern "C" int pqc_dense_apply_batch(const pqc_c128* d_in, int64_t S, int n, const pqc_c128* d_M,
                                     const double* d_lambda, const double* d_theta,
                                     int64_t theta_stride, int deriv, pqc_c128* d_out,
                                     void* stream) {
  if (S <= 0) return 0;
  if (n < 1 || n > 13) PQC_FAIL(-1, "dense operators are limited to 13 qubits (a 2^n x 2^n matrix)");
  if (!d_in || !d_M || !d_out) PQC_FAIL(-1, "null buffer");
  if (d_in == d_out) PQC_FAIL(-1, "dense apply is out of place");
  if (d_lambda && !d_theta) PQC_FAIL(-1, "eigenvalues without angles");
  const long long D = 1ll << n;
  dim3 grid((unsigned)((D + DN_T - 1) / DN_T), (unsigned)((S + DN_T - 1) / DN_T));
  if ((S + DN_T - 1) / DN_T > 65535) PQC_FAIL(-1, "dense apply: batch too large; split it");
  k_dense_apply<<<grid, 256, 0, (cudaStream_t)stream>>>((const c128*)d_in, S, n, (const c128*)d_M,
                                                        d_lambda, d_theta, theta_stride, deriv,
                                                        (c128*)d_out);
  PQC_LAUNCH_CHECK();
  return 0;
}
