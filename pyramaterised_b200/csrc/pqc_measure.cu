// pqc_measure.cu -- capacity-measure kernels over batches of statevectors.
// Each entry point cites the reference lines it replaces (/root/reference/pyramaterised/).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "pqc_common.cuh"

// ---------------------------------------------------------------------------------
// Qobj.overlap: out[i] = <a_i|b_i>   (measure.py:52,58,135; circuit.py:141)
// One CTA per pair -> fixed summation order -> bitwise reproducible run to run.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_overlap(const c128* __restrict__ A, long long sa,
                                                 const c128* __restrict__ B, long long sb,
                                                 long long dim, c128* __restrict__ out) {
  __shared__ double red[32];
  const c128* a = A + (long long)blockIdx.x * sa;
  const c128* b = B + (long long)blockIdx.x * sb;
  double re = 0.0, im = 0.0;
  for (long long i = threadIdx.x; i < dim; i += 256) {
    const c128 x = a[i], y = b[i];
    re += x.x * y.x + x.y * y.y;
    im += x.x * y.y - x.y * y.x;
  }
  re = block_sum<256>(re, red);
  im = block_sum<256>(im, red);
  if (threadIdx.x == 0) out[blockIdx.x] = make_double2(re, im);
}

extern "C" int pqc_overlap_batch(const pqc_c128* d_a, int64_t stride_a, const pqc_c128* d_b,
                                 int64_t stride_b, int64_t dim, int64_t count, pqc_c128* d_out,
                                 void* stream) {
  if (count <= 0) return 0;
  if (count > 0x7fffffffLL) PQC_FAIL(-1, "too many overlaps in one call");
  k_overlap<<<(unsigned)count, 256, 0, (cudaStream_t)stream>>>(
      (const c128*)d_a, stride_a, (const c128*)d_b, stride_b, dim, (c128*)d_out);
  PQC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------
// Meyer-Wallach (measure.py:226-249): per qubit k the 1-qubit reduced density matrix
// rho_k = [[r00, r01],[conj r01, r11]] and Q = 2 (1 - 1/n sum_k Tr rho_k^2),
// Tr rho^2 = r00^2 + r11^2 + 2|r01|^2.
// acc layout [S][n][4] = (r00, r11, Re r01, Im r01).
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_mw_accumulate(const c128* __restrict__ states, int n,
                                                       int chunk_log2, double* __restrict__ acc) {
  __shared__ double red[32];
  const long long D = 1ll << n;
  const int chunks_log2 = n - chunk_log2;
  const long long s = blockIdx.x >> chunks_log2;
  const long long c0 = ((long long)blockIdx.x & ((1ll << chunks_log2) - 1)) << chunk_log2;
  const c128* psi = states + s * D;
  const long long csize = 1ll << chunk_log2;
  for (int b = 0; b < n; ++b) {
    const long long bit = 1ll << b;
    double r00 = 0, r11 = 0, xr = 0, xi = 0;
    for (long long i = threadIdx.x; i < csize; i += 256) {
      const long long x = c0 + i;
      if (x & bit) continue;
      const c128 a0 = psi[x], a1 = psi[x | bit];
      r00 += a0.x * a0.x + a0.y * a0.y;
      r11 += a1.x * a1.x + a1.y * a1.y;
      xr += a0.x * a1.x + a0.y * a1.y;      // a0 * conj(a1)
      xi += a0.y * a1.x - a0.x * a1.y;
    }
    r00 = block_sum<256>(r00, red);
    r11 = block_sum<256>(r11, red);
    xr = block_sum<256>(xr, red);
    xi = block_sum<256>(xi, red);
    if (threadIdx.x == 0) {
      double* o = acc + (s * n + (n - 1 - b)) * 4;   // store by QUBIT index
      if (chunks_log2 == 0) {
        o[0] = r00; o[1] = r11; o[2] = xr; o[3] = xi;
      } else {
        atomicAdd(o + 0, r00); atomicAdd(o + 1, r11); atomicAdd(o + 2, xr); atomicAdd(o + 3, xi);
      }
    }
  }
}

__global__ void k_mw_finalize(const double* __restrict__ acc, long long S, int n,
                              double* __restrict__ Q) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  double tot = 0.0;
  for (int k = 0; k < n; ++k) {
    const double* o = acc + (s * n + k) * 4;
    tot += o[0] * o[0] + o[1] * o[1] + 2.0 * (o[2] * o[2] + o[3] * o[3]);
  }
  Q[s] = 2.0 * (1.0 - (1.0 / n) * tot);
}

// ---------------------------------------------------------------------------------
// Meyer-Wallach for n > 14: every qubit's (r00, Re r01, Im r01) from ONE read of the state per
// group of tile bits -- all 12 bits of a 4096-amplitude tile in pass 0, 8 more bits per further
// pass (3 reads at 28 qubits instead of the n / 2 of k_mw_accumulate) -- and bitwise
// reproducible: a CTA walks a fixed range of tiles of one state, every tile's 16 values per
// sweep are reduced over the warp by a fixed halving exchange and accumulated per lane, CTA and
// chunk sums are added in a fixed order (no floating-point atomics).
// Sweep A holds tile positions 8-11 in registers (loaded straight from global memory, lanes on
// amplitude bits 0-3), B positions 0-3, C positions 4-7 (through swizzled shared memory).
// r11 = N - r00 with the norm N taken in pass 0.
// ---------------------------------------------------------------------------------
struct MWPass {
  int tb[12];                // amplitude bit of tile position p
  int ob[PQC_MAX_QUBITS];    // amplitude bits outside the tile, ascending
  int doB, doC;              // sweeps B / C carry target bits (A always does)
};

__device__ __forceinline__ uint32_t mw_swz(uint32_t i) {
  return i ^ (((i >> 3) ^ (i >> 6) ^ (i >> 9)) & 7u);
}

// (r00, Re r01, Im r01) of register bit K over the thread's 16 amplitudes
template <int K>
__device__ __forceinline__ void mw_bit(const c128 (&a)[16], double& r00, double& xr, double& xi) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j & (1 << K)) continue;
    const c128 x = a[j], y = a[j | (1 << K)];
    s0 = fma(x.x, x.x, fma(x.y, x.y, s0));
    s1 = fma(x.x, y.x, fma(x.y, y.y, s1));      // a0 * conj(a1)
    s2 = fma(x.y, y.x, fma(-x.x, y.y, s2));
  }
  r00 = s0; xr = s1; xi = s2;
}

// warp reduce-scatter of 16 values: afterwards lane l holds the warp total of value (l >> 1) & 15
__device__ __forceinline__ double mw_reduce16(double (&v)[16], int lane) {
#pragma unroll
  for (int step = 0; step < 4; ++step) {
    const int m = 16 >> step, half = 8 >> step;      // lane-xor distance, values kept
    const bool hi = (lane & m) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < half) {
        const double send = hi ? v[i] : v[half + i];
        const double keep = hi ? v[half + i] : v[i];
        v[i] = keep + __shfl_xor_sync(0xffffffffu, send, m);
      }
    }
  }
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <bool DOB, bool DOC>
__global__ void __launch_bounds__(256, 2) k_mw_tiles(const c128* __restrict__ states, int n,
                                                     const MWPass P, int tiles_per_cta, int cps,
                                                     double* __restrict__ part) {
  extern __shared__ __align__(16) c128 mw_sm[];
  __shared__ double wred[3][8][16];
  const int tid = threadIdx.x, lo = tid & 15, hi = tid >> 4, lane = tid & 31, warp = tid >> 5;
  const long long s = blockIdx.x / cps;
  const int chunk = blockIdx.x % cps;
  const int tiles_log2 = n - 12;
  const c128* psi = states + (s << n);
  uint32_t ampA = (uint32_t)lo;
#pragma unroll
  for (int i = 0; i < 4; ++i) ampA |= (((uint32_t)hi >> i) & 1u) << P.tb[4 + i];
  const uint32_t g0 = 1u << P.tb[8], g1 = 1u << P.tb[9], g2 = 1u << P.tb[10], g3 = 1u << P.tb[11];
  const uint32_t sbA = mw_swz((uint32_t)tid), sbB = mw_swz((uint32_t)tid << 4),
                 sbC = mw_swz((uint32_t)lo | ((uint32_t)hi << 8));
  double accA = 0.0, accB = 0.0, accC = 0.0;
#define MW_SEL4(j, a0, a1, a2, a3) \
  ((((j)&1) ? (a0) : 0u) | (((j)&2) ? (a1) : 0u) | (((j)&4) ? (a2) : 0u) | (((j)&8) ? (a3) : 0u))
#define MW_CA(j) ((((j) << 8) ^ ((((j) << 2) ^ ((j) >> 1)) & 7)))   /* swz(j << 8) */
#define MW_CB(j) (((j) ^ ((j) >> 3)))                               /* swz(j)      */
#define MW_CC(j) ((((j) << 4) ^ ((((j) << 1) ^ ((j) >> 2)) & 7)))   /* swz(j << 4) */
  for (int ti = 0; ti < tiles_per_cta; ++ti) {
    const uint32_t tile = (uint32_t)chunk * (uint32_t)tiles_per_cta + (uint32_t)ti;
    uint32_t tbase = 0;
    for (int j = 0; j < tiles_log2; ++j) tbase |= ((tile >> j) & 1u) << P.ob[j];
    c128 a[16];
    {
      const c128* sp = psi + (tbase | ampA);
#pragma unroll
      for (int j = 0; j < 16; ++j) a[j] = sp[MW_SEL4(j, g0, g1, g2, g3)];
    }
    double v[16];
    {
      mw_bit<0>(a, v[0], v[1], v[2]);
      mw_bit<1>(a, v[3], v[4], v[5]);
      mw_bit<2>(a, v[6], v[7], v[8]);
      mw_bit<3>(a, v[9], v[10], v[11]);
      // the thread's share of the norm: r00 + r11 of register bit 0
      double nr = v[0];
#pragma unroll
      for (int j = 1; j < 16; j += 2) nr = fma(a[j].x, a[j].x, fma(a[j].y, a[j].y, nr));
      v[12] = nr;
      v[13] = v[14] = v[15] = 0.0;
      accA += mw_reduce16(v, lane);
    }
    if (DOB || DOC) {
      if (ti > 0) __syncthreads();                // the previous tile's readers are done
#pragma unroll
      for (int j = 0; j < 16; ++j) mw_sm[sbA ^ MW_CA(j)] = a[j];
      __syncthreads();
      if (DOB) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = mw_sm[sbB ^ MW_CB(j)];
        mw_bit<0>(a, v[0], v[1], v[2]);
        mw_bit<1>(a, v[3], v[4], v[5]);
        mw_bit<2>(a, v[6], v[7], v[8]);
        mw_bit<3>(a, v[9], v[10], v[11]);
        v[12] = v[13] = v[14] = v[15] = 0.0;
        accB += mw_reduce16(v, lane);
      }
      if (DOC) {
#pragma unroll
        for (int j = 0; j < 16; ++j) a[j] = mw_sm[sbC ^ MW_CC(j)];
        mw_bit<0>(a, v[0], v[1], v[2]);
        mw_bit<1>(a, v[3], v[4], v[5]);
        mw_bit<2>(a, v[6], v[7], v[8]);
        mw_bit<3>(a, v[9], v[10], v[11]);
        v[12] = v[13] = v[14] = v[15] = 0.0;
        accC += mw_reduce16(v, lane);
      }
    }
  }
#undef MW_CA
#undef MW_CB
#undef MW_CC
  // warps in a fixed order; lane 2 i of a warp holds value i
  if ((lane & 1) == 0) {
    wred[0][warp][lane >> 1] = accA;
    wred[1][warp][lane >> 1] = accB;
    wred[2][warp][lane >> 1] = accC;
  }
  __syncthreads();
  if (tid < 48) {
    const int sw = tid >> 4, i = tid & 15;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += wred[sw][w][i];
    part[((long long)blockIdx.x) * 48 + tid] = t;
  }
}

struct MWFin {
  int npass, cps;
  int src_pass[PQC_MAX_QUBITS];   // per amplitude bit: pass and value slot (sweep * 16 + 3 k)
  int src_slot[PQC_MAX_QUBITS];
};

// part: [pass][state][chunk][48]; chunks added in order; Q = 2 (1 - 1/n sum_k Tr rho_k^2)
__global__ void k_mw_tiles_fin(const double* __restrict__ part, long long S, int n, const MWFin F,
                               double* __restrict__ Q, double* __restrict__ acc_out) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= S) return;
  auto total = [&](int pass, int slot) {
    const double* p = part + (((long long)pass * S + s) * F.cps) * 48 + slot;
    double t = 0.0;
    for (int c = 0; c < F.cps; ++c) t += p[(long long)c * 48];
    return t;
  };
  const double N = total(0, 12);
  double tot = 0.0;
  for (int b = 0; b < n; ++b) {
    const double r00 = total(F.src_pass[b], F.src_slot[b]);
    const double xr = total(F.src_pass[b], F.src_slot[b] + 1);
    const double xi = total(F.src_pass[b], F.src_slot[b] + 2);
    const double r11 = N - r00;
    tot += r00 * r00 + r11 * r11 + 2.0 * (xr * xr + xi * xi);
    if (acc_out) {                               // k_mw_accumulate's layout, by QUBIT index
      double* o = acc_out + (s * n + (n - 1 - b)) * 4;
      o[0] = r00; o[1] = r11; o[2] = xr; o[3] = xi;
    }
  }
  if (Q) Q[s] = 2.0 * (1.0 - (1.0 / n) * tot);
}

static bool mw_tiles_enabled() {                 // PQC_MW=generic keeps k_mw_accumulate
  const char* e = getenv("PQC_MW");
  return !(e && strcmp(e, "generic") == 0);
}

// d_Q and / or d_acc ([S][n][4], the layout of k_mw_accumulate) may be null
static int mw_tiles(const c128* d_states, long long S, int n, double* d_Q, double* d_acc,
                    cudaStream_t st) {
  const int T = 1 << (n - 12);
  // chunks per state: enough CTAs to fill the GPU twice over, a power of two dividing T
  int cps = 1;
  while (cps < T && S * cps < 2 * 148 * 2) cps *= 2;
  const int tpc = T / cps;
  if (S * cps > 0x7fffffffLL) PQC_FAIL(-1, "Meyer-Wallach grid too large; split the batch");
  MWPass passes[4];
  MWFin fin;
  memset(&fin, 0, sizeof(fin));
  fin.cps = cps;
  int np = 0;
  {
    // pass 0: bits 0-11; later passes: bits 0-3 and the next 8 (or 4, padded) bits
    int next = 0;
    while (next < n) {
      if (np >= 4) PQC_FAIL(-5, "internal: too many Meyer-Wallach passes");
      MWPass& p = passes[np];
      memset(&p, 0, sizeof(p));
      bool in[32] = {false};
      if (np == 0) {
        for (int i = 0; i < 12; ++i) { p.tb[i] = i; in[i] = true; }
        p.doB = p.doC = 1;
        for (int i = 0; i < 12; ++i) {
          const int sw = i >= 8 ? 0 : (i < 4 ? 1 : 2), k = i & 3;
          fin.src_pass[i] = 0;
          fin.src_slot[i] = sw * 16 + 3 * k;
        }
        next = 12;
      } else {
        const int left = n - next;
        for (int i = 0; i < 4; ++i) { p.tb[i] = i; in[i] = true; }
        const int nA = std::min(4, left);          // new bits on positions 8-11 first
        int filler = 4;                            // bits 4-11 were measured in pass 0
        auto take_filler = [&]() {
          while (in[filler]) ++filler;
          in[filler] = true;
          return filler;
        };
        // positions must ascend with the amplitude bit inside the tile only for readability; any
        // assignment works because every address is built from per-position masks
        for (int k = 0; k < 4; ++k) {
          if (k < nA) {
            p.tb[8 + k] = next + k; in[next + k] = true;
            fin.src_pass[next + k] = np;
            fin.src_slot[next + k] = 0 * 16 + 3 * k;
          } else {
            p.tb[8 + k] = -1;
          }
        }
        const int nC = std::min(4, left - nA);
        for (int k = 0; k < 4; ++k) {
          if (k < nC) {
            p.tb[4 + k] = next + nA + k; in[next + nA + k] = true;
            fin.src_pass[next + nA + k] = np;
            fin.src_slot[next + nA + k] = 2 * 16 + 3 * k;
          } else {
            p.tb[4 + k] = -1;
          }
        }
        for (int i = 4; i < 12; ++i)
          if (p.tb[i] < 0) p.tb[i] = take_filler();
        p.doB = 0;
        p.doC = nC > 0 ? 1 : 0;
        next += nA + nC;
      }
      int o = 0;
      for (int b = 0; b < n; ++b)
        if (!in[b]) p.ob[o++] = b;
      ++np;
    }
  }
  fin.npass = np;
  double* part = nullptr;
  PQC_CUDA(cudaMallocAsync(&part, sizeof(double) * 48 * (size_t)np * S * cps, st));
  static bool attr[64] = {false};
  int dev = 0;
  PQC_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && !attr[dev]) {
    PQC_CUDA(cudaFuncSetAttribute(k_mw_tiles<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    PQC_CUDA(cudaFuncSetAttribute(k_mw_tiles<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    PQC_CUDA(cudaFuncSetAttribute(k_mw_tiles<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    attr[dev] = true;
  }
  int rc = 0;
  for (int q = 0; q < np && rc == 0; ++q) {
    double* pp = part + (size_t)q * S * cps * 48;
    const unsigned grid = (unsigned)(S * cps);
    if (passes[q].doB)
      k_mw_tiles<true, true><<<grid, 256, 65536, st>>>(d_states, n, passes[q], tpc, cps, pp);
    else if (passes[q].doC)
      k_mw_tiles<false, true><<<grid, 256, 65536, st>>>(d_states, n, passes[q], tpc, cps, pp);
    else
      k_mw_tiles<false, false><<<grid, 256, 0, st>>>(d_states, n, passes[q], tpc, cps, pp);
    ++g_pqc_launches;
    if (cudaGetLastError() != cudaSuccess) rc = -2;
  }
  if (rc == 0) {
    k_mw_tiles_fin<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(part, S, n, fin, d_Q, d_acc);
    ++g_pqc_launches;
    if (cudaGetLastError() != cudaSuccess) rc = -2;
  }
  cudaFreeAsync(part, st);
  if (rc) pqc_set_error("Meyer-Wallach tile kernel launch failed");
  return rc;
}

// ---------------------------------------------------------------------------------
// Meyer-Wallach for n <= 11 (BASELINE configs 1 and 2: 4 and 10 qubits, up to 1e5 states): one
// CTA of 128 threads per state.  The state is read ONCE into shared memory; every thread then
// accumulates (r00, Re r01, Im r01) of four qubits at a time over its share of the amplitude
// pairs, the 16 values of a group are reduced over the warp by the fixed halving exchange of
// k_mw_tiles and over the four warps in order.  Output in k_mw_tiles' partial layout (48 values
// per state: group g = bits 4g .. 4g+3 at 16 g + 3 k, the norm at 12), finished by
// k_mw_tiles_fin -- bitwise reproducible.  (k_mw_accumulate, which this replaces for small n,
// re-read the state per qubit and ran four block reductions per qubit: 3.3 ms per 1e5 states of
// 10 qubits.)
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_mw_small(const c128* __restrict__ states, int n,
                                                  double* __restrict__ part) {
  extern __shared__ __align__(16) c128 ms_sm[];
  __shared__ double wred[3][4][16];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int D = 1 << n, half = D >> 1;
  const c128* psi = states + ((long long)blockIdx.x << n);
  double nr = 0.0;
  for (int i = tid; i < D; i += 128) {
    const c128 v = psi[i];
    ms_sm[i] = v;
    nr = fma(v.x, v.x, fma(v.y, v.y, nr));
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    double v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0.0;
    if (4 * g < n) {                                   // uniform over the grid
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int b = 4 * g + k;
        if (b < n) {
          const int lowmask = (1 << b) - 1;
          double s0 = 0.0, s1 = 0.0, s2 = 0.0;
          for (int p = tid; p < half; p += 128) {
            const int x = ((p & ~lowmask) << 1) | (p & lowmask);
            const c128 a0 = ms_sm[x], a1 = ms_sm[x | (1 << b)];
            s0 = fma(a0.x, a0.x, fma(a0.y, a0.y, s0));
            s1 = fma(a0.x, a1.x, fma(a0.y, a1.y, s1));      // a0 * conj(a1)
            s2 = fma(a0.y, a1.x, fma(-a0.x, a1.y, s2));
          }
          v[3 * k] = s0; v[3 * k + 1] = s1; v[3 * k + 2] = s2;
        }
      }
      if (g == 0) v[12] = nr;
    }
    const double r = mw_reduce16(v, lane);             // lane 2 i holds value i of this warp
    if ((lane & 1) == 0) wred[g][warp][lane >> 1] = r;
  }
  __syncthreads();
  if (tid < 48) {
    const int g = tid >> 4, i = tid & 15;
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 4; ++w) t += wred[g][w][i];
    part[(long long)blockIdx.x * 48 + tid] = t;
  }
}

static int mw_small(const c128* d_states, long long S, int n, double* d_Q, double* d_acc,
                    cudaStream_t st) {
  if (S > 0x7fffffffLL) PQC_FAIL(-1, "Meyer-Wallach grid too large; split the batch");
  MWFin fin;
  memset(&fin, 0, sizeof(fin));
  fin.npass = 1;
  fin.cps = 1;
  for (int b = 0; b < n; ++b) {
    fin.src_pass[b] = 0;
    fin.src_slot[b] = (b >> 2) * 16 + 3 * (b & 3);
  }
  double* part = nullptr;
  PQC_CUDA(cudaMallocAsync(&part, sizeof(double) * 48 * (size_t)S, st));
  int rc = 0;
  k_mw_small<<<(unsigned)S, 128, sizeof(c128) << n, st>>>(d_states, n, part);
  ++g_pqc_launches;
  if (cudaGetLastError() != cudaSuccess) rc = -2;
  if (rc == 0) {
    k_mw_tiles_fin<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(part, S, n, fin, d_Q, d_acc);
    ++g_pqc_launches;
    if (cudaGetLastError() != cudaSuccess) rc = -2;
  }
  cudaFreeAsync(part, st);
  if (rc) pqc_set_error("Meyer-Wallach small-state kernel launch failed");
  return rc;
}

// which Meyer-Wallach path: n >= 12 the tile kernel, n <= 11 the one-CTA-per-state kernel;
// PQC_MW=generic keeps k_mw_accumulate (the independent cross-check of the tests)
static int mw_dispatch(const c128* d_states, long long S, int n, double* d_Q, double* d_acc,
                       cudaStream_t st, bool* handled) {
  *handled = mw_tiles_enabled();
  if (!*handled) return 0;
  return n >= 12 ? mw_tiles(d_states, S, n, d_Q, d_acc, st) : mw_small(d_states, S, n, d_Q, d_acc, st);
}

static int mw_accumulate(const c128* d_states, long long S, int n, double* d_acc, cudaStream_t st) {
  const int chunk_log2 = std::min(n, 14);
  const long long grid = S << (n - chunk_log2);
  if (grid > 0x7fffffffLL) PQC_FAIL(-1, "Meyer-Wallach grid too large; split the batch");
  if (n > chunk_log2) PQC_CUDA(cudaMemsetAsync(d_acc, 0, sizeof(double) * 4 * n * S, st));
  k_mw_accumulate<<<(unsigned)grid, 256, 0, st>>>(d_states, n, chunk_log2, d_acc);
  PQC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pqc_meyer_wallach(const pqc_c128* d_states, int64_t S, int n, double* d_Q,
                                 void* stream) {
  if (S <= 0) return 0;
  if (n < 1 || n > PQC_MAX_QUBITS) PQC_FAIL(-1, "bad qubit count");
  cudaStream_t st = (cudaStream_t)stream;
  {
    bool handled = false;
    const int rc = mw_dispatch((const c128*)d_states, S, n, d_Q, nullptr, st, &handled);
    if (handled) return rc;
  }
  double* acc = nullptr;
  PQC_CUDA(cudaMallocAsync(&acc, sizeof(double) * 4 * n * S, st));
  int rc = mw_accumulate((const c128*)d_states, S, n, acc, st);
  if (rc == 0) {
    k_mw_finalize<<<(unsigned)((S + 127) / 128), 128, 0, st>>>(acc, S, n, d_Q);
    ++g_pqc_launches;
    if (cudaGetLastError() != cudaSuccess) rc = -2;
  }
  cudaFreeAsync(acc, st);
  return rc;
}

__global__ void k_rho_from_acc(const double* __restrict__ acc, int n, int qubit,
                               c128* __restrict__ rho) {
  const double* o = acc + qubit * 4;
  rho[0] = make_double2(o[0], 0.0);
  rho[1] = make_double2(o[2], o[3]);
  rho[2] = make_double2(o[2], -o[3]);
  rho[3] = make_double2(o[1], 0.0);
}

extern "C" int pqc_ptrace_1q(const pqc_c128* d_state, int n, int qubit, pqc_c128* d_rho,
                             void* stream) {
  if (qubit < 0 || qubit >= n) PQC_FAIL(-1, "ptrace qubit out of range");
  cudaStream_t st = (cudaStream_t)stream;
  double* acc = nullptr;
  PQC_CUDA(cudaMallocAsync(&acc, sizeof(double) * 4 * n, st));
  bool handled = false;
  int rc = mw_dispatch((const c128*)d_state, 1, n, nullptr, acc, st, &handled);
  if (!handled) rc = mw_accumulate((const c128*)d_state, 1, n, acc, st);
  if (rc == 0) {
    k_rho_from_acc<<<1, 1, 0, st>>>(acc, n, qubit, (c128*)d_rho);
    ++g_pqc_launches;
    if (cudaGetLastError() != cudaSuccess) rc = -2;
  }
  cudaFreeAsync(acc, st);
  return rc;
}

// ---------------------------------------------------------------------------------
// Pauli sums: H|psi> and <psi|H|psi>  (circuit.py:28-31,132-137; measure.py:468; and the
// generator sums of gates.py:454-457 used for derivative states).
// (P psi)[y] = i^{ny} (-1)^{popcount((y^xm) & zm)} psi[y ^ xm]
// ---------------------------------------------------------------------------------
__device__ __forceinline__ c128 pauli_sum_at(const c128* __restrict__ psi, long long y,
                                             const GenTerm* __restrict__ terms, int nterms) {
  double re = 0.0, im = 0.0;
  for (int t = 0; t < nterms; ++t) {
    const GenTerm g = terms[t];
    const long long x = y ^ (long long)g.xmask;
    const c128 v = psi[x];
    int ph = __popc(g.xmask & g.zmask) + 2 * __popc((uint32_t)x & g.zmask);   // power of i
    ph &= 3;
    c128 w;
    if (ph == 0) w = v;
    else if (ph == 1) w = make_double2(-v.y, v.x);
    else if (ph == 2) w = make_double2(-v.x, -v.y);
    else w = make_double2(v.y, -v.x);
    re += g.re * w.x - g.im * w.y;
    im += g.re * w.y + g.im * w.x;
  }
  return make_double2(re, im);
}

// dst item i <- sum_t coef_t P_t src item i.  Items are (sample, slot) pairs addressed
// like the tile kernel: offset = (sample*slots_total + slot) << n.
__global__ void __launch_bounds__(256) k_pauli_apply(const c128* __restrict__ src, c128* __restrict__ dst,
                                                     int n, long long n_samples, int slots_total,
                                                     int src_slot, int dst_slot,
                                                     const GenTerm* __restrict__ terms, int nterms) {
  const long long D = 1ll << n;
  const long long total = n_samples * D;
  for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < total;
       g += (long long)gridDim.x * 256) {
    const long long s = g >> n, y = g & (D - 1);
    const c128* psi = src + ((s * slots_total + src_slot) << n);
    dst[((s * slots_total + dst_slot) << n) + y] = pauli_sum_at(psi, y, terms, nterms);
  }
}

int pqc_pauli_apply_slots(const c128* src, c128* dst, int n, long long S, int slots_total,
                          int src_slot, int dst_slot, const GenTerm* d_terms, int nterms,
                          cudaStream_t st) {
  if (S <= 0) return 0;
  const long long total = S << n;
  const long long grid = std::min<long long>((total + 255) / 256, 148 * 32);
  k_pauli_apply<<<(unsigned)grid, 256, 0, st>>>(src, dst, n, S, slots_total, src_slot, dst_slot,
                                                d_terms, nterms);
  PQC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------
// fSim-family derivative spawn: dst <- M psi with M = diag(1, [[d, o],[o, d]], corner) on the
// pair (b0, b1): the element-wise "d/dtheta" / "d/dphi" matrices of gates.py:609-648,719-737
// (quirk Q3: [0,0] stays 1 and [3,3] is not differentiated).  M commutes with the fSim gate
// (both are polynomials of the same 2x2 block), so it may be applied right after the gate.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pair_spawn(c128* __restrict__ buf, int n, long long S,
                                                    int slots_total, int dst_slot, ParamSpawn ps,
                                                    const double* __restrict__ angles, long long ld) {
  const long long D = 1ll << n, total = S * D;
  const long long ma = 1ll << ps.b0, mb = 1ll << ps.b1;
  for (long long g = (long long)blockIdx.x * 256 + threadIdx.x; g < total;
       g += (long long)gridDim.x * 256) {
    const long long s = g >> n, x = g & (D - 1);
    const c128* psi = buf + ((s * slots_total) << n);
    double th = ps.offset + (ps.p_theta >= 0 ? angles[s * ld + ps.p_theta] : 0.0);
    const double ph = ps.which == 3 ? 0.0 : (ps.p_phi >= 0 ? angles[s * ld + ps.p_phi] : ps.phi_fixed);
    double sn, cs, sp, cp;
    sincos(th, &sn, &cs);
    sincos(ph, &sp, &cp);
    const bool ba = (x & ma) != 0, bb = (x & mb) != 0;
    const c128 v = psi[x];
    c128 out;
    if (!ba && !bb) {
      out = v;
    } else if (ba && bb) {
      // corner: e^{-i phi} (which 1), -i e^{-i phi} (which 2), 1 (which 3)
      const double cr = ps.which == 3 ? 1.0 : (ps.which == 1 ? cp : -sp);
      const double ci = ps.which == 3 ? 0.0 : (ps.which == 1 ? -sp : -cp);
      out = make_double2(v.x * cr - v.y * ci, v.y * cr + v.x * ci);
    } else {
      const c128 w = psi[x ^ ma ^ mb];
      // which 2: diag cos, off -i sin;  which 1 / 3: diag -sin, off -i cos
      const double d = ps.which == 2 ? cs : -sn, o = ps.which == 2 ? sn : cs;
      out = make_double2(d * v.x + o * w.y, d * v.y - o * w.x);
    }
    buf[((s * slots_total + dst_slot) << n) + x] = out;
  }
}

int pqc_pair_spawn(c128* buf, int n, long long S, int slots_total, int dst_slot,
                   const ParamSpawn& ps, const double* d_angles, long long ld, cudaStream_t st) {
  const long long total = S << n;
  const long long grid = std::min<long long>((total + 255) / 256, 148 * 32);
  k_pair_spawn<<<(unsigned)grid, 256, 0, st>>>(buf, n, S, slots_total, dst_slot, ps, d_angles, ld);
  PQC_LAUNCH_CHECK();
  return 0;
}

__global__ void __launch_bounds__(256) k_pauli_expect(const c128* __restrict__ states, int n,
                                                      const GenTerm* __restrict__ terms, int nterms,
                                                      c128* __restrict__ out) {
  __shared__ double red[32];
  const long long D = 1ll << n;
  const c128* psi = states + (long long)blockIdx.x * D;
  double re = 0.0, im = 0.0;
  for (long long y = threadIdx.x; y < D; y += 256) {
    const c128 h = pauli_sum_at(psi, y, terms, nterms);
    const c128 a = psi[y];
    re += a.x * h.x + a.y * h.y;
    im += a.x * h.y - a.y * h.x;
  }
  re = block_sum<256>(re, red);
  im = block_sum<256>(im, red);
  if (threadIdx.x == 0) out[blockIdx.x] = make_double2(re, im);
}

static int upload_terms(const pqc_pauli_term* h_terms, int n_terms, GenTerm** d_terms,
                        cudaStream_t st) {
  std::vector<GenTerm> t(n_terms);
  for (int i = 0; i < n_terms; ++i) {
    t[i].xmask = h_terms[i].xmask;
    t[i].zmask = h_terms[i].zmask;
    t[i].re = h_terms[i].re;
    t[i].im = h_terms[i].im;
  }
  PQC_CUDA(cudaMallocAsync(d_terms, sizeof(GenTerm) * std::max(1, n_terms), st));
  // pageable source: the copy is staged before the call returns, so `t` may die here
  PQC_CUDA(cudaMemcpyAsync(*d_terms, t.data(), sizeof(GenTerm) * n_terms, cudaMemcpyHostToDevice,
                           st));
  return 0;
}

extern "C" int pqc_pauli_expect_batch(const pqc_c128* d_states, int64_t S, int n, int n_terms,
                                      const pqc_pauli_term* h_terms, pqc_c128* d_out,
                                      void* stream) {
  if (S <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GenTerm* d_terms = nullptr;
  int rc = upload_terms(h_terms, n_terms, &d_terms, st);
  if (rc) return rc;
  k_pauli_expect<<<(unsigned)S, 256, 0, st>>>((const c128*)d_states, n, d_terms, n_terms,
                                              (c128*)d_out);
  ++g_pqc_launches;
  if (cudaGetLastError() != cudaSuccess) rc = -2;
  cudaFreeAsync(d_terms, st);
  return rc;
}

extern "C" int pqc_pauli_apply_batch(const pqc_c128* d_states, int64_t S, int n, int n_terms,
                                     const pqc_pauli_term* h_terms, pqc_c128* d_out,
                                     void* stream) {
  if (S <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  GenTerm* d_terms = nullptr;
  int rc = upload_terms(h_terms, n_terms, &d_terms, st);
  if (rc) return rc;
  rc = pqc_pauli_apply_slots((const c128*)d_states, (c128*)d_out, n, S, 1, 0, 0, d_terms,
                             n_terms, st);
  cudaFreeAsync(d_terms, st);
  return rc;
}

// ---------------------------------------------------------------------------------
// np.histogram(F, bins=B, range=(0,1)) binning (measure.py:153-155), restated exactly:
// idx = int(F*B); idx==B -> B-1; then the two edge corrections numpy applies against
// edges = linspace(0,1,B+1) (edge_i = i*(1/B), edge_B = 1); values outside [0,1] dropped.
// ---------------------------------------------------------------------------------
__device__ __forceinline__ long long np_hist_bin(double f, long long B, double step) {
  if (!(f >= 0.0) || !(f <= 1.0)) return -1;
  long long idx = (long long)(f * (double)B);
  if (idx == B) idx = B - 1;
  const double lo = __dmul_rn((double)idx, step);
  if (f < lo) {
    idx -= 1;
  } else if (idx != B - 1) {
    const double hi = (idx + 1 == B) ? 1.0 : __dmul_rn((double)(idx + 1), step);
    if (f >= hi) idx += 1;
  }
  return idx;
}

__global__ void k_hist_f64(const double* __restrict__ F, long long count, long long B,
                           double step, unsigned long long* __restrict__ hist) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += (long long)gridDim.x * blockDim.x) {
    const long long b = np_hist_bin(F[i], B, step);
    if (b >= 0) atomicAdd(hist + b, 1ull);
  }
}

extern "C" int pqc_hist_f64(const double* d_F, int64_t count, int64_t bins, long long* d_hist,
                            void* stream) {
  if (bins <= 0) PQC_FAIL(-1, "`bins` must be positive, when an integer");
  if (count <= 0) return 0;
  const long long grid = std::min<long long>((count + 255) / 256, 148 * 16);
  k_hist_f64<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      d_F, count, bins, 1.0 / (double)bins, (unsigned long long*)d_hist);
  PQC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------
// pairwise fidelities + histogram (measure.py:123-159).  v0: FP64 CUDA-core tiled
// ZHERK-style kernel, 64x64 pair tile per CTA, 4x4 complex accumulators per thread,
// K staged through shared memory 16 amplitudes at a time.
// ---------------------------------------------------------------------------------
#define FT 64
#define FK 16
__global__ void __launch_bounds__(256) k_fidelity(const c128* __restrict__ A, long long SA,
                                                  const c128* __restrict__ Bm, long long SB,
                                                  int n, int triangular, long long bins,
                                                  double step, unsigned long long* __restrict__ hist,
                                                  double* __restrict__ Fout) {
  // block -> (bi, bj) tile coordinates
  long long bi, bj;
  const long long nbj = (SB + FT - 1) / FT;
  if (triangular) {
    // linear index over the upper triangle of the tile grid (bj >= bi)
    long long t = blockIdx.x;
    const long long nb = nbj;
    // row bi has (nb - bi) tiles; solve by search (few iterations in double then fix up)
    double disc = (2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)t;
    bi = (long long)(((2.0 * nb + 1.0) - sqrt(disc)) * 0.5);
    if (bi < 0) bi = 0;
    while (bi > 0 && bi * nb - bi * (bi - 1) / 2 > t) --bi;
    while ((bi + 1) * nb - (bi + 1) * bi / 2 <= t) ++bi;
    bj = bi + (t - (bi * nb - bi * (bi - 1) / 2));
  } else {
    bi = blockIdx.x / nbj;
    bj = blockIdx.x % nbj;
  }
  __shared__ c128 sa[FK][FT + 1];
  __shared__ c128 sb[FK][FT + 1];
  const long long D = 1ll << n;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double are[4][4], aim[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) are[r][c] = aim[r][c] = 0.0;

  for (long long k0 = 0; k0 < D; k0 += FK) {
    // each thread loads 4 elements of each operand tile: FT rows x FK cols = 1024 elements
    for (int e = threadIdx.x; e < FT * FK; e += 256) {
      const int r = e / FK, kk = e % FK;
      const long long ia = bi * FT + r, ib = bj * FT + r;
      const long long k = k0 + kk;
      c128 va = make_double2(0, 0), vb = make_double2(0, 0);
      if (k < D) {
        if (ia < SA) va = A[ia * D + k];
        if (ib < SB) vb = Bm[ib * D + k];
      }
      sa[kk][r] = va;
      sb[kk][r] = vb;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < FK; ++kk) {
      c128 ra[4], rb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) ra[r] = sa[kk][ty * 4 + r];
#pragma unroll
      for (int c = 0; c < 4; ++c) rb[c] = sb[kk][tx * 4 + c];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          // conj(a) * b
          are[r][c] += ra[r].x * rb[c].x + ra[r].y * rb[c].y;
          aim[r][c] += ra[r].x * rb[c].y - ra[r].y * rb[c].x;
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const long long i = bi * FT + ty * 4 + r, j = bj * FT + tx * 4 + c;
      if (i >= SA || j >= SB) continue;
      if (triangular && j <= i) continue;
      const double re = are[r][c], im = aim[r][c];
      const double mag = hypot(re, im);               // np.abs(z) ** 2 (measure.py:135)
      const double f = mag * mag;
      if (hist) {
        const long long b = np_hist_bin(f, bins, step);
        if (b >= 0) atomicAdd(hist + b, 1ull);
      }
      if (Fout) {
        const long long idx = triangular ? (i * (2 * SA - i - 1) / 2 + (j - i - 1)) : (i * SB + j);
        Fout[idx] = f;
      }
    }
}

// ---------------------------------------------------------------------------------
// FP64 tensor-core version of the pair-fidelity block (the one genuinely dense contraction
// of the path): C = conj(A) B^T as four real DMMA streams per 8x8x4 step,
//   Re C = Ar Br^T + Ai Bi^T,   Im C = Ar Bi^T - Ai Br^T
// (mma.sync.aligned.m8n8k4.f64 -- tcgen05 has no FP64 kind).  64x64 pair tile per CTA, 8
// warps of 32x16 pairs, K streamed 16 amplitudes at a time through cp.async double-buffered
// shared memory (row stride 320 B so the 16-byte fragment loads are bank-conflict free).
// Epilogue identical to k_fidelity: np.abs(z)**2, numpy-exact binning, int64 atomics.
// ---------------------------------------------------------------------------------
#define DT 64
#define DK 16
#define DROW 20                      // complex per padded row (16 + 4)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(sa), "l"(gmem));
}

__global__ void __launch_bounds__(256, 2) k_fidelity_dmma(
    const c128* __restrict__ A, long long SA, const c128* __restrict__ Bm, long long SB, int n,
    int triangular, long long bins, double step, unsigned long long* __restrict__ hist,
    double* __restrict__ Fout) {
  extern __shared__ __align__(16) unsigned char smraw[];
  c128* sA = reinterpret_cast<c128*>(smraw);                 // [2][DT][DROW]
  c128* sB = sA + 2 * DT * DROW;
  long long bi, bj;
  const long long nbj = (SB + DT - 1) / DT;
  if (triangular) {
    const long long t = blockIdx.x, nb = nbj;
    const double disc = (2.0 * nb + 1.0) * (2.0 * nb + 1.0) - 8.0 * (double)t;
    bi = (long long)(((2.0 * nb + 1.0) - sqrt(disc)) * 0.5);
    if (bi < 0) bi = 0;
    while (bi > 0 && bi * nb - bi * (bi - 1) / 2 > t) --bi;
    while ((bi + 1) * nb - (bi + 1) * bi / 2 <= t) ++bi;
    bj = bi + (t - (bi * nb - bi * (bi - 1) / 2));
  } else {
    bi = blockIdx.x / nbj;
    bj = blockIdx.x % nbj;
  }
  const long long D = 1ll << n;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int wr = (warp >> 2) * 32, wc = (warp & 3) * 16;       // warp tile origin (rows x cols)

  // each thread moves 4 + 4 16-byte pieces per stage: row = e / 16, k = e % 16
  auto stage_load = [&](int buf, long long k0) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int e = tid + q * 256, r = e >> 4, kk = e & 15;
      long long ra = bi * DT + r, rb = bj * DT + r;
      if (ra >= SA) ra = SA - 1;                                // clamp: masked in the epilogue
      if (rb >= SB) rb = SB - 1;
      cp_async16(sA + (buf * DT + r) * DROW + kk, A + ra * D + k0 + kk);
      cp_async16(sB + (buf * DT + r) * DROW + kk, Bm + rb * D + k0 + kk);
    }
    asm volatile("cp.async.commit_group;");
  };

  double re[4][2][2], im[4][2][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j) re[i][j][0] = re[i][j][1] = im[i][j][0] = im[i][j][1] = 0.0;

  const long long nk = D / DK;
  stage_load(0, 0);
  for (long long ks = 0; ks < nk; ++ks) {
    const int buf = (int)(ks & 1);
    if (ks + 1 < nk) {
      stage_load(buf ^ 1, (ks + 1) * DK);
      asm volatile("cp.async.wait_group 1;");
    } else {
      asm volatile("cp.async.wait_group 0;");
    }
    __syncthreads();
    const c128* a_s = sA + buf * DT * DROW;
    const c128* b_s = sB + buf * DT * DROW;
#pragma unroll
    for (int kk = 0; kk < DK / 4; ++kk) {
      c128 fa[4], fb[2];
#pragma unroll
      for (int i = 0; i < 4; ++i) fa[i] = a_s[(wr + 8 * i + g) * DROW + kk * 4 + t4];
#pragma unroll
      for (int j = 0; j < 2; ++j) fb[j] = b_s[(wc + 8 * j + g) * DROW + kk * 4 + t4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          dmma(re[i][j][0], re[i][j][1], fa[i].x, fb[j].x);
          dmma(re[i][j][0], re[i][j][1], fa[i].y, fb[j].y);
          dmma(im[i][j][0], im[i][j][1], fa[i].x, fb[j].y);
          dmma(im[i][j][0], im[i][j][1], -fa[i].y, fb[j].x);
        }
    }
    __syncthreads();
  }
  // accumulator element c of thread (g, t4): row g, column 2 * t4 + c of its 8x8 tile
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const long long row = bi * DT + wr + 8 * i + g, col = bj * DT + wc + 8 * j + 2 * t4 + c;
        if (row >= SA || col >= SB) continue;
        if (triangular && col <= row) continue;
        const double mag = hypot(re[i][j][c], im[i][j][c]);
        const double f = mag * mag;
        if (hist) {
          const long long b = np_hist_bin(f, bins, step);
          if (b >= 0) atomicAdd(hist + b, 1ull);
        }
        if (Fout) {
          const long long idx = triangular ? (row * (2 * SA - row - 1) / 2 + (col - row - 1))
                                           : (row * SB + col);
          Fout[idx] = f;
        }
      }
}

// Split-K form for a few pairs of very long vectors: CTA (slice, pair) takes the partial inner
// product over its K slice; the slices are then added in a fixed order (reproducible).
__global__ void __launch_bounds__(256) k_fid_splitk(const c128* __restrict__ A, long long SA,
                                                    const c128* __restrict__ B, long long SB, int n,
                                                    int tri, int ksplit, double2* __restrict__ part) {
  __shared__ double red[32];
  long long p = blockIdx.y, i, j;
  if (tri) {                                  // itertools.combinations order
    i = 0;
    while (p >= SA - 1 - i) { p -= SA - 1 - i; ++i; }
    j = i + 1 + p;
  } else {
    i = p / SB;
    j = p - i * SB;
  }
  const long long D = 1ll << n, len = D / ksplit, k0 = (long long)blockIdx.x * len;
  const c128* a = A + i * D + k0;
  const c128* b = B + j * D + k0;
  double re = 0.0, im = 0.0;
  for (long long k = threadIdx.x; k < len; k += 256) {
    const c128 x = a[k], y = b[k];
    re += x.x * y.x + x.y * y.y;
    im += x.x * y.y - x.y * y.x;
  }
  re = block_sum<256>(re, red);
  im = block_sum<256>(im, red);
  if (threadIdx.x == 0) part[(long long)blockIdx.y * ksplit + blockIdx.x] = make_double2(re, im);
}

__global__ void k_fid_splitk_fin(const double2* __restrict__ part, long long npairs, int ksplit,
                                 long long bins, double step, unsigned long long* __restrict__ hist,
                                 double* __restrict__ F) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= npairs) return;
  double re = 0.0, im = 0.0;
  for (int k = 0; k < ksplit; ++k) {
    re += part[p * ksplit + k].x;
    im += part[p * ksplit + k].y;
  }
  const double f = re * re + im * im;
  if (F) F[p] = f;
  if (hist) {
    const long long b = np_hist_bin(f, bins, step);
    if (b >= 0) atomicAdd(hist + b, 1ull);
  }
}

// Blocked split-K form (config 5: a block of resident 4 GiB states against a few travelling
// ones): CTA (K slice, pair tile) streams its slice of FB_TA = 32 A-rows and FB_TB = 4 B-rows
// through a double-buffered shared-memory ring (cp.async, FB_KC amplitudes per row and stage) and
// every warp keeps a 4 x 4 tile of pair accumulators in registers, so each state is read ONCE per
// pair tile instead of once per pair: 16 * 2^n * (rows_A + rows_B) bytes per tile instead of
// 32 * 2^n per pair.  Lane partials are folded by a fixed shuffle tree and the K slices are added
// in order by k_fid_block_fin (reproducible; no floating-point atomics).
#define FB_TA 32
#define FB_TB 4
#define FB_KC 128
#define FB_ROWS (FB_TA + FB_TB)
#define FB_STAGE_BYTES (FB_ROWS * FB_KC * 16)

__global__ void __launch_bounds__(256, 1) k_fid_block(const c128* __restrict__ A, long long SA,
                                                      const c128* __restrict__ B, long long SB, int n,
                                                      int ksplit, long long tiles_b,
                                                      double2* __restrict__ part) {
  extern __shared__ __align__(16) unsigned char fb_sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long ta = blockIdx.y / tiles_b, tb = blockIdx.y % tiles_b;
  const long long a0 = ta * FB_TA, b0 = tb * FB_TB;
  const long long D = 1ll << n, len = D / ksplit, k0 = (long long)blockIdx.x * len;
  const int nchunk = (int)(len / FB_KC);
  // stage loader: row r of the tile (A rows first), FB_KC amplitudes = 2 KB = 128 x 16 B
  auto load_stage = [&](int c, int buf) {
    unsigned char* dst = fb_sm + (size_t)buf * FB_STAGE_BYTES;
    const long long kk = k0 + (long long)c * FB_KC;
    for (int e = tid; e < FB_ROWS * FB_KC; e += 256) {
      const int r = e / FB_KC, k = e % FB_KC;
      const c128* src;
      if (r < FB_TA) {
        if (a0 + r >= SA) continue;                // ragged tile: the row's products are discarded
        src = A + (a0 + r) * D;
      } else {
        if (b0 + (r - FB_TA) >= SB) continue;
        src = B + (b0 + (r - FB_TA)) * D;
      }
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                       (unsigned)__cvta_generic_to_shared(dst + (size_t)e * 16)), "l"(src + kk + k) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  double re[4][4], im[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) re[i][j] = im[i][j] = 0.0;
  for (int e = tid; e < 2 * FB_STAGE_BYTES / 16; e += 256)      // rows a ragged tile never loads
    reinterpret_cast<double2*>(fb_sm)[e] = make_double2(0.0, 0.0);
  __syncthreads();
  if (nchunk > 0) load_stage(0, 0);
  for (int c = 0; c < nchunk; ++c) {
    if (c + 1 < nchunk) {
      load_stage(c + 1, (c + 1) & 1);
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const c128* st = reinterpret_cast<const c128*>(fb_sm + (size_t)(c & 1) * FB_STAGE_BYTES);
#pragma unroll
    for (int q = 0; q < FB_KC / 32; ++q) {
      const int k = lane + 32 * q;
      c128 av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) av[i] = st[(warp * 4 + i) * FB_KC + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) bv[j] = st[(FB_TA + j) * FB_KC + k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {             // conj(a) * b
          re[i][j] = fma(av[i].x, bv[j].x, fma(av[i].y, bv[j].y, re[i][j]));
          im[i][j] = fma(av[i].x, bv[j].y, fma(-av[i].y, bv[j].x, im[i][j]));
        }
    }
    __syncthreads();                              // the stage may be overwritten two loads later
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      double r = re[i][j], m = im[i][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        r += __shfl_xor_sync(0xffffffffu, r, o);
        m += __shfl_xor_sync(0xffffffffu, m, o);
      }
      const long long ia = a0 + warp * 4 + i, jb = b0 + j;
      if (lane == 0 && ia < SA && jb < SB)
        part[(ia * SB + jb) * ksplit + blockIdx.x] = make_double2(r, m);
    }
}

// K slices in order, |.|^2, binning; tri: only the pairs i < j of the full SA x SA matrix, emitted
// in itertools.combinations order
__global__ void k_fid_block_fin(const double2* __restrict__ part, long long SA, long long SB, int tri,
                                int ksplit, long long bins, double step,
                                unsigned long long* __restrict__ hist, double* __restrict__ F) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= SA * SB) return;
  const long long i = e / SB, j = e % SB;
  if (tri && j <= i) return;
  double re = 0.0, im = 0.0;
  for (int k = 0; k < ksplit; ++k) {
    re += part[e * ksplit + k].x;
    im += part[e * ksplit + k].y;
  }
  const double f = re * re + im * im;
  if (F) {
    const long long idx = tri ? (i * (2 * SA - i - 1) / 2 + (j - i - 1)) : e;
    F[idx] = f;
  }
  if (hist) {
    const long long b = np_hist_bin(f, bins, step);
    if (b >= 0) atomicAdd(hist + b, 1ull);
  }
}

static int fid_block(const c128* A, long long SA, const c128* B, long long SB, int n, int tri,
                     long long bins, unsigned long long* hist, double* F, cudaStream_t st) {
  const long long tiles_a = (SA + FB_TA - 1) / FB_TA, tiles_b = (SB + FB_TB - 1) / FB_TB;
  const long long tiles = tiles_a * tiles_b;
  if (tiles > 65535) PQC_FAIL(-1, "fidelity block kernel: too many pair tiles; split the block");
  const long long D = 1ll << n;
  int ksplit = 1;                                  // >= 2 CTAs per SM in total, slices >= 4 chunks
  while (ksplit < 4096 && D / (2 * ksplit) >= 4 * FB_KC && tiles * ksplit < 148 * 2) ksplit *= 2;
  double2* part = nullptr;
  PQC_CUDA(cudaMallocAsync(&part, sizeof(double2) * SA * SB * ksplit, st));
  static PqcDeviceOnce once;
  if (once.first())
    PQC_CUDA(cudaFuncSetAttribute(k_fid_block, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  2 * FB_STAGE_BYTES));
  k_fid_block<<<dim3((unsigned)ksplit, (unsigned)tiles), 256, 2 * FB_STAGE_BYTES, st>>>(
      A, SA, B, SB, n, ksplit, tiles_b, part);
  k_fid_block_fin<<<(unsigned)((SA * SB + 127) / 128), 128, 0, st>>>(
      part, SA, SB, tri, ksplit, bins, bins > 0 ? 1.0 / (double)bins : 0.0, hist, F);
  g_pqc_launches += 2;
  const cudaError_t e = cudaGetLastError();
  cudaFreeAsync(part, st);
  if (e != cudaSuccess) PQC_FAIL(-2, std::string("fidelity block kernel launch: ") + cudaGetErrorString(e));
  return 0;
}

extern "C" int pqc_fidelity_hist(const pqc_c128* d_A, int64_t n_a, const pqc_c128* d_B,
                                 int64_t n_b, int n, int triangular, int64_t bins,
                                 long long* d_hist, double* d_F, void* stream) {
  if (n_a <= 0 || n_b <= 0) return 0;
  if (d_hist && bins <= 0) PQC_FAIL(-1, "`bins` must be positive, when an integer");
  if (triangular && (d_A != d_B || n_a != n_b))
    PQC_FAIL(-1, "triangular mode needs the same block on both sides");
  const long long nbi = (n_a + FT - 1) / FT, nbj = (n_b + FT - 1) / FT;
  const long long grid = triangular ? nbi * (nbi + 1) / 2 : nbi * nbj;
  if (grid > 0x7fffffffLL) PQC_FAIL(-1, "fidelity grid too large; split the block");
  const char* force = getenv("PQC_FIDELITY");
  {
    // a handful of huge states (config 5: 4 GiB each): the pair-tile kernels would run on a few
    // CTAs, each walking all 2^n amplitudes -- split the inner products over K instead
    const long long npairs = triangular ? n_a * (n_a - 1) / 2 : n_a * n_b;
    // PQC_FIDELITY=splitk keeps the one-pair-per-CTA form (tests compare the two)
    const bool blocked = (n >= 20 && grid < 64 && !force) ||
                         (force && strcmp(force, "block") == 0 && n >= 12);
    if (blocked && npairs > 0)
      return fid_block((const c128*)d_A, n_a, (const c128*)d_B, n_b, n, triangular, bins,
                       (unsigned long long*)d_hist, d_F, (cudaStream_t)stream);
    const bool splitk = force && strcmp(force, "splitk") == 0 && n >= 12 && npairs <= 65535;
    if (splitk && npairs > 0) {
      cudaStream_t st = (cudaStream_t)stream;
      int ksplit = 1;
      while (ksplit < 1024 && (1ll << n) / (2 * ksplit) >= 4096 && npairs * ksplit < 148 * 16)
        ksplit *= 2;
      double2* part = nullptr;
      PQC_CUDA(cudaMallocAsync(&part, sizeof(double2) * npairs * ksplit, st));
      k_fid_splitk<<<dim3((unsigned)ksplit, (unsigned)npairs), 256, 0, st>>>(
          (const c128*)d_A, n_a, (const c128*)d_B, n_b, n, triangular, ksplit, part);
      k_fid_splitk_fin<<<(unsigned)((npairs + 127) / 128), 128, 0, st>>>(
          part, npairs, ksplit, bins, bins > 0 ? 1.0 / (double)bins : 0.0,
          (unsigned long long*)d_hist, d_F);
      g_pqc_launches += 2;
      const cudaError_t e = cudaGetLastError();
      cudaFreeAsync(part, st);
      if (e != cudaSuccess) PQC_FAIL(-2, std::string("fidelity split-K launch: ") + cudaGetErrorString(e));
      return 0;
    }
  }
  const bool use_dmma = n >= 4 && !(force && strcmp(force, "fma") == 0);
  if (use_dmma) {
    static PqcDeviceOnce attr_once;
    const size_t smem = 4 * DT * DROW * sizeof(c128);
    if (attr_once.first()) {
      PQC_CUDA(cudaFuncSetAttribute(k_fidelity_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)smem));
    }
    k_fidelity_dmma<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(
        (const c128*)d_A, n_a, (const c128*)d_B, n_b, n, triangular, bins,
        bins > 0 ? 1.0 / (double)bins : 0.0, (unsigned long long*)d_hist, d_F);
    PQC_LAUNCH_CHECK();
    return 0;
  }
  k_fidelity<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>(
      (const c128*)d_A, n_a, (const c128*)d_B, n_b, n, triangular, bins,
      bins > 0 ? 1.0 / (double)bins : 0.0, (unsigned long long*)d_hist, d_F);
  PQC_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------
// Measurements.expr on histogram counts (measure.py:161-180):
//   P_pqc = counts / sum(counts); F_mid = bin midpoints; haar = (N-1)(1-F)^(N-2);
//   P_haar = haar / sum(haar); KL = sum kl_div(P_pqc, P_haar)   (scipy.special.kl_div)
// scratch[0] = sum counts, scratch[1] = sum haar, scratch[2] = KL
// ---------------------------------------------------------------------------------
__device__ __forceinline__ double bin_mid(long long i, long long B, double step) {
  const double lo = __dmul_rn((double)i, step);
  const double hi = (i + 1 == B) ? 1.0 : __dmul_rn((double)(i + 1), step);
  return (lo + hi) / 2.0;
}

__global__ void __launch_bounds__(256) k_kl_sums(const long long* __restrict__ hist, long long B,
                                                 double step, double N, double* __restrict__ part) {
  __shared__ double red[32];
  double sc = 0.0, sh = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < B; i += (long long)gridDim.x * 256) {
    sc += (double)hist[i];
    sh += (N - 1.0) * pow(1.0 - bin_mid(i, B, step), N - 2.0);
  }
  sc = block_sum<256>(sc, red);
  sh = block_sum<256>(sh, red);
  if (threadIdx.x == 0) {                     // per-CTA partials, added in CTA order by k_kl_fold
    part[2 * blockIdx.x + 0] = sc;
    part[2 * blockIdx.x + 1] = sh;
  }
}

// out[c] = sum over CTAs of part[cta * ncomp + c], in a fixed order (one CTA)
__global__ void __launch_bounds__(256) k_kl_fold(const double* __restrict__ part, int nparts,
                                                 int ncomp, double* __restrict__ out) {
  __shared__ double red[32];
  for (int c = 0; c < ncomp; ++c) {
    double t = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) t += part[i * ncomp + c];
    t = block_sum<256>(t, red);
    if (threadIdx.x == 0) out[c] = t;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_kl_terms(const long long* __restrict__ hist, long long B,
                                                  double step, double N,
                                                  const double* __restrict__ scratch,
                                                  double* __restrict__ part) {
  __shared__ double red[32];
  const double tc = scratch[0], th = scratch[1];
  double kl = 0.0;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < B; i += (long long)gridDim.x * 256) {
    const double x = (double)hist[i] / tc;
    const double y = (N - 1.0) * pow(1.0 - bin_mid(i, B, step), N - 2.0) / th;
    double t;                                  // scipy.special.kl_div, NaN first
    if (isnan(x) || isnan(y)) t = NAN;         // e.g. every Haar weight underflowed: 0 / 0
    else if (x > 0.0 && y > 0.0) t = x * log(x / y) - x + y;
    else if (x == 0.0 && y >= 0.0) t = y;
    else t = INFINITY;
    kl += t;
  }
  kl = block_sum<256>(kl, red);
  if (threadIdx.x == 0) part[blockIdx.x] = kl;
}

__global__ void k_copy1(const double* src, double* dst) { *dst = *src; }

extern "C" int pqc_kl_haar(const long long* d_hist, int64_t bins, double N, double* d_out,
                           double* d_scratch, void* stream) {
  if (bins <= 0) PQC_FAIL(-1, "`bins` must be positive, when an integer");
  cudaStream_t st = (cudaStream_t)stream;
  PQC_CUDA(cudaMemsetAsync(d_scratch, 0, 4 * sizeof(double), st));
  const long long grid = std::min<long long>((bins + 255) / 256, 148 * 8);
  const double step = 1.0 / (double)bins;
  // fixed-order reductions (no floating-point atomics): the result is bitwise reproducible
  double* part = nullptr;
  PQC_CUDA(cudaMallocAsync(&part, sizeof(double) * 2 * (size_t)grid, st));
  k_kl_sums<<<(unsigned)grid, 256, 0, st>>>(d_hist, bins, step, N, part);
  k_kl_fold<<<1, 256, 0, st>>>(part, (int)grid, 2, d_scratch);
  k_kl_terms<<<(unsigned)grid, 256, 0, st>>>(d_hist, bins, step, N, d_scratch, part);
  k_kl_fold<<<1, 256, 0, st>>>(part, (int)grid, 1, d_scratch + 2);
  g_pqc_launches += 4;
  k_copy1<<<1, 1, 0, st>>>(d_scratch + 2, d_out);
  const cudaError_t le = cudaGetLastError();
  cudaFreeAsync(part, st);
  if (le != cudaSuccess) PQC_FAIL(-2, cudaGetErrorString(le));
  ++g_pqc_launches;
  return 0;
}

// ---------------------------------------------------------------------------------
// Renyi / GKP magic (measure.py:318-368).  For X-mask k the reference's column M[:,k] is
// the Walsh-Hadamard transform over j of v_j = conj(c_j) c_{j^k}.  Because
// v_{j^k} = conj(v_j), WHT(Re v)_i vanishes when i.k is odd and WHT(Im v)_i when it is
// even, so |M[i,k]| = |WHT(Re v + Im v)_i|: one REAL in-shared-memory FWHT per mask
// instead of the reference's 2^n x 2^n x 2^n complex GEMM.
// One CTA per (sample, slice of masks); state cached in shared memory.
// out[a][s] accumulates sum_{i,k} |W|^{2 alpha_a}; finalised by k_magic_finalize.
// ---------------------------------------------------------------------------------
#define MAGIC_MAX_ALPHA 4
struct MagicArgs {
  const c128* states;
  int n;
  int masks_per_cta;
  int n_alpha;
  double alpha[MAGIC_MAX_ALPHA];
  double* sums;      // [n_alpha][S][slices] per-CTA partial sums (folded by k_magic_finalize)
  long long S;
  int slices;
};

__global__ void __launch_bounds__(256) k_magic(const MagicArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int n = a.n;
  const uint32_t D = 1u << n;
  double* cr = reinterpret_cast<double*>(smraw);   // [D] real parts
  double* ci = cr + D;                             // [D] imaginary parts
  double* u = ci + D;                              // [D] FWHT buffer
  __shared__ double red[32];
  const int slices = (int)((D + a.masks_per_cta - 1) / a.masks_per_cta);
  const long long s = blockIdx.x / slices;
  const int slice = blockIdx.x % slices;
  const c128* psi = a.states + s * (long long)D;
  for (uint32_t i = threadIdx.x; i < D; i += 256) {
    const c128 v = psi[i];
    cr[i] = v.x;
    ci[i] = v.y;
  }
  __syncthreads();
  double acc[MAGIC_MAX_ALPHA];
#pragma unroll
  for (int q = 0; q < MAGIC_MAX_ALPHA; ++q) acc[q] = 0.0;
  const uint32_t k_begin = (uint32_t)slice * a.masks_per_cta;
  const uint32_t k_end = min(D, k_begin + (uint32_t)a.masks_per_cta);
  for (uint32_t k = k_begin; k < k_end; ++k) {
    for (uint32_t j = threadIdx.x; j < D; j += 256) {
      const double xr = cr[j], xi = ci[j], yr = cr[j ^ k], yi = ci[j ^ k];
      // v = conj(c_j) c_{j^k};  u = Re v + Im v
      u[j] = (xr * yr + xi * yi) + (xr * yi - xi * yr);
    }
    __syncthreads();
    for (uint32_t h = 1; h < D; h <<= 1) {
      for (uint32_t p = threadIdx.x; p < (D >> 1); p += 256) {
        const uint32_t i0 = ((p & ~(h - 1)) << 1) | (p & (h - 1));
        const double x = u[i0], y = u[i0 + h];
        u[i0] = x + y;
        u[i0 + h] = x - y;
      }
      __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < D; i += 256) {
      const double w = fabs(u[i]);
#pragma unroll
      for (int q = 0; q < MAGIC_MAX_ALPHA; ++q) {
        if (q < a.n_alpha) {
          const double al = a.alpha[q];
          double t;
          if (al == 2.0) { const double w2 = w * w; t = w2 * w2; }
          else if (al == 0.5) t = w;
          else if (al == 1.0) t = w * w;
          else t = (w == 0.0) ? 0.0 : pow(w, 2.0 * al);
          acc[q] += t;
        }
      }
    }
    __syncthreads();
  }
  for (int q = 0; q < a.n_alpha; ++q) {
    const double t = block_sum<256>(acc[q], red);
    if (threadIdx.x == 0) a.sums[((long long)q * a.S + s) * a.slices + slice] = t;
  }
}

// n = 12 (BASELINE config 4): register-blocked FWHT.  256 threads x 16 points; the three
// sweeps hold index bits 8-11, 0-3 and 4-7 in registers (four butterfly stages each, no memory
// traffic) with two shared-memory transposes in between instead of twelve in-memory stages.
// The thread's own amplitudes c_j stay in registers for all masks; only the partners c_{j^k}
// are read from the shared-memory copy of the state.  The transpose buffer is XOR-swizzled,
// sigma(i) = i ^ ((i >> 4) & 15), which makes all three access patterns conflict free for
// 8-byte elements.
template <int K>
__device__ __forceinline__ void mg_bfly(double (&v)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    if (j & (1 << K)) continue;
    const double x = v[j], y = v[j | (1 << K)];
    v[j] = x + y;
    v[j | (1 << K)] = x - y;
  }
}

__global__ void __launch_bounds__(256, 2) k_magic12(const MagicArgs a) {
  extern __shared__ __align__(16) unsigned char smraw[];
  c128* sc = reinterpret_cast<c128*>(smraw);            // [4096] the state
  double* u = reinterpret_cast<double*>(sc + 4096);     // [4096] transpose buffer (swizzled)
  __shared__ double red[32];
  const uint32_t D = 4096u;
  const uint32_t tid = threadIdx.x, lo = tid & 15u, hi = tid >> 4;
  const int slices = (int)((D + a.masks_per_cta - 1) / a.masks_per_cta);
  const long long s = blockIdx.x / slices;
  const int slice = blockIdx.x % slices;
  const c128* psi = a.states + s * (long long)D;
  c128 own[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) own[j] = psi[tid | ((uint32_t)j << 8)];
#pragma unroll
  for (int j = 0; j < 16; ++j) sc[tid | ((uint32_t)j << 8)] = own[j];
  __syncthreads();
  double acc[MAGIC_MAX_ALPHA];
#pragma unroll
  for (int q = 0; q < MAGIC_MAX_ALPHA; ++q) acc[q] = 0.0;
  // swizzled transpose-buffer index of register j in the three sweeps
  const uint32_t ia = tid ^ hi;                 // | j << 8          (index bits 8-11 in registers)
  const uint32_t ib = tid << 4;                 // | (j ^ lo)        (bits 0-3)
  const uint32_t ic = hi << 8;                  // | j << 4 | lo ^ j (bits 4-7)
  const uint32_t k_begin = (uint32_t)slice * a.masks_per_cta;
  const uint32_t k_end = min(D, k_begin + (uint32_t)a.masks_per_cta);
  for (uint32_t k = k_begin; k < k_end; ++k) {
    const uint32_t pl = tid ^ (k & 255u), kh = k >> 8;
    double v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const c128 y = sc[pl | (((uint32_t)j ^ kh) << 8)];
      const double xr = own[j].x, xi = own[j].y;
      // v = conj(c_j) c_{j^k};  u = Re v + Im v
      v[j] = (xr * y.x + xi * y.y) + (xr * y.y - xi * y.x);
    }
    mg_bfly<0>(v); mg_bfly<1>(v); mg_bfly<2>(v); mg_bfly<3>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) u[ia | ((uint32_t)j << 8)] = v[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = u[ib | ((uint32_t)j ^ lo)];
    mg_bfly<0>(v); mg_bfly<1>(v); mg_bfly<2>(v); mg_bfly<3>(v);
#pragma unroll
    for (int j = 0; j < 16; ++j) u[ib | ((uint32_t)j ^ lo)] = v[j];
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = u[ic | ((uint32_t)j << 4) | (lo ^ (uint32_t)j)];
    __syncthreads();                 // the buffer is free for the next mask
    mg_bfly<0>(v); mg_bfly<1>(v); mg_bfly<2>(v); mg_bfly<3>(v);
#pragma unroll
    for (int q = 0; q < MAGIC_MAX_ALPHA; ++q) {
      if (q < a.n_alpha) {
        const double al = a.alpha[q];
        double t = 0.0;
        if (al == 2.0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) { const double w2 = v[j] * v[j]; t = fma(w2, w2, t); }
        } else if (al == 0.5) {
#pragma unroll
          for (int j = 0; j < 16; ++j) t += fabs(v[j]);
        } else if (al == 1.0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) t = fma(v[j], v[j], t);
        } else {
          for (int j = 0; j < 16; ++j) {
            const double w = fabs(v[j]);
            t += (w == 0.0) ? 0.0 : pow(w, 2.0 * al);
          }
        }
        acc[q] += t;
      }
    }
  }
  for (int q = 0; q < a.n_alpha; ++q) {
    const double t = block_sum<256>(acc[q], red);
    if (threadIdx.x == 0) a.sums[((long long)q * a.S + s) * a.slices + slice] = t;
  }
}

__global__ void k_magic_finalize(double* __restrict__ out, long long S, int n, int n_alpha,
                                 MagicArgs a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S * n_alpha) return;
  const int q = (int)(e / S);
  const double al = a.alpha[q];
  // slices in a fixed order; sum |2^{-n/2} W|^{2 alpha} = 2^{-n alpha} sum |W|^{2 alpha}
  double acc = 0.0;
  for (int c = 0; c < a.slices; ++c) acc += a.sums[e * a.slices + c];
  const double tot = acc * exp2(-(double)n * al);
  out[e] = 1.0 / (1.0 - al) * log(tot) - (double)n * 0.69314718055994530942;
}

extern "C" int pqc_magic_batch(const pqc_c128* d_states, int64_t S, int n, int n_alpha,
                               const double* h_alphas, double* d_out, void* stream) {
  if (S <= 0 || n_alpha <= 0) return 0;
  if (n_alpha > MAGIC_MAX_ALPHA) PQC_FAIL(-1, "at most 4 alphas per call");
  if (n < 1 || n > 13) PQC_FAIL(-1, "magic kernel keeps the state in shared memory: n <= 13");
  cudaStream_t st = (cudaStream_t)stream;
  MagicArgs a;
  a.states = (const c128*)d_states;
  a.n = n;
  a.n_alpha = n_alpha;
  for (int q = 0; q < n_alpha; ++q) {
    if (h_alphas[q] == 1.0) PQC_FAIL(-1, "alpha = 1 is singular (1/(1-alpha))");
    a.alpha[q] = h_alphas[q];
  }
  a.sums = nullptr;
  a.S = S;
  a.slices = 1;
  const long long D = 1ll << n;
  // enough CTAs to fill the machine: >= 4 * 148 CTAs when the batch is small
  long long slices = std::max<long long>(1, std::min<long long>(D, (148 * 4 + S - 1) / S));
  a.masks_per_cta = (int)((D + slices - 1) / slices);
  slices = (D + a.masks_per_cta - 1) / a.masks_per_cta;
  static PqcDeviceOnce attr_once;
  if (attr_once.first()) {
    PQC_CUDA(cudaFuncSetAttribute(k_magic, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PQC_CUDA(cudaFuncSetAttribute(k_magic12, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
  }
  const char* mg = getenv("PQC_MAGIC");       // PQC_MAGIC=generic: in-memory FWHT for every n
  const bool generic_only = mg && strcmp(mg, "generic") == 0;   // (read per call: tests compare)
  if (n == 12 && !generic_only) {
    // register-blocked kernel, 2 CTAs per SM: cut every sample's masks into enough slices for
    // >= 8 waves so the tail of the last wave stays small
    slices = std::max<long long>(1, std::min<long long>(64, (148 * 2 * 8 + S - 1) / S));
    a.masks_per_cta = (int)((D + slices - 1) / slices);
    slices = (D + a.masks_per_cta - 1) / a.masks_per_cta;
    if (S * slices > 0x7fffffffLL) PQC_FAIL(-1, "magic grid too large; split the batch");
    a.slices = (int)slices;
    PQC_CUDA(cudaMallocAsync(&a.sums, sizeof(double) * S * n_alpha * slices, st));
    k_magic12<<<(unsigned)(S * slices), 256, D * (sizeof(c128) + sizeof(double)), st>>>(a);
    k_magic_finalize<<<(unsigned)((S * n_alpha + 127) / 128), 128, 0, st>>>(d_out, S, n, n_alpha, a);
    g_pqc_launches += 2;
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(a.sums, st);
    if (le != cudaSuccess) PQC_FAIL(-2, cudaGetErrorString(le));
    return 0;
  }
  const size_t smem = 3 * D * sizeof(double);
  if (S * slices > 0x7fffffffLL) PQC_FAIL(-1, "magic grid too large; split the batch");
  a.slices = (int)slices;
  PQC_CUDA(cudaMallocAsync(&a.sums, sizeof(double) * S * n_alpha * slices, st));
  k_magic<<<(unsigned)(S * slices), 256, smem, st>>>(a);
  k_magic_finalize<<<(unsigned)((S * n_alpha + 127) / 128), 128, 0, st>>>(d_out, S, n, n_alpha, a);
  g_pqc_launches += 2;
  const cudaError_t le = cudaGetLastError();
  cudaFreeAsync(a.sums, st);
  if (le != cudaSuccess) PQC_FAIL(-2, cudaGetErrorString(le));
  return 0;
}

// ---------------------------------------------------------------------------------
// get_QFI (measure.py:33-71) from explicit states + derivative states.
// One CTA per (sample, p, q>=p) would be wasteful; use one CTA per (sample, p) row and
// loop q, after first computing s_p = <psi|d_p>.  Deterministic summation order.
// G scratch: [S][P+1][P] complex: row 0 = s_q, row 1+p = <d_p|d_q>.
// ---------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gram_rows(const c128* __restrict__ states,
                                                   const c128* __restrict__ grads, int n, int P,
                                                   c128* __restrict__ G) {
  __shared__ double red[32];
  const long long D = 1ll << n;
  const long long s = blockIdx.x / (P + 1);
  const int row = blockIdx.x % (P + 1);
  const c128* a = row == 0 ? states + s * D : grads + (s * P + (row - 1)) * D;
  const int q0 = row == 0 ? 0 : row - 1;
  for (int q = q0; q < P; ++q) {
    const c128* b = grads + (s * P + q) * D;
    double re = 0.0, im = 0.0;
    for (long long i = threadIdx.x; i < D; i += 256) {
      const c128 x = a[i], y = b[i];
      re += x.x * y.x + x.y * y.y;
      im += x.x * y.y - x.y * y.x;
    }
    re = block_sum<256>(re, red);
    im = block_sum<256>(im, red);
    if (threadIdx.x == 0) G[(s * (P + 1) + row) * P + q] = make_double2(re, im);
  }
}

// F_pq = 4 Re(G_pq - conj(s_p) s_q), p <= q, mirrored (measure.py:55-70)
__global__ void k_qfim_finalize(const c128* __restrict__ G, long long S, int P,
                                double* __restrict__ F) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= S * P * P) return;
  const long long s = e / ((long long)P * P);
  const int r = (int)((e / P) % P), c = (int)(e % P);
  const int p = r < c ? r : c, q = r < c ? c : r;
  const c128* g = G + s * (long long)(P + 1) * P;
  const c128 sp = g[p], sq = g[q], d = g[(long long)(1 + p) * P + q];
  // np.conjugate(s_p) * s_q, real part
  const double rhs = sp.x * sq.x + sp.y * sq.y;
  F[e] = 4.0 * (d.x - rhs);
}

int pqc_qfim_finalize(const c128* d_G, long long S, int P, double* d_F, cudaStream_t st) {
  const long long tot = S * P * P;
  if (tot <= 0) return 0;
  k_qfim_finalize<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(d_G, S, P, d_F);
  PQC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pqc_qfim_from_grads(const pqc_c128* d_states, const pqc_c128* d_grads, int n,
                                   int P, int64_t S, double* d_qfim, void* stream) {
  if (S <= 0 || P <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  c128* G = nullptr;
  PQC_CUDA(cudaMallocAsync(&G, sizeof(c128) * S * (P + 1) * P, st));
  int rc = 0;
  if (S * (P + 1) > 0x7fffffffLL) {
    pqc_set_error("qfim grid too large");
    rc = -1;
  } else {
    k_gram_rows<<<(unsigned)(S * (P + 1)), 256, 0, st>>>((const c128*)d_states,
                                                        (const c128*)d_grads, n, P, G);
    ++g_pqc_launches;
    if (cudaGetLastError() != cudaSuccess) rc = -2;
    if (rc == 0) rc = pqc_qfim_finalize(G, S, P, d_qfim, st);
  }
  cudaFreeAsync(G, st);
  return rc;
}

// ---------------------------------------------------------------------------------
// eigenvalues of S symmetric PxP matrices (scipy.linalg.eigh, measure.py:73-75,84):
// parallel cyclic Jacobi in shared memory, one CTA per matrix, round-robin pairing so
// P/2 disjoint rotations run concurrently; ascending output.
// ---------------------------------------------------------------------------------
template <bool WITH_V>
__global__ void __launch_bounds__(128) k_jacobi_eigvals(const double* __restrict__ mats, int P,
                                                        int Pp, double* __restrict__ eig,
                                                        double* __restrict__ vecs) {
  extern __shared__ double smd[];
  double* A = smd;                        // [Pp][Pp+1]
  const int ld = Pp + 1;
  double* cs = A + (size_t)Pp * ld;       // [Pp/2][2]
  int* pr = reinterpret_cast<int*>(cs + Pp);   // [Pp] current pairing
  double* V = reinterpret_cast<double*>(pr + Pp + (Pp & 1));   // [Pp][Pp+1] accumulated rotations
  __shared__ double red[32];
  __shared__ double s_off, s_norm;
  double prev_off = 1e300;
  const double* M = mats + (long long)blockIdx.x * P * P;
  const int tid = threadIdx.x;
  for (int e = tid; e < Pp * Pp; e += 128) {
    const int r = e / Pp, c = e % Pp;
    double v = 0.0;
    if (r < P && c < P) v = 0.5 * (M[(long long)r * P + c] + M[(long long)c * P + r]);
    else if (r == c) v = 1e300;            // padding: decoupled sentinel, dropped at the end
    A[r * ld + c] = v;
  }
  for (int i = tid; i < Pp; i += 128) pr[i] = i;
  if (WITH_V)
    for (int e = tid; e < Pp * Pp; e += 128) V[(e / Pp) * ld + (e % Pp)] = (e / Pp == e % Pp) ? 1.0 : 0.0;
  __syncthreads();
  const int half = Pp / 2;
  for (int sweep = 0; sweep < 40; ++sweep) {
    // convergence test: off-diagonal Frobenius norm vs diagonal
    double off = 0.0, nrm = 0.0;
    for (int e = tid; e < P * P; e += 128) {
      const int r = e / P, c = e % P;
      const double v = A[r * ld + c];
      if (r == c) nrm += v * v; else off += v * v;
    }
    off = block_sum<128>(off, red);
    nrm = block_sum<128>(nrm, red);
    if (tid == 0) { s_off = off; s_norm = nrm; }
    __syncthreads();
    // stop at the rounding floor: either negligible, or small and no longer shrinking
    const double cur_off = s_off;
    if (cur_off == 0.0 || cur_off <= 1e-33 * s_norm ||
        (cur_off <= 1e-26 * s_norm && cur_off >= 0.25 * prev_off))
      break;
    prev_off = cur_off;
    for (int round = 0; round < Pp - 1; ++round) {
      // rotation angles for the `half` disjoint pairs (p = pr[i], q = pr[Pp-1-i])
      for (int i = tid; i < half; i += 128) {
        int p = pr[i], q = pr[Pp - 1 - i];
        if (p > q) { const int t = p; p = q; q = t; }
        const double apq = A[p * ld + q];
        double c = 1.0, s = 0.0;
        if (apq != 0.0 && p < P && q < P) {
          const double tau = (A[q * ld + q] - A[p * ld + p]) / (2.0 * apq);
          const double t = (tau >= 0.0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
          c = 1.0 / sqrt(1.0 + t * t);
          s = t * c;
        }
        cs[2 * i] = c;
        cs[2 * i + 1] = s;
      }
      __syncthreads();
      // rows: A <- J^T A
      for (int e = tid; e < half * Pp; e += 128) {
        const int i = e / Pp, col = e % Pp;
        int p = pr[i], q = pr[Pp - 1 - i];
        if (p > q) { const int t = p; p = q; q = t; }
        const double c = cs[2 * i], s = cs[2 * i + 1];
        const double x = A[p * ld + col], y = A[q * ld + col];
        A[p * ld + col] = c * x - s * y;
        A[q * ld + col] = s * x + c * y;
      }
      __syncthreads();
      // columns: A <- A J
      for (int e = tid; e < half * Pp; e += 128) {
        const int i = e / Pp, row = e % Pp;
        int p = pr[i], q = pr[Pp - 1 - i];
        if (p > q) { const int t = p; p = q; q = t; }
        const double c = cs[2 * i], s = cs[2 * i + 1];
        const double x = A[row * ld + p], y = A[row * ld + q];
        A[row * ld + p] = c * x - s * y;
        A[row * ld + q] = s * x + c * y;
        if (WITH_V) {
          const double vx = V[row * ld + p], vy = V[row * ld + q];
          V[row * ld + p] = c * vx - s * vy;
          V[row * ld + q] = s * vx + c * vy;
        }
      }
      __syncthreads();
      // round-robin tournament: keep pr[0], rotate the rest
      if (tid == 0) {
        const int last = pr[Pp - 1];
        for (int i = Pp - 1; i > 1; --i) pr[i] = pr[i - 1];
        pr[1] = last;
      }
      __syncthreads();
    }
  }
  // rank sort of the P real diagonal entries (ties broken by index)
  for (int i = tid; i < P; i += 128) {
    const double v = A[i * ld + i];
    int rank = 0;
    for (int j = 0; j < P; ++j) {
      const double w = A[j * ld + j];
      rank += (w < v) || (w == v && j < i);
    }
    eig[(long long)blockIdx.x * P + rank] = v;
    if (WITH_V)     // column `rank` of the output = eigenvector of the rank-th eigenvalue
      for (int r = 0; r < P; ++r)
        vecs[((long long)blockIdx.x * P + r) * P + rank] = V[r * ld + i];
  }
}

static int eigh_launch(const double* d_mats, int64_t n_mats, int dim, double* d_eig,
                       double* d_vecs, cudaStream_t st) {
  if (n_mats <= 0 || dim <= 0) return 0;
  const int Pp = (dim + 1) & ~1;
  size_t smem = ((size_t)Pp * (Pp + 1) + Pp) * sizeof(double) + (Pp + (Pp & 1)) * sizeof(int);
  if (d_vecs) smem += (size_t)Pp * (Pp + 1) * sizeof(double);
  if (smem > 200 * 1024)
    PQC_FAIL(-1, "eigh: matrix too large for the shared-memory Jacobi (dim <= 158, or 110 with vectors)");
  if (n_mats > 0x7fffffffLL) PQC_FAIL(-1, "too many matrices");
  static PqcDeviceOnce attr_once;
  if (attr_once.first()) {
    PQC_CUDA(cudaFuncSetAttribute(k_jacobi_eigvals<false>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    PQC_CUDA(cudaFuncSetAttribute(k_jacobi_eigvals<true>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  if (d_vecs)
    k_jacobi_eigvals<true><<<(unsigned)n_mats, 128, smem, st>>>(d_mats, dim, Pp, d_eig, d_vecs);
  else
    k_jacobi_eigvals<false><<<(unsigned)n_mats, 128, smem, st>>>(d_mats, dim, Pp, d_eig, nullptr);
  PQC_LAUNCH_CHECK();
  return 0;
}

extern "C" int pqc_eigvalsh_batch(const double* d_mats, int64_t n_mats, int dim, double* d_eig,
                                  void* stream) {
  return eigh_launch(d_mats, n_mats, dim, d_eig, nullptr, (cudaStream_t)stream);
}

extern "C" int pqc_eigh_batch(const double* d_mats, int64_t n_mats, int dim, double* d_eig,
                              double* d_vecs, void* stream) {
  return eigh_launch(d_mats, n_mats, dim, d_eig, d_vecs, (cudaStream_t)stream);
}

__global__ void k_count_greater(const double* __restrict__ v, long long rows, int cols,
                                double cutoff, int* __restrict__ out) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  int c = 0;
  for (int j = 0; j < cols; ++j) c += v[r * cols + j] > cutoff;
  out[r] = c;
}

extern "C" int pqc_count_greater(const double* d_vals, int64_t rows, int cols, double cutoff,
                                 int32_t* d_counts, void* stream) {
  if (rows <= 0) return 0;
  k_count_greater<<<(unsigned)((rows + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      d_vals, rows, cols, cutoff, d_counts);
  PQC_LAUNCH_CHECK();
  return 0;
}
