// membench.cu -- what does HBM give for the pass kernel's access pattern?  (developer tool)
// Each CTA of 256 threads moves one 4096-amplitude tile (64 KB) of a 2^n-amplitude vector:
// the tile's 12 index bits are the `lr` lowest bits plus the (12 - lr) highest bits of the
// vector index; the remaining bits select the tile.  Every thread issues 16 x 16-byte loads,
// then 16 x 16-byte stores (in place or to a second buffer), like k_sweep_pass.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/membench.cu -o gpurun_out/membench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

template <int OCC>
__global__ void __launch_bounds__(256, OCC) k_move(const double2* __restrict__ src,
                                                   double2* __restrict__ dst, int n, int lr,
                                                   int reg_hi) {
  extern __shared__ double2 sm[];
  const int tid = threadIdx.x;
  const int tiles_log2 = n - 12;
  const long long vec = blockIdx.x >> tiles_log2;
  const unsigned tile = blockIdx.x & ((1u << tiles_log2) - 1u);
  // local 12-bit index l -> amplitude: low lr bits stay, upper (12 - lr) bits go to the top
  // thread holds 16 amplitudes: register bits are the 4 highest local bits (reg_hi) or
  // local bits 4..7 (else); thread bits fill the rest, lowest first
  unsigned amp[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    unsigned l;
    if (reg_hi) l = (unsigned)tid | ((unsigned)j << 8);
    else l = ((unsigned)tid & 15u) | ((unsigned)j << 4) | (((unsigned)tid >> 4) << 8);
    const unsigned lo = l & ((1u << lr) - 1u), hi = l >> lr;
    amp[j] = lo | (tile << lr) | (hi << (lr + tiles_log2));
  }
  const double2* s = src + (vec << n);
  double2* d = dst + (vec << n);
  double2 a[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) a[j] = s[amp[j]];
#pragma unroll
  for (int j = 0; j < 16; ++j) { a[j].x += 1.0; }
  if (tid == 9999) sm[0] = a[0];
#pragma unroll
  for (int j = 0; j < 16; ++j) d[amp[j]] = a[j];
}

int main(int argc, char** argv) {
  const int n = 16;
  const long long nvec = 6656;                   // 6.5 GiB per buffer, like a QFIM chunk
  const size_t bytes = (size_t)nvec << (n + 4);
  double2 *a, *b;
  cudaMalloc(&a, bytes);
  cudaMalloc(&b, bytes);
  cudaMemset(a, 0, bytes);
  cudaMemset(b, 0, bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const unsigned grid = (unsigned)(nvec << (n - 12));
  cudaFuncSetAttribute(k_move<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  cudaFuncSetAttribute(k_move<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  for (int occ = 2; occ <= 3; ++occ)
    for (int inplace = 0; inplace <= 1; ++inplace)
      for (int reg_hi = 0; reg_hi <= 1; ++reg_hi)
        for (int lr : {4, 5, 6, 8, 12}) {
          const size_t smem = occ == 2 ? 83 * 1024 : 64 * 1024;
          float best = 1e30f;
          for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            if (occ == 2) k_move<2><<<grid, 256, smem>>>(a, inplace ? a : b, n, lr, reg_hi);
            else k_move<3><<<grid, 256, smem>>>(a, inplace ? a : b, n, lr, reg_hi);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep && ms < best) best = ms;
          }
          printf("{\"occ\": %d, \"inplace\": %d, \"reg_hi\": %d, \"run_bytes\": %d, \"ms\": %.3f, \"GBps\": %.1f}\n",
                 occ, inplace, reg_hi, 16 << lr, best, 2.0 * bytes / best / 1e6);
        }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
  return 0;
}
