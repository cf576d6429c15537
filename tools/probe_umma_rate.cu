// Hardware probe (not part of libtod.so): tcgen05.mma issue/execute rate for M=128, K=16 bf16 as a function of N and of
// the A-operand descriptor geometry (aligned vs row-shifted start, SBO = 8 rows vs 10 rows), SWIZZLE_128B, K-major.
// One CTA per SM on every SM (so shared-memory/tensor contention is per SM as in the conv kernel); smem content is
// whatever is there (timing only).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_rate probe_umma_rate.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../transparent_object_detection_b200/csrc/tod_common.cuh"

using namespace tod;

struct Params {
  int row_bytes;   // 128: SWIZZLE_128B (4 K-steps per row), 64: SWIZZLE_64B (2 K-steps per row)
  int n, shift_rows, sbo_rows, iters, distinct_a;  // distinct_a: number of different A start rows cycled through
  int b_shift_rows;
  long long* cycles;  // per CTA
};

__global__ void __launch_bounds__(128, 1) rate(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  // zero the operand area so the tensor core sees finite numbers
  for (uint32_t i = threadIdx.x; i < (160u * 1024u) / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16), "r"(0) : "memory");
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (warp == 1) {
    const uint32_t sa = base, sb = base + 96 * 1024;
    const uint32_t rb = p.row_bytes, lt = rb == 128 ? 2u : 4u;
    const uint32_t hi_a = ((p.sbo_rows * rb) >> 4) | (1u << 14) | (lt << 29);
    const uint32_t hi_b = ((8u * rb) >> 4) | (1u << 14) | (lt << 29);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(p.n) >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_lo = umma_desc_lo(sa + p.shift_rows * rb), b_lo = umma_desc_lo(sb + p.b_shift_rows * rb);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      // distinct_a views: A start moves by 3 rows per view (like conv taps); every k4 call = 4 MMAs
      if (rb == 128) {
        for (int it = 0; it < p.iters / 4; it += 2) {
          umma_bf16_k4(tmem, a_lo, hi_a, b_lo, hi_b, idesc, 1);
          umma_bf16_k4(tmem + p.n, a_lo + (p.distinct_a > 1 ? 24u : 0u), hi_a, b_lo, hi_b, idesc, 1);
        }
      } else {   // 64-byte rows: 2 K-steps per row, four accumulators in turn like the conv's m = 4 sub-tiles
        for (int it = 0; it < p.iters / 2; it += 4) {
          umma_bf16_k2(tmem, a_lo, hi_a, b_lo, hi_b, idesc, 1);
          umma_bf16_k2(tmem + p.n, a_lo + 768u, hi_a, b_lo, hi_b, idesc, 1);
          umma_bf16_k2(tmem + 2 * p.n, a_lo + 1536u, hi_a, b_lo, hi_b, idesc, 1);
          umma_bf16_k2(tmem + 3 * p.n, a_lo + 2304u, hi_a, b_lo, hi_b, idesc, 1);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (t0 != 0) p.cycles[blockIdx.x] = clock64() - t0;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  long long* dc;
  cudaMalloc(&dc, sms * sizeof(long long));
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  std::vector<long long> hc(sms);
  const int iters = 4096;
  struct Cfg { int n, shift, sbo, da, bshift; const char* name; int rb = 128; };
  const Cfg cfgs[] = {
      {256, 0, 8, 1, 0, "N256 aligned sbo8"},   {256, 1, 8, 1, 0, "N256 shift1 sbo8"},   {256, 0, 10, 1, 0, "N256 aligned sbo10"},
      {256, 11, 10, 1, 0, "N256 shift11 sbo10"}, {256, 0, 16, 1, 0, "N256 aligned sbo16"}, {256, 1, 16, 1, 0, "N256 shift1 sbo16"},
      {128, 0, 8, 1, 0, "N128 aligned sbo8"},   {128, 1, 8, 1, 0, "N128 shift1 sbo8"},   {128, 0, 10, 1, 0, "N128 aligned sbo10"},
      {128, 11, 10, 1, 0, "N128 shift11 sbo10"}, {128, 11, 10, 9, 0, "N128 shift11 sbo10 9 views"}, {128, 0, 8, 1, 3, "N128 aligned, B shift3"},
      {64, 0, 8, 1, 0, "N64 aligned sbo8"},     {64, 11, 10, 1, 0, "N64 shift11 sbo10"},
      {32, 0, 8, 1, 0, "N32 aligned sbo8"},     {32, 11, 10, 1, 0, "N32 shift11 sbo10"},
      {16, 0, 8, 1, 0, "N16 aligned sbo8"},
      {32, 0, 8, 1, 0, "N32 64B rows aligned sbo8", 64},   {32, 11, 10, 1, 0, "N32 64B rows shift11 sbo10", 64},
      {64, 0, 8, 1, 0, "N64 64B rows aligned sbo8", 64},   {64, 11, 10, 1, 0, "N64 64B rows shift11 sbo10", 64},
      {128, 11, 10, 1, 0, "N128 64B rows shift11 sbo10", 64},
  };
  for (const Cfg& c : cfgs) {
    Params p{c.rb, c.n, c.shift, c.sbo, iters, c.da, c.bshift, dc};
    for (int rep = 0; rep < 2; ++rep) {
      rate<<<sms, 128, smem>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", c.name, cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(hc.data(), dc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; long long mx = 0;
    for (int i = 0; i < sms; ++i) { avg += hc[i]; if (hc[i] > mx) mx = hc[i]; }
    avg /= sms;
    const double per = avg / iters;
    printf("%-30s cycles/MMA avg %.1f (max CTA %.1f)  ideal %.1f  -> %.0f%% of tensor peak\n", c.name, per, (double)mx / iters,
           c.n / 2.0, 100.0 * (c.n / 2.0) / per);
  }
  return 0;
}
