// Hardware probe (not part of libtod.so): throughput of the instructions the conv epilogue is made of, per SM, with
// one CTA per SM on every SM:  MUFU (tanh / ex2 / rcp, f32 and packed 16-bit forms) and tcgen05.ld (32x32b x16/x32/x64)
// with 4 and 8 warps.  Answers: is SiLU's one MUFU per element or the TMEM read the floor of an HBM-bound conv epilogue?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_epilogue probe_epilogue.cu
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../transparent_object_detection_b200/csrc/tod_common.cuh"

using namespace tod;

constexpr int kIters = 512;

template <int OP>
__device__ __forceinline__ uint32_t sfu(uint32_t x) {
  uint32_t y;
  if (OP == 0) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 3) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 4) asm volatile("tanh.approx.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 5) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  if (OP == 6) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

template <int OP>
__global__ void mufu_rate(long long* cycles, uint32_t* sink) {
  uint32_t v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0x3c003c00u + threadIdx.x * 8 + j;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = sfu<OP>(v[j]);
  }
  __syncthreads();
  const long long t1 = clock64();
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) s ^= v[j];
  if (s == 0x12345678u) sink[0] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int X>
__global__ void __launch_bounds__(256, 1) ldtm_rate(long long* cycles, uint32_t* sink, int iters) {
  __shared__ uint32_t tmem_base_smem;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  const long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
    const uint32_t col = (it * X) & 255;
    if (X == 16) {
      uint32_t r[16];
      tmem_ld_32x32b_x16(tmem + col, r);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r[j];
    } else if (X == 32) {
      uint32_t r0[16], r1[16];
      tmem_ld_32x32b_x16(tmem + col, r0);
      tmem_ld_32x32b_x16(tmem + col + 16, r1);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r0[j] ^ r1[j];
    } else {
      uint32_t r0[16], r1[16], r2[16], r3[16];
      tmem_ld_32x32b_x16(tmem + col, r0);
      tmem_ld_32x32b_x16(tmem + col + 16, r1);
      tmem_ld_32x32b_x16(tmem + col + 32, r2);
      tmem_ld_32x32b_x16(tmem + col + 48, r3);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) acc ^= r0[j] ^ r1[j] ^ r2[j] ^ r3[j];
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if (acc == 0x12345678u) sink[0] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base_smem, 512);
  }
}

static double avg(const std::vector<long long>& v) {
  double s = 0;
  for (auto x : v) s += x;
  return s / v.size();
}

int main() {
  long long* d_cycles;
  uint32_t* d_sink;
  cudaMalloc(&d_cycles, 148 * sizeof(long long));
  cudaMalloc(&d_sink, 64);
  std::vector<long long> h(148);
  const char* names[7] = {"tanh.approx.f32", "ex2.approx.f32", "rcp.approx.f32", "tanh.approx.f16x2", "tanh.approx.bf16x2",
                          "ex2.approx.bf16x2", "ex2.approx.f16x2"};
  for (int warps : {4, 8, 16}) {
    for (int op = 0; op < 7; ++op) {
      for (int rep = 0; rep < 2; ++rep) {
        switch (op) {
          case 0: mufu_rate<0><<<148, warps * 32>>>(d_cycles, d_sink); break;
          case 1: mufu_rate<1><<<148, warps * 32>>>(d_cycles, d_sink); break;
          case 2: mufu_rate<2><<<148, warps * 32>>>(d_cycles, d_sink); break;
          case 3: mufu_rate<3><<<148, warps * 32>>>(d_cycles, d_sink); break;
          case 4: mufu_rate<4><<<148, warps * 32>>>(d_cycles, d_sink); break;
          case 5: mufu_rate<5><<<148, warps * 32>>>(d_cycles, d_sink); break;
          case 6: mufu_rate<6><<<148, warps * 32>>>(d_cycles, d_sink); break;
        }
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h.data(), d_cycles, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      const double cyc = avg(h);
      const double warp_instr = static_cast<double>(warps) * 8 * kIters;
      printf("MUFU %-20s %2d warps/SM: %.2f cycles per warp-instruction per SMSP  -> %.1f lane-ops/clk/SM\n", names[op], warps,
             cyc * 4 / warp_instr, warp_instr * 32 / cyc);
    }
  }
  for (int warps : {4, 8}) {
    for (int x : {16, 32, 64}) {
      const int iters = 4096;
      for (int rep = 0; rep < 2; ++rep) {
        if (x == 16) ldtm_rate<16><<<148, warps * 32>>>(d_cycles, d_sink, iters);
        if (x == 32) ldtm_rate<32><<<148, warps * 32>>>(d_cycles, d_sink, iters);
        if (x == 64) ldtm_rate<64><<<148, warps * 32>>>(d_cycles, d_sink, iters);
        cudaDeviceSynchronize();
      }
      cudaMemcpy(h.data(), d_cycles, 148 * sizeof(long long), cudaMemcpyDeviceToHost);
      const double cyc = avg(h);
      const double bytes = static_cast<double>(warps) * 32 * x * 4 * iters;
      printf("LDTM 32x32b, %2d columns per wait, %d warps/SM: %.1f B/clk/SM, %.1f cycles per (ld..wait) round\n", x, warps,
             bytes / cyc, cyc / iters);
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : 1;
}
