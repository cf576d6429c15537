// Hardware probe (not part of libtod.so): how far can the MMA-issuing thread run ahead of the tensor pipe, and what do
// concurrent TMEM loads / shared-memory stores from other warps cost the MMA stream?
//   E1  issue n MMAs (M=128, N=128, K=16): cycles until the last issue returns vs cycles until completion
//   E2  4096 MMAs while warps 4..7 loop on tcgen05.ld of the other half of TMEM
//   E3  4096 MMAs while warps 4..7 stream st.shared.v4 into an unused shared-memory region
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_queue probe_umma_queue.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

#include "../transparent_object_detection_b200/csrc/tod_common.cuh"

using namespace tod;

struct Params {
  int n_mma, mode, nblk;   // mode 0: plain, 1: concurrent tcgen05.ld, 2: concurrent st.shared
  long long* out;          // per CTA: [issue cycles, total cycles]
};

__global__ void __launch_bounds__(256, 1) probe(const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  __shared__ volatile int stop_flag;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
    stop_flag = 0;
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 512);
    tmem_relinquish();
  }
  for (uint32_t i = threadIdx.x; i < (96u * 1024u) / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16), "r"(0x3f803f80) : "memory");   // bf16 1.0
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (warp == 1) {
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(p.nblk) >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_lo = umma_desc_lo(base), b_lo = umma_desc_lo(base + 32 * 1024);
    long long t0 = 0, t1 = 0;
    if (elect_one()) {
      t0 = clock64();
      for (int it = 0; it < p.n_mma; it += 4) umma_bf16_k4(tmem + ((it >> 2) & 1) * p.nblk, a_lo, hi, b_lo, hi, idesc, 1);
      t1 = clock64();
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (t0 != 0) {
      p.out[blockIdx.x * 2] = t1 - t0;
      p.out[blockIdx.x * 2 + 1] = clock64() - t0;
      stop_flag = 1;
    }
  } else if (warp >= 4 && p.mode == 1) {
    // TMEM loads of columns [256, 512) (never written by the MMAs) until the MMA stream is done
    uint32_t sink = 0;
    while (!stop_flag) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16) + 256 + ((sink & 7) * 16), v);
      tmem_ld_wait();
      sink += v[0] + 1;
    }
    if (sink == 0xdeadbeef) p.out[0] = sink;
  } else if (warp >= 4 && p.mode == 2) {
    uint32_t i = 0;
    const uint32_t dst = base + 100 * 1024 + (warp - 4) * 16384;
    while (!stop_flag) {
      asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(dst + ((i * 32 + lane) & 1023) * 16), "r"(i) : "memory");
      ++i;
    }
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
  long long* d;
  cudaMalloc(&d, sms * 2 * sizeof(long long));
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  std::vector<long long> h(sms * 2);
  auto run = [&](int n, int mode, int nblk, const char* tag) {
    Params p{n, mode, nblk, d};
    for (int rep = 0; rep < 2; ++rep) {
      probe<<<sms, 256, smem>>>(p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("%s: CUDA error %s\n", tag, cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h.data(), d, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
    double iss = 0, tot = 0;
    for (int i = 0; i < sms; ++i) { iss += h[2 * i]; tot += h[2 * i + 1]; }
    printf("%-34s N=%3d n=%5d  issue %8.0f cyc (%.1f/MMA)  total %8.0f cyc (%.1f/MMA)\n", tag, nblk, n, iss / sms, iss / sms / n,
           tot / sms, tot / sms / n);
  };
  for (int n : {4, 8, 16, 32, 64, 128, 256}) run(n, 0, 128, "E1 queue depth");
  for (int n : {4, 8, 16, 32, 64}) run(n, 0, 256, "E1 queue depth");
  run(4096, 0, 128, "E2 baseline");
  run(4096, 1, 128, "E2 + concurrent tcgen05.ld");
  run(4096, 2, 128, "E3 + concurrent st.shared");
  run(4096, 0, 256, "E2 baseline");
  run(4096, 1, 256, "E2 + concurrent tcgen05.ld");
  run(4096, 2, 256, "E3 + concurrent st.shared");
  run(4096, 0, 64, "E2 baseline");
  run(4096, 1, 64, "E2 + concurrent tcgen05.ld");
  return 0;
}
