// Hardware probe (not part of libtod.so): how much does background shared-memory traffic slow a stream of tcgen05 MMAs?
// M = 128 (or a CTA pair, M = 256), K = 16 bf16 MMAs read A (128 x 32 B) and B (N x 32 B, half of it per CTA in a pair) from
// shared memory: at N = 128 that is 8 KB per 64 cycles = the SM's whole 128 B/clk of shared-memory bandwidth.  The conv
// kernel's TMA fills (~40 B/clk/SM) and epilogue staging (~14 B/clk/SM) land in the same shared memory.  Here warp 2
// streams TMA loads from global memory into a ring at a chosen pace while warp 1 issues MMAs; warp 3 optionally writes
// and reads a staging buffer with st.shared / ld.shared like the epilogue does.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include -o probe_smem tools/probe_smem_contention.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../transparent_object_detection_b200/csrc/tod_common.cuh"

using namespace tod;

struct Params {
  CUtensorMap tm;      // [rows, 64] bf16 global tensor, box {64, 128} = 16 KB
  int n, iters, pair;
  int tma_every;       // one 16 KB TMA load per `tma_every` MMAs (0 = none): bytes/clk = 16384 / (tma_every * cycles_per_mma)
  int lsu;             // 1: a warp group also does st.shared + ld.shared of 16 KB per `tma_every` MMAs
  int rows;
  long long* cycles;
};

__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | (1ull << 16) | ((addr >> 4) & 0x3FFFu);
}
__device__ __forceinline__ void umma2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d), "l"(da),
               "l"(db), "r"(idesc)
               : "memory");
}

template <bool PAIR>
__global__ void __launch_bounds__(256, 1) contention(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t done_bar, ring_full[4];
  __shared__ uint32_t tmem_base_smem;
  __shared__ volatile int stop_flag;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(&done_bar, 1);
    for (int i = 0; i < 4; ++i) mbar_init(&ring_full[i], 1);
    fence_mbar_init();
    stop_flag = 0;
  }
  if (warp == 0) {
    if (PAIR) { tmem_alloc2(&tmem_base_smem, 512); tmem_relinquish2(); }
    else { tmem_alloc(&tmem_base_smem, 512); tmem_relinquish(); }
  }
  for (uint32_t i = threadIdx.x; i < (192u * 1024u) / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16), "r"(0) : "memory");
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t sa = base, sb = base + 48 * 1024, ring = base + 96 * 1024, stage = base + 160 * 1024;   // ring: 4 x 16 KB
  if (warp == 1) {
    const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(p.n) >> 3) << 17) | (((PAIR ? 256u : 128u) >> 4) << 24);
    long long t0 = 0;
    if (rank == 0 && elect_one()) {
      t0 = clock64();
      for (int it = 0; it < p.iters; it += 8) {
        for (int k = 0; k < 4; ++k) {
          if (PAIR) umma2(tmem, mk_desc(sa + k * 32, hi), mk_desc(sb + k * 32, hi), idesc);
          else umma_bf16(tmem, mk_desc(sa + k * 32, hi), mk_desc(sb + k * 32, hi), idesc, 1);
        }
        for (int k = 0; k < 4; ++k) {
          if (PAIR) umma2(tmem + p.n, mk_desc(sa + 16384 + k * 32, hi), mk_desc(sb + k * 32, hi), idesc);
          else umma_bf16(tmem + p.n, mk_desc(sa + 16384 + k * 32, hi), mk_desc(sb + k * 32, hi), idesc, 1);
        }
      }
      if (PAIR) umma2_commit_both(smem_u32(&done_bar));
      else umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);          // (pair: the leader's commit arrives on both CTAs' barriers)
    if (t0 != 0) p.cycles[blockIdx.x] = clock64() - t0;
    stop_flag = 1;
  } else if (warp == 2 && p.tma_every > 0) {
    // background TMA loads paced against the clock: one 16 KB box per tma_every * (n / 2) cycles
    if (elect_one()) {
      const long long period = static_cast<long long>(p.tma_every) * (p.n / 2);
      long long next = clock64();
      int slot = 0, row = (blockIdx.x * 977) % (p.rows - 128);
      uint32_t phase[4] = {0, 0, 0, 0};
      int inflight = 0;
      while (!stop_flag) {
        if (clock64() >= next) {
          if (inflight == 4) {                         // ring full: wait for the oldest
            mbar_wait(&ring_full[slot], phase[slot]);
            phase[slot] ^= 1u;
            --inflight;
          }
          mbar_arrive_expect_tx(&ring_full[slot], 16384u);
          tma_load_2d(&p.tm, &ring_full[slot], ring + slot * 16384, 0, row);
          row = (row + 128) % (p.rows - 128);
          slot = (slot + 1) & 3;
          ++inflight;
          next += period;
        }
      }
      while (inflight > 0) {                            // drain before exit
        const int s_ = (slot - inflight) & 3;
        mbar_wait(&ring_full[s_], phase[s_]);
        phase[s_] ^= 1u;
        --inflight;
      }
    }
    __syncwarp();
  } else if (warp >= 4 && p.lsu) {
    // epilogue-like traffic: each of 128 threads writes and reads back its 128-byte row (16 KB per round), paced like the TMA
    const int r = threadIdx.x - 128;
    const long long period = static_cast<long long>(p.tma_every > 0 ? p.tma_every : 8) * (p.n / 2);
    long long next = clock64();
    uint32_t acc = 0;
    while (!stop_flag) {
      if (clock64() >= next) {
        for (int j = 0; j < 8; ++j) {
          const uint32_t a = stage + r * 128 + ((j ^ (r & 7)) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(a), "r"(acc) : "memory");
        }
        for (int j = 0; j < 8; ++j) {
          uint32_t x, y, z, w;
          const uint32_t a = stage + r * 128 + ((j ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a));
          acc += x + y + z + w;
        }
        next += period;
      }
    }
    if (acc == 0x12345678u) p.cycles[0] = acc;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  if (warp == 0) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc2(tmem, 512);
    else tmem_dealloc(tmem, 512);
  }
}

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int rows = 1 << 20;                      // 128 MB of bf16 [rows, 64]: larger than L2 is not needed, L2-resident is the conv's case too
  __nv_bfloat16* g;
  cudaMalloc(&g, static_cast<size_t>(rows) * 64 * 2);
  cudaMemset(g, 0, static_cast<size_t>(rows) * 64 * 2);
  long long* dc;
  cudaMalloc(&dc, sms * sizeof(long long));
  Params p;
  memset(&p, 0, sizeof(p));
  {
    cuuint64_t gd[2] = {64, static_cast<cuuint64_t>(rows)}, gs[1] = {128};
    cuuint32_t bx[2] = {64, 128}, es[2] = {1, 1};
    CUresult r = cuTensorMapEncodeTiled(&p.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  }
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(contention<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  cudaFuncSetAttribute(contention<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  std::vector<long long> hc(sms);
  const int iters = 8192;
  struct Cfg { int n, pair, tma_every, lsu; };
  const Cfg cfgs[] = {{128, 0, 0, 0}, {128, 0, 8, 0}, {128, 0, 4, 0}, {128, 0, 2, 0}, {128, 0, 4, 1}, {128, 0, 0, 1},
                      {128, 1, 0, 0}, {128, 1, 8, 0}, {128, 1, 4, 0}, {128, 1, 2, 0}, {128, 1, 4, 1},
                      {256, 0, 0, 0}, {256, 0, 2, 0}, {256, 1, 0, 0}, {256, 1, 2, 0}, {256, 1, 1, 1},
                      {64, 0, 0, 0}, {64, 0, 8, 0}, {64, 0, 4, 1}};
  for (const Cfg& c : cfgs) {
    p.n = c.n; p.iters = iters; p.pair = c.pair; p.tma_every = c.tma_every; p.lsu = c.lsu; p.rows = rows; p.cycles = dc;
    cudaMemset(dc, 0, sms * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(sms - (c.pair ? sms % 2 : 0));
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = smem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c.pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      if (c.pair) cudaLaunchKernelEx(&cfg, contention<true>, p);
      else cudaLaunchKernelEx(&cfg, contention<false>, p);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(hc.data(), dc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; int cnt = 0;
    for (int i = 0; i < sms; ++i) if (hc[i] > 0) { avg += hc[i]; ++cnt; }
    avg /= cnt;
    const double per = avg / iters, ideal = c.n / 2.0;
    const double tma_bpc = c.tma_every ? 16384.0 / (c.tma_every * ideal) : 0.0;
    printf("%s N=%3d  background TMA %5.1f B/clk/SM (nominal), st/ld.shared %s   cycles/MMA %.1f (ideal %.1f) -> %.0f%% of tensor peak\n",
           c.pair ? "pair M=256" : "single M=128", c.n, tma_bpc, c.lsu ? "yes" : "no ", per, ideal, 100.0 * ideal / per);
    fflush(stdout);
  }
  return 0;
}
