// Hardware probe (not part of libtod.so): tcgen05 cta_group::2 (CTA pair, M = 256) -- every primitive the pair mode of
// the conv kernel relies on, checked numerically against the host, then timed:
//   * cluster launch (2,1,1), barrier.cluster, tcgen05.alloc / dealloc .cta_group::2 issued by the same warp of both CTAs
//   * TMA tile loads with .cta_group::2 whose completion bytes land on the LEADER CTA's mbarrier (peer-bit-masked address)
//   * tcgen05.mma.cta_group::2: A rows [128 r, 128 r + 128) and B rows [N/2 r, N/2 r + N/2) live in CTA r's shared
//     memory at the same offsets; D rows [128 r, +128) in CTA r's TMEM
//   * tcgen05.commit.cta_group::2 ... multicast::cluster to the same barrier offset in both CTAs
//   * a remote mbarrier arrive from CTA 1 on a barrier of CTA 0 (mapa + mbarrier.arrive.shared::cluster)
// and the MMA rate: cta_group::2 at N = 64 / 128 / 256, and cta_group::1 with ONE accumulator (dependent chain) against
// two alternating accumulators.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -I include -o probe_pair tools/probe_pair.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../transparent_object_detection_b200/csrc/tod_common.cuh"

using namespace tod;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// TMA load issued by either CTA of a pair; the completion bytes are signalled on the barrier at `bar_addr` of the LEADER
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t bar_addr, uint32_t dst_smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_addr & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar_addr, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar_addr), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ uint64_t mk_desc(uint32_t addr, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | (1ull << 16) | ((addr >> 4) & 0x3FFFu);
}

struct FParams {
  CUtensorMap tm_a, tm_b;
  int n;
  float* d;   // [256, n]
  int* flag;
};

// ------------------------------------------------------------------ functional test: D[256, n] = A[256, 64] . B[n, 64]^T
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(192, 1) pair_func(const __grid_constant__ FParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full_bar, done_bar, remote_bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  if (threadIdx.x == 0) {
    mbar_init(&full_bar, 1);
    mbar_init(&done_bar, 1);
    mbar_init(&remote_bar, 2);   // one local arrive + one from the peer CTA
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc2(&tmem_base_smem, 256);
    tmem_relinquish2();
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers are initialised before anything can signal them
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  const uint32_t sa = base, sb = base + 16384;
  const uint32_t half_n = p.n / 2;
  if (warp == 0 && elect_one()) {
    if (rank == 0) mbar_arrive_expect_tx(&full_bar, 2u * (16384u + half_n * 128u));   // both CTAs' bytes land on the leader's barrier
    tma_load_2d_pair(&p.tm_a, smem_u32(&full_bar), sa, 0, static_cast<int>(rank) * 128);
    tma_load_2d_pair(&p.tm_b, smem_u32(&full_bar), sb, 0, static_cast<int>(rank * half_n));
  }
  if (warp == 1 && rank == 0 && elect_one()) {
    mbar_wait(&full_bar, 0);
    tcgen05_fence_after();
    const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(p.n) >> 3) << 17) | ((256u >> 4) << 24);
    for (int k = 0; k < 4; ++k) umma2_bf16(tmem, mk_desc(sa + k * 32, hi), mk_desc(sb + k * 32, hi), idesc, k != 0);
    umma2_commit_mc(&done_bar, 0b11);
  }
  if (warp >= 2) {   // four epilogue warps: lane quarter q = warp & 3
    mbar_wait(&done_bar, 0);
    tcgen05_fence_after();
    const int q = warp & 3, r = q * 32 + lane;
    for (int c0 = 0; c0 < p.n; c0 += 16) {
      uint32_t v[16];
      tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(q * 32) << 16) + c0, v);
      tmem_ld_wait();
      for (int i = 0; i < 16; ++i) p.d[static_cast<size_t>(rank * 128 + r) * p.n + c0 + i] = __uint_as_float(v[i]);
    }
    tcgen05_fence_before();
  }
  // remote arrive: both CTAs arrive on CTA 0's remote_bar; CTA 0 waits for both
  if (threadIdx.x == 0) {
    mbar_arrive_remote(smem_u32(&remote_bar), 0);
    if (rank == 0) {
      mbar_wait(&remote_bar, 0);
      *p.flag = 1;
    }
  }
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc2(tmem, 256);
  }
}

// ------------------------------------------------------------------ rate
struct RParams {
  int n, iters, pair, accs;   // accs: 1 = every MMA accumulates into the same TMEM tile, 2 = two tiles alternately
  long long* cycles;
};

template <bool PAIR>
__global__ void __launch_bounds__(128, 1) rate(const RParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int warp = threadIdx.x >> 5;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    if (PAIR) {
      tmem_alloc2(&tmem_base_smem, 512);
      tmem_relinquish2();
    } else {
      tmem_alloc(&tmem_base_smem, 512);
      tmem_relinquish();
    }
  }
  for (uint32_t i = threadIdx.x; i < (160u * 1024u) / 16; i += blockDim.x)
    asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 16), "r"(0) : "memory");
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (warp == 1 && rank == 0) {
    const uint32_t sa = base, sb = base + 96 * 1024;
    const uint32_t hi = ((8u * 128u) >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((static_cast<uint32_t>(p.n) >> 3) << 17) |
                           (((PAIR ? 256u : 128u) >> 4) << 24);
    long long t0 = 0;
    if (elect_one()) {
      t0 = clock64();
      const uint32_t second = p.accs == 2 ? static_cast<uint32_t>(p.n) : 0u;
      for (int it = 0; it < p.iters; it += 8) {
        for (int k = 0; k < 4; ++k) {
          if (PAIR) umma2_bf16(tmem, mk_desc(sa + k * 32, hi), mk_desc(sb + k * 32, hi), idesc, 1);
          else umma_bf16(tmem, mk_desc(sa + k * 32, hi), mk_desc(sb + k * 32, hi), idesc, 1);
        }
        for (int k = 0; k < 4; ++k) {
          if (PAIR) umma2_bf16(tmem + second, mk_desc(sa + 16384 + k * 32, hi), mk_desc(sb + k * 32, hi), idesc, 1);
          else umma_bf16(tmem + second, mk_desc(sa + 16384 + k * 32, hi), mk_desc(sb + k * 32, hi), idesc, 1);
        }
      }
      if (PAIR) umma2_commit_mc(&bar, 0b01);
      else umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if (t0 != 0) p.cycles[blockIdx.x] = clock64() - t0;
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

static float bf16_round(float x) { return __bfloat162float(__float2bfloat16(x)); }

int main() {
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  // ---------------- functional
  for (int n : {64, 128, 256}) {
    std::vector<__nv_bfloat16> ha(256 * 64), hb(n * 64);
    std::vector<float> fa(256 * 64), fb(n * 64);
    srand(7 + n);
    for (size_t i = 0; i < ha.size(); ++i) { fa[i] = bf16_round((rand() % 2001 - 1000) / 500.0f); ha[i] = __float2bfloat16(fa[i]); }
    for (size_t i = 0; i < hb.size(); ++i) { fb[i] = bf16_round((rand() % 2001 - 1000) / 500.0f); hb[i] = __float2bfloat16(fb[i]); }
    __nv_bfloat16 *da, *db;
    float* dd;
    int* dflag;
    cudaMalloc(&da, ha.size() * 2); cudaMalloc(&db, hb.size() * 2); cudaMalloc(&dd, 256 * n * 4); cudaMalloc(&dflag, 4);
    cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, 256 * n * 4); cudaMemset(dflag, 0, 4);
    FParams p;
    memset(&p, 0, sizeof(p));
    p.n = n; p.d = dd; p.flag = dflag;
    {
      cuuint64_t gd[2] = {64, 256}, gs[1] = {128};
      cuuint32_t bx[2] = {64, 128}, es[2] = {1, 1};
      CUresult r = cuTensorMapEncodeTiled(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      cuuint64_t gd2[2] = {64, static_cast<cuuint64_t>(n)};
      cuuint32_t bx2[2] = {64, static_cast<cuuint32_t>(n / 2)};
      CUresult r2 = cuTensorMapEncodeTiled(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, gd2, gs, bx2, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS || r2 != CUDA_SUCCESS) { printf("tensor map encode failed %d %d\n", (int)r, (int)r2); return 1; }
    }
    cudaFuncSetAttribute(pair_func, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    pair_func<<<2, 192, 64 * 1024>>>(p);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("pair_func N=%d: CUDA error %s\n", n, cudaGetErrorString(e)); return 1; }
    std::vector<float> hd(256 * n);
    int hflag = 0;
    cudaMemcpy(hd.data(), dd, hd.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hflag, dflag, 4, cudaMemcpyDeviceToHost);
    double maxerr = 0;
    for (int i = 0; i < 256; ++i)
      for (int j = 0; j < n; ++j) {
        double s = 0;
        for (int k = 0; k < 64; ++k) s += static_cast<double>(fa[i * 64 + k]) * fb[j * 64 + k];
        maxerr = fmax(maxerr, fabs(s - hd[i * n + j]));
      }
    printf("pair functional N=%3d: max |err| = %.3e  remote-arrive flag %d  -> %s\n", n, maxerr, hflag,
           (maxerr < 1e-3 && hflag == 1) ? "OK" : "MISMATCH");
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dflag);
  }
  // ---------------- rate
  long long* dc;
  cudaMalloc(&dc, sms * sizeof(long long));
  const size_t smem = 200 * 1024;
  cudaFuncSetAttribute(rate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  cudaFuncSetAttribute(rate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  std::vector<long long> hc(sms);
  const int iters = 4096;
  struct Cfg { int n, pair, accs; };
  const Cfg cfgs[] = {{64, 0, 1}, {64, 0, 2}, {128, 0, 1}, {128, 0, 2}, {256, 0, 1}, {32, 0, 1}, {32, 0, 2},
                      {64, 1, 1}, {64, 1, 2}, {128, 1, 1}, {128, 1, 2}, {256, 1, 1}, {256, 1, 2}, {32, 1, 2}};
  for (const Cfg& c : cfgs) {
    RParams p{c.n, iters, c.pair, c.accs, dc};
    cudaMemset(dc, 0, sms * sizeof(long long));
    for (int rep = 0; rep < 2; ++rep) {
      if (c.pair) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sms - sms % 2);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, rate<true>, p);
      } else {
        rate<false><<<sms, 128, smem>>>(p);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("rate: CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    }
    cudaMemcpy(hc.data(), dc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
    double avg = 0; int cnt = 0;
    for (int i = 0; i < sms; ++i) if (hc[i] > 0) { avg += hc[i]; ++cnt; }
    avg /= cnt;
    const double per = avg / iters;
    const double ideal = c.n / 2.0;   // per SM: 128 x N x 16 per MMA (a pair does 256 x N x 16 on two SMs)
    printf("%s N=%3d accs=%d   cycles/MMA %.1f   ideal %.1f -> %.0f%% of tensor peak (%d issuing CTAs)\n", c.pair ? "cta_group::2 M=256" : "cta_group::1 M=128",
           c.n, c.accs, per, ideal, 100.0 * ideal / per, cnt);
  }
  return 0;
}
