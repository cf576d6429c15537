// Hardware probe (not part of libtod.so): does tcgen05.mma accept a K-major SWIZZLE_128B shared-memory descriptor whose
// start address is shifted by whole 128-byte rows (not 1024-byte aligned), and a stride-byte-offset that is not a
// multiple of 1024?  The answer decides whether a 3x3 conv can read all nine taps out of ONE halo patch in smem.
//
//   A (smem) : R rows x 64 bf16, written by TMA with SWIZZLE_128B (row r at linear offset r*128)
//   B (smem) : 64 x 64 identity  ->  D[i][j] = A[row(i)][j]
//   row(i) = shift + (i / 8) * sbo_rows + (i % 8)
// For each (shift, sbo_rows, base_offset mode) the kernel computes D and the host checks it against A.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_rowshift probe_umma_rowshift.cu -lcuda
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../transparent_object_detection_b200/csrc/tod_common.cuh"

using namespace tod;

constexpr int kRows = 512;  // rows of A held in smem (64 KB)

struct Params {
  CUtensorMap tm_a, tm_b;
  int shift, sbo_rows, base_mode;
  float* out;  // [128][64]
};

__global__ void __launch_bounds__(128, 1) probe(const __grid_constant__ Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_load, bar_mma;
  __shared__ uint32_t tmem_base_smem;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sa = base, sb = base + kRows * 128;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bar_load, 1);
    mbar_init(&bar_mma, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(&tmem_base_smem, 64);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_smem;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bar_load, kRows * 128 + 64 * 128);
    for (int r = 0; r < kRows; r += 256) tma_load_2d(&p.tm_a, &bar_load, sa + r * 128, 0, r);
    tma_load_2d(&p.tm_b, &bar_load, sb, 0, 0);
    mbar_wait(&bar_load, 0);
    tcgen05_fence_after();
    const uint32_t start = sa + p.shift * 128;
    uint32_t base_off = 0;
    if (p.base_mode == 1) base_off = (start >> 7) & 7;
    const uint32_t sbo = p.sbo_rows * 128;
    const uint32_t hi_a = (sbo >> 4) | (1u << 14) | (base_off << 17) | (2u << 29);
    const uint32_t hi_b = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = (static_cast<uint64_t>(hi_a) << 32) | (1ull << 16) | (((start + 32 * k) >> 4) & 0x3FFFu);
      const uint64_t db = (static_cast<uint64_t>(hi_b) << 32) | (1ull << 16) | (((sb + 32 * k) >> 4) & 0x3FFFu);
      umma_bf16(tmem, da, db, idesc, k != 0);
    }
    umma_commit(&bar_mma);
  }
  __syncthreads();
  mbar_wait(&bar_mma, 0);
  tcgen05_fence_after();
  for (int c = 0; c < 64; c += 16) {
    uint32_t v[16];
    tmem_ld_32x32b_x16(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) p.out[(warp * 32 + lane) * 64 + c + j] = __uint_as_float(v[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) {
    tcgen05_fence_after();
    tmem_dealloc(tmem, 64);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return 2;
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(sym);
  std::vector<__nv_bfloat16> ha(kRows * 64), hb(64 * 64);
  for (int r = 0; r < kRows; ++r)
    for (int c = 0; c < 64; ++c) ha[r * 64 + c] = __float2bfloat16(static_cast<float>((r * 7 + c * 3) % 251) - 125.f);
  for (int r = 0; r < 64; ++r)
    for (int c = 0; c < 64; ++c) hb[r * 64 + c] = __float2bfloat16(r == c ? 1.f : 0.f);
  __nv_bfloat16 *da, *db;
  float* dout;
  cudaMalloc(&da, ha.size() * 2);
  cudaMalloc(&db, hb.size() * 2);
  cudaMalloc(&dout, 128 * 64 * 4);
  cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice);
  Params p;
  memset(&p, 0, sizeof(p));
  {
    cuuint64_t dims[2] = {64, kRows};
    cuuint64_t str[1] = {128};
    cuuint32_t box[2] = {64, 256}, es[2] = {1, 1};
    if (enc(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, da, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 3;
    cuuint64_t dimsb[2] = {64, 64};
    cuuint32_t boxb[2] = {64, 64};
    if (enc(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, db, dimsb, str, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 3;
  }
  p.out = dout;
  const size_t smem = kRows * 128 + 64 * 128 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  std::vector<float> ho(128 * 64);
  const int shifts[] = {0, 1, 2, 3, 5, 8, 9, 10, 11, 20, 22, 23, 43};
  const int sbos[] = {8, 10, 12, 16, 18, 22, 24};
  for (int mode = 0; mode < 2; ++mode)
    for (int sbo : sbos) {
      printf("base_mode %d sbo_rows %2d :", mode, sbo);
      for (int shift : shifts) {
        p.shift = shift; p.sbo_rows = sbo; p.base_mode = mode;
        cudaMemset(dout, 0, 128 * 64 * 4);
        probe<<<1, 128, smem>>>(p);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" shift %d: CUDA error %s\n", shift, cudaGetErrorString(e)); return 4; }
        cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int i = 0; i < 128; ++i) {
          const int r = shift + (i / 8) * sbo + (i % 8);
          for (int c = 0; c < 64; ++c)
            if (ho[i * 64 + c] != __bfloat162float(ha[r * 64 + c])) ++bad;
        }
        printf(" s%d=%s", shift, bad == 0 ? "OK" : "bad");
        if (bad) printf("(%d)", bad);
      }
      printf("\n");
    }
  return 0;
}
