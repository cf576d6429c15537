// Host-side helpers shared by the conv kernels: tensor-map encoding through the driver entry point, SM count.
#pragma once

#include <cstdlib>
#include <mutex>
#include <utility>

#include "tod_common.cuh"

namespace tod {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

// dims / box are innermost first; strides_bytes[i] is the byte stride of dimension i+1.
inline int encode_map(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, CUtensorMapSwizzle swz,
                      CUtensorMapDataType dtype = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                      CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled entry point not available");
    return TOD_ERR_CUDA;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i - 1];
  }
  CUresult r = fn(map, dtype, rank, const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                  promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u] "
              "stride0 %llu base %p",
              (int)r, rank, (unsigned long long)dims[0], (unsigned long long)dims[1],
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1],
              rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0, (unsigned long long)strides_bytes[0], base);
    return TOD_ERR_CUDA;
  }
  return TOD_OK;
}

// Programmatic dependent launch is on unless TOD_PDL=0 (A/B measurements).  Kernels launched through launch_pdl call
// pdl_wait() before they touch anything an earlier kernel of the stream wrote or still reads.
inline bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TOD_PDL");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// EXPERIMENT ONLY (results are wrong: races): TOD_PDL_NOWAIT=1 makes the conv kernels skip griddepcontrol.wait, which bounds
// from above what tile-granular inter-layer dependencies could gain over the grid-wide wait.
inline int pdl_nowait() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TOD_PDL_NOWAIT");
    on = (e != nullptr && e[0] == '1') ? 1 : 0;
  }
  return on;
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// Same, as (2,1,1) clusters: CTA pairs for the tcgen05 cta_group::2 kernels (grid.x must be even).
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_pair(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = 2;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// CTA-pair conv kernels: TOD_PAIR=0 never, 2 wherever the kernel supports it, default (1) where the plan's rule says so.
inline int pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("TOD_PAIR");
    mode = (e != nullptr) ? atoi(e) : 1;
  }
  return mode;
}

// L2 promotion (the granularity at which a TMA load's misses are fetched into L2) for a tensor whose innermost box covers
// `window_bytes` of a `pitch_bytes` pixel: a channel WINDOW of a wider concat buffer (C2f's chunk / bottleneck inputs,
// model/blocks.py:104-108) must not be promoted past its own width, or every miss drags the neighbouring windows' bytes
// out of DRAM as well (ncu, round 1: 2.0x the algorithmic DRAM reads on the 32-of-96-channel layers at 160^2).
// TOD_L2PROMO=256 restores the fixed 256-byte promotion for A/B measurements.
inline CUtensorMapL2promotion l2_promotion_for(uint64_t window_bytes, uint64_t pitch_bytes) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("TOD_L2PROMO");
    mode = (e != nullptr) ? atoi(e) : 0;
  }
  if (window_bytes >= pitch_bytes || mode == 256) return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;   // dense tensor
  if (mode == 128) return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (mode == 64) return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
  if (mode == 1) return CU_TENSOR_MAP_L2_PROMOTION_NONE;
  if (window_bytes >= 256) return CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (window_bytes >= 128) return CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (window_bytes >= 64) return CU_TENSOR_MAP_L2_PROMOTION_L2_64B;
  return CU_TENSOR_MAP_L2_PROMOTION_NONE;
}

// Row-flat tiles for 3x3 stride-1 convs on small maps (conv_halo_tcgen05.cu); TOD_FLAT=0 keeps 16x8 patches / the per-tap
// kernel for A/B measurements.
inline bool flat_tiles_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("TOD_FLAT");
    on = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

// SM budget of one launch (tod_set_sm_budget / TOD_SM_BUDGET; 0 = the whole device).  The persistent kernels size their
// grids by it.  K independent plans (batches in flight on K streams) with a budget of #SMs / K each run side by side on
// disjoint SMs: every CTA then lives K times longer, so what a CTA pays once per launch -- prologue, the wait for the
// previous grid, the latency of its first loads, the drain of its last accumulator: ~8-10 us of a ~35 us CTA, during which
// its SM's tensor pipe idles -- is amortised over K times more tiles.
int sm_budget_value();   // capi.cu

inline int num_sms() {
  static int sms[64] = {0};
  const int dev = current_device_index();
  if (sms[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    sms[dev] = n;
  }
  const int budget = sm_budget_value();
  return (budget > 0 && budget < sms[dev]) ? budget : sms[dev];
}

inline int pick_block_k(int cin, int hint) {
  if (hint == 16 || hint == 32 || hint == 64) return hint;
  if (cin % 64 == 0) return 64;
  if (cin % 32 == 0) return 32;
  return 16;
}

// Entry of the halo-patch kernel (conv_halo_tcgen05.cu); the descriptor has already been validated.
int conv_halo_launch(const tod_conv_desc* d, void* stream, const tod_head_fuse_desc* fuse = nullptr);
int conv_halo_launch_tail(const tod_conv_desc* d, const tod_conv_tail_desc* t, void* stream,
                          const tod_head_fuse_desc* fuse = nullptr);

}  // namespace tod
