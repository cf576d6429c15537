// Row softmax f32 -> bf16: the attention weights of the reference's SelfAttention block (model/blocks.py:246-247,
// nn.Softmax(dim=-1) over the key axis) -- SURVEY.md section 8 row f1.  One block per row; the row (<= 32 KB) is read
// three times (maximum, sum of exponentials, normalised write) and stays in L1 between the passes.
#include "tod_common.cuh"

namespace tod {

constexpr int kSoftmaxThreads = 256;

__device__ __forceinline__ float block_reduce(float v, bool is_max, float* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();               // sh may still be read from the previous reduction
  if (lane == 0) sh[warp] = v;
  __syncthreads();
  float r = sh[0];
  for (int i = 1; i < kSoftmaxThreads / 32; ++i) r = is_max ? fmaxf(r, sh[i]) : r + sh[i];   // same order in every thread
  return r;
}

__global__ void __launch_bounds__(kSoftmaxThreads) softmax_rows_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                                      int cols, long long in_pitch, long long out_pitch) {
  __shared__ float sh[kSoftmaxThreads / 32];
  const float* row = in + blockIdx.x * in_pitch;
  __nv_bfloat16* orow = out + blockIdx.x * out_pitch;
  float m = -INFINITY;
  for (int i = threadIdx.x * 4; i < cols; i += kSoftmaxThreads * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
  }
  m = block_reduce(m, true, sh);
  float s = 0.0f;
  for (int i = threadIdx.x * 4; i < cols; i += kSoftmaxThreads * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    s += (__expf(v.x - m) + __expf(v.y - m)) + (__expf(v.z - m) + __expf(v.w - m));
  }
  s = block_reduce(s, false, sh);
  const float inv = 1.0f / s;
  for (int i = threadIdx.x * 4; i < cols; i += kSoftmaxThreads * 4) {
    const float4 v = *reinterpret_cast<const float4*>(row + i);
    uint2 o;
    o.x = pack_bf16x2(__expf(v.x - m) * inv, __expf(v.y - m) * inv);
    o.y = pack_bf16x2(__expf(v.z - m) * inv, __expf(v.w - m) * inv);
    *reinterpret_cast<uint2*>(orow + i) = o;
  }
}

// [batch][rows][cols] bf16 -> [batch][cols][rows] bf16 through a 64 x 64 shared-memory tile: the value projection of
// SelfAttention comes out of the batched 1x1 conv as [B, N, C] and the attention kernel wants V^T [B, C, N] (K-major for
// P . V^T).  One launch instead of one small GEMM per image.
__global__ void __launch_bounds__(256) transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                            int rows, int cols, long long in_pitch, long long out_pitch) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const __nv_bfloat16* src = in + static_cast<long long>(blockIdx.z) * rows * in_pitch;
  __nv_bfloat16* dst = out + static_cast<long long>(blockIdx.z) * cols * out_pitch;
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {          // 64 rows x 32 bf16 pairs
    const int r = i >> 5, cp = (i & 31) * 2;
    if (r0 + r < rows && c0 + cp < cols) {
      const __nv_bfloat162 v = *reinterpret_cast<const __nv_bfloat162*>(src + (r0 + r) * in_pitch + c0 + cp);
      tile[r][cp] = v.x;
      tile[r][cp + 1] = v.y;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * 32; i += 256) {          // 64 output rows (= input columns) x 32 pairs of input rows
    const int c = i >> 5, rp = (i & 31) * 2;
    if (c0 + c < cols && r0 + rp < rows) {
      __nv_bfloat162 v;
      v.x = tile[rp][c];
      v.y = tile[rp + 1][c];
      *reinterpret_cast<__nv_bfloat162*>(dst + (c0 + c) * out_pitch + r0 + rp) = v;
    }
  }
}

}  // namespace tod

using namespace tod;

extern "C" int tod_transpose_bf16(const void* d_in, void* d_out, int32_t batch, int32_t rows, int32_t cols, int64_t in_pitch,
                                  int64_t out_pitch, void* stream) {
  TOD_CHECK_ARG(d_in != nullptr && d_out != nullptr, "transpose: null pointer");
  TOD_CHECK_ARG(batch > 0 && batch <= 65535 && rows > 0 && cols > 0 && rows % 2 == 0 && cols % 2 == 0 && in_pitch >= cols &&
                    out_pitch >= rows && in_pitch % 2 == 0 && out_pitch % 2 == 0,
                "transpose: batch %d rows %d cols %d (even sizes and pitches)", batch, rows, cols);
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_in) & 3) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 3) == 0, "transpose: alignment");
  transpose_bf16_kernel<<<dim3((rows + 63) / 64, (cols + 63) / 64, batch), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(d_in), reinterpret_cast<__nv_bfloat16*>(d_out), rows, cols, in_pitch, out_pitch);
  TOD_CHECK_LAUNCH("transpose_bf16_kernel launch");
  return TOD_OK;
}

extern "C" int tod_softmax_rows_f32_bf16(const float* d_in, void* d_out, int32_t rows, int32_t cols, int64_t in_pitch,
                                         int64_t out_pitch, void* stream) {
  TOD_CHECK_ARG(d_in != nullptr && d_out != nullptr, "softmax: null pointer");
  TOD_CHECK_ARG(rows > 0 && cols > 0 && cols % 4 == 0 && in_pitch >= cols && out_pitch >= cols && in_pitch % 4 == 0 && out_pitch % 4 == 0,
                "softmax: rows %d cols %d (cols and pitches must be multiples of 4)", rows, cols);
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(d_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 7) == 0, "softmax: alignment");
  softmax_rows_kernel<<<rows, kSoftmaxThreads, 0, static_cast<cudaStream_t>(stream)>>>(d_in, reinterpret_cast<__nv_bfloat16*>(d_out), cols,
                                                                                      in_pitch, out_pitch);
  TOD_CHECK_LAUNCH("softmax_rows_kernel launch");
  return TOD_OK;
}
