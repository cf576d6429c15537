// Fused self-attention on tcgen05 / TMEM (sm_100a): the reference's SelfAttention block (model/blocks.py:236-254,
// instance model/backbone.py:33) without ever materialising the N x N score matrix -- SURVEY.md section 8 row f1.
//
//   out[i, :] = sum_j softmax_j(q_i . k_j) * v[j, :]  + bias + x[i, :]          (gamma is folded into v / bias by the host,
//                                                                                 log2(e) into q: the scores are base-2 logits)
//
// One CTA owns 128 queries of one image and walks the keys in tiles of 128, twice:
//   pass 1  S = Q K_j^T (one tcgen05.mma, K = d <= 64) -> the softmax threads read S from TMEM and keep the row maximum;
//   pass 2  S again -> p = exp(s - m) (exact maximum: no rescaling of the accumulator is ever needed), row sum in
//           registers, p as bf16 into shared memory in the canonical K-major A layout (two SWIZZLE_128B panels of 64 keys)
//           -> O += P V_j^T (8 tcgen05.mma, accumulator O in TMEM) ; the epilogue divides by the row sum.
// Recomputing S is cheap (K = 16 against K = 128 for P V); what it buys is one exponential per score instead of two and no
// TMEM read-modify-write of O.  Shared memory ~108 KB and 256 TMEM columns at C = 128: two CTAs per SM, so one CTA's
// softmax (MUFU-bound: 128 exponentials per row per tile) overlaps the other's MMAs without intra-CTA pipelining.
// Warp roles (320 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..9 softmax + epilogue -- two warps
// per TMEM lane quarter, each taking one half of a tile's keys (= one of the two P panels); with four softmax warps per
// CTA ncu showed MUFU at 40 % and issue slots at 35 %: latency-bound (profiles/r1_attention_ncu_full_summary.txt).
#include <mutex>

#include "tma_host.cuh"

namespace tod {

constexpr int kAttnThreads = 320;   // TMA warp, MMA warp, 8 softmax warps (two per TMEM lane quarter: one per half of the keys)
constexpr int kAttnTile = 128;   // queries per CTA = keys per tile

struct __align__(64) AttnParams {
  CUtensorMap tm_q, tm_k, tm_v;
  const __nv_bfloat16* x;   // residual, [B][N][x_pitch]
  __nv_bfloat16* out;       // [B][N][out_pitch] (may alias x)
  const float* bias;        // [C] or null
  int n, c, d16, x_pitch, out_pitch, tiles;
  uint32_t qk_bytes, v_panel_bytes;      // one Q / K tile; one 64-key panel of V^T
  uint32_t off_k, off_v, off_p;
  uint32_t hi_qk, hi_p, hi_v, idesc_s, idesc_s2, idesc_o, tmem_cols;
  int tiles2;   // pass 1 walks the keys in tiles of 256 (the O columns of TMEM are free until pass 2) when c >= 128
  int ksteps;
};

__device__ __forceinline__ void attn_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, "
      "%18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float attn_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads) attention_tcgen05(const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full, s_free, p_ready, p_free, o_done;
  __shared__ uint32_t tmem_base_smem;
  __shared__ float xch[2][kAttnTile];          // row maximum / row sum exchange between the two halves of a row
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int q0 = blockIdx.x * kAttnTile, img = blockIdx.y;
  const int T = p.tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_q);
    tma_prefetch_desc(&p.tm_k);
    tma_prefetch_desc(&p.tm_v);
    mbar_init(&q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(&s_full, 1);
    mbar_init(&s_free, 8);     // one arrival per softmax warp
    mbar_init(&p_ready, 8);
    mbar_init(&p_free, 1);
    mbar_init(&o_done, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(&tmem_base_smem, p.tmem_cols);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_s = tmem_base_smem, tmem_o = tmem_base_smem + kAttnTile;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      mbar_arrive_expect_tx(&q_full, p.qk_bytes);
      tma_load_3d(&p.tm_q, &q_full, smem_base, 0, q0, img);
      uint32_t use[2] = {0, 0}, vc = 0;      // completed uses of each K slot (barrier phases)
      if (p.tiles2 > 0) {
        // pass 1 on 256-key tiles: both K slots as one buffer behind slot 0's barriers
        for (int j = 0; j < p.tiles2; ++j, ++use[0]) {
          mbar_wait(&k_empty[0], (use[0] & 1) ^ 1u);
          mbar_arrive_expect_tx(&k_full[0], 2 * p.qk_bytes);
          tma_load_3d(&p.tm_k, &k_full[0], smem_base + p.off_k, 0, j * 2 * kAttnTile, img);
          tma_load_3d(&p.tm_k, &k_full[0], smem_base + p.off_k + p.qk_bytes, 0, j * 2 * kAttnTile + kAttnTile, img);
        }
        mbar_wait(&k_empty[0], (use[0] & 1) ^ 1u);   // the last 256-key MMA has read slot 1's memory too
      }
      for (int pass = (p.tiles2 > 0 ? 1 : 0); pass < 2; ++pass)
        for (int j = 0; j < T; ++j) {
          const uint32_t ks = j & 1;
          mbar_wait(&k_empty[ks], (use[ks] & 1) ^ 1u);
          mbar_arrive_expect_tx(&k_full[ks], p.qk_bytes);
          tma_load_3d(&p.tm_k, &k_full[ks], smem_base + p.off_k + ks * p.qk_bytes, 0, j * kAttnTile, img);
          ++use[ks];
          if (pass == 1) {
            const uint32_t vs = vc & 1;
            mbar_wait(&v_empty[vs], ((vc >> 1) & 1) ^ 1u);
            mbar_arrive_expect_tx(&v_full[vs], 2 * p.v_panel_bytes);
            const uint32_t dst = smem_base + p.off_v + vs * 2 * p.v_panel_bytes;
            tma_load_3d(&p.tm_v, &v_full[vs], dst, j * kAttnTile, 0, img);
            tma_load_3d(&p.tm_v, &v_full[vs], dst + p.v_panel_bytes, j * kAttnTile + 64, 0, img);
            ++vc;
          }
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      mbar_wait(&q_full, 0);
      const uint32_t q_lo = umma_desc_lo(smem_base);
      uint32_t use[2] = {0, 0}, vc = 0, sc = 0;
      if (p.tiles2 > 0) {
        const uint32_t k_lo = umma_desc_lo(smem_base + p.off_k);
        for (int j = 0; j < p.tiles2; ++j, ++use[0], ++sc) {
          mbar_wait(&k_full[0], use[0] & 1);
          mbar_wait(&s_free, (sc & 1) ^ 1u);
          tcgen05_fence_after();
          for (int s = 0; s < p.ksteps; ++s)     // S (256 columns: the S and O regions) = Q . K^T for 256 keys
            umma_bf16_k1(tmem_s, q_lo + 2 * s, p.hi_qk, k_lo + 2 * s, p.hi_qk, p.idesc_s2, s != 0 ? 1u : 0u);
          umma_commit(&k_empty[0]);
          umma_commit(&s_full);
        }
      }
      auto issue_s = [&](int j) {              // S = Q . K_j^T once the softmax threads have drained the previous S
        const uint32_t ks = j & 1;
        mbar_wait(&k_full[ks], use[ks] & 1);
        ++use[ks];
        mbar_wait(&s_free, (sc & 1) ^ 1u);
        ++sc;
        tcgen05_fence_after();
        const uint32_t k_lo = umma_desc_lo(smem_base + p.off_k + ks * p.qk_bytes);
        for (int s = 0; s < p.ksteps; ++s)
          umma_bf16_k1(tmem_s, q_lo + 2 * s, p.hi_qk, k_lo + 2 * s, p.hi_qk, p.idesc_s, s != 0 ? 1u : 0u);
        umma_commit(&k_empty[ks]);
        umma_commit(&s_full);
      };
      if (p.tiles2 == 0)
        for (int j = 0; j < T; ++j) issue_s(j);                       // pass 1 on 128-key tiles (c < 128)
      // pass 2: the softmax threads release S as soon as they have READ it, so S_{j+1} is issued before P_j V_j^T and its
      // latency runs under the exponentials of tile j
      issue_s(0);
      for (int j = 0; j < T; ++j) {
        if (j + 1 < T) issue_s(j + 1);
        const uint32_t vs = vc & 1;
        mbar_wait(&p_ready, j & 1);
        mbar_wait(&v_full[vs], (vc >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll 1
        for (int pn = 0; pn < 2; ++pn) {
          const uint32_t a_lo = umma_desc_lo(smem_base + p.off_p + pn * 16384u);
          const uint32_t b_lo = umma_desc_lo(smem_base + p.off_v + (vs * 2 + pn) * p.v_panel_bytes);
          umma_bf16_k4(tmem_o, a_lo, p.hi_p, b_lo, p.hi_v, p.idesc_o, (j | pn) != 0 ? 1u : 0u);
        }
        umma_commit(&v_empty[vs]);
        umma_commit(&p_free);
        if (j == T - 1) umma_commit(&o_done);
        ++vc;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax + epilogue: thread <-> query row
    const int qd = warp & 3;                    // TMEM lane quarter this warp may touch
    const int half = (warp - 2) >> 2;           // which half of a tile's keys (and of the output channels) this thread takes
    const int r = qd * 32 + lane;
    const uint32_t lane_sel = static_cast<uint32_t>(qd * 32) << 16;
    const uint32_t p_base = smem_base + p.off_p;
    float m = -INFINITY;
    uint32_t sc = 0;
    // pass 1: row maximum (256 keys per tile when the O columns can hold the second half of S)
    const int t1 = p.tiles2 > 0 ? p.tiles2 : T, w1 = p.tiles2 > 0 ? 2 * kAttnTile : kAttnTile;
    for (int j = 0; j < t1; ++j, ++sc) {
      mbar_wait(&s_full, sc & 1);
      tcgen05_fence_after();
      const int valid = min(w1, p.n - j * w1);
#pragma unroll 1
      for (int ch = half * (w1 / 64); ch < (half + 1) * (w1 / 64); ++ch) {
        uint32_t v[32];
        attn_ld32(tmem_s + lane_sel + ch * 32, v);
        tmem_ld_wait();
        if (valid == w1) {                   // full tile (all but the last one): no per-element predicates
#pragma unroll
          for (int i = 0; i < 32; i += 2) m = fmaxf(m, fmaxf(__uint_as_float(v[i]), __uint_as_float(v[i + 1])));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ch * 32 + i < valid) m = fmaxf(m, __uint_as_float(v[i]));
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free);
    }
    xch[half][r] = m;                           // the row maximum over both halves
    named_bar_sync(1, 256);
    m = fmaxf(xch[0][r], xch[1][r]);
    named_bar_sync(1, 256);                     // (xch is reused for the row sums)
    // pass 2: p = exp(s - m), row sum, bf16 panels for the P V^T MMAs
    float l = 0.0f;
    for (int j = 0; j < T; ++j, ++sc) {
      mbar_wait(&s_full, sc & 1);
      tcgen05_fence_after();
      const int valid = min(kAttnTile, p.n - j * kAttnTile);
      mbar_wait(&p_free, (j & 1) ^ 1u);          // the MMAs of the previous tile have finished reading the panels
      // one chunk's TMEM load is in flight while the previous chunk's exponentials run
      auto process = [&](int ch, const uint32_t (&v)[32]) {
        uint32_t o[16];
        // the host folds log2(e) into the query projection, so the scores are base-2 logits: one MUFU.EX2 per score
        if (valid == kAttnTile) {
          float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float e0 = attn_ex2(__uint_as_float(v[i]) - m), e1 = attn_ex2(__uint_as_float(v[i + 1]) - m);
            l0 += e0;
            l1 += e1;
            o[i >> 1] = pack_bf16x2(e0, e1);
          }
          l += l0 + l1;
        } else {
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float e0 = ch * 32 + i < valid ? attn_ex2(__uint_as_float(v[i]) - m) : 0.0f;
            const float e1 = ch * 32 + i + 1 < valid ? attn_ex2(__uint_as_float(v[i + 1]) - m) : 0.0f;
            l += e0 + e1;
            o[i >> 1] = pack_bf16x2(e0, e1);
          }
        }
        // keys [ch*32, ch*32+32) -> panel ch/2, 16-byte chunks (ch%2)*4 .. +3 of this row, SWIZZLE_128B
        const uint32_t row = p_base + (ch >> 1) * 16384u + static_cast<uint32_t>(r) * 128u;
#pragma unroll
        for (int c4 = 0; c4 < 4; ++c4) {
          const uint32_t chunk = static_cast<uint32_t>((ch & 1) * 4 + c4) ^ (static_cast<uint32_t>(r) & 7u);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row + chunk * 16u), "r"(o[4 * c4]), "r"(o[4 * c4 + 1]),
                       "r"(o[4 * c4 + 2]), "r"(o[4 * c4 + 3])
                       : "memory");
        }
      };
      {
        uint32_t va[32], vb[32];                // this thread's 64 keys = panel `half`
        attn_ld32(tmem_s + lane_sel + half * 64, va);
        attn_ld32(tmem_s + lane_sel + half * 64 + 32, vb);
        tmem_ld_wait();
        tcgen05_fence_before();                 // S has been read completely: hand it back before the exponentials
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free);
        process(2 * half, va);
        process(2 * half + 1, vb);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready);
    }
    // epilogue: out = O / l + bias + x
    mbar_wait(&o_done, 0);
    tcgen05_fence_after();
    xch[half][r] = l;
    named_bar_sync(1, 256);
    l = xch[0][r] + xch[1][r];
    const int q = q0 + r;
    const float inv = 1.0f / l;
    const size_t rowi = static_cast<size_t>(img) * p.n + q;
#pragma unroll 1
    for (int c0 = half * 32; c0 < p.c; c0 += 64) {          // 32-channel chunks alternate between the two halves
      uint32_t v[32];
      attn_ld32(tmem_o + lane_sel + c0, v);
      tmem_ld_wait();
      if (q < p.n) {
        const uint4* xr = reinterpret_cast<const uint4*>(p.x + rowi * p.x_pitch + c0);
        uint4* orow = reinterpret_cast<uint4*>(p.out + rowi * p.out_pitch + c0);
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint4 xv = xr[g];
          const uint32_t xu[4] = {xv.x, xv.y, xv.z, xv.w};
          uint32_t ou[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int cc = c0 + g * 8 + e * 2;
            const float b0 = p.bias ? __ldg(p.bias + cc) : 0.0f, b1 = p.bias ? __ldg(p.bias + cc + 1) : 0.0f;
            const float a0 = __uint_as_float(v[g * 8 + e * 2]) * inv + b0 + __uint_as_float(xu[e] << 16);
            const float a1 = __uint_as_float(v[g * 8 + e * 2 + 1]) * inv + b1 + __uint_as_float(xu[e] & 0xffff0000u);
            ou[e] = pack_bf16x2(a0, a1);
          }
          orow[g] = make_uint4(ou[0], ou[1], ou[2], ou[3]);
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base_smem, p.tmem_cols);
  }
}

static uint32_t attn_desc_hi(uint32_t sbo_bytes, int row_bytes) {
  const uint32_t layout_type = row_bytes == 128 ? 2u : (row_bytes == 64 ? 4u : 6u);  // SWIZZLE_128B / 64B / 32B
  return (sbo_bytes >> 4) | (1u << 14) | (layout_type << 29);
}

}  // namespace tod

using namespace tod;

extern "C" int tod_attention_fused(const tod_attention_desc* d, void* stream) {
  TOD_CHECK_ARG(d != nullptr && d->d_q && d->d_k && d->d_vt && d->d_x && d->d_out, "attention: null pointer");
  TOD_CHECK_ARG(d->batch > 0 && d->batch <= 65535 && d->n > 0 && d->n % 16 == 0, "attention: batch %d tokens %d (multiple of 16)", d->batch, d->n);
  TOD_CHECK_ARG(d->d16 == 16 || d->d16 == 32 || d->d16 == 64, "attention: q/k width %d (16, 32 or 64)", d->d16);
  TOD_CHECK_ARG(d->c >= 32 && d->c % 32 == 0 && d->c <= 256, "attention: channels %d (multiple of 32, <= 256)", d->c);
  TOD_CHECK_ARG(d->x_pitch >= d->c && d->out_pitch >= d->c && d->x_pitch % 8 == 0 && d->out_pitch % 8 == 0, "attention: pitches");
  static PerDeviceOnce attr_once;   // the attribute is per device
  int rc;
  if (attr_once.needed()) {
    if ((rc = check_cuda(cudaFuncSetAttribute(attention_tcgen05, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                         "cudaFuncSetAttribute(attention_tcgen05)")) != TOD_OK)
      return rc;
    attr_once.done();
  }
  AttnParams p;
  memset(&p, 0, sizeof(p));
  const int rb = d->d16 * 2;
  const CUtensorMapSwizzle swz_qk = rb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (rb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(d->d16), static_cast<uint64_t>(d->n), static_cast<uint64_t>(d->batch)};
    const uint64_t str[2] = {static_cast<uint64_t>(rb), static_cast<uint64_t>(rb) * d->n};
    const uint32_t box[3] = {static_cast<uint32_t>(d->d16), kAttnTile, 1};
    if ((rc = encode_map(&p.tm_q, d->d_q, 3, dims, str, box, swz_qk)) != TOD_OK) return rc;
    if ((rc = encode_map(&p.tm_k, d->d_k, 3, dims, str, box, swz_qk)) != TOD_OK) return rc;
  }
  {
    const uint64_t dims[3] = {static_cast<uint64_t>(d->n), static_cast<uint64_t>(d->c), static_cast<uint64_t>(d->batch)};
    const uint64_t str[2] = {static_cast<uint64_t>(d->n) * 2, static_cast<uint64_t>(d->n) * 2 * d->c};
    const uint32_t box[3] = {64, static_cast<uint32_t>(d->c), 1};
    if ((rc = encode_map(&p.tm_v, d->d_vt, 3, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) != TOD_OK) return rc;
  }
  p.x = reinterpret_cast<const __nv_bfloat16*>(d->d_x);
  p.out = reinterpret_cast<__nv_bfloat16*>(d->d_out);
  p.bias = d->d_bias;
  p.n = d->n;
  p.c = d->c;
  p.d16 = d->d16;
  p.x_pitch = d->x_pitch;
  p.out_pitch = d->out_pitch;
  p.tiles = (d->n + kAttnTile - 1) / kAttnTile;
  p.ksteps = d->d16 / 16;
  p.qk_bytes = static_cast<uint32_t>(kAttnTile) * rb;
  p.v_panel_bytes = static_cast<uint32_t>(d->c) * 128u;
  const uint32_t qk_slot = (p.qk_bytes + 1023u) & ~1023u;
  TOD_CHECK_ARG(qk_slot == p.qk_bytes, "attention: tile size");   // 128 rows x 32/64/128 B is always a multiple of 1024
  p.off_k = qk_slot;
  p.off_v = p.off_k + 2 * qk_slot;
  p.off_p = p.off_v + 4 * p.v_panel_bytes;
  const size_t smem = static_cast<size_t>(p.off_p) + 2 * 16384 + 1024;
  TOD_CHECK_ARG(smem <= 200 * 1024, "attention: shared-memory plan (%zu bytes)", smem);
  p.hi_qk = attn_desc_hi(8 * rb, rb);
  p.hi_p = attn_desc_hi(1024, 128);
  p.hi_v = attn_desc_hi(1024, 128);
  p.idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(kAttnTile >> 3) << 17) | ((128u >> 4) << 24);
  p.idesc_s2 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(256 >> 3) << 17) | ((128u >> 4) << 24);
  p.tiles2 = d->c >= kAttnTile ? (d->n + 2 * kAttnTile - 1) / (2 * kAttnTile) : 0;   // needs 256 TMEM columns: S + (idle) O
  p.idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(d->c >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(kAttnTile + d->c)) cols <<= 1;
  p.tmem_cols = cols;
  attention_tcgen05<<<dim3(p.tiles, d->batch), kAttnThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  TOD_CHECK_LAUNCH("attention_tcgen05 launch");
  return TOD_OK;
}
