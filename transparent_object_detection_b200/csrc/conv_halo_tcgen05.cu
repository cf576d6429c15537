// Halo-patch implicit-GEMM convolution on tcgen05 / TMEM (sm_100a): the main conv kernel of libtod.so.
//
// Same contract as conv_tcgen05.cu (reference Conv.forward model/blocks.py:52-54 with fuse_conv :160-187 folded,
// Bottleneck add :80-82, head 1x1 convs model/head.py:31,42, concat-by-offset), different data movement:
//
//   * A operand (activations).  tcgen05 shared-memory descriptors address rows as
//         row(i) = start + (i / 8) * SBO + (i % 8) * row_bytes
//     and the 128B/64B/32B swizzle is a function of the ABSOLUTE shared-memory address (measured:
//     profiles/r1_probe_umma_rowshift.txt), so `start` may be shifted by whole rows and SBO may be any row multiple.
//     A 3x3 conv therefore loads ONE (16+2) x (8+2) pixel halo patch per 64-channel chunk and issues its nine taps as
//     nine descriptor views of that patch (tap (kh,kw): start += (kh*10 + kw) rows, SBO = 10 rows): L2->SM traffic for
//     activations drops from 9x to 1.4x.  Stride 2 uses the four parity planes of the input the same way; 1x1 is the
//     degenerate case (one tap, 128 consecutive pixels).
//   * B operand (weights).  One weight tile [block_n x block_k] per (tap, chunk) feeds `m` M sub-tiles (m accumulators
//     in TMEM); when the whole [block_n x K] weight panel fits in shared memory it is loaded once per CTA and stays
//     resident across the persistent tile loop.
//   * Epilogue.  TMEM -> registers -> bias / upsample-add / SiLU / residual -> swizzled staging panel in shared memory
//     -> TMA store (clipped at the tensor edges, so partial tiles and channel windows need no guards).
//
// Warp roles (320 threads, 1 CTA / SM, persistent): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer,
// warps 2..9 two epilogue groups that alternate over the two TMEM accumulator stages.
#include <cstring>
#include <mutex>
#include <type_traits>

#include "decode_math.cuh"
#include "tma_host.cuh"

namespace tod {

constexpr int kHaloThreads = 320;
constexpr int kHaloThreadsTail = 352;   // + one warp that issues the second GEMM of the fused 1x1 tail (EXTRA 7)
constexpr int kMaxA = 4;      // A-ring slots
constexpr int kMaxB = 40;     // B-ring slots (>= taps * chunks when the weights are resident)
constexpr int kPatchH = 16, kPatchW = 8;
constexpr uint32_t kStageBytes = 16384;              // one epilogue staging panel: 128 rows x 128 B
constexpr uint32_t kHaloSmemLimit = 225u * 1024u;    // dynamic; static barriers + bias take < 2 KB of the 227 KB

// n / d by multiplication: exact while n * d < 2^40 (checked on the host).  A runtime integer division is a ~40
// instruction dependent chain, which a single-thread role (producer, MMA issuer) cannot afford per tile.
struct FastDiv {
  uint32_t d;
  uint64_t mul;   // ceil(2^40 / d)
  __device__ __forceinline__ int div(int n) const { return static_cast<int>((static_cast<uint64_t>(n) * mul) >> 40); }
};

struct TileSched {
  int full_rounds, G, g, m, tail_begin, tail_cnt;
  __device__ __forceinline__ TileSched(int num_subtiles, int m_, int G_, int g_) : G(G_), g(g_), m(m_) {
    full_rounds = num_subtiles / (G * m);
    const int base = full_rounds * G * m, rem = num_subtiles - base;
    tail_begin = base + static_cast<int>(static_cast<long long>(g) * rem / G);
    tail_cnt = base + static_cast<int>(static_cast<long long>(g + 1) * rem / G) - tail_begin;
  }
  __device__ __forceinline__ int iters() const { return full_rounds + (tail_cnt > 0 ? 1 : 0); }
  __device__ __forceinline__ void get(int it, int& s0, int& m_cur) const {
    if (it < full_rounds) {
      s0 = (it * G + g) * m;
      m_cur = m;
    } else {
      s0 = tail_begin;
      m_cur = tail_cnt;
    }
  }
};

struct __align__(64) HaloParams {
  CUtensorMap tm_a[4];
  CUtensorMap tm_w;
  CUtensorMap tm_out;
  CUtensorMap tm_res;        // residual panels (EXTRA 5: TMA ring in shared memory), same boxes / swizzle as tm_out
  // geometry
  int patch_mode;            // 1: 16x8 pixel patches (3x3), 0: 128 consecutive pixels (1x1), 2: row-flat tiles (3x3 s1 on small maps)
  int hout, wout;
  int tiles_w, tiles_per_img;
  int tile_w_step, tile_h_step;   // first output column / row of tile (tx, ty) = tx * tile_w_step, ty * tile_h_step
  int patch_cols, rows_valid;     // accumulator row r <-> pixel (r / patch_cols, r % patch_cols) of the tile, r < rows_valid
  FastDiv fd_patch_cols;
  FastDiv fd_tiles_per_img, fd_tiles_w, fd_wout, fd_hw;
  long long mtot;            // batch * hout * wout
  int batch;
  int num_subtiles, m, num_super;
  int num_units;             // scheduled units: sub-tiles, or PAIRS of adjacent sub-tiles (2u, 2u + 1) in CTA-pair mode
  int oob_row;               // flat 1x1 mode: a first row >= mtot (what the odd CTA of a pair loads when it has no sub-tile)
  uint32_t b_half_rows;      // CTA-pair mode: weight rows per CTA (block_n / 2)
  int reverse;               // TOD_CONV_REVERSE: sub-tile s stands for sub-tile num_subtiles - 1 - s (last image first)
  int nowait;                // experiment: skip the programmatic-launch wait (TOD_PDL_NOWAIT=1; results are wrong)
  int dyn_w;                 // TOD_CONV_DYNAMIC_W: the weights are an earlier kernel's output (no prefetch before pdl_wait)
  int n_tiles, block_n, cout;
  int block_k, ksteps, chunks, num_taps;
  // A loads per (sub-tile, chunk)
  int n_aloads;
  int al_map[4], al_dw[4], al_dh[4];
  uint32_t al_off[4];
  uint32_t a_tx_bytes;       // bytes landed per sub-tile per chunk
  uint32_t b_tx_bytes;
  uint32_t sub_bytes, a_slot_bytes, b_slot_bytes;
  int sa, sb, stationary;
  uint32_t tap_a_off[9];
  uint32_t tap_hi_a[9];      // upper descriptor word per tap (SBO differs between parity planes)
  uint32_t hi_b;
  uint32_t idesc, tmem_cols;
  uint32_t off_b, off_stage, off_res;
  uint32_t res_tx_bytes;     // bytes of one residual panel load
  // epilogue
  const float* bias;
  const __nv_bfloat16* residual;
  const float* upadd;
  int res_pitch;
  int act, out_f32;
  int pc, pb, smask;         // panel columns, panel row bytes, swizzle XOR mask (7 / 3 / 1)
  // fused head decode (EXTRA 3: box tower output -> cand_box, EXTRA 4: class tower output -> cand_conf / cand_cls)
  int dec_nc, dec_level_off, dec_anchors;
  float dec_stride, dec_in_w, dec_in_h;
  float* cand_box;
  float* cand_conf;
  int* cand_cls;
  // fused 1x1 tail (EXTRA 7): y2 = act(W2 . act(conv) + b2); W2 [64 x 64] bf16 resident in shared memory
  CUtensorMap tm_w2;
  const float* bias2;
  uint32_t off_w2, idesc2, hi_w2, hi_stage;
  unsigned long long* prof;  // optional per-CTA wait counters (tools/conv_profile.py), null in production
  TimelineTag tl;            // optional per-CTA start / end records (tools/timeline.py), null in production
};

// Profiling helper: accumulates the cycles one role's single issuing thread spends inside a wait.
struct WaitClock {
  unsigned long long acc[4];
  bool on;
  __device__ __forceinline__ explicit WaitClock(bool enabled) : on(enabled) { acc[0] = acc[1] = acc[2] = acc[3] = 0; }
  __device__ __forceinline__ long long begin() const { return on ? clock64() : 0; }
  __device__ __forceinline__ void end(int i, long long t0) { if (on) acc[i] += clock64() - t0; }
};

__device__ __forceinline__ uint64_t halo_desc(uint32_t smem_addr, uint32_t hi) {
  return (static_cast<uint64_t>(hi) << 32) | (1ull << 16) | ((smem_addr >> 4) & 0x3FFFu);
}

__device__ __forceinline__ uint32_t swz(uint32_t off, uint32_t smask) { return off ^ (((off >> 7) & smask) << 4); }

// ---- epilogue arithmetic.  The epilogue warps are issue/latency bound on the HBM-bound layers (ncu: ~220 SASS
// instructions per 32 columns, IPC 0.2 per SM sub-partition), so the kernel is specialised on (activation, output type,
// extra operand) and the per-element work is FFMA2 (packed f32x2) + one MUFU.TANH + half an F2FP.
__device__ __forceinline__ uint64_t pk2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pk2u(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void upk2(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
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

// Extra epilogue operand of one 32-column chunk of this thread's row, held in registers so that its global loads are
// issued one chunk ahead of their use.  EXTRA 1: bf16 residual (added after the activation, 64 B = 4 x 16 B);
// EXTRA 2: f32 upsample-add (added before it, 128 B = 8 x 16 B).
template <int EXTRA>
struct ExtraRegs {
  uint4 q[EXTRA == 2 ? 8 : (EXTRA == 1 ? 4 : 1)];
  __device__ __forceinline__ void load(const void* row_ptr, int col, int n) {   // n = 16 or 32 columns
    if (EXTRA == 1) {
      const uint4* g = reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(row_ptr) + col);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (i * 8 < n) q[i] = __ldg(g + i);
    } else if (EXTRA == 2) {
      const uint4* g = reinterpret_cast<const uint4*>(reinterpret_cast<const float*>(row_ptr) + col);
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i * 4 < n) q[i] = __ldg(g + i);
    }
  }
};

// N (16 or 32) accumulator columns of this thread's row -> bias / upsample-add / SiLU / residual -> N output values
// packed for the staging panel: bf16 pairs (N/2 words) or f32 (N words) in o[].  With SiLU the bias in shared memory is
// pre-halved:  h = acc/2 + b/2 is one FMA and silu(x) = h + h * tanh(h).
template <int N, bool SILU, bool OUT_F32, int EXTRA>
__device__ __forceinline__ void epilogue_math(const uint32_t* v, const float* bias_col, const ExtraRegs<EXTRA>& ex, bool ex_valid,
                                              uint32_t* o) {
  const uint64_t sc2 = SILU ? pk2(0.5f, 0.5f) : pk2(1.0f, 1.0f);
#pragma unroll
  for (int j = 0; j < N; j += 4) {
    const float4 b = *reinterpret_cast<const float4*>(bias_col + j);
    uint64_t f0 = ffma2(pk2u(v[j], v[j + 1]), sc2, pk2(b.x, b.y));
    uint64_t f1 = ffma2(pk2u(v[j + 2], v[j + 3]), sc2, pk2(b.z, b.w));
    if (EXTRA == 2) {
      const uint4 u = ex.q[j >> 2];
      if (ex_valid) {
        f0 = ffma2(pk2u(u.x, u.y), sc2, f0);
        f1 = ffma2(pk2u(u.z, u.w), sc2, f1);
      }
    }
    if (SILU) {
      float h0, h1, h2, h3, t0, t1, t2, t3;
      upk2(f0, h0, h1);
      upk2(f1, h2, h3);
      asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(h0));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(h1));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t2) : "f"(h2));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t3) : "f"(h3));
      f0 = ffma2(f0, pk2(t0, t1), f0);
      f1 = ffma2(f1, pk2(t2, t3), f1);
    }
    if (EXTRA == 1) {
      const uint4 rq = ex.q[j >> 3];
      const uint32_t r0 = (j & 4) ? rq.z : rq.x, r1 = (j & 4) ? rq.w : rq.y;
      if (ex_valid) {   // bf16 -> f32 is a 16-bit shift / mask
        f0 = fadd2(f0, pk2u(r0 << 16, r0 & 0xffff0000u));
        f1 = fadd2(f1, pk2u(r1 << 16, r1 & 0xffff0000u));
      }
    }
    float a0, a1, a2, a3;
    upk2(f0, a0, a1);
    upk2(f1, a2, a3);
    if (OUT_F32) {
      o[j] = __float_as_uint(a0);
      o[j + 1] = __float_as_uint(a1);
      o[j + 2] = __float_as_uint(a2);
      o[j + 3] = __float_as_uint(a3);
    } else {
      o[j >> 1] = pack_bf16x2(a0, a1);
      o[(j >> 1) + 1] = pack_bf16x2(a2, a3);
    }
  }
}

// ---- compact wait primitives.  Every role's steady-state code must stay small: one warp walking tens of KB of
// unrolled code stalls on instruction fetch (measured: stall_no_inst dominated the MMA warp), so loops are kept rolled
// and the cold timeout path is out of line.
__device__ __noinline__ void halo_wait_timeout(uint32_t bar, uint32_t parity) {
  printf("tod: conv_halo mbarrier wait timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
  __trap();
}
__device__ __forceinline__ bool try_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a system-dependent time before it reports failure)
__device__ __forceinline__ bool test_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(bar), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_addr(uint32_t bar, uint32_t parity) {
  if (try_wait_addr(bar, parity)) return;
  const long long t0 = clock64();
#pragma unroll 1
  while (!try_wait_addr(bar, parity))
    if (clock64() - t0 > 4000000000ll) halo_wait_timeout(bar, parity);
}
// Up to four barriers at once (mask bit i enables bars[i]): the try_waits overlap instead of serialising ~100 cycles
// each on the MMA issuer's critical path.
__device__ __forceinline__ void wait_set(const uint32_t (&bars)[4], const uint32_t (&pars)[4], uint32_t mask) {
  uint32_t ready = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i)   // independent of one another: the four try_waits are in flight together
    if ((mask >> i) & 1u) ready |= static_cast<uint32_t>(try_wait_addr(bars[i], pars[i])) << i;
  uint32_t pending = mask & ~ready;
  if (pending == 0) return;
  const long long t0 = clock64();
#pragma unroll 1
  while (pending != 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if ((pending >> i) & 1u)
        if (try_wait_addr(bars[i], pars[i])) pending &= ~(1u << i);
    if (clock64() - t0 > 4000000000ll) halo_wait_timeout(bars[0], pending);
  }
}

// PAIR: the two CTAs of a (2,1,1) cluster work on adjacent sub-tiles as ONE M = 256 MMA (tcgen05 cta_group::2): each CTA
// loads its own activation patches and only HALF of every weight tile (rows [block_n/2 r, +block_n/2)), so the weight
// bytes an SM pulls out of L2 per output pixel halve -- the layers that stream their weights (3x3 with >= 128 channels,
// stride 2, 20x20 maps) run at the chip's L2 -> SM limit otherwise (ncu: 38-48 B/clk/SM of xbar2l1tex reads).  The
// leader (rank 0) issues every MMA; both producers signal the leader's "full" barriers; the leader's commits arrive on
// the "empty" / "accumulator full" barriers of both CTAs; both CTAs' epilogue warps arrive on the leader's "accumulator
// empty" barrier.  Primitives checked by tools/probe_pair.cu.
template <bool SILU, bool OUT_F32, int EXTRA, bool PAIR = false>
__global__ void __launch_bounds__((EXTRA == 7 || EXTRA == 8) ? kHaloThreadsTail : kHaloThreads, 1)
conv_halo_tcgen05(const __grid_constant__ HaloParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t a_full[kMaxA], a_empty[kMaxA];
  __shared__ __align__(8) uint64_t b_full[kMaxB], b_empty[kMaxB];
  __shared__ __align__(8) uint64_t tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ __align__(8) uint64_t res_full_bar[4];   // residual ring: 2 slots per epilogue group
  __shared__ __align__(16) float bias_s[256];
  // Fused 1x1 tail: EXTRA 7 stores act(W2 . y + b2); EXTRA 8 (last two convs of a box tower, model/head.py:36-42) decodes
  // W2 . y + b2 straight into the NMS box candidates like EXTRA 3.  (Tail-only shared memory: static + dynamic of the
  // other variants already sits 252 bytes under the 227 KB limit.)
  constexpr bool TAIL = EXTRA == 7 || EXTRA == 8;
  static_assert(!PAIR || EXTRA == 0 || EXTRA == 1 || EXTRA == 5 || EXTRA == 6, "CTA-pair mode: plain / residual / upsample-add convs only");
  __shared__ __align__(16) float bias2_s[TAIL ? 64 : 4];
  __shared__ __align__(8) uint64_t tail_bars[TAIL ? 5 : 1];   // tail weights landed / panel ready x2 / tail done x2
  uint64_t& w2_full_bar = tail_bars[0];
  uint64_t* const p_full_bar = &tail_bars[TAIL ? 1 : 0];
  uint64_t* const d2_full_bar = &tail_bars[TAIL ? 3 : 0];
  __shared__ uint32_t tmem_base_smem;
  __shared__ unsigned long long tl_marks[7];   // timeline: programmatic wait returned / first accumulator complete / MMA role done / last accumulator complete

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const unsigned long long tl_t0 = (p.tl.buf != nullptr && threadIdx.x == 0) ? global_timer_ns() : 0ull;
  if (p.tl.buf != nullptr && threadIdx.x == 0)
    for (int i = 0; i < 7; ++i) tl_marks[i] = 0;

  // persistent schedule: this CTA owns N tile `nt`.  Full rounds are interleaved (CTA g takes super-tile r * G + g, so
  // the grid streams through adjacent memory together); what is left after the last full round is split evenly at
  // SUB-tile granularity, so the tail imbalance is one sub-tile instead of one super-tile of m.
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;            // CTA-pair mode: 0 = leader
  const int cta = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int ncta = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int nt = cta % p.n_tiles;
  const TileSched sched(p.num_units, p.m, ncta / p.n_tiles, cta / p.n_tiles);
  const int n0 = nt * p.block_n;
  // TMA coordinates of the sub-tile this CTA takes of scheduled unit u; false (and coordinates wholly outside the tensor:
  // loads deliver zeros, stores are clipped away) when the odd CTA of a pair has no sub-tile left
  auto tile_coords = [&](int u, int& c1, int& c2, int& c3) -> bool {
    int s = PAIR ? 2 * u + static_cast<int>(rank) : u;
    if (p.reverse) s = p.num_subtiles - 1 - s;
    const bool valid = !PAIR || static_cast<unsigned>(s) < static_cast<unsigned>(p.num_subtiles);
    if (p.patch_mode) {
      const int img = p.fd_tiles_per_img.div(valid ? s : 0);
      const int rem = s - img * p.tiles_per_img;
      const int ti = p.fd_tiles_w.div(valid ? rem : 0);
      c1 = valid ? (rem - ti * p.tiles_w) * p.tile_w_step : 0;
      c2 = valid ? ti * p.tile_h_step : 0;
      c3 = valid ? img : p.batch;
    } else {
      c1 = valid ? s * 128 : p.oob_row;
      c2 = 0;
      c3 = 0;
    }
    return valid;
  };
  const int acc_cols = p.m * p.block_n;   // TMEM columns of one accumulator stage
  const uint32_t a_full0 = smem_u32(&a_full[0]), a_empty0 = smem_u32(&a_empty[0]);
  const uint32_t b_full0 = smem_u32(&b_full[0]), b_empty0 = smem_u32(&b_empty[0]);   // barrier i lives at base + 8 * i

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tm_a[0]);
    tma_prefetch_desc(&p.tm_w);
    tma_prefetch_desc(&p.tm_out);
    for (int s = 0; s < p.sa; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < p.sb; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], PAIR ? 8 : 4);   // the leader's barrier collects both CTAs' epilogue warps
    }
    for (int a = 0; a < 4; ++a) mbar_init(&res_full_bar[a], 1);
    if (TAIL) {
      mbar_init(&w2_full_bar, 1);
      for (int a = 0; a < 2; ++a) {
        mbar_init(&p_full_bar[a], 1);
        mbar_init(&d2_full_bar[a], 1);
      }
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if (PAIR) {
      tmem_alloc2(&tmem_base_smem, p.tmem_cols);
      tmem_relinquish2();
    } else {
      tmem_alloc(&tmem_base_smem, p.tmem_cols);
      tmem_relinquish();
    }
  }
  if (warp >= 2) {
    for (int i = threadIdx.x - 64; i < p.block_n; i += kHaloThreads - 64)
      bias_s[i] = (p.bias != nullptr && n0 + i < p.cout) ? __ldg(p.bias + n0 + i) * (SILU ? 0.5f : 1.0f) : 0.0f;
    if (TAIL && threadIdx.x - 64 < 64)   // (pre-halved for the SiLU of EXTRA 7; the box logits of EXTRA 8 take it as it is)
      bias2_s[threadIdx.x - 64] = p.bias2 != nullptr ? __ldg(p.bias2 + threadIdx.x - 64) * (EXTRA == 7 && SILU ? 0.5f : 1.0f) : 0.0f;
  }
  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // the peer's barriers are initialised before anything can signal them
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (threadIdx.x == 0) pdl_launch_dependents();   // the next kernel's prologue may overlap this grid's tail

  // CTA-pair mode: this CTA's half of the weight tile / its own activation patches go to its own shared memory, the
  // bytes are counted on the LEADER's barrier, which the leader arms for both CTAs
  const int w_row0 = n0 + (PAIR ? static_cast<int>(rank * p.b_half_rows) : 0);
  auto load_w = [&](int i, int k) {   // weight tile of (tap, chunk) k-offset `k` into ring / resident slot i
    if (!PAIR || rank == 0) mbar_arrive_expect_tx(&b_full[i], PAIR ? 2 * p.b_tx_bytes : p.b_tx_bytes);
    if (PAIR) tma_load_2d_pair(&p.tm_w, pair_leader_addr(b_full0 + 8 * i), smem_base + p.off_b + i * p.b_slot_bytes, k, w_row0);
    else tma_load_2d(&p.tm_w, &b_full[i], smem_base + p.off_b + i * p.b_slot_bytes, k, w_row0);
  };
  auto load_resident_weights = [&]() {
#pragma unroll 1
    for (int c = 0; c < p.chunks; ++c)
#pragma unroll 1
      for (int t = 0; t < p.num_taps; ++t) load_w(c * p.num_taps + t, (t * p.chunks + c) * p.block_k);
  };
  // Streamed weights: the first pass through the ring is requested ahead of the programmatic-launch wait as well (the tile
  // sequence of a super-tile -- chunk-major, taps inside -- repeats for every super-tile, so ring slot q holds tile q of
  // the sequence; never more than this CTA consumes): these bytes then cross the L2 -> SM fabric while the previous grid
  // drains instead of in the burst every CTA starts with.
  const int b_per_tile = p.num_taps * p.chunks;
  const long long b_wanted = static_cast<long long>(sched.iters()) * b_per_tile;
  const int b_prefetched = (!p.stationary && !p.dyn_w) ? (b_wanted < p.sb ? static_cast<int>(b_wanted) : p.sb) : 0;
  // Constant weights do not depend on earlier kernels.  They are requested by the MMA warp's thread (idle until the first
  // operands land) so that the producer reaches its wait -- and the first activation loads behind it -- without first
  // issuing up to 40 weight loads (~0.13 us each).
  auto prefetch_weights = [&]() {
    if (p.dyn_w) return;
    if (p.stationary) {
      load_resident_weights();
      return;
    }
#pragma unroll 1
    for (int q = 0; q < b_prefetched; ++q) {
      const int qq = q % b_per_tile;
      const int c = qq / p.num_taps, t = qq - c * p.num_taps;
      load_w(q, (t * p.chunks + c) * p.block_k);
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one elected lane)
    if (elect_one()) {
      WaitClock wc(p.prof != nullptr);
      const long long role_t0 = wc.begin();
      int b_pre = b_prefetched;
      if (TAIL) {
        mbar_arrive_expect_tx(&w2_full_bar, 64u * 128u);
        tma_load_2d(&p.tm_w2, &w2_full_bar, smem_base + p.off_w2, 0, 0);
      }
      // Everything above (and the weights just requested) is independent of earlier kernels.  Activations, residual
      // and upsample-add operands and the output buffer are not: every access to them in this grid is ordered after
      // this wait through the mbarrier chain that starts at the first A load below.
      if (!p.nowait) pdl_wait();
      if (p.tl.buf != nullptr) tl_marks[0] = global_timer_ns();
      if (p.stationary && p.dyn_w) load_resident_weights();   // "weights" produced by the previous kernel (q . k^T)
      int ai = 0, bi = 0;
      uint32_t pha = 0, phb = 0;   // ring phase bits
      const int n_it = sched.iters();
#pragma unroll 1
      for (int it = 0; it < n_it; ++it) {
        int s0, m_cur;
        sched.get(it, s0, m_cur);
#pragma unroll 1
        for (int c = 0; c < p.chunks; ++c) {
          long long tw = wc.begin();
          wait_addr(a_empty0 + 8 * ai, pha ^ 1u);
          wc.end(1, tw);
          if (!PAIR || rank == 0) mbar_arrive_expect_tx(&a_full[ai], (PAIR ? 2 : 1) * m_cur * p.a_tx_bytes);
          const uint32_t slot = smem_base + ai * p.a_slot_bytes;
#pragma unroll 1
          for (int mt = 0; mt < m_cur; ++mt) {
            int c1, c2, c3;
            tile_coords(s0 + mt, c1, c2, c3);
#pragma unroll 1
            for (int a = 0; a < p.n_aloads; ++a) {
              if (PAIR)
                tma_load_4d_pair(&p.tm_a[p.al_map[a]], pair_leader_addr(a_full0 + 8 * ai), slot + mt * p.sub_bytes + p.al_off[a],
                                 c * p.block_k, c1 + p.al_dw[a], c2 + p.al_dh[a], c3);
              else
                tma_load_4d(&p.tm_a[p.al_map[a]], &a_full[ai], slot + mt * p.sub_bytes + p.al_off[a], c * p.block_k,
                            c1 + p.al_dw[a], c2 + p.al_dh[a], c3);
            }
          }
          if (!p.stationary) {
#pragma unroll 1
            for (int t = 0; t < p.num_taps; ++t) {
              if (b_pre > 0) {
                --b_pre;             // requested before the programmatic-launch wait
              } else {
                tw = wc.begin();
                wait_addr(b_empty0 + 8 * bi, phb ^ 1u);
                wc.end(2, tw);
                load_w(bi, (t * p.chunks + c) * p.block_k);
              }
              if (++bi == p.sb) {
                bi = 0;
                phb ^= 1u;
              }
            }
          }
          if (++ai == p.sa) {
            ai = 0;
            pha ^= 1u;
          }
        }
      }
      if (p.prof) {
        wc.end(0, role_t0);
        for (int i = 0; i < 3; ++i) p.prof[blockIdx.x * 16 + i] = wc.acc[i];
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer: ONE elected lane runs the whole role
    // (waits and issue).  Warp-level waits + elect + __syncwarp around every tap group cost ~430 cycles per group in
    // which the 8-deep MMA queue (<= 512 tensor cycles at N = 128) drained: the tensor pipe was busy 66 % on the
    // weight-streaming layers.
    const bool mma_thread = elect_one();
    if (mma_thread) prefetch_weights();   // (both CTAs of a pair: each its own halves)
    if ((!PAIR || rank == 0) && mma_thread) {
    WaitClock wc(p.prof != nullptr || p.tl.buf != nullptr);
    const long long role_t0 = wc.begin();
    // The whole role is instantiated per K-step count (4 / 2 / generic) and the sub-tile loop is unrolled (m <= 4): at
    // N <= 64 one MMA is <= 48 tensor cycles and the ~25 uniform-datapath instructions (constant reloads, three K-step
    // branch tests, loop control) that the rolled version spent per 2-MMA call made the issuing thread the limiter
    // (measured with tools/conv_profile.py: the role was busy issuing 92 % of a 32->32 3x3 layer at ~85 cycles per MMA
    // against the 40-cycle tensor floor of profiles/r1_probe_umma_rate.txt).
    auto commit = [&](uint32_t bar_addr) {   // CTA-pair mode: the barrier at this offset in BOTH CTAs
      if (PAIR) umma2_commit_both(bar_addr);
      else asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
    };
    auto run = [&](auto ksc) {
    constexpr int KS = decltype(ksc)::value;
    uint32_t lt = 0;
    int ai = 0, rbi = 0;
    uint32_t pha = 0, phb = 0;
    const uint32_t sub16 = p.sub_bytes >> 4;
    const uint32_t b_ring = smem_base + p.off_b;
    const uint32_t hi_b = p.hi_b, idesc = p.idesc, block_n = p.block_n;
    const int n_it = sched.iters();
#pragma unroll 1
    for (int it = 0; it < n_it; ++it, ++lt) {
      int s0, m_cur;
      sched.get(it, s0, m_cur);
      const uint32_t acc = lt & 1;
      long long tw = wc.begin();
      wait_addr(smem_u32(&tmem_empty_bar[acc]), ((lt >> 1) & 1) ^ 1u);
      wc.end(1, tw);
      tcgen05_fence_after();
      const uint32_t tmem_acc = tmem_base + acc * acc_cols;
      const bool wait_b = !p.stationary || lt == 0;   // resident weights are only awaited on the first pass
#pragma unroll 1
      for (int c = 0; c < p.chunks; ++c) {
        const uint32_t a_base = smem_base + ai * p.a_slot_bytes;
        // One synchronisation group = taps [T0, T0 + GSZ) of this chunk: an overlapped wait for the group's weight tiles
        // (and the A slot at the start of a chunk), then GSZ * m * ksteps MMAs back to back.  T0 / GSZ are compile-time
        // so that the per-tap descriptor offsets are constant-bank operands (statically indexed kernel parameters); read
        // through a runtime tap index they were dependent ~50-cycle loads in front of every tap.
        auto group = [&](auto t0c, auto gszc) {
          constexpr int T0 = decltype(t0c)::value, GSZ = decltype(gszc)::value;
          // (the host makes the ring size a multiple of the group size, so a group's slots are consecutive and share a phase)
          int slot0;
          uint32_t par_b = 0;
          if (p.stationary) {
            slot0 = c * p.num_taps + T0;
          } else {
            slot0 = rbi;
            par_b = phb;
            rbi += GSZ;
            if (rbi == p.sb) {
              rbi = 0;
              phb ^= 1u;
            }
          }
          uint32_t bars[4], pars[4];
          bars[0] = a_full0 + 8 * ai;
          pars[0] = pha;
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            bars[j + 1] = b_full0 + 8 * (slot0 + j);
            pars[j + 1] = par_b;
          }
          const uint32_t mask = (T0 == 0 ? 1u : 0u) | (wait_b ? (GSZ == 3 ? 14u : 2u) : 0u);
          long long tw2 = wc.begin();
          if (mask != 0) wait_set(bars, pars, mask);
          wc.end(3, tw2);
          tcgen05_fence_after();
          uint32_t b_lo = umma_desc_lo(b_ring + slot0 * p.b_slot_bytes);
          const uint32_t b_step = p.b_slot_bytes >> 4;
#pragma unroll
          for (int j = 0; j < GSZ; ++j) {
            const int t = T0 + j;                                   // compile-time after unrolling
            const uint32_t a_lo = umma_desc_lo(a_base + p.tap_a_off[t]);
            const uint32_t hi_a = p.tap_hi_a[t];
            const uint32_t accf = (c | t) != 0 ? 1u : 0u;
#pragma unroll
            for (int mt = 0; mt < 4; ++mt) {
              if (mt < m_cur) {
                const uint32_t d_t = tmem_acc + mt * block_n;
                const uint32_t al = a_lo + mt * sub16;
                if (KS == 4) {
                  if (PAIR) umma2_bf16_k4(d_t, al, hi_a, b_lo, hi_b, idesc, accf);
                  else umma_bf16_k4(d_t, al, hi_a, b_lo, hi_b, idesc, accf);
                } else if (KS == 2) {   // 32-channel chunks
                  if (PAIR) umma2_bf16_k2(d_t, al, hi_a, b_lo, hi_b, idesc, accf);
                  else umma_bf16_k2(d_t, al, hi_a, b_lo, hi_b, idesc, accf);
                } else {
#pragma unroll 1
                  for (int k = 0; k < p.ksteps; ++k) {
                    if (PAIR) umma2_bf16_k1(d_t, al + 2 * k, hi_a, b_lo + 2 * k, hi_b, idesc, accf | (k != 0 ? 1u : 0u));
                    else umma_bf16_k1(d_t, al + 2 * k, hi_a, b_lo + 2 * k, hi_b, idesc, accf | (k != 0 ? 1u : 0u));
                  }
                }
              }
            }
            if (!p.stationary) commit(b_empty0 + 8 * (slot0 + j));
            b_lo += b_step;
          }
          if (T0 + GSZ >= p.num_taps) {
            commit(a_empty0 + 8 * ai);
            if (c == p.chunks - 1) commit(smem_u32(&tmem_full_bar[acc]));
          }
        };
        using std::integral_constant;
        if (p.num_taps == 9) {
          if (wait_b) {
            group(integral_constant<int, 0>{}, integral_constant<int, 3>{});
            group(integral_constant<int, 3>{}, integral_constant<int, 3>{});
            group(integral_constant<int, 6>{}, integral_constant<int, 3>{});
          } else {   // resident weights that have landed: nothing to wait for between taps
            group(integral_constant<int, 0>{}, integral_constant<int, 9>{});
          }
        } else {
          group(integral_constant<int, 0>{}, integral_constant<int, 1>{});
        }
        if (++ai == p.sa) {
          ai = 0;
          pha ^= 1u;
        }
      }
    }
    };   // run
    if (p.ksteps == 4) run(std::integral_constant<int, 4>{});
    else if (p.ksteps == 2) run(std::integral_constant<int, 2>{});
    else run(std::integral_constant<int, 0>{});
    if (p.tl.buf != nullptr) tl_marks[2] = global_timer_ns();
    if (wc.on) {
      wc.end(0, role_t0);
      if (p.prof != nullptr)
        for (int i = 0; i < 4; ++i) p.prof[blockIdx.x * 16 + 3 + i] = wc.acc[i];
      if (p.tl.buf != nullptr) {
        tl_marks[4] = wc.acc[0];   // role lifetime
        tl_marks[5] = wc.acc[3];   // waiting for operands (A / B full)
        tl_marks[6] = wc.acc[1];   // waiting for a free accumulator
      }
    }
    }   // elected lane
    __syncwarp();
  } else if (warp >= 10) {
    // ------------------------------------------------------------------ EXTRA 7: issuer of the fused 1x1 tail.  An
    // epilogue group turns a sub-tile's accumulator into the activated bf16 panel in its staging buffer -- 128 rows x
    // 128 B, SWIZZLE_128B: exactly the canonical K-major A operand for K = 64 -- and signals p_full; this thread issues
    // D2[group] = panel . W2^T (four K steps) and commits d2_full; the group reads D2 back, applies bias2 / activation
    // and stores through the same staging buffer.  The intermediate never reaches memory.
    if (TAIL && elect_one()) {
      int total[2] = {0, 0};
      const int n_it = sched.iters();
      for (int it = 0; it < n_it; ++it) {
        int s0, m_cur;
        sched.get(it, s0, m_cur);
        total[it & 1] += m_cur;
      }
      wait_addr(smem_u32(&w2_full_bar), 0);
      const uint32_t w2_lo = umma_desc_lo(smem_base + p.off_w2);
      int done[2] = {0, 0};
      const long long t_start = clock64();
      while (done[0] < total[0] || done[1] < total[1]) {
        if (clock64() - t_start > 8000000000ll) halo_wait_timeout(smem_u32(&p_full_bar[0]), 77u);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          if (done[g] < total[g] && test_wait_addr(smem_u32(&p_full_bar[g]), done[g] & 1)) {   // serve whichever group is ready
            tcgen05_fence_after();
            const uint32_t a_lo = umma_desc_lo(smem_base + p.off_stage + g * kStageBytes);
            umma_bf16_k4(tmem_base + 2 * acc_cols + g * 64, a_lo, p.hi_stage, w2_lo, p.hi_w2, p.idesc2, 0u);
            umma_commit(&d2_full_bar[g]);
            ++done[g];
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: 2 groups x 4 warps
    static_assert(!(OUT_F32 && (EXTRA == 1 || EXTRA == 2 || EXTRA == 5 || EXTRA == 6 || EXTRA == 7 || EXTRA == 8)), "extra operands are only combined with bf16 output");
    // EXTRA 5 = the bf16 residual of EXTRA 1, but its panels arrive through a two-slot TMA ring in shared memory (requested
    // by the group's leader two panels ahead) instead of per-thread register loads: with 8 epilogue warps the register
    // path keeps only ~16 KB in flight per SM and the residual read ran at ~2 TB/s next to everything else.
    // EXTRA 6 = the f32 upsample-add of EXTRA 2 through the same ring (1x1 conv run on 16x8 pixel tiles): a panel's operand
    // is the 8x4 low-resolution pixels under the tile, two [32 px x 32 f32] SWIZZLE_128B boxes, and the thread of output
    // pixel (y, x) reads low-resolution row (y/2)*4 + x/2.  The register path issued 8 LDG.128 per 32 columns per thread
    // whose lanes hit 16 different 512-byte-apart rows; the epilogue, not HBM, bounded those layers (MMA role waited for
    // a free accumulator 77 % of the kernel).
    constexpr int MX = EXTRA == 5 ? 1 : (EXTRA == 6 ? 2 : ((EXTRA == 7 || EXTRA == 8) ? 0 : EXTRA));   // arithmetic flavour of the extra operand
    constexpr bool RING = EXTRA == 5 || EXTRA == 6;
    static_assert((EXTRA != 3 && EXTRA != 4) || (OUT_F32 && !SILU), "the fused head decode consumes the f32 logits of a bare 1x1 conv");
    constexpr int kPanelCols = OUT_F32 ? 32 : 64;      // a full staging panel row is 128 bytes
    constexpr int kChunks = kPanelCols / 32;           // 32-column chunks per panel
    constexpr int kWordsPerChunk = OUT_F32 ? 32 : 16;  // packed output words of one chunk
    const int group = (warp - 2) >> 2;
    const int q = warp & 3;                 // TMEM lane quarter this warp may read
    const int r = q * 32 + lane;            // accumulator row = pixel within the sub-tile
    // pixel of row r inside a tile: (r / 8, r % 8) in a 16 x 8 patch, (r / patch_cols, r % patch_cols) in a row-flat tile
    const int rh = p.patch_mode == 2 ? p.fd_patch_cols.div(r) : (r >> 3);
    const int rw = p.patch_mode == 2 ? r - rh * p.patch_cols : (r & 7);
    const bool r_valid = r < p.rows_valid;
    const bool leader = q == 0 && lane == 0;   // one thread per group issues the TMA stores
    const uint32_t stage = smem_base + p.off_stage + group * kStageBytes;
    const uint32_t bar_id = 1 + group;
    const int npanels = ceil_div(p.block_n, p.pc);
    const uint32_t row_off = static_cast<uint32_t>(r) * p.pb;
    const int sts_chunks = p.pb >> 4;       // 16-byte chunks per staging row (2, 4 or 8)
    WaitClock wc(p.prof != nullptr && leader);
    const long long role_t0 = wc.begin();
    if constexpr (EXTRA == 3 || EXTRA == 4) {
      // ---------------------------------------------------------------- fused head decode (flat 1x1 mode only)
      // The last conv of a head tower writes no logits: this thread owns one anchor row of the accumulator and reduces it
      // on the spot -- DFL + anchor/stride decode -> NMS corners (EXTRA 3), or class max on the logits -> (score, class)
      // (EXTRA 4) -- with the arithmetic of head_decode_kernel (decode_math.cuh), so the candidates are bit-identical to
      // the unfused path.  No staging panel, no TMA store, no block barriers.
      uint32_t lt = 0;
      const int n_it = sched.iters();
      const int hw = p.hout * p.wout;
#pragma unroll 1
      for (int it = 0; it < n_it; ++it, ++lt) {
        if ((lt & 1) != static_cast<uint32_t>(group)) continue;
        int s0, m_cur;
        sched.get(it, s0, m_cur);
        wait_addr(smem_u32(&tmem_full_bar[group]), (lt >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll 1
        for (int mt = 0; mt < m_cur; ++mt) {
          const int sr = p.reverse ? p.num_subtiles - 1 - (s0 + mt) : s0 + mt;
          const long long pix = static_cast<long long>(sr) * 128 + r;
          const bool valid = pix < p.mtot;
          const int ipix = valid ? static_cast<int>(pix) : 0;
          const int img = p.fd_hw.div(ipix);
          const int a = ipix - img * hw;                      // anchor inside the level
          const size_t g = static_cast<size_t>(img) * p.dec_anchors + p.dec_level_off + a;
          const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + group * acc_cols + mt * p.block_n;
          if (EXTRA == 3) {
            float dist[4];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(taddr0 + ch * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int sd = 0; sd < 2; ++sd) {
                float l[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) l[i] = __fadd_rn(__uint_as_float(v[sd * 16 + i]), bias_s[ch * 32 + sd * 16 + i]);
                dist[ch * 2 + sd] = dfl_side(l);
              }
            }
            if (mt == m_cur - 1) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[group]);
            }
            if (valid) {
              const int gy = p.fd_wout.div(a), gx = a - gy * p.wout;
              const float4 bpx = box_xywh_px(dist[0], dist[1], dist[2], dist[3], gx, gy, p.dec_stride);
              reinterpret_cast<float4*>(p.cand_box)[g] = box_corners_norm(bpx, p.dec_in_w, p.dec_in_h);
            }
          } else {
            // fn(logit, class, j) over this anchor's nc classes; j = class & 3 lets the caller keep four independent
            // dependency chains (a single running top-2 is a serial chain of ~5 instructions per class and the eight
            // epilogue warps of a CTA cannot hide it)
            auto scan = [&](auto&& fn) {
#pragma unroll 1
              for (int c0 = 0; c0 < p.block_n; c0 += 32) {
                uint32_t v[32];
                if (p.block_n - c0 >= 32) {
                  tmem_ld_32x32b_x32(taddr0 + c0, v);
                } else {
                  tmem_ld_32x32b_x16(taddr0 + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
#pragma unroll
                  for (int i = 16; i < 32; ++i) v[i] = 0;
                }
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i += 4) {
                  const float4 b4 = *reinterpret_cast<const float4*>(bias_s + c0 + i);
                  const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                  for (int e = 0; e < 4; ++e)
                    if (c0 + i + e < p.dec_nc) fn(__fadd_rn(__uint_as_float(v[i + e]), bb[e]), c0 + i + e, e);
                }
              }
            };
            LogitTop2 tops[4];
            scan([&](float x, int c, int j) { tops[j].add(x, c); });
            tops[0].merge(tops[1]);
            tops[2].merge(tops[3]);
            tops[0].merge(tops[2]);
            const LogitTop2 top = tops[0];
            const float thr_logit = tie_window_threshold(top.x1);
            float best = -1.0f;
            int bi = 0x7fffffff;
            // rare: several logits of some row inside its tie window -> evaluate them all.  tcgen05.ld is warp-collective,
            // so the rescan is decided per warp; rows that do not need it keep their single-candidate result.
            const bool rescan = top.x2 >= thr_logit;
            if (__any_sync(0xffffffffu, rescan))
              scan([&](float x, int c, int) { if (rescan && x >= thr_logit) better(best, bi, sigmoid_ref(x), c); });
            if (!rescan && top.x1 >= thr_logit) {
              best = sigmoid_ref(top.x1);
              bi = top.c1;
            }
            if (mt == m_cur - 1) {
              tcgen05_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(&tmem_empty_bar[group]);
            }
            if (valid) {
              p.cand_conf[g] = best;
              p.cand_cls[g] = bi;
            }
          }
        }
      }
    } else {
    // tile coordinates of sub-tile s and this thread's row of the extra operand (null: row outside the tensor)
    auto locate = [&](int s_sched, int& c1, int& c2, int& c3, const void*& ex_row) -> bool {
      ex_row = nullptr;
      const bool tile_valid = tile_coords(s_sched, c1, c2, c3);
      if (p.patch_mode) {
        const int img = c3;
        if (EXTRA == 1 || EXTRA == 2) {
          const int h = c2 + rh, w = c1 + rw;
          if (tile_valid && r_valid && h < p.hout && w < p.wout) {
            if (EXTRA == 1)
              ex_row = p.residual + ((static_cast<long long>(img) * p.hout + h) * p.wout + w) * p.res_pitch + n0;
            else
              ex_row = p.upadd + ((static_cast<long long>(img) * (p.hout >> 1) + (h >> 1)) * (p.wout >> 1) + (w >> 1)) * p.cout + n0;
          }
        }
      } else {
        if (EXTRA == 1 || EXTRA == 2) {
          const long long pix = static_cast<long long>(c1) + r;
          if (tile_valid && pix < p.mtot) {
            if (EXTRA == 1) {
              ex_row = p.residual + pix * p.res_pitch + n0;
            } else {   // only the upsample-add needs (img, h, w) of a flat pixel index
              const int ipix = static_cast<int>(pix);
              const int img = p.fd_hw.div(ipix);
              const int rem = ipix - img * (p.hout * p.wout);
              const int h = p.fd_wout.div(rem);
              const int w = rem - h * p.wout;
              ex_row = p.upadd + ((static_cast<long long>(img) * (p.hout >> 1) + (h >> 1)) * (p.wout >> 1) + (w >> 1)) * p.cout + n0;
            }
          }
        }
      }
      return tile_valid;
    };
    // The extra operand was written by an earlier kernel and its first chunk is requested BEFORE the accumulator wait
    // (the loads do not depend on this grid's MMAs), so these threads order themselves after the earlier grids directly.
    if (EXTRA != 0 && EXTRA != 7 && EXTRA != 8 && !p.nowait) pdl_wait();
    const int first_cols = min(32, p.block_n);
    int tj = 0;   // EXTRA 7: sub-tiles this group has sent through the tail
    uint32_t lt = 0;
    const int n_it = sched.iters();
    // ---- residual ring (EXTRA 5): leader-side prefetch cursor over this group's panels, consumer-side panel counter
    const uint32_t res_base = smem_base + p.off_res + group * 2 * kStageBytes;
    const uint32_t res_bar0 = smem_u32(&res_full_bar[group * 2]);
    int pf_it = group, pf_mt = 0, pf_pn = 0, pf_s0 = 0, pf_m = 0, pf_j = 0;
    auto pf_issue = [&]() {   // request the cursor's panel into slot (pf_j & 1) and advance
      if (pf_it >= n_it) return;
      if (pf_mt == 0 && pf_pn == 0) sched.get(pf_it, pf_s0, pf_m);
      int c1, c2, c3;
      const void* unused;
      locate(pf_s0 + pf_mt, c1, c2, c3, unused);
      const int slot = pf_j & 1;
      mbar_arrive_expect_tx(&res_full_bar[group * 2 + slot], p.res_tx_bytes);
      if (EXTRA == 6) {
        tma_load_4d(&p.tm_res, &res_full_bar[group * 2 + slot], res_base + slot * kStageBytes, n0 + pf_pn * p.pc, c1 >> 1, c2 >> 1, c3);
        tma_load_4d(&p.tm_res, &res_full_bar[group * 2 + slot], res_base + slot * kStageBytes + 4096, n0 + pf_pn * p.pc + 32,
                    c1 >> 1, c2 >> 1, c3);
      } else {
        tma_load_4d(&p.tm_res, &res_full_bar[group * 2 + slot], res_base + slot * kStageBytes, n0 + pf_pn * p.pc, c1, c2, c3);
      }
      ++pf_j;
      if (++pf_pn == npanels) {
        pf_pn = 0;
        if (++pf_mt == pf_m) {
          pf_mt = 0;
          pf_it += 2;
        }
      }
    };
    if (RING && leader) {
      pf_issue();
      pf_issue();
    }
    const uint32_t lr_off = static_cast<uint32_t>(((r >> 4) << 2) + ((r & 7) >> 1)) * 128u;   // EXTRA 6: this row's low-res row
    int rj = 0;   // panels consumed by this group
    // The LAST accumulator of a CTA is drained by BOTH groups (alternate panels): nothing is left to overlap with, and the
    // time from the last MMA to the CTA's exit is a tensor-pipe bubble on this SM (measured 4 us of a ~35 us CTA).
    constexpr bool kSplitLast = EXTRA == 0;
#pragma unroll 1
    for (int it = 0; it < n_it; ++it, ++lt) {
      const uint32_t acc_stage = lt & 1;   // accumulator stage of this iteration (its owner is group acc_stage)
      const bool split = kSplitLast && it == n_it - 1;
      if (acc_stage != static_cast<uint32_t>(group) && !split) continue;
      const uint32_t split_sel = acc_stage == static_cast<uint32_t>(group) ? 0u : 1u;   // owner: even panels, helper: odd panels
      uint32_t item = 0;
      int s0, m_cur;
      sched.get(it, s0, m_cur);
      int nc1 = 0, nc2 = 0, nc3 = 0;
      const void* nrow = nullptr;
      ExtraRegs<MX> exn;                     // first chunk of the coming sub-tile's extra operand
      bool nvalid = locate(s0, nc1, nc2, nc3, nrow);
      if (EXTRA != 0 && !RING && nrow != nullptr) exn.load(nrow, 0, first_cols);
      long long tw = wc.begin();
      wait_addr(smem_u32(&tmem_full_bar[acc_stage]), (lt >> 1) & 1);
      wc.end(1, tw);
      if (p.tl.buf != nullptr && leader) {
        const unsigned long long now = global_timer_ns();
        if (it == 0) tl_marks[1] = now;
        tl_marks[3] = now;   // (the two groups alternate: the later write is the later accumulator)
      }
      tcgen05_fence_after();
#pragma unroll 1
      for (int mt = 0; mt < m_cur; ++mt) {
        const int c1 = nc1, c2 = nc2, c3 = nc3;
        const bool tile_valid = nvalid;     // false: the odd CTA of a pair past the last sub-tile (nothing to store)
        const void* ex_row = nrow;
        const bool ex_valid = ex_row != nullptr;
        const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc_stage * acc_cols + mt * p.block_n;
        ExtraRegs<MX> ex[kChunks];
        ex[0] = exn;
        bool ex0_ready = true;              // ex[0] already holds the first chunk of the coming panel
        if (mt + 1 < m_cur) {               // request the next sub-tile's first chunk a whole sub-tile ahead
          nvalid = locate(s0 + mt + 1, nc1, nc2, nc3, nrow);
          if (EXTRA != 0 && !RING && nrow != nullptr) exn.load(nrow, 0, first_cols);
        }
#pragma unroll 1
        for (int pn = 0; pn < npanels; ++pn) {
          if (kSplitLast && split && ((item++ & 1u) != split_sel)) continue;   // the other group's panel
          const int col0 = pn * p.pc;
          const int ncols = min(p.pc, p.block_n - col0);
          if (EXTRA != 0 && !RING && ex_valid && !ex0_ready) ex[0].load(ex_row, col0, min(32, ncols));
          ex0_ready = false;
          uint32_t res_slot = 0;
          if (RING) {                       // this panel's residual / upsample-add operand has landed in the ring
            res_slot = res_base + (rj & 1) * kStageBytes;
            tw = wc.begin();
            wait_addr(res_bar0 + 8 * (rj & 1), (rj >> 1) & 1);
            wc.end(3, tw);   // (tools) ring waits are reported together with the staging-written barrier
            ++rj;
          }
          uint32_t o[kChunks * kWordsPerChunk];
#pragma unroll
          for (int ch = 0; ch < kChunks; ++ch) {
            const int cc = ch * 32;
            if (cc < ncols) {               // warp-uniform
              const bool full = ncols - cc >= 32;
              uint32_t v[32];
              if (full) {
                tmem_ld_32x32b_x32(taddr0 + col0 + cc, v);
              } else {
                tmem_ld_32x32b_x16(taddr0 + col0 + cc, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
              }
              if (EXTRA == 5) {               // this row's 64 bytes of the chunk from the swizzled ring slot
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  if (cc * 2 + i * 16 < ncols * 2) {
                    const uint32_t off = swz(row_off + cc * 2 + i * 16, p.smask);
                    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(ex[ch].q[i].x), "=r"(ex[ch].q[i].y), "=r"(ex[ch].q[i].z), "=r"(ex[ch].q[i].w)
                                 : "r"(res_slot + off));
                  }
              }
              if (EXTRA == 6) {               // 128 bytes (32 f32) of this row's low-resolution pixel from the chunk's box
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                  const uint32_t off = swz(lr_off + i * 16, 7u);
                  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                               : "=r"(ex[ch].q[i].x), "=r"(ex[ch].q[i].y), "=r"(ex[ch].q[i].z), "=r"(ex[ch].q[i].w)
                               : "r"(res_slot + ch * 4096 + off));
                }
              }
              if (EXTRA != 0 && !RING && ex_valid) {   // issue the next chunk's extra-operand loads one chunk ahead of their use
                if (ch + 1 < kChunks) {
                  if (cc + 32 < ncols) ex[ch + 1].load(ex_row, col0 + cc + 32, min(32, ncols - cc - 32));
                } else if (pn + 1 < npanels) {
                  ex[0].load(ex_row, col0 + p.pc, min(32, p.block_n - col0 - p.pc));
                  ex0_ready = true;
                }
              }
              tmem_ld_wait();
              if (full)
                epilogue_math<32, SILU, OUT_F32, MX>(v, bias_s + col0 + cc, ex[ch], ex_valid || RING, &o[ch * kWordsPerChunk]);
              else
                epilogue_math<16, SILU, OUT_F32, MX>(v, bias_s + col0 + cc, ex[ch], ex_valid || RING, &o[ch * kWordsPerChunk]);
            }
          }
          if (mt == m_cur - 1 && pn == npanels - 1 && !split) {
            // last TMEM read of this accumulator stage: hand it back to the MMA warp (nobody waits for the CTA's last one)
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(smem_u32(&tmem_empty_bar[acc_stage]), 0);   // the leader's MMA issuer owns both halves
              else mbar_arrive(&tmem_empty_bar[acc_stage]);
            }
          }
          // the previous panel's TMA store must have finished reading the staging buffer
          tw = wc.begin();
          if (leader) bulk_wait_read_all();
          named_bar_sync(bar_id, 128);
          wc.end(2, tw);
          if (RING && leader) pf_issue();   // every thread has read this panel's ring slot: refill it (two panels ahead)
          const int nvalid16 = (ncols * (OUT_F32 ? 4 : 2)) >> 4;   // 16-byte chunks of this row that hold data
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (j < sts_chunks && j < nvalid16) {
              const uint32_t off = swz(row_off + j * 16, p.smask);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + off), "r"(o[4 * j]), "r"(o[4 * j + 1]),
                           "r"(o[4 * j + 2]), "r"(o[4 * j + 3])
                           : "memory");
            }
          }
          fence_proxy_async_smem();
          tw = wc.begin();
          named_bar_sync(bar_id, 128);
          wc.end(3, tw);
          if constexpr (EXTRA == 8) {
            // box tower: D2 = the 4 x 16 DFL logits of this row's anchor -> distances -> NMS corners (EXTRA 3's arithmetic)
            if (leader) mbar_arrive(&p_full_bar[group]);
            wait_addr(smem_u32(&d2_full_bar[group]), tj & 1);
            ++tj;
            tcgen05_fence_after();
            const uint32_t taddr2 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * acc_cols + group * 64;
            float dist[4];
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(taddr2 + ch * 32, v);
              tmem_ld_wait();
#pragma unroll
              for (int sd = 0; sd < 2; ++sd) {
                float l[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) l[i] = __fadd_rn(__uint_as_float(v[sd * 16 + i]), bias2_s[ch * 32 + sd * 16 + i]);
                dist[ch * 2 + sd] = dfl_side(l);
              }
            }
            tcgen05_fence_before();
            const int ah = c2 + rh, aw = c1 + rw;
            if (r_valid && ah < p.hout && aw < p.wout) {
              const size_t g = static_cast<size_t>(c3) * p.dec_anchors + p.dec_level_off + ah * p.wout + aw;
              const float4 bpx = box_xywh_px(dist[0], dist[1], dist[2], dist[3], aw, ah, p.dec_stride);
              reinterpret_cast<float4*>(p.cand_box)[g] = box_corners_norm(bpx, p.dec_in_w, p.dec_in_h);
            }
          }
          if constexpr (EXTRA == 7) {
            // the activated panel is the A operand of the tail GEMM: hand it to the tail issuer, read D2 back
            if (leader) mbar_arrive(&p_full_bar[group]);
            wait_addr(smem_u32(&d2_full_bar[group]), tj & 1);
            ++tj;
            tcgen05_fence_after();
            const uint32_t taddr2 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + 2 * acc_cols + group * 64;
#pragma unroll
            for (int ch = 0; ch < 2; ++ch) {
              uint32_t v[32];
              tmem_ld_32x32b_x32(taddr2 + ch * 32, v);
              tmem_ld_wait();
              epilogue_math<32, SILU, OUT_F32, 0>(v, bias2_s + ch * 32, ex[0], false, &o[ch * kWordsPerChunk]);
            }
            tcgen05_fence_before();
            // d2_full also means the tail MMAs have finished reading the staging panel: overwrite it with the output
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t off = swz(row_off + j * 16, p.smask);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + off), "r"(o[4 * j]), "r"(o[4 * j + 1]),
                           "r"(o[4 * j + 2]), "r"(o[4 * j + 3])
                           : "memory");
            }
            fence_proxy_async_smem();
            named_bar_sync(bar_id, 128);
          }
          if (EXTRA != 8 && leader && tile_valid) {
            tma_store_4d(&p.tm_out, stage, n0 + col0, c1, c2, c3);
            bulk_commit_group();
          }
        }
      }
    }
    // the staging buffers must outlive the stores' READS; their writes are this grid's memory operations, which grid
    // completion (and the dependent grid's programmatic-launch wait) covers
    if (leader) bulk_wait_read_all();
    if (wc.on) {
      wc.end(0, role_t0);
      for (int i = 0; i < 4; ++i) p.prof[blockIdx.x * 16 + 8 + group * 4 + i] = wc.acc[i];
    }
    }   // EXTRA < 3
  }

  tcgen05_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();   // no CTA of a pair leaves while the other may still signal its barriers / read its operands
  if (warp == 1) {
    tcgen05_fence_after();
    if (PAIR) tmem_dealloc2(tmem_base, p.tmem_cols);
    else tmem_dealloc(tmem_base, p.tmem_cols);
  }
  if (threadIdx.x == 0) timeline_write(p.tl, tl_t0, tl_marks[0], tl_marks[1], tl_marks[2], tl_marks[3], tl_marks[4], tl_marks[5], tl_marks[6]);
}

// ------------------------------------------------------------------------------------------------ host
static unsigned long long* g_prof = nullptr;   // set by tod_debug_set_conv_profile (tools only)

static uint32_t desc_hi(uint32_t sbo_bytes, int bk) {
  const uint32_t layout_type = bk == 64 ? 2u : (bk == 32 ? 4u : 6u);  // SWIZZLE_128B / 64B / 32B
  return (sbo_bytes >> 4) | (1u << 14) | (layout_type << 29);
}

// Fills `p` for K-chunk width bk (cin_pad is fixed by the packed weights).  *fits = false when no shared-memory plan
// exists for this bk (the caller retries with a narrower chunk).
// res_ring: 0 none, 1 bf16 residual panels, 2 f32 upsample-add panels (1x1 conv on 16x8 pixel tiles) through the
// shared-memory TMA ring of the epilogue groups.
static int build_params(const tod_conv_desc* d, int bk, int cin_pad, HaloParams& p, size_t* smem_bytes, bool* fits,
                        int res_ring = 0, bool tail = false, double* cost_out = nullptr, bool pair = false) {
  int rc;
  *fits = true;
  memset(&p, 0, sizeof(p));
  const int hout = d->hin / d->stride, wout = d->win / d->stride;
  const int taps = d->ksize * d->ksize;
  const int k_total = taps * cin_pad;
  const uint32_t rb = bk * 2;  // bytes per smem row
  const CUtensorMapSwizzle swz_in =
      bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const uint64_t px = static_cast<uint64_t>(d->x_pitch) * 2;
  const long long mtot = static_cast<long long>(d->batch) * hout * wout;
  // (stride 2: the four parity planes of a tile together read every pixel, so the pitch that matters is still one pixel)
  const CUtensorMapL2promotion promo_in = l2_promotion_for(static_cast<uint64_t>(d->cin) * 2, px);

  p.hout = hout;
  p.wout = wout;
  p.mtot = mtot;
  p.batch = d->batch;
  p.cout = d->cout;
  p.block_k = bk;
  p.ksteps = bk >> 4;
  p.chunks = cin_pad / bk;
  p.num_taps = taps;
  const int n_cap = (d->reserved[3] >= 16 && d->reserved[3] <= 256) ? d->reserved[3] : 256;   // tools: cap on the N tile
  p.n_tiles = ceil_div(d->cout, n_cap);
  p.block_n = round_up(ceil_div(d->cout, p.n_tiles), 16);
  // CTA-pair mode: each CTA of the pair holds (and loads) half of the rows of every weight tile
  p.b_half_rows = p.block_n / 2;
  const int b_rows = pair ? p.block_n / 2 : p.block_n;
  p.b_tx_bytes = b_rows * rb;
  p.b_slot_bytes = round_up(b_rows * rb, 1024);
  p.hi_b = desc_hi(8 * rb, bk);

  TOD_CHECK_ARG(mtot < (1ll << 31) - 256, "conv: too many output pixels");
  p.oob_row = static_cast<int>((mtot + 127) / 128 * 128);
  // ---- A operand: tensor maps, loads per (sub-tile, chunk), per-tap descriptor views
  uint32_t sub_bytes = 0;
  if (d->ksize == 1 && res_ring != 2) {
    p.patch_mode = 0;
    const uint32_t rows = mtot < 128 ? static_cast<uint32_t>(mtot) : 128u;
    const uint64_t dims[4] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(mtot), 1, 1};
    const uint64_t str[3] = {px, px * mtot, px * mtot};
    const uint32_t box[4] = {static_cast<uint32_t>(bk), rows, 1, 1};
    if ((rc = encode_map(&p.tm_a[0], d->d_x, 4, dims, str, box, swz_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_in)) != TOD_OK) return rc;
    p.n_aloads = 1;
    p.al_map[0] = 0;
    p.a_tx_bytes = rows * rb;
    p.tap_a_off[0] = 0;
    p.tap_hi_a[0] = desc_hi(8 * rb, bk);
    sub_bytes = 128 * rb;
    p.num_subtiles = static_cast<int>((mtot + 127) / 128);
    p.tiles_w = 1;
    p.tiles_per_img = 1;
    p.patch_cols = kPatchW;
    p.rows_valid = 128;
  } else {
    // Row-flat tiles (3x3 stride 1 on small maps): a tile is `rt` whole output rows of (wout + 2) positions -- the two extra
    // positions per row are the halo columns, computed and never stored -- so that accumulator row r is simply the r-th
    // position in row-major order and a tap (kh, kw) is the SAME flat run of shared-memory rows shifted by kh * (wout + 2)
    // + kw: any map whose width is not a multiple of 8 / height not a multiple of 16 tiles far better this way
    // (20x20: 4 tiles of 5 rows per image instead of 6 patches of 16x8, i.e. 78 % instead of 52 % useful MMA rows;
    // 40x40: 14 tiles of 3 rows, 89 % instead of 83 %).
    const int flat_cols = wout + 2, flat_rt = flat_cols <= 128 ? 128 / flat_cols : 0;
    bool flat = false;
    if (d->ksize == 3 && d->stride == 1 && res_ring != 2 && flat_rt >= 1 && flat_tiles_enabled() && !(d->flags & TOD_CONV_PATCH_TILES)) {
      // taken when it saves at least a fifth of the tiles: a flat tile's patch is ~17 % larger than a 16x8 one, and at
      // 40x40 (14 tiles instead of 15) the shared memory that costs the weight ring made the layer slower (49 vs 37 us)
      const long long patch_tiles = static_cast<long long>(ceil_div(wout, kPatchW)) * ceil_div(hout, kPatchH);
      flat = ceil_div(hout, flat_rt) * 100ll <= patch_tiles * 80ll;
    }
    p.patch_mode = flat ? 2 : 1;
    p.tiles_w = flat ? 1 : ceil_div(wout, kPatchW);
    p.tiles_per_img = flat ? ceil_div(hout, flat_rt) : p.tiles_w * ceil_div(hout, kPatchH);
    p.tile_w_step = flat ? 0 : kPatchW;
    p.tile_h_step = flat ? flat_rt : kPatchH;
    p.patch_cols = flat ? flat_cols : kPatchW;
    p.rows_valid = flat ? flat_rt * flat_cols : 128;
    const long long nsub = static_cast<long long>(d->batch) * p.tiles_per_img;
    TOD_CHECK_ARG(nsub < (1ll << 30), "conv: too many tiles");
    p.num_subtiles = static_cast<int>(nsub);
    if (d->ksize == 1) {   // 1x1 on pixel tiles (no halo): one box = 128 rows in the canonical 8-row-group layout
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->win),
                                static_cast<uint64_t>(d->hin), static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {px, px * d->win, px * d->win * d->hin};
      const uint32_t box[4] = {static_cast<uint32_t>(bk), kPatchW, kPatchH, 1};
      if ((rc = encode_map(&p.tm_a[0], d->d_x, 4, dims, str, box, swz_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_in)) != TOD_OK) return rc;
      p.n_aloads = 1;
      p.al_map[0] = 0;
      p.a_tx_bytes = 128 * rb;
      p.tap_a_off[0] = 0;
      p.tap_hi_a[0] = desc_hi(8 * rb, bk);
      sub_bytes = 128 * rb;
    } else if (flat) {
      const int ph = flat_rt + 2;                      // rows of the loaded box: the tile's rows plus one halo row each side
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->win),
                                static_cast<uint64_t>(d->hin), static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {px, px * d->win, px * d->win * d->hin};
      const uint32_t box[4] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(flat_cols), static_cast<uint32_t>(ph), 1};
      if ((rc = encode_map(&p.tm_a[0], d->d_x, 4, dims, str, box, swz_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_in)) != TOD_OK) return rc;
      p.n_aloads = 1;
      p.al_map[0] = 0;
      p.al_dw[0] = -1;
      p.al_dh[0] = -1;
      p.al_off[0] = 0;
      p.a_tx_bytes = flat_cols * ph * rb;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          p.tap_a_off[kh * 3 + kw] = (kh * flat_cols + kw) * rb;
          p.tap_hi_a[kh * 3 + kw] = desc_hi(8 * rb, bk);     // consecutive rows: the canonical 8-row group stride
        }
      // the 128 accumulator rows of the last tap read rows [2 * cols + 2, 2 * cols + 130): past the loaded box for the
      // unused rows >= rows_valid (stale shared memory, never stored) -- the slot must cover them
      sub_bytes = (2 * flat_cols + 2 + 128) * rb;
      if (sub_bytes < p.a_tx_bytes) sub_bytes = p.a_tx_bytes;
    } else if (d->stride == 1) {
      const int pw = kPatchW + 2, ph = kPatchH + 2;
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->win),
                                static_cast<uint64_t>(d->hin), static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {px, px * d->win, px * d->win * d->hin};
      const uint32_t box[4] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(pw), static_cast<uint32_t>(ph), 1};
      if ((rc = encode_map(&p.tm_a[0], d->d_x, 4, dims, str, box, swz_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_in)) != TOD_OK) return rc;
      p.n_aloads = 1;
      p.al_map[0] = 0;
      p.al_dw[0] = -1;
      p.al_dh[0] = -1;
      p.al_off[0] = 0;
      p.a_tx_bytes = pw * ph * rb;
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          p.tap_a_off[kh * 3 + kw] = (kh * pw + kw) * rb;
          p.tap_hi_a[kh * 3 + kw] = desc_hi(pw * rb, bk);
        }
      sub_bytes = pw * ph * rb;
    } else {
      // input row 2*oh + kh - 1:  kh=0 -> plane row oh-1 of parity 1, kh=1 -> row oh of parity 0, kh=2 -> row oh of
      // parity 1 (same for columns).  Parity-1 planes are loaded with one extra leading row / column.
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cin), static_cast<uint64_t>(d->win / 2),
                                static_cast<uint64_t>(d->hin / 2), static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {px * 2, px * d->win * 2, px * d->win * d->hin};
      uint32_t plane_off[2][2];
      int n = 0;
      uint32_t off = 0;
      p.a_tx_bytes = 0;
      for (int ph = 1; ph >= 0; --ph)
        for (int pw = 1; pw >= 0; --pw) {
          const int bw = kPatchW + pw, bh = kPatchH + ph;
          const uint8_t* base = reinterpret_cast<const uint8_t*>(d->d_x) + (static_cast<uint64_t>(ph) * d->win + pw) * px;
          const uint32_t box[4] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(bw), static_cast<uint32_t>(bh), 1};
          if ((rc = encode_map(&p.tm_a[n], base, 4, dims, str, box, swz_in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_in)) != TOD_OK) return rc;
          p.al_map[n] = n;
          p.al_dw[n] = -pw;
          p.al_dh[n] = -ph;
          p.al_off[n] = off;
          plane_off[ph][pw] = off;
          p.a_tx_bytes += bw * bh * rb;
          off += round_up(bw * bh * rb, 1024);
          ++n;
        }
      p.n_aloads = n;
      sub_bytes = off;
      const int par[3] = {1, 0, 1};
      for (int kh = 0; kh < 3; ++kh)
        for (int kw = 0; kw < 3; ++kw) {
          const int ph = par[kh], pw = par[kw];
          const int bw = kPatchW + pw;
          const int ro = kh == 2 ? 1 : 0, co = kw == 2 ? 1 : 0;
          p.tap_a_off[kh * 3 + kw] = plane_off[ph][pw] + (ro * bw + co) * rb;
          p.tap_hi_a[kh * 3 + kw] = desc_hi(bw * rb, bk);
        }
    }
  }
  p.sub_bytes = round_up(sub_bytes, 1024);
  p.num_units = pair ? (p.num_subtiles + 1) / 2 : p.num_subtiles;
  {
    auto make_fd = [](long long dd, long long nmax, FastDiv* f) {
      f->d = static_cast<uint32_t>(dd);
      f->mul = ((1ull << 40) + dd - 1) / dd;
      return dd > 0 && nmax * dd < (1ll << 40);
    };
    const bool ok = make_fd(p.tiles_per_img, p.num_subtiles, &p.fd_tiles_per_img) &&
                    make_fd(p.tiles_w, p.tiles_per_img, &p.fd_tiles_w) &&
                    make_fd(p.patch_cols > 0 ? p.patch_cols : 1, 128, &p.fd_patch_cols) &&
                    make_fd(wout, static_cast<long long>(hout) * wout, &p.fd_wout) &&
                    make_fd(static_cast<long long>(hout) * wout, mtot, &p.fd_hw);
    TOD_CHECK_ARG(ok, "conv: problem too large for the tile index arithmetic");
  }
  {
    const uint64_t dims[2] = {static_cast<uint64_t>(k_total), static_cast<uint64_t>(d->cout)};
    const uint64_t str[1] = {static_cast<uint64_t>(k_total) * 2};
    const uint32_t box[2] = {static_cast<uint32_t>(bk), static_cast<uint32_t>(b_rows)};
    if ((rc = encode_map(&p.tm_w, d->d_w, 2, dims, str, box, swz_in)) != TOD_OK) return rc;
  }

  // ---- output: staging panels of 128 rows x pb bytes, TMA store
  p.out_f32 = d->out_dtype == TOD_OUT_F32;
  const int esize = p.out_f32 ? 4 : 2;
  int pc = 128 / esize;
  while (pc > p.block_n) pc >>= 1;
  if (p.n_tiles > 1)
    while (p.block_n % pc) pc >>= 1;
  p.pc = pc;
  p.pb = pc * esize;
  p.smask = p.pb >= 128 ? 7 : (p.pb == 64 ? 3 : 1);
  TOD_CHECK_ARG(p.pb >= 32, "conv: output panel narrower than 32 bytes (cout %d)", d->cout);
  {
    const CUtensorMapSwizzle swz_out =
        p.pb >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.pb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const CUtensorMapDataType dt = p.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    const uint64_t opx = static_cast<uint64_t>(d->out_pitch) * esize;
    if (p.patch_mode) {
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cout), static_cast<uint64_t>(wout), static_cast<uint64_t>(hout),
                                static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {opx, opx * wout, opx * wout * hout};
      const uint32_t box[4] = {static_cast<uint32_t>(pc), static_cast<uint32_t>(p.patch_mode == 2 ? p.patch_cols : kPatchW),
                               static_cast<uint32_t>(p.patch_mode == 2 ? p.tile_h_step : kPatchH), 1};
      if ((rc = encode_map(&p.tm_out, d->d_out, 4, dims, str, box, swz_out, dt)) != TOD_OK) return rc;
    } else {
      const uint32_t rows = mtot < 128 ? static_cast<uint32_t>(mtot) : 128u;
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cout), static_cast<uint64_t>(mtot), 1, 1};
      const uint64_t str[3] = {opx, opx * mtot, opx * mtot};
      const uint32_t box[4] = {static_cast<uint32_t>(pc), rows, 1, 1};
      if ((rc = encode_map(&p.tm_out, d->d_out, 4, dims, str, box, swz_out, dt)) != TOD_OK) return rc;
    }
  }

  if (res_ring == 2) {   // low-resolution f32 operand: [32 f32 x 4 x 8 px] boxes, two per 64-column panel
    if (p.pc != 64 || !p.patch_mode || (hout & 1) || (wout & 1)) {
      *fits = false;
      return TOD_OK;
    }
    const uint64_t upx = static_cast<uint64_t>(d->cout) * 4;
    const uint64_t dims[4] = {static_cast<uint64_t>(d->cout), static_cast<uint64_t>(wout / 2), static_cast<uint64_t>(hout / 2),
                              static_cast<uint64_t>(d->batch)};
    const uint64_t str[3] = {upx, upx * (wout / 2), upx * (wout / 2) * (hout / 2)};
    const uint32_t box[4] = {32, kPatchW / 2, kPatchH / 2, 1};
    if ((rc = encode_map(&p.tm_res, d->d_upadd, 4, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_DATA_TYPE_FLOAT32)) !=
        TOD_OK)
      return rc;
    p.res_tx_bytes = 2u * 4096u;
  } else if (res_ring) {   // residual panels: the output store's boxes and swizzle on the residual tensor
    const CUtensorMapSwizzle swz_res =
        p.pb >= 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (p.pb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const uint64_t rpx = static_cast<uint64_t>(d->res_pitch) * 2;
    const CUtensorMapL2promotion promo_res = l2_promotion_for(static_cast<uint64_t>(d->cout) * 2, rpx);
    if (p.patch_mode) {
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cout), static_cast<uint64_t>(wout), static_cast<uint64_t>(hout),
                                static_cast<uint64_t>(d->batch)};
      const uint64_t str[3] = {rpx, rpx * wout, rpx * wout * hout};
      const uint32_t bw = p.patch_mode == 2 ? p.patch_cols : kPatchW, bh = p.patch_mode == 2 ? p.tile_h_step : kPatchH;
      const uint32_t box[4] = {static_cast<uint32_t>(p.pc), bw, bh, 1};
      if ((rc = encode_map(&p.tm_res, d->d_residual, 4, dims, str, box, swz_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_res)) != TOD_OK) return rc;
      p.res_tx_bytes = static_cast<uint32_t>(p.pb) * bw * bh;
    } else {
      const uint32_t rows = mtot < 128 ? static_cast<uint32_t>(mtot) : 128u;
      const uint64_t dims[4] = {static_cast<uint64_t>(d->cout), static_cast<uint64_t>(mtot), 1, 1};
      const uint64_t str[3] = {rpx, rpx * mtot, rpx * mtot};
      const uint32_t box[4] = {static_cast<uint32_t>(p.pc), rows, 1, 1};
      if ((rc = encode_map(&p.tm_res, d->d_residual, 4, dims, str, box, swz_res, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, promo_res)) != TOD_OK) return rc;
      p.res_tx_bytes = static_cast<uint32_t>(p.pb) * rows;
    }
  }

  // ---- shared-memory plan: m sub-tiles per weight tile, A ring, B ring or resident weights
  // (+ four residual panels when the residual goes through the shared-memory ring; that plan needs resident weights)
  const uint32_t staging = (res_ring ? 6 : 2) * kStageBytes + (tail ? 8192u : 0u);   // tail: + the resident [64 x 64] W2
  const uint32_t budget = kHaloSmemLimit - 1024 - staging - (tail ? 1024u : 0u);   // tail: its extra static shared memory
  const uint32_t b_total = static_cast<uint32_t>(taps) * p.chunks * p.b_slot_bytes;
  const bool may_station = taps * p.chunks <= kMaxB && d->reserved[2] != 1;
  int m_max = (512 - (tail ? 128 : 0)) / (2 * p.block_n);   // tail: 2 x 64 TMEM columns for the second accumulators
  if (m_max > 4) m_max = 4;
  if (m_max < 1) m_max = 1;
  if (m_max > p.num_units) m_max = p.num_units;
  if (res_ring == 1 && m_max > 2) m_max = 2;   // measured (32->32 3x3 @160^2 + residual): m = 2 91 us, m = 4 101 us
  if (d->reserved[1] > 0 && d->reserved[1] < m_max) m_max = d->reserved[1];
  double best_cost = -1.0;
  for (int m = m_max; m >= 1; --m) {
    const uint32_t a_slot = m * p.sub_bytes;
    for (int stn = 1; stn >= (res_ring == 1 ? 1 : 0); --stn) {
      int sa, sb;
      if (stn) {
        if (!may_station || b_total + 2 * a_slot > budget) continue;
        sb = taps * p.chunks;
        sa = static_cast<int>((budget - b_total) / a_slot);
      } else {
        const int sb_min = taps >= 3 ? 3 : 2;
        uint32_t used = 2 * a_slot + sb_min * p.b_slot_bytes;
        if (used > budget) continue;
        sa = 2;
        sb = sb_min;
        // grow whichever ring currently gives the shorter look-ahead (an A slot lasts `taps` weight tiles)
        for (;;) {
          const bool grow_a = (sa - 1) * taps <= (sb - 1);
          if (grow_a && sa < kMaxA && used + a_slot <= budget) {
            ++sa;
            used += a_slot;
          } else if (sb < kMaxB && used + p.b_slot_bytes <= budget) {
            ++sb;
            used += p.b_slot_bytes;
          } else if (!grow_a && sa < kMaxA && used + a_slot <= budget) {
            ++sa;
            used += a_slot;
          } else {
            break;
          }
        }
      }
      if (sa > kMaxA) sa = kMaxA;
      if (sb > kMaxB) sb = kMaxB;
      if (!stn && taps == 9) sb -= sb % 3;   // the MMA issuer consumes weight tiles in groups of three consecutive slots
      // L2 -> SM bytes per output pixel
      const double cost = (static_cast<double>(p.a_tx_bytes) * p.chunks + (stn ? 0.0 : static_cast<double>(b_total) / m)) / 128.0;
      if (best_cost < 0 || cost < best_cost * 0.98) {
        best_cost = cost;
        p.m = m;
        p.sa = sa;
        p.sb = sb;
        p.stationary = stn;
      }
    }
  }
  if (best_cost < 0) {
    *fits = false;
    return TOD_OK;
  }
  if (cost_out != nullptr) *cost_out = best_cost * p.n_tiles;   // L2 -> SM bytes per output pixel over all N tiles
  if (d->num_stages > 0 && d->num_stages < p.sa) p.sa = d->num_stages;
  p.a_slot_bytes = p.m * p.sub_bytes;
  p.off_b = p.sa * p.a_slot_bytes;
  p.off_stage = p.off_b + p.sb * p.b_slot_bytes;
  p.off_res = p.off_stage + 2 * kStageBytes;
  p.off_w2 = p.off_stage + (res_ring ? 6 : 2) * kStageBytes;
  const size_t smem = static_cast<size_t>(p.off_stage) + staging + 1024;   // staging includes the residual ring
  TOD_CHECK_ARG(smem <= kHaloSmemLimit - (tail ? 1024u : 0u), "conv: shared-memory plan overflows (%zu bytes)", smem);
  p.num_super = ceil_div(p.num_units, p.m);

  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(p.block_n >> 3) << 17) | (((pair ? 256u : 128u) >> 4) << 24);
  uint32_t cols = 32;
  while (cols < static_cast<uint32_t>(2 * p.m * p.block_n + (tail ? 128 : 0))) cols <<= 1;
  p.tmem_cols = cols;

  p.bias = d->d_bias;
  p.residual = reinterpret_cast<const __nv_bfloat16*>(d->d_residual);
  p.upadd = d->d_upadd;
  p.res_pitch = d->res_pitch;
  p.act = d->act;
  p.reverse = (d->flags & TOD_CONV_REVERSE) ? 1 : 0;
  p.dyn_w = (d->flags & TOD_CONV_DYNAMIC_W) ? 1 : 0;
  p.nowait = pdl_nowait();
  p.prof = g_prof;
  {
    char nm[48];
    snprintf(nm, sizeof(nm), "conv %d>%d k%d s%d @%dx%d%s", d->cin, d->cout, d->ksize, d->stride, hout, wout, tail ? " +tail" : "");
    p.tl = timeline_tag(nm);
  }
  *smem_bytes = smem;
  return TOD_OK;
}

// kernel variants: [0..2] bf16 out, no activation, extra 0/1/2; [3..5] bf16 out, SiLU, extra 0/1/2; [6] f32 out, no
// activation; [7] f32 out, SiLU; [8] fused box decode; [9] fused class decode; [10] bf16 out, SiLU, residual through
// the shared-memory TMA ring; [11] bf16 out, SiLU, upsample-add through the ring; [12] bf16 out, SiLU, fused 1x1 tail
using HaloKernel = void (*)(HaloParams);
constexpr int kHaloVariants = 14;
// CTA-pair instantiations of the variants that stream weights: [0] / [3] plain, [4] residual (registers), [6] f32 out,
// [10] residual through the ring, [11] upsample-add through the ring
static HaloKernel halo_kernel_pair(int i) {
  switch (i) {
    case 0: return conv_halo_tcgen05<false, false, 0, true>;
    case 3: return conv_halo_tcgen05<true, false, 0, true>;
    case 4: return conv_halo_tcgen05<true, false, 1, true>;
    case 6: return conv_halo_tcgen05<false, true, 0, true>;
    case 10: return conv_halo_tcgen05<true, false, 5, true>;
    case 11: return conv_halo_tcgen05<true, false, 6, true>;
    default: return nullptr;
  }
}
static HaloKernel halo_kernel(int i) {
  switch (i) {
    case 0: return conv_halo_tcgen05<false, false, 0>;
    case 1: return conv_halo_tcgen05<false, false, 1>;
    case 2: return conv_halo_tcgen05<false, false, 2>;
    case 3: return conv_halo_tcgen05<true, false, 0>;
    case 4: return conv_halo_tcgen05<true, false, 1>;
    case 5: return conv_halo_tcgen05<true, false, 2>;
    case 6: return conv_halo_tcgen05<false, true, 0>;
    case 7: return conv_halo_tcgen05<true, true, 0>;
    case 8: return conv_halo_tcgen05<false, true, 3>;
    case 9: return conv_halo_tcgen05<false, true, 4>;
    case 10: return conv_halo_tcgen05<true, false, 5>;
    case 11: return conv_halo_tcgen05<true, false, 6>;
    case 12: return conv_halo_tcgen05<true, false, 7>;
    default: return conv_halo_tcgen05<true, false, 8>;
  }
}

// Where the plan uses CTA pairs by itself.  Measured per layer at batch 64 (profiles/r2_pair_vs_single_per_layer.txt): the
// weight stream is NOT what bounds the single-CTA kernel (xbar2l1tex reads 20-27 B/clk/SM on the 3x3 layers, half the
// chip's L2 -> SM limit), so halving it buys nothing there and the cluster launch costs 1-2 us.  What pairs do buy is
// shared memory: with half-size weight slots the residual layers of the 64-channel bottlenecks keep resident weights, the
// residual ring AND two sub-tiles per weight tile (m = 2; m = 1 alone costs 33 % at N = 64): 71 -> 60 us.
static bool pair_rule(const tod_conv_desc* d) {
  return d->ksize == 3 && d->stride == 1 && d->d_residual != nullptr && d->act == TOD_ACT_SILU && d->cin == 64 && d->cout == 64;
}

int conv_halo_launch(const tod_conv_desc* d, void* stream, const tod_head_fuse_desc* fuse) {
  static PerDeviceOnce attr_once;   // the attribute is per device
  int rc;
  if (attr_once.needed()) {
    for (int i = 0; i < kHaloVariants - 2; ++i) {   // (the tail variants set their own, smaller limit)
      if ((rc = check_cuda(cudaFuncSetAttribute(halo_kernel(i), cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemLimit),
                           "cudaFuncSetAttribute(conv_halo_tcgen05)")) != TOD_OK)
        return rc;
      if (halo_kernel_pair(i) != nullptr &&
          (rc = check_cuda(cudaFuncSetAttribute(halo_kernel_pair(i), cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemLimit),
                           "cudaFuncSetAttribute(conv_halo_tcgen05 pair)")) != TOD_OK)
        return rc;
    }
    attr_once.done();
  }
  const bool silu = d->act == TOD_ACT_SILU, f32 = d->out_dtype == TOD_OUT_F32;
  const int extra = d->d_residual != nullptr ? 1 : (d->d_upadd != nullptr ? 2 : 0);
  TOD_CHECK_ARG(!(d->d_residual != nullptr && d->d_upadd != nullptr), "conv: residual and upsample-add are mutually exclusive");
  TOD_CHECK_ARG(!(f32 && extra != 0), "conv: residual / upsample-add need bf16 output");
  int kvar = f32 ? (silu ? 7 : 6) : ((silu ? 3 : 0) + extra);
  if (fuse != nullptr) {
    TOD_CHECK_ARG(d->ksize == 1 && !silu && extra == 0, "fused head decode: needs a bare 1x1 conv (no activation, no extra operand)");
    TOD_CHECK_ARG(fuse->mode == TOD_FUSE_BOX || fuse->mode == TOD_FUSE_CLS, "fused head decode: bad mode %d", fuse->mode);
    TOD_CHECK_ARG(fuse->mode != TOD_FUSE_BOX || (d->cout == 64 && fuse->d_cand_box != nullptr &&
                                                 (reinterpret_cast<uintptr_t>(fuse->d_cand_box) & 15) == 0),
                  "fused box decode: needs cout 64 and a 16-byte aligned cand_box");
    TOD_CHECK_ARG(fuse->mode != TOD_FUSE_CLS || (fuse->nc > 0 && fuse->nc <= d->cout && d->cout <= 256 &&
                                                 fuse->d_cand_conf != nullptr && fuse->d_cand_cls != nullptr),
                  "fused class decode: needs 0 < nc <= cout <= 256 and cand_conf / cand_cls");
    TOD_CHECK_ARG(fuse->anchors > 0 && fuse->level_off >= 0 && fuse->level_off + d->hin * d->win <= fuse->anchors &&
                      fuse->in_h > 0 && fuse->in_w > 0 && fuse->stride > 0.0f,
                  "fused head decode: bad anchor geometry");
    kvar = fuse->mode == TOD_FUSE_BOX ? 8 : 9;
  }

  // CTA pairs (cta_group::2): forced by the descriptor (tools / tests), else by the plan's rule -- the layers whose weight
  // stream is what bounds them: 3x3 convs with >= 128 output channels (every stride) and 1x1 convs that cannot keep
  // their weights resident
  bool pair = false;
  if (fuse == nullptr && !(d->flags & TOD_CONV_PAIR_OFF) && pair_mode() != 0 && num_sms() >= 2) {
    if ((d->flags & TOD_CONV_PAIR_ON) || pair_mode() == 2) pair = true;
    else if (pair_mode() == 3) pair = pair_rule(d) || (d->ksize == 3 && d->stride == 1 && d->cin >= 128 && d->cout >= 128);   // A/B
    else pair = pair_rule(d);
  }
  HaloParams p;
  size_t smem = 0;
  bool fits = false;
  const int bk0 = pick_block_k(d->cin, d->block_k);
  const int cin_pad = round_up(d->cin, bk0);
  if (extra == 1 && silu && fuse == nullptr && d->reserved[2] != 1) {
    // residual through the shared-memory TMA ring when the weights can stay resident next to it
    if ((rc = build_params(d, bk0, cin_pad, p, &smem, &fits, 1, false, nullptr, pair)) != TOD_OK) return rc;
    if (fits) kvar = 10;
  }
  if (extra == 2 && silu && fuse == nullptr && d->ksize == 1 && d->stride == 1 && d->cout % 64 == 0 && d->reserved[2] != 1) {
    // upsample-add operand through the ring (1x1 conv on pixel tiles)
    if ((rc = build_params(d, bk0, cin_pad, p, &smem, &fits, 2, false, nullptr, pair)) != TOD_OK) return rc;
    if (fits) kvar = 11;
  }
  if (pair && halo_kernel_pair(kvar) == nullptr) {   // no pair instantiation of this flavour
    pair = false;
    fits = false;
  }
  double cost = 0.0;
  for (int bk = bk0; bk >= 16 && !fits; bk >>= 1)
    if ((rc = build_params(d, bk, cin_pad, p, &smem, &fits, 0, false, &cost, pair)) != TOD_OK) return rc;
  TOD_CHECK_ARG(fits, "conv: no shared-memory plan fits (cin %d cout %d ksize %d stride %d)", d->cin, d->cout, d->ksize,
                d->stride);
  if (kvar != 10 && kvar != 11 && fuse == nullptr && d->ksize == 3 && d->cout > 128 && d->reserved[3] == 0) {
    // wide 3x3 layers: two N tiles of <= 128 columns let TWO sub-tiles share every streamed weight tile (m = 2 within the
    // 512 TMEM columns) at the price of reading the activations twice -- taken when the plan's L2 -> SM bytes per output
    // pixel are lower (256 -> 256 3x3 @20x20: 774 vs 1272 KB per 128 pixels; measured 56 vs 73 us)
    tod_conv_desc d2 = *d;
    d2.reserved[3] = 128;
    HaloParams p2;
    size_t smem2 = 0;
    bool fits2 = false;
    double cost2 = 0.0;
    if ((rc = build_params(&d2, p.block_k, cin_pad, p2, &smem2, &fits2, 0, false, &cost2, pair)) != TOD_OK) return rc;
    if (fits2 && cost2 < cost * 0.9) {
      p = p2;
      smem = smem2;
    }
  }

  if (fuse != nullptr) {
    TOD_CHECK_ARG(p.n_tiles == 1 && !p.patch_mode, "fused head decode: unexpected tiling");
    p.dec_nc = fuse->nc;
    p.dec_level_off = fuse->level_off;
    p.dec_anchors = fuse->anchors;
    p.dec_stride = fuse->stride;
    p.dec_in_w = static_cast<float>(fuse->in_w);
    p.dec_in_h = static_cast<float>(fuse->in_h);
    p.cand_box = fuse->d_cand_box;
    p.cand_conf = fuse->d_cand_conf;
    p.cand_cls = fuse->d_cand_cls;
  }
  const long long work = static_cast<long long>(p.num_super) * p.n_tiles;
  long long grid = pair ? num_sms() / 2 : num_sms();   // CTAs, or CTA pairs
  if (grid > work) grid = work;
  grid -= grid % p.n_tiles;
  if (grid < p.n_tiles) grid = p.n_tiles;
  if (pair) {
    if ((rc = check_cuda(launch_pdl_pair(halo_kernel_pair(kvar), dim3(static_cast<unsigned>(2 * grid)), dim3(kHaloThreads), smem,
                                         static_cast<cudaStream_t>(stream), p),
                         "conv_halo_tcgen05 (CTA pairs) launch")) != TOD_OK)
      return rc;
    return TOD_OK;
  }
  HaloKernel kern = halo_kernel(kvar);
  if ((rc = check_cuda(launch_pdl(kern, dim3(static_cast<unsigned>(grid)), dim3(kHaloThreads), smem,
                                  static_cast<cudaStream_t>(stream), p),
                       "conv_halo_tcgen05 launch")) != TOD_OK)
    return rc;
  return TOD_OK;
}

// Conv (64 output channels, SiLU) + a 1x1 conv 64 -> 64 (SiLU) on its output, the intermediate kept on chip (EXTRA 7).
int conv_halo_launch_tail(const tod_conv_desc* d, const tod_conv_tail_desc* t, void* stream, const tod_head_fuse_desc* fuse) {
  static PerDeviceOnce attr_once;
  int rc;
  if (attr_once.needed()) {
    for (int i = 12; i <= 13; ++i)
      if ((rc = check_cuda(cudaFuncSetAttribute(halo_kernel(i), cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemLimit - 1024),
                           "cudaFuncSetAttribute(conv_halo_tcgen05 tail)")) != TOD_OK)
        return rc;
    attr_once.done();
  }
  TOD_CHECK_ARG(t != nullptr && t->d_w2 != nullptr && (t->d_out2 != nullptr || fuse != nullptr), "conv tail: null pointer");
  TOD_CHECK_ARG(d->cout == 64 && t->cout2 == 64, "conv tail: needs 64 -> 64 (got %d -> %d)", d->cout, t->cout2);
  TOD_CHECK_ARG(d->act == TOD_ACT_SILU && t->act2 == (fuse ? TOD_ACT_NONE : TOD_ACT_SILU),
                "conv tail: SiLU conv, then a SiLU 1x1 conv (or the bare box-logit conv when decoding)");
  TOD_CHECK_ARG(d->d_residual == nullptr && d->d_upadd == nullptr && d->out_dtype == TOD_OUT_BF16, "conv tail: plain bf16 conv first");
  TOD_CHECK_ARG((reinterpret_cast<uintptr_t>(t->d_w2) & 15) == 0, "conv tail: weight alignment");
  tod_conv_desc c = *d;          // the tensor map of the store describes the TAIL's output
  if (fuse != nullptr) {         // nothing is stored: candidates only
    TOD_CHECK_ARG(fuse->mode == TOD_FUSE_BOX && fuse->d_cand_box != nullptr && (reinterpret_cast<uintptr_t>(fuse->d_cand_box) & 15) == 0 &&
                      d->ksize == 3 && d->stride == 1 && fuse->anchors > 0 && fuse->level_off >= 0 &&
                      fuse->level_off + d->hin * d->win <= fuse->anchors && fuse->in_h > 0 && fuse->in_w > 0 && fuse->stride > 0.0f,
                  "conv tail + box decode: bad descriptor");
    c.d_out = const_cast<void*>(d->d_x);
    c.out_pitch = d->x_pitch >= 64 ? d->x_pitch : 64;
  } else {
    TOD_CHECK_ARG(t->out2_pitch >= 64 && t->out2_pitch % 8 == 0 && (reinterpret_cast<uintptr_t>(t->d_out2) & 15) == 0,
                  "conv tail: output pitch / alignment");
    c.d_out = t->d_out2;
    c.out_pitch = t->out2_pitch;
  }
  HaloParams p;
  size_t smem = 0;
  bool fits = false;
  const int bk0 = pick_block_k(c.cin, c.block_k);
  const int cin_pad = round_up(c.cin, bk0);
  for (int bk = bk0; bk >= 16 && !fits; bk >>= 1)
    if ((rc = build_params(&c, bk, cin_pad, p, &smem, &fits, 0, true)) != TOD_OK) return rc;
  TOD_CHECK_ARG(fits, "conv tail: no shared-memory plan fits");
  TOD_CHECK_ARG(p.n_tiles == 1 && p.block_n == 64 && p.pc == 64, "conv tail: unexpected tiling");
  {
    const uint64_t dims[2] = {64, 64};
    const uint64_t str[1] = {128};
    const uint32_t box[2] = {64, 64};
    if ((rc = encode_map(&p.tm_w2, t->d_w2, 2, dims, str, box, CU_TENSOR_MAP_SWIZZLE_128B)) != TOD_OK) return rc;
  }
  p.bias2 = t->d_bias2;
  p.hi_w2 = desc_hi(8 * 128, 64);
  p.hi_stage = desc_hi(8 * 128, 64);
  p.idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(64 >> 3) << 17) | ((128u >> 4) << 24);
  if (fuse != nullptr) {
    TOD_CHECK_ARG(p.patch_mode, "conv tail + box decode: unexpected tiling");
    p.dec_nc = fuse->nc;
    p.dec_level_off = fuse->level_off;
    p.dec_anchors = fuse->anchors;
    p.dec_stride = fuse->stride;
    p.dec_in_w = static_cast<float>(fuse->in_w);
    p.dec_in_h = static_cast<float>(fuse->in_h);
    p.cand_box = fuse->d_cand_box;
  }
  const long long work = p.num_super;
  long long grid = num_sms();
  if (grid > work) grid = work;
  return check_cuda(launch_pdl(halo_kernel(fuse != nullptr ? 13 : 12), dim3(static_cast<unsigned>(grid)), dim3(kHaloThreadsTail), smem,
                               static_cast<cudaStream_t>(stream), p),
                    "conv_halo_tcgen05 (tail) launch");
}

}  // namespace tod

// Tools only: d_buf (>= 16 * 8 bytes per CTA, 148 CTAs) receives per-CTA wait-cycle counters of the next conv launches:
// [0..2] producer total / wait A-empty / wait B-empty, [3..6] MMA total / wait TMEM-empty / wait A-full / wait B-full,
// [8..11] epilogue group 0 total / wait TMEM-full / wait staging free / wait panel written, [12..15] group 1.
extern "C" int tod_debug_set_conv_profile(void* d_buf) {
  tod::g_prof = reinterpret_cast<unsigned long long*>(d_buf);
  return TOD_OK;
}
