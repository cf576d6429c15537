// Shared helpers for libtod.so (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "tod.h"

namespace tod {

void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

#define TOD_CHECK_ARG(cond, ...)        \
  do {                                  \
    if (!(cond)) {                      \
      ::tod::set_error(__VA_ARGS__);    \
      return TOD_ERR_INVALID;           \
    }                                   \
  } while (0)

#define TOD_CHECK_LAUNCH(what)                                  \
  do {                                                          \
    int _rc = ::tod::check_cuda(cudaGetLastError(), what);      \
    if (_rc != TOD_OK) return _rc;                              \
  } while (0)

// cudaFuncSetAttribute and the SM count are per DEVICE, and one process may drive several GPUs (BaseModel.engine(device=)):
// one-time setup is therefore keyed by the current device.  Returns true until mark_device_ready() has been called for it;
// two threads racing through the (idempotent) setup is harmless.
inline int current_device_index() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev > 63) dev = 0;
  return dev;
}
struct PerDeviceOnce {
  unsigned long long ready = 0;   // bit per device ordinal
  bool needed() const { return ((__atomic_load_n(&ready, __ATOMIC_ACQUIRE) >> current_device_index()) & 1ull) == 0; }
  void done() { __atomic_fetch_or(&ready, 1ull << current_device_index(), __ATOMIC_RELEASE); }
};

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int round_up(int a, int b) { return ceil_div(a, b) * b; }

// ---------------------------------------------------------------------------------- timeline (tools only)
// tod_debug_set_timeline(d_buf): every instrumented kernel launched afterwards appends one record per CTA --
// {launch id | blockIdx << 32, smid, globaltimer at CTA start, at CTA end, after the programmatic-launch wait, at the
// first complete accumulator} -- to d_buf (header: cursor,
// capacity; then 12 x u64 records).  The tag is baked into a launch at enqueue / capture time, so a CUDA graph captured
// while the timeline is set keeps recording on every replay: that is how tools/timeline.py sees which kernels overlap
// INSIDE the captured graph (ncu serialises kernels, and nsys is not in this image).  Null tag = no code executed.
struct TimelineTag {
  unsigned long long* buf;
  int id;
};
TimelineTag timeline_tag(const char* name);   // capi.cu: current buffer + a fresh launch id (name kept host-side)

// ---------------------------------------------------------------------------------- PTX wrappers
#ifdef __CUDACC__

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// t_ready: the programmatic-launch wait returned (0 if the kernel has none); t_acc: the first accumulator was complete;
// t_mma_end: the MMA role issued its last commit; t_last_acc: the last accumulator was complete (what follows is drain)
// mma_cycles / wait_operand_cycles / wait_acc_cycles: the MMA role's lifetime and the cycles it spent waiting for operands
// (A / B full barriers) and for a free accumulator, so that stalls can be attributed INSIDE the captured graph
__device__ __forceinline__ void timeline_write(const TimelineTag& tag, unsigned long long t0, unsigned long long t_ready = 0,
                                               unsigned long long t_acc = 0, unsigned long long t_mma_end = 0,
                                               unsigned long long t_last_acc = 0, unsigned long long mma_cycles = 0,
                                               unsigned long long wait_operand_cycles = 0, unsigned long long wait_acc_cycles = 0) {
  if (tag.buf == nullptr) return;
  unsigned smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const unsigned long long slot = atomicAdd(tag.buf, 1ull);
  if (slot < tag.buf[1]) {
    unsigned long long* r = tag.buf + 2 + 12 * slot;
    r[0] = static_cast<unsigned long long>(static_cast<unsigned>(tag.id)) | (static_cast<unsigned long long>(blockIdx.x) << 32);
    r[1] = smid | (static_cast<unsigned long long>(clock64()) << 16);   // low 48 bits of the SM cycle counter at CTA end
    r[2] = t0;
    r[3] = global_timer_ns();
    r[4] = t_ready;
    r[5] = t_acc;
    r[6] = t_mma_end;
    r[7] = t_last_acc;
    r[8] = mma_cycles;
    r[9] = wait_operand_cycles;
    r[10] = wait_acc_cycles;
    r[11] = 0;
  }
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 1023u) == 0 && clock64() - t0 > 4000000000ll) {  // ~2 s at 2 GHz
      printf("tod: mbarrier wait timeout (block %d,%d thread %d parity %u)\n", blockIdx.x, blockIdx.y,
             threadIdx.x, parity);
      __trap();
    }
  }
}

// TMA tiled loads (global -> shared), completion on an mbarrier.
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, uint32_t dst_smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, uint32_t dst_smem, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, uint32_t dst_smem, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
// TMA tiled store (shared -> global), bulk-group completion.  Out-of-bounds parts of the box are clipped.
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src_smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(map)), "r"(src_smem), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all of this thread's bulk groups have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all of this thread's bulk groups have completed (writes performed)
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// named barrier among `count` threads (count % 32 == 0); id 0 is __syncthreads
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// Programmatic dependent launch.  pdl_wait(): every grid this launch depends on has completed and its writes are visible
// to this grid -- one thread suffices when every other thread of the grid reaches dependent data only through
// synchronisation with it (mbarrier chains).  pdl_launch_dependents(): this CTA no longer holds back the launch of the
// next kernel in the stream, whose prologue (barrier init, TMEM alloc, weight loads) then overlaps this grid's tail.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// tcgen05 / TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> f32, single CTA.
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// One lane of a converged warp.  Branching directly on this predicate lets ptxas treat the guarded code as
// single-threaded: tcgen05 / TMA operands are then computed in uniform registers with no per-lane broadcast loop
// (a plain `lane == 0` test costs ~28 SASS instructions per MMA instead of ~3).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// K-unrolled issue: KS consecutive 16-wide K steps of one (A tile, B tile) pair; descriptors are passed as (lo, hi)
// halves so that stepping along K is a 32-bit add of 2 (= 32 bytes >> 4) on the low word.
__device__ __forceinline__ void umma_bf16_k4(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred p, one;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
      "setp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 one, 0, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u32 al, %1, 2;\n\tadd.u32 bl, %3, 2;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, one;\n\t"
      "add.u32 al, %1, 4;\n\tadd.u32 bl, %3, 4;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, one;\n\t"
      "add.u32 al, %1, 6;\n\tadd.u32 bl, %3, 6;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, one;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_k2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred p, one;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
      "setp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 one, 0, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u32 al, %1, 2;\n\tadd.u32 bl, %3, 2;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, one;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_k1(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                             uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first)
      : "memory");
}
// low word of a K-major swizzled smem matrix descriptor: start address >> 4, LBO field = 1 (unused)
__device__ __forceinline__ uint32_t umma_desc_lo(uint32_t smem_addr) { return (1u << 16) | ((smem_addr >> 4) & 0x3FFFu); }
// mbarrier arrives once all tcgen05 ops previously issued by this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// ---- CTA pair (cta_group::2): the two CTAs of a (2,1,1) cluster drive one M = 256 MMA.  A rows [128 r, 128 r + 128) and B
// rows [N/2 r, N/2 r + N/2) live in CTA r's shared memory at the SAME offsets, D rows [128 r, +128) in CTA r's TMEM; only
// the leader (rank 0) issues MMAs and waits on the "full" barriers, which both CTAs' TMA loads signal; its commits arrive
// on the same barrier offset in both CTAs.  Every primitive below is exercised by tools/probe_pair.cu.
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
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// shared::cluster address of the LEADER CTA's copy of a barrier, from either CTA of the pair (bit 24 selects the peer)
__device__ __forceinline__ uint32_t pair_leader_addr(uint32_t smem_addr) { return smem_addr & 0xFEFFFFFFu; }
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst_smem, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const CUtensorMap* map, uint32_t leader_bar, uint32_t dst_smem, int c0, int c1,
                                                 int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst_smem), "l"(reinterpret_cast<uint64_t>(map)), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t local_bar_addr, uint32_t cta) {
  uint32_t remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_bar_addr), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ void umma2_bf16_k4(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred p, one;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
      "setp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 one, 0, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u32 al, %1, 2;\n\tadd.u32 bl, %3, 2;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, one;\n\t"
      "add.u32 al, %1, 4;\n\tadd.u32 bl, %3, 4;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, one;\n\t"
      "add.u32 al, %1, 6;\n\tadd.u32 bl, %3, 6;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, one;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_k2(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred p, one;\n\t.reg .b64 da, db;\n\t.reg .b32 al, bl;\n\t"
      "setp.ne.b32 p, %6, 0;\n\tsetp.eq.u32 one, 0, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t"
      "add.u32 al, %1, 2;\n\tadd.u32 bl, %3, 2;\n\tmov.b64 da, {al, %2};\n\tmov.b64 db, {bl, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, one;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first)
      : "memory");
}
__device__ __forceinline__ void umma2_bf16_k1(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi,
                                              uint32_t idesc, uint32_t acc_first) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
      ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc_first)
      : "memory");
}
// arrives (once all tcgen05 ops previously issued by this thread have completed) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma2_commit_both(uint32_t bar_addr) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar_addr),
               "h"(static_cast<uint16_t>(3))
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 16 consecutive f32 columns -> 16 registers per thread (thread t <-> TMEM lane base+t).
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.0f + __expf(-x)); }
// SiLU from the half argument h = x / 2:  x * sigmoid(x) = h + h * tanh(h).  One SFU op (tanh.approx.f32, max relative
// error 2^-11) and one FMA instead of ex2 + rcp with their range fix-ups: |error| <= |x| * 2.5e-4, below bf16 rounding.
__device__ __forceinline__ float silu_from_half(float h) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// packed f32x2 arithmetic (FFMA2 on sm_100): two lanes per issue slot for the epilogues' bias / SiLU FMAs
__device__ __forceinline__ uint64_t pk2f(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ uint64_t pk2u32(uint32_t a, uint32_t b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ void upk2f(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t ffma2_rn(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

#endif  // __CUDACC__
}  // namespace tod
