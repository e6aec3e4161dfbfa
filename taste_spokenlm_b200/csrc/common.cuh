// Shared device helpers for the sm_100a kernels: mbarrier, TMA, tcgen05/TMEM PTX wrappers, small math.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

// Library flavour (compile time): the 16-bit type of every tensor-core operand the path produces and consumes
// (activations, packed weights, softmax probabilities).  0 = bf16 (libtaste_b200.so, BASELINE config 2's dtype),
// 1 = fp16 (libtaste_b200_f16.so: the reference's own GPU dtype - `torch.cuda.amp.autocast()` defaults to fp16, JES:133,
// JES:336 - with 3 more mantissa bits at the same tensor-core rate).  Accumulators, the residual stream, LayerNorm
// statistics, softmax state, the aggregator output and the RVQ are fp32 in both.  The log-mel DFT always runs on split
// bf16 planes (see logmel.cu), whichever flavour.
#ifndef TASTE_F16
#define TASTE_F16 0
#endif
#include <cuda.h>
#include <stdint.h>

#define TASTE_DEVINL __device__ __forceinline__

namespace taste {

// ------------------------------------------------------------------------------------------------
// shared-memory addresses
// ------------------------------------------------------------------------------------------------
TASTE_DEVINL uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ------------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------------
TASTE_DEVINL void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
TASTE_DEVINL void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
TASTE_DEVINL void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
TASTE_DEVINL bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// A wait that can never complete (a lost TMA transaction, a wrong phase) traps after ~10 s instead of hanging the
// device: the launch then fails with cudaErrorLaunchFailure and the caller sees a non-zero return.
#ifndef TASTE_WAIT_LIMIT_NS
#define TASTE_WAIT_LIMIT_NS 4000000000ull
#endif
TASTE_DEVINL uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
TASTE_DEVINL void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 0xFFFFu) == 0) {               // rare: keep the spin body to a handful of instructions
      const uint64_t t = global_timer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > TASTE_WAIT_LIMIT_NS) __trap();
    }
  }
}
// Same, for roles with slack (TMA producers): backs off between polls so the spin does not take issue slots from
// the compute warps that share the scheduler.
TASTE_DEVINL void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  uint64_t t0 = 0;
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if ((++spins & 0xFFFu) == 0) {
      const uint64_t t = global_timer_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > TASTE_WAIT_LIMIT_NS) __trap();
    }
  }
}
// One lane of a fully converged warp (warp-uniform control flow around it lets the compiler emit tcgen05 / TMA
// instructions once, instead of a per-active-lane ELECT loop as inside an `if (lane == 0)` region).
TASTE_DEVINL bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
TASTE_DEVINL void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
TASTE_DEVINL void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), tile mode, completion on an mbarrier
// ------------------------------------------------------------------------------------------------
TASTE_DEVINL void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
TASTE_DEVINL void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
TASTE_DEVINL void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
TASTE_DEVINL void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// TMA store (shared -> global, bulk async-group completion).  The issuing thread commits a group per store and waits
// for the reads of its earlier groups before the staging buffer is rewritten; out-of-range box rows are clipped.
TASTE_DEVINL void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
TASTE_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
TASTE_DEVINL void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
TASTE_DEVINL void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------------
TASTE_DEVINL void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
TASTE_DEVINL void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
TASTE_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
TASTE_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
TASTE_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16/f16 inputs, fp32 accumulate.  One thread issues for the CTA.
TASTE_DEVINL void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
TASTE_DEVINL void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Make an mbarrier track completion of all prior tcgen05 ops of this thread (implies fence::before_thread_sync).
TASTE_DEVINL void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// CTA pairs (cta_group::2): cluster rank, peer addresses, remote barrier arrives, 2-SM TMA / MMA / commit
// ------------------------------------------------------------------------------------------------
TASTE_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
TASTE_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `smem_addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
TASTE_DEVINL uint32_t mapa_shared(uint32_t smem_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
  return r;
}
TASTE_DEVINL void mbar_expect_tx_cluster(uint32_t cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(cluster_addr), "r"(bytes)
               : "memory");
}
TASTE_DEVINL void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion bytes are signalled on a barrier of the PAIR LEADER (cluster address)
TASTE_DEVINL void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
TASTE_DEVINL void tma_load_4d_2sm(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2,
                                  int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
TASTE_DEVINL void tmem_alloc_2sm(uint32_t* smem_dst, uint32_t ncols) {   // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
TASTE_DEVINL void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
TASTE_DEVINL void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[smem, 128 rows per CTA] * B[smem, N/2 rows per CTA]; issued by the leader CTA only
TASTE_DEVINL void umma_ss_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this smem offset in every CTA of `cta_mask` once all prior MMAs of this thread retire
TASTE_DEVINL void umma_commit_2sm(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// K-major, 128-byte-swizzled operand tile (rows of 64 bf16 = 128 B, 8-row swizzle atoms 1024 B apart).
// Field layout follows the sm_100 shared-memory matrix descriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout type [61,64) (2 = SWIZZLE_128B).
TASTE_DEVINL uint64_t umma_desc_k_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>(1) << 16;                 // LBO (unused for swizzled K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;         // SBO: 8 rows * 128 B
  d |= static_cast<uint64_t>(1) << 46;                 // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                 // SWIZZLE_128B
  return d;
}
// MN-major, 128-byte-swizzled operand tile: rows = K index, each row holds 64 contiguous MN elements (128 B);
// 8-row (8 K) swizzle atoms 1024 B apart (SBO); LBO = stride between 64-wide MN atoms (unused when MN extent is 64).
TASTE_DEVINL uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
// Instruction descriptor, kind::f16: D=f32, A/B = bf16 (1) or f16 (0).  a_mn/b_mn: 1 = MN-major operand.
constexpr int kActBf16 = TASTE_F16 ? 0 : 1;
__host__ __device__ constexpr uint32_t umma_idesc(int m, int n, int ab_fmt_bf16, int a_mn, int b_mn) {
  return (1u << 4) | (static_cast<uint32_t>(ab_fmt_bf16) << 7) | (static_cast<uint32_t>(ab_fmt_bf16) << 10) |
         (static_cast<uint32_t>(a_mn) << 15) | (static_cast<uint32_t>(b_mn) << 16) |
         (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}

// TMEM -> registers: this thread's lane, 32 consecutive 32-bit columns.
TASTE_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// registers -> TMEM: this thread's lane, 32 consecutive 32-bit columns.
TASTE_DEVINL void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
TASTE_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
// named barrier among `count` threads (a multiple of 32) of the CTA; id 1..15 (0 is __syncthreads)
TASTE_DEVINL void named_bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
TASTE_DEVINL void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
TASTE_DEVINL void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// math
// ------------------------------------------------------------------------------------------------
TASTE_DEVINL float fast_rcp(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
TASTE_DEVINL float fast_exp2_(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// erf to ~2e-7 absolute (Abramowitz & Stegun 7.1.26): two MUFU (rcp, ex2) + 9 FMA-pipe ops.  Far below bf16 resolution.
TASTE_DEVINL float fast_erf(float x) {
  const float ax = fabsf(x);
  const float t = fast_rcp(fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  const float e = fast_exp2_((ax * -1.4426950408889634f) * ax);
  const float r = fmaf(-p, e, 1.0f);
  return copysignf(r, x);
}
// exact-erf GELU as the reference's ACT2FN['gelu'] (CW:699); `fast` selects the polynomial erf.
template <bool kFast>
TASTE_DEVINL float gelu_erf(float x) {
  const float e = kFast ? fast_erf(x * 0.70710678118654752f) : erff(x * 0.70710678118654752f);
  const float hx = 0.5f * x;
  return fmaf(hx, e, hx);
}

TASTE_DEVINL float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- packed fp32 x 2 arithmetic (FFMA2 / FADD2 on sm_100): halves the FMA-pipe issue slots of elementwise code ----
TASTE_DEVINL uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
TASTE_DEVINL void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
TASTE_DEVINL uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
TASTE_DEVINL uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
TASTE_DEVINL uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
TASTE_DEVINL uint64_t f2_splat(float v) { return f2_pack(v, v); }
// exact-erf GELU (CW:699) of two values with packed FMA-pipe arithmetic: same Abramowitz & Stegun 7.1.26 erf as
// fast_erf (|err| < 2e-7), 12 packed ops + 4 MUFU + 4 logic ops per PAIR instead of 15 ops per element.
TASTE_DEVINL void gelu_erf_pair(float x0, float x1, float& y0, float& y1) {
  const uint64_t x2 = f2_pack(x0, x1);
  float z0, z1;
  f2_unpack(f2_mul(x2, f2_splat(0.70710678118654752f)), z0, z1);
  const uint64_t ax2 = f2_pack(fabsf(z0), fabsf(z1));
  float d0, d1;
  f2_unpack(f2_fma(ax2, f2_splat(0.3275911f), f2_splat(1.0f)), d0, d1);
  const uint64_t t2 = f2_pack(fast_rcp(d0), fast_rcp(d1));
  uint64_t p = f2_fma(f2_splat(-1.061405429f), t2, f2_splat(1.453152027f));      // negated polynomial: r = 1 + p * e
  p = f2_fma(p, t2, f2_splat(-1.421413741f));
  p = f2_fma(p, t2, f2_splat(0.284496736f));
  p = f2_fma(p, t2, f2_splat(-0.254829592f));
  p = f2_mul(p, t2);
  float q0, q1;
  f2_unpack(f2_mul(f2_mul(ax2, f2_splat(-1.4426950408889634f)), ax2), q0, q1);
  const uint64_t e2 = f2_pack(fast_exp2_(q0), fast_exp2_(q1));
  float r0, r1;
  f2_unpack(f2_fma(p, e2, f2_splat(1.0f)), r0, r1);
  const uint64_t erf2 = f2_pack(copysignf(r0, z0), copysignf(r1, z1));
  const uint64_t hx2 = f2_mul(x2, f2_splat(0.5f));
  f2_unpack(f2_fma(hx2, erf2, hx2), y0, y1);
}

// 2^t for two values on the FMA pipe (no MUFU): t = n + f, n = round(t) via the 1.5 * 2^23 magic add, 2^f by a
// degree-3 minimax polynomial on [-0.5, 0.5] (max relative error 7.5e-5, far below bf16 resolution), 2^n by adding n
// to the exponent field.  t is clamped at -126 (results below 2^-126 flush towards 1.2e-38, harmless for softmax).
TASTE_DEVINL void exp2_poly2(uint64_t t2, float& p0, float& p1) {
  float t0, t1;
  f2_unpack(t2, t0, t1);
  t0 = fmaxf(t0, -126.0f);
  t1 = fmaxf(t1, -126.0f);
  const uint64_t t = f2_pack(t0, t1);
  const uint64_t magic = f2_pack(12582912.0f, 12582912.0f);
  const uint64_t nmagic = f2_pack(-12582912.0f, -12582912.0f);
  const uint64_t u = f2_add(t, magic);                        // low mantissa bits = round(t)
  const uint64_t n = f2_add(u, nmagic);
  const uint64_t f = f2_fma(n, f2_pack(-1.0f, -1.0f), t);     // t - n  in [-0.5, 0.5]
  uint64_t p = f2_fma(f, f2_pack(5.517166712e-02f, 5.517166712e-02f), f2_pack(2.426111220e-01f, 2.426111220e-01f));
  p = f2_fma(p, f, f2_pack(6.932609858e-01f, 6.932609858e-01f));
  p = f2_fma(p, f, f2_pack(9.999280736e-01f, 9.999280736e-01f));
  float q0, q1, u0, u1;
  f2_unpack(p, q0, q1);
  f2_unpack(u, u0, u1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(u0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(u1) << 23));
}

// Same polynomial, fed with raw scores: 2^(s * log2e - m) with the argument clamped on BOTH sides at no extra cost.
// t' = sat(s * c1 + c0) with c1 = log2e / 252, c0 = (126 - m) / 252 maps t = s * log2e - m from [-126, 126] onto [0, 1]
// (FFMA.SAT; -inf and NaN give 0), u = 252 t' - 126 + magic rounds t to the nearest integer n in the low mantissa bits,
// f = 252 t' - 126 - n.  Ten instructions per pair, as exp2_poly2 + its scale FFMA2.  The upper clamp makes an argument
// beyond the fp32 range produce >= 2^125 instead of a wrapped exponent, so the caller can detect it in the row sum.
TASTE_DEVINL void exp2_poly2_sat(float s0, float s1, float c1, float c0, float& p0, float& p1) {
  float a0, a1;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(a0) : "f"(s0), "f"(c1), "f"(c0));
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(a1) : "f"(s1), "f"(c1), "f"(c0));
  const uint64_t a = f2_pack(a0, a1);
  const uint64_t k252 = f2_pack(252.0f, 252.0f);
  const uint64_t mg = f2_pack(12582786.0f, 12582786.0f);                 // 1.5 * 2^23 - 126
  const uint64_t u = f2_fma(a, k252, mg);                                // low mantissa bits = round(t)
  const uint64_t nb = f2_fma(u, f2_pack(-1.0f, -1.0f), mg);              // -(n + 126), exact
  const uint64_t f = f2_fma(a, k252, nb);                                // t - n  in [-0.5, 0.5]
  uint64_t p = f2_fma(f, f2_pack(5.517166712e-02f, 5.517166712e-02f), f2_pack(2.426111220e-01f, 2.426111220e-01f));
  p = f2_fma(p, f, f2_pack(6.932609858e-01f, 6.932609858e-01f));
  p = f2_fma(p, f, f2_pack(9.999280736e-01f, 9.999280736e-01f));
  float q0, q1, u0, u1;
  f2_unpack(p, q0, q1);
  f2_unpack(u, u0, u1);
  p0 = __uint_as_float(__float_as_uint(q0) + (__float_as_uint(u0) << 23));
  p1 = __uint_as_float(__float_as_uint(q1) + (__float_as_uint(u1) << 23));
}

TASTE_DEVINL uint32_t pack_true_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
// Two fp32 values -> one packed pair of the library flavour's 16-bit operand type (see kActBf16).
TASTE_DEVINL uint32_t pack_act2(float lo, float hi) {
#if TASTE_F16
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
#else
  return pack_true_bf16x2(lo, hi);
#endif
}

TASTE_DEVINL float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
TASTE_DEVINL float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace taste
