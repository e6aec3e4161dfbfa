// Non-causal flash attention on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 64, bf16 in / out, fp32
// scores and accumulators.  Serves the encoder self-attention (CW:377-394 eager semantics: softmax(q k^T) v with q
// pre-scaled, NO mask — JES:198-203), which is 16 % of the path's FLOPs.
//
// Persistent: one CTA per SM walks over work items of 256 queries of one (utterance, head), as two 128-row tiles:
//   S_i = Q_i K_j^T   tcgen05.mma  SS  (M 128, N 128 keys, K 64)      -> TMEM, 128 fp32 columns per tile
//   softmax           one thread per query row reads its whole S row ONCE with tcgen05.ld (no shuffles) and hands S
//                     back at once, keeps the running max / sum in registers, exponentiates in packed FFMA2 pairs
//                     (MUFU ex2 for 10 of 16 pairs, a degree-3 polynomial on the FMA pipe for the other 6) and
//                     writes P (bf16) back to TMEM with tcgen05.st
//   O_i += P_i V_j    tcgen05.mma  TS  (A = P from TMEM, B = V tile, MN-major; M 128, N 64, K 128 keys)
// Each tile has its own MMA-issuing warp, so the next block's QK^T of a tile runs under that tile's exponentials and
// under the other tile's GEMMs.  Q (double buffered) and K/V (4-stage mbarrier ring) arrive by TMA (128-byte swizzle)
// straight from the packed [rows, 3*D] QKV activation.  The accumulator is rescaled only when a row's maximum grows by
// more than 2^8 (the stale maximum keeps exp2 arguments <= 8, exact in fp32 / bf16 range), so the O read-modify-write
// in TMEM is rare after the first key block.
//
// Roles (384 threads, three warpgroups): warps 0-3 softmax tile 0, warps 4-7 softmax tile 1 (warp w owns TMEM lanes
// 32*(w%4)..+31), warp 8 TMA producer, warp 9 MMA issuer of tile 0 + TMEM allocator, warp 10 MMA issuer of tile 1,
// warp 11 idle (it completes the third warpgroup so that setmaxnreg can move registers: softmax threads 232, the
// rest 40 - the softmax thread holds a 128-score row and schedules its exp2 phase far better with the extra 64).
// The two tiles' exp2 phases are ping-ponged with named barriers: left alone they drift into lock-step (measured with
// the in-kernel timeline), where both warps of a scheduler stall on the MUFU queue and the block period grows by a
// third.
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace taste {

constexpr int FA_BQ = 128;          // query rows per tile (2 tiles per work item)
constexpr int FA_BK = 128;          // keys per block
constexpr int FA_HD = 64;
constexpr int FA_STAGES = 4;
constexpr int FA_THREADS = 384;          // 12 warps: three full warpgroups (setmaxnreg is per warpgroup)
#ifndef FA_POLY
#define FA_POLY 6                   // of every 16 score pairs, this many take the polynomial exp2 path
#endif
// VAR bits: 2 = setmaxnreg (softmax warpgroups 216, others 64), 8 = 232 / 40 instead, 16 = ping-pong of the two
// tiles' exp2 phases, 32 = hand the turn over one chunk early, 4 = timeline trace, 64 = split phases: the MUFU pairs
// first (one tile's warp saturates the XU pipe), the turn is handed over, then the polynomial pairs (FMA pipe) run
// under the other tile's MUFU phase, 128 = rolling prefetch: as soon as a 32-score chunk has been exponentiated its
// registers are refilled with the NEXT block's scores (tcgen05.ld under the exponentials), so the s_full wait and the
// TMEM load latency leave the per-tile serial loop, 256 = eight row-maximum chains instead of four, 512 = MMA issuers
// wait parked (try_wait with a suspend hint) and the producer polls every ~1 us (their polling loops took issue slots
// from the softmax warps on three of the four schedulers)
constexpr int FA_VAR_DEFAULT = 2 | 8 | 16 | 32;
constexpr int FA_VAR_SPLIT = FA_VAR_DEFAULT | 64;
constexpr uint32_t FA_TILE_BYTES = FA_BQ * FA_HD * 2;      // 16 KB: one Q, K or V tile
// Q is double buffered (2 x 2 tiles) so the next work item's queries load under the current item's last blocks
constexpr size_t FA_SMEM = 1024 + size_t(4 + 2 * FA_STAGES) * FA_TILE_BYTES + 256;

// TMEM columns
constexpr uint32_t FA_COL_S = 0;      // S0 [0,128)  S1 [128,256)
constexpr uint32_t FA_COL_O = 256;    // O0 [256,320) O1 [320,384)
constexpr uint32_t FA_COL_P = 384;    // P0 [384,448) P1 [448,512)   (bf16 pairs: 64 columns per 128 keys)

struct FaParams {
  long long* dbg;      // timeline trace (VAR bit 2 only)
  __nv_bfloat16* o;
  int ldo;
  int q_len, kv_len;
  int heads, q_blocks, n_items;      // work item w = (b * heads + head) * q_blocks + qb
};

#define FA_TRACE(slot, idx)                                                      \
  do {                                                                           \
    if ((VAR & 4) && p.dbg && blockIdx.x == 0 && lane == 0 && (idx) < 256)       \
      p.dbg[(slot) * 256 + (idx)] = clock64();                                   \
  } while (0)

// Persistent: one CTA per SM loops over work items (256 queries of one (utterance, head)); TMEM, barriers and the
// K/V ring live across items, the next item's Q / K / V loads and first QK^T run under the current item's tail, so
// the ~4 us of per-CTA set-up and drain measured on the non-persistent version is paid once per launch.
template <int VAR, int POLY>
__global__ void __launch_bounds__(FA_THREADS, 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                         const __grid_constant__ CUtensorMap tma_v, const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;                                   // 2 buffers x 2 tiles
  uint8_t* sKV = smem + 4 * FA_TILE_BYTES;              // FA_STAGES x {K, V}
  uint64_t* bars = reinterpret_cast<uint64_t*>(sKV + size_t(2 * FA_STAGES) * FA_TILE_BYTES);
  uint64_t* q_full = bars;                              // [2]
  uint64_t* q_empty = bars + 2;                         // [2]  all QK^T of the item retired (MMA -> producer)
  uint64_t* kv_full = bars + 4;                         // [FA_STAGES]
  uint64_t* kv_empty = kv_full + FA_STAGES;             // [FA_STAGES]
  uint64_t* s_full = kv_empty + FA_STAGES;              // [2]  S_i ready               (MMA -> softmax)
  uint64_t* s_free = s_full + 2;                        // [2]  S_i read from TMEM      (softmax -> MMA), 4 warp arrivals
  uint64_t* p_full = s_free + 2;                        // [2]  P_i written             (softmax -> MMA), 4 warp arrivals
  uint64_t* o_full = p_full + 2;                        // [2]  P_i V done              (MMA -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blocks = (p.kv_len + FA_BK - 1) / FA_BK;
  const int my_items = (p.n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int total_g = my_items * n_blocks;              // key blocks this CTA walks through, over all its items

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], 2);            // one tcgen05.commit per MMA issuer
    }
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], 2);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run warp-uniform loops and elect one lane around the TMA / tcgen05 instructions (inside an
  // `if (lane == 0)` region every UTCHMMA is wrapped in an ELECT loop: ~90 cycles per MMA, measured with FA_TRACE).
  // VAR bit 1: register reallocation between warpgroups (the softmax threads hold a whole 128-score row + the prefetched
  // chunk of the next block; the TMA / MMA warps need almost nothing).  256 x 216 + 128 x 64 <= 384 x 168.
  // (the instruction sits at the head of each role's branch: ptxas budgets registers per region it dominates)
#define FA_REGS_SMALL() do { if constexpr ((VAR & 2) != 0) { if constexpr ((VAR & 8) != 0) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;"); else asm volatile("setmaxnreg.dec.sync.aligned.u32 64;"); } } while (0)
#define FA_REGS_LARGE() do { if constexpr ((VAR & 2) != 0) { if constexpr ((VAR & 8) != 0) asm volatile("setmaxnreg.inc.sync.aligned.u32 232;"); else asm volatile("setmaxnreg.inc.sync.aligned.u32 216;"); } } while (0)
  if (warp == 11) {
    FA_REGS_SMALL();        // idle: only completes the third warpgroup
  } else if (warp == 8) {
    // ===================== TMA producer =====================
    FA_REGS_SMALL();
    int stage = 0;
    uint32_t phase = 0;
    for (int it = 0; it < my_items; ++it) {
      const int w = int(blockIdx.x) + it * int(gridDim.x);
      const int qb = w % p.q_blocks;
      const int bh = w / p.q_blocks;
      const int head = bh % p.heads;
      const int b = bh / p.heads;
      const int q0 = qb * (2 * FA_BQ);
      const int qbuf = it & 1;
      if constexpr ((VAR & 512) != 0) mbar_wait_parked(&q_empty[qbuf], uint32_t(((it >> 1) & 1) ^ 1));
      else mbar_wait_relaxed(&q_empty[qbuf], uint32_t(((it >> 1) & 1) ^ 1));
      if (elect_one()) {
        uint8_t* sq = sQ + size_t(2 * qbuf) * FA_TILE_BYTES;
        mbar_expect_tx(&q_full[qbuf], 2 * FA_TILE_BYTES);
        tma_load_3d(sq, &tma_q, &q_full[qbuf], head * FA_HD, q0, b);
        tma_load_3d(sq + FA_TILE_BYTES, &tma_q, &q_full[qbuf], head * FA_HD, q0 + FA_BQ, b);
      }
      __syncwarp();
      for (int j = 0; j < n_blocks; ++j) {
        if constexpr ((VAR & 512) != 0) mbar_wait_parked(&kv_empty[stage], phase ^ 1);
        else mbar_wait_relaxed(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = sKV + size_t(2 * stage) * FA_TILE_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&kv_full[stage], 2 * FA_TILE_BYTES);
          tma_load_3d(sk, &tma_k, &kv_full[stage], head * FA_HD, j * FA_BK, b);
          tma_load_3d(sk + FA_TILE_BYTES, &tma_v, &kv_full[stage], head * FA_HD, j * FA_BK, b);
        }
        __syncwarp();
        if (++stage == FA_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 9 || warp == 10) {
    // ===================== MMA issuers: warp 9 drives query tile 0, warp 10 tile 1 =====================
    // One issuing thread per tile, so neither tile's GEMMs ever queue behind a barrier that only the other tile's
    // softmax warps can satisfy (with a single issuer and a fixed wait order the two warpgroups throttled each other).
    FA_REGS_SMALL();
    const int i = warp - 9;
    constexpr uint32_t idesc_qk = umma_idesc(FA_BQ, FA_BK, 1, 0, 0);        // A, B K-major
    constexpr uint32_t idesc_pv = umma_idesc(FA_BQ, FA_HD, 1, 0, 1);        // A from TMEM, B (V) MN-major
    // S_i = Q_i K^T for global block index g (item g / n_blocks, key block g % n_blocks)
    auto issue_qk = [&](int g) {
      const int qbuf = (g / n_blocks) & 1;
      const int stage = g % FA_STAGES;
      const uint64_t da = umma_desc_k_sw128(smem_u32(sQ + size_t(2 * qbuf + i) * FA_TILE_BYTES));
      const uint64_t db = umma_desc_k_sw128(smem_u32(sKV + size_t(2 * stage) * FA_TILE_BYTES));
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FA_HD / 16; ++k)
          umma_ss(tmem_base + FA_COL_S + uint32_t(i * FA_BK), da + uint64_t(k * 2), db + uint64_t(k * 2), idesc_qk,
                  k != 0 ? 1u : 0u);
        umma_commit(&s_full[i]);
      }
      __syncwarp();
    };
    // everything block g needs from the producer: its K/V stage and, on the first block of an item, the item's Q
    auto wait_bar = [&](uint64_t* bar, uint32_t parity) {
      if constexpr ((VAR & 512) != 0) mbar_wait_parked(bar, parity);
      else mbar_wait(bar, parity);
    };
    auto wait_inputs = [&](int g) {
      const int it = g / n_blocks;
      if (g % n_blocks == 0) wait_bar(&q_full[it & 1], uint32_t((it >> 1) & 1));
      wait_bar(&kv_full[g % FA_STAGES], uint32_t((g / FA_STAGES) & 1));
      tc_fence_after();
    };
    if (total_g > 0) {
      wait_inputs(0);
      // tile 1 starts half a block behind tile 0 (when tile 0's first S tile has been read), so that one warpgroup
      // is in its exp2-heavy pass while the other reads / reduces scores instead of both hitting the MUFU together
      if (i == 1) wait_bar(&s_free[0], 0);
      issue_qk(0);
    }
    for (int g = 0; g < total_g; ++g) {
      const int j = g % n_blocks;
      const int stage = g % FA_STAGES;
      const bool more = g + 1 < total_g;
      const uint64_t dv = umma_desc_mn_sw128(smem_u32(sKV + size_t(2 * stage + 1) * FA_TILE_BYTES), 0);
      if (more) wait_inputs(g + 1);
      wait_bar(&s_free[i], uint32_t(g & 1));
      FA_TRACE(2 + i, g * 8 + 0);
      tc_fence_after();
      if (more) issue_qk(g + 1);          // next block's scores first: the softmax warps wait on these
      FA_TRACE(2 + i, g * 8 + 1);
      wait_bar(&p_full[i], uint32_t(g & 1));
      FA_TRACE(2 + i, g * 8 + 2);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FA_BK / 16; ++k)     // 16 keys per MMA: 8 TMEM columns of P, 2048 B of V
          umma_ts(tmem_base + FA_COL_O + uint32_t(i * FA_HD), tmem_base + FA_COL_P + uint32_t(i * 64 + k * 8),
                  dv + uint64_t(k * (2048 >> 4)), idesc_pv, (j | k) != 0 ? 1u : 0u);
        umma_commit(&o_full[i]);
        umma_commit(&kv_empty[stage]);                                   // needs both tiles' commits
        if (j == n_blocks - 1) umma_commit(&q_empty[(g / n_blocks) & 1]);   // likewise
      }
      __syncwarp();
      FA_TRACE(2 + i, g * 8 + 3);
    }
  } else {
    // ===================== softmax + output (warps 0-7) =====================
    FA_REGS_LARGE();
    const int i = warp >> 2;                        // query tile
    const int q = warp & 3;                         // TMEM lane quarter
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t t_s = lane_base + FA_COL_S + uint32_t(i * FA_BK);
    const uint32_t t_o = lane_base + FA_COL_O + uint32_t(i * FA_HD);
    const uint32_t t_p = lane_base + FA_COL_P + uint32_t(i * 64);
    const float kLog2e = 1.4426950408889634f;
    if constexpr ((VAR & 16) != 0) {
      if (i == 1 && total_g > 0) asm volatile("bar.arrive 1, 256;" ::: "memory");      // tile 0 goes first
    }
    int g = 0;
    uint32_t r[4][32];       // one whole S row (128 scores)
    bool pending = false;    // P of the previous block written but not yet signalled
    if constexpr ((VAR & 128) != 0) {
      if (total_g > 0) {
        mbar_wait(&s_full[i], 0);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32b_x32(t_s + uint32_t(c * 32), r[c]);
      }
    }
    for (int it = 0; it < my_items; ++it) {
      const int w = int(blockIdx.x) + it * int(gridDim.x);
      const int qb = w % p.q_blocks;
      const int bh = w / p.q_blocks;
      const int head = bh % p.heads;
      const int b = bh / p.heads;
      const int row = qb * (2 * FA_BQ) + i * FA_BQ + q * 32 + lane;   // query index within the utterance
      float m_used = -INFINITY;      // stale running maximum (raw score units)
      float l_run = 0.f;

      for (int j = 0; j < n_blocks; ++j, ++g) {
        const int valid = p.kv_len - j * FA_BK;       // keys of this block that exist
        if constexpr ((VAR & 128) != 0) {
          // rolling prefetch: this block's scores were requested chunk by chunk under the previous block's
          // exponentials (or by the prologue); P of the previous block is signalled together with the S hand-back
          FA_TRACE(i, g * 8 + 0);
          tmem_ld_wait();
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (pending) mbar_arrive(&p_full[i]);
            mbar_arrive(&s_free[i]);
          }
          pending = false;
          FA_TRACE(i, g * 8 + 1);
        } else {
        FA_TRACE(i, g * 8 + 0);
        mbar_wait(&s_full[i], uint32_t(g & 1));
        FA_TRACE(i, g * 8 + 1);
        tc_fence_after();
        // The whole S row (128 fp32) is read into registers ONCE: one exposed TMEM latency per block, and the S tile
        // can be handed back to the MMA warp immediately, so the next block's QK^T runs under this block's entire
        // softmax.  P goes back to TMEM in four 16-column stores as soon as each 32-score chunk is exponentiated,
        // which keeps the live set at 128 scores + 16 packed probabilities.
        // (Tried and rejected, with measurements in DESIGN.md: two passes over TMEM with 64 live scores; a speculative
        // single pass against the stale maximum; 16 softmax warps with two threads per row.)
        tmem_ld_32x32b_x32(t_s, r[0]);
        tmem_ld_32x32b_x32(t_s + 32, r[1]);
        if (pending) {
          // P of the previous block: its stores are waited for only now, under the latency of the loads above
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[i]);
          pending = false;
        }
        tmem_ld_32x32b_x32(t_s + 64, r[2]);
        tmem_ld_32x32b_x32(t_s + 96, r[3]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[i]);       // every read of S has landed: the next QK^T may overwrite it
        }
        FA_TRACE(i, g * 8 + 3);
        if (valid < FA_BK) {
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int e = 0; e < 32; ++e)
              if (c * 32 + e >= valid) r[c][e] = 0xff800000u;      // -inf
        }
        // four independent chains (a single chain of 64 dependent FMNMX3 costs ~400 cycles per block)
        float mxc[8] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            constexpr int kHalf = (VAR & 256) ? 16 : 32;      // 8 chains: each chunk's halves reduce separately
            const int ch = 2 * c + (e / kHalf) % 2;
            mxc[ch] = fmaxf(mxc[ch], __uint_as_float(r[c][e]));
          }
        const float mx = fmaxf(fmaxf(fmaxf(mxc[0], mxc[1]), fmaxf(mxc[2], mxc[3])),
                               fmaxf(fmaxf(mxc[4], mxc[5]), fmaxf(mxc[6], mxc[7])));
        FA_TRACE(i, g * 8 + 2);
        const bool grow = mx > m_used + 5.545177f;     // 8 in log2 units; first block: m_used = -inf
        const bool any_grow = __any_sync(0xffffffffu, grow);
        float alpha = 1.0f;
        if (any_grow) {
          const float m_new = grow ? mx : m_used;
          alpha = (m_used == -INFINITY) ? 0.f : fast_exp2((m_used - m_new) * kLog2e);
          l_run *= alpha;
          m_used = m_new;
        }
        const float neg_m = -m_used * kLog2e;
        const uint64_t negm2 = f2_pack(neg_m, neg_m);
        const uint64_t log2e2 = f2_pack(kLog2e, kLog2e);
        uint64_t sum2 = f2_pack(0.f, 0.f);
        // VAR bit 4: ping-pong.  The two tiles' exp2 phases alternate (named barriers 1 / 2, FA3-style) instead of
        // drifting into lock-step, where both warps of a scheduler fight for the MUFU queue and neither feeds the FMA
        // pipe; the other tile's TMEM loads, row maximum and waits run under this tile's exponentials.
        if constexpr ((VAR & 16) != 0) {
          if (i == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
          else asm volatile("bar.sync 2, 256;" ::: "memory");
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
          // Exponentials in pairs (packed FFMA2 / FADD2).  16/clk/SM of MUFU.EX2 would cap the tensor pipe at 50 %,
          // so POLY of every 16 pairs are evaluated on the FMA pipe instead (exp2_poly2).
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            constexpr int kMuPairs = 64 - 4 * POLY;           // split phases: pairs [0, kMuPairs) on the MUFU
            if constexpr ((VAR & 64) != 0) {
              if (c * 16 + e == kMuPairs) {                   // MUFU phase over: the other tile's turn
                if (i == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
                else if (g + 1 < total_g) asm volatile("bar.arrive 1, 256;" ::: "memory");
              }
            }
            const uint64_t t2 = f2_fma(f2_pack(__uint_as_float(r[c][2 * e]), __uint_as_float(r[c][2 * e + 1])), log2e2, negm2);
            float p0, p1;
            const bool use_poly = (VAR & 64) ? (c * 16 + e >= kMuPairs) : (((e * POLY) & 15) < POLY && POLY > 0);
            if (use_poly) {       // evenly spread POLY of 16 (split phases: the last 4 * POLY pairs)
              exp2_poly2(t2, p0, p1);
            } else {
              f2_unpack(t2, p0, p1);
              p0 = fast_exp2(p0);
              p1 = fast_exp2(p1);
            }
            sum2 = f2_add(sum2, f2_pack(p0, p1));
            pk[e] = pack_bf16x2(p0, p1);
          }
          if (c == 0 && j > 0) {
            // Only now is the previous block's P V needed: P_i has been consumed (it may be overwritten) and O_i is
            // stable (it may be rescaled).  On the first block of an item the output pass below already waited.
            FA_TRACE(i, g * 8 + 4);
            mbar_wait(&o_full[i], uint32_t((g - 1) & 1));
            FA_TRACE(i, g * 8 + 5);
            tc_fence_after();
            if (any_grow) {
#pragma unroll
              for (int cc = 0; cc < FA_HD / 16; ++cc) {
                uint32_t o[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
                    "%13, %14, %15}, [%16];"
                    : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7]),
                      "=r"(o[8]), "=r"(o[9]), "=r"(o[10]), "=r"(o[11]), "=r"(o[12]), "=r"(o[13]), "=r"(o[14]),
                      "=r"(o[15])
                    : "r"(t_o + uint32_t(cc * 16))
                    : "memory");
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 16; ++e) o[e] = __float_as_uint(__uint_as_float(o[e]) * alpha);
                tmem_st_32x32b_x16(t_o + uint32_t(cc * 16), o);
              }
            }
          }
          tmem_st_32x32b_x16(t_p + uint32_t(c * 16), pk);
          if constexpr ((VAR & 128) != 0) {
            // chunk c's score registers are dead: refill them with the next block's scores (the next QK^T was issued
            // when this block's S was handed back, a whole max + exp2 chunk ago)
            if (g + 1 < total_g && c >= 1) {
              if (c == 1) {
                mbar_wait(&s_full[i], uint32_t((g + 1) & 1));
                tc_fence_after();
                tmem_ld_32x32b_x32(t_s, r[0]);
              }
              tmem_ld_32x32b_x32(t_s + uint32_t(c * 32), r[c]);
            }
          }
          if constexpr ((VAR & 16) != 0 && ((VAR & 64) == 0 || POLY == 0)) {
            if (c == ((VAR & 32) ? 2 : 3)) {            // hand the turn over (bit 5: one chunk early)
              if (i == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
              else if (g + 1 < total_g) asm volatile("bar.arrive 1, 256;" ::: "memory");
            }
          }
        }
        float sum0, sum1;
        f2_unpack(sum2, sum0, sum1);
        l_run += sum0 + sum1;
        if (j + 1 < n_blocks) {
          pending = true;              // the P stores are waited for (and p_full signalled) under the next block's loads
        } else {
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[i]);
        }
        FA_TRACE(i, g * 8 + 6);
      }

      // ---- output: O_i / l ----
      mbar_wait(&o_full[i], uint32_t((g - 1) & 1));
      tc_fence_after();
      const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;
      __nv_bfloat16* orow = p.o + (int64_t(b) * p.q_len + row) * p.ldo + head * FA_HD;
#pragma unroll
      for (int c = 0; c < FA_HD / 32; ++c) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(t_o + uint32_t(c * 32), r);
        tmem_ld_wait();
        if (row < p.q_len) {
          uint4* dst = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint4 u;
            u.x = pack_bf16x2(__uint_as_float(r[8 * e + 0]) * inv, __uint_as_float(r[8 * e + 1]) * inv);
            u.y = pack_bf16x2(__uint_as_float(r[8 * e + 2]) * inv, __uint_as_float(r[8 * e + 3]) * inv);
            u.z = pack_bf16x2(__uint_as_float(r[8 * e + 4]) * inv, __uint_as_float(r[8 * e + 5]) * inv);
            u.w = pack_bf16x2(__uint_as_float(r[8 * e + 6]) * inv, __uint_as_float(r[8 * e + 7]) * inv);
            dst[e] = u;
          }
        }
      }
      tc_fence_before();      // O_i has been read: the next item's first P V (accumulate = 0) may overwrite it; that MMA
                              // is issued only after this warpgroup's next p_full arrive, which follows in program order
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// {64 * heads columns, rows, batch} view of a [batch * rows, ld] bf16 activation; box = 64 x 128 x 1
static int make_map(EncodeTiledFn enc, CUtensorMap* m, const void* base, int heads, int rows, int batch, int ld) {
  cuuint64_t dims[3] = {(cuuint64_t)heads * FA_HD, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {FA_HD, FA_BQ, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error((int)r, "attention: tensor map encode failed (%d)", (int)r);
  return 0;
}

static long long* g_fa_trace = nullptr;
extern "C" void taste_dbg_attention_trace(void* dev_buf) { g_fa_trace = static_cast<long long*>(dev_buf); }

bool attention_tcgen05_eligible(const AttnDesc& d) {
  if (d.cu_q || d.cu_kv || d.causal) return false;
  if (d.q_len < 2 * FA_BQ || d.kv_len < FA_BK) return false;           // small problems: the mma.sync kernel
  if ((reinterpret_cast<uintptr_t>(d.q) | reinterpret_cast<uintptr_t>(d.k) | reinterpret_cast<uintptr_t>(d.v)) & 15)
    return false;
  if ((d.ldq | d.ldk | d.ldv | d.ldo) % 8 != 0) return false;
  return true;
}

int launch_attention_tcgen05(const AttnDesc& d, cudaStream_t stream) {
  EncodeTiledFn enc = get_tensor_map_encoder();
  if (!enc) return set_error(TASTE_E_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap mq, mk, mv;
  int rc;
  if ((rc = make_map(enc, &mq, d.q, d.heads, d.q_len, d.batch, d.ldq))) return rc;
  if ((rc = make_map(enc, &mk, d.k, d.heads, d.kv_len, d.batch, d.ldk))) return rc;
  if ((rc = make_map(enc, &mv, d.v, d.heads, d.kv_len, d.batch, d.ldv))) return rc;
  // compiled variants (VAR, POLY); FA_VAR_SPLIT with 5 polynomial pairs of 16 is the default, the others are A/B knobs (TASTE_FA_VAR / TASTE_FA_POLY)
#define FA_VARIANTS(X)                                                                                          \
  X(FA_VAR_DEFAULT, FA_POLY) X(FA_VAR_DEFAULT, 0) X(FA_VAR_DEFAULT | 4, FA_POLY) X(0, FA_POLY)                  \
  X(FA_VAR_SPLIT, 5) X(FA_VAR_SPLIT, 6) X(FA_VAR_SPLIT | 512, 5) X(FA_VAR_SPLIT | 128, 5)                      \
  X(FA_VAR_SPLIT | 128 | 256, 5) X(FA_VAR_SPLIT | 128 | 512, 5) X(FA_VAR_SPLIT | 128 | 256 | 512, 5)           \
  X(FA_VAR_SPLIT | 128 | 256 | 512, 6) X(FA_VAR_SPLIT | 128 | 256 | 512, 4) X(FA_VAR_DEFAULT | 128 | 256 | 512, FA_POLY)
  static bool configured = false;
  if (!configured) {
#define FA_CFG(V, P) TASTE_CUDA_OK(cudaFuncSetAttribute(attention_tcgen05_kernel<(V), (P)>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FA_SMEM));
    FA_VARIANTS(FA_CFG)
#undef FA_CFG
    configured = true;
  }
  const char* ev = getenv("TASTE_FA_VAR");
  const int var = ev ? atoi(ev) : -1;
  FaParams p;
  p.dbg = g_fa_trace;
  p.o = static_cast<__nv_bfloat16*>(d.o);
  p.ldo = d.ldo;
  p.q_len = d.q_len;
  p.kv_len = d.kv_len;
  p.heads = d.heads;
  p.q_blocks = (d.q_len + 2 * FA_BQ - 1) / (2 * FA_BQ);
  p.n_items = p.q_blocks * d.heads * d.batch;
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (n_sm <= 0) n_sm = 148;
  }
  dim3 grid(p.n_items < n_sm ? p.n_items : n_sm);
  const double pairs = double(d.batch) * d.q_len * d.kv_len;
  ProfScope ps(stream, d.kclass == KC_ATTN_ENC ? KC_ATTN_ENC : KC_ATTN_TC, 4.0 * pairs * FA_HD * d.heads,
               2.0 * FA_HD * d.heads * double(d.batch) * (2.0 * d.q_len + 2.0 * d.kv_len));
  const char* ep = getenv("TASTE_FA_POLY");          // A/B knob: "0" = every exponential on the MUFU
  const int want_var = var >= 0 ? var : FA_VAR_SPLIT;
  const int want_poly = ep ? atoi(ep) : ((want_var & 64) ? 5 : FA_POLY);
  bool launched = false;
  // variants without register reallocation run 11 warps (no idle twelfth warp)
#define FA_GO(V, P)                                                                                                   \
  if (!launched && want_var == (V) && want_poly == (P)) {                                                             \
    attention_tcgen05_kernel<(V), (P)><<<grid, ((V) & 2) ? FA_THREADS : FA_THREADS - 32, FA_SMEM, stream>>>(mq, mk, mv, p); \
    launched = true;                                                                                                  \
  }
  FA_VARIANTS(FA_GO)
#undef FA_GO
#undef FA_VARIANTS
  if (!launched) return set_error(TASTE_E_ARG, "attention: variant %d / poly %d is not compiled", want_var, want_poly);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace taste
