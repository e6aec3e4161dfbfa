// Non-causal flash attention on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 64, bf16 in / out, fp32
// scores and accumulators.  Serves the encoder self-attention (CW:377-394 eager semantics: softmax(q k^T) v with q
// pre-scaled, NO mask — JES:198-203), which is 16 % of the path's FLOPs.
//
// Persistent: one CTA per SM walks over work items of NT x 128 queries of one (utterance, head), as NT 128-row tiles,
// over key blocks of BK keys (default NT = 3, BK = 64; 2 x 128 is the other compiled geometry):
//   S_i = Q_i K_j^T   tcgen05.mma  SS  (M 128, N BK keys, K 64)       -> TMEM, BK fp32 columns per tile
//   softmax           one thread per query row reads its whole S row ONCE with tcgen05.ld (no shuffles) and hands S
//                     back at once, keeps the running max / sum in registers, exponentiates in packed FFMA2 pairs
//                     (MUFU ex2 for 12 of 16 pairs, a degree-3 polynomial on the FMA pipe for the other 4) and
//                     writes P (bf16) back to TMEM with tcgen05.st
//   O_i += P_i V_j    tcgen05.mma  TS  (A = P from TMEM, B = V tile, MN-major; M 128, N 64, K BK keys)
//   output            O_i / l staged per warp in shared memory (128-byte swizzle) and written with TMA stores
// Each tile has its own MMA-issuing warp, so the next block's QK^T of a tile runs under that tile's exponentials and
// under the other tiles' GEMMs.  Q (double buffered) and K/V (4-stage mbarrier ring) arrive by TMA (128-byte swizzle)
// straight from the packed [rows, 3*D] QKV activation.  The accumulator is rescaled only when a row's maximum grows by
// more than 2^8 (the stale maximum keeps exp2 arguments <= 8, exact in fp32 / bf16 range), so the O read-modify-write
// in TMEM is rare after the first key block.
//
// Roles (NT x 128 + 128 threads): warps 0 .. 4 NT - 1 softmax (warp w: tile w / 4, TMEM lanes 32 * (w % 4) .. + 31),
// then one TMA producer warp and NT MMA issuers (the first also allocates TMEM); with NT = 2 a twelfth, idle warp
// completes the last warpgroup so that setmaxnreg can move registers to the softmax threads.
// Scheduling of the softmax warpgroups, all measured on B200:
//   3 x 64, free-running: three softmax warps per scheduler cover each other's MUFU / TMEM / mbarrier latencies
//                         (0.95 ms per encoder layer, batch 64);
//   2 x 128, ping-pong  : with two warps per scheduler the tiles drift into lock-step when left alone, so their exp2
//                         phases alternate under named barriers, MUFU pairs first, polynomial pairs after the
//                         hand-over (1.00 ms).
// Nothing in the per-block loops may touch the XU pipe except the exponentials: an integer division (I2F, MUFU.RCP,
// F2I) in an MMA issuer queues behind them for hundreds of cycles.
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"
#include "internal.h"

namespace taste {

constexpr int FA_BQ = 128;          // query rows per tile
constexpr int FA_HD = 64;
// Geometry (template parameters NT, BK): NT query tiles of 128 rows per work item, each with its own softmax warpgroup
// and MMA-issuing warp, walking over key blocks of BK keys.
//   2 x 128: one thread holds a 128-score row (232 registers after setmaxnreg), two softmax warps per scheduler
//   3 x 64 : 64-score rows (152 registers), three softmax warps per scheduler to hide each other's latencies
template <int NT, int BK>
struct FaCfg {
  static constexpr int kThreads = NT * 128 + 128;       // softmax warpgroups + {TMA, NT MMA issuers, idle} warpgroup
  static constexpr int kStages = 4;                     // K/V ring (a fifth 64-key stage fits but does not help)
  static constexpr uint32_t kQTile = FA_BQ * FA_HD * 2;      // 16 KB
  static constexpr uint32_t kKvTile = BK * FA_HD * 2;        // one K or V block
  static constexpr uint32_t kOutBytes = 32 * FA_HD * 2;      // one warp's 32-row output tile
  // Q is double buffered (2 x NT tiles) so the next work item's queries load under the current item's last blocks;
  // + one 4 KB output staging tile per softmax warp (TMA-stored).  2 x 128: 230 656 B of the 232 448 available.
  static constexpr size_t kSmem = 1024 + size_t(2 * NT) * kQTile + size_t(2 * kStages) * kKvTile + size_t(4 * NT) * kOutBytes + 256;
  // TMEM columns: S_i fp32 [BK], O_i fp32 [64], P_i bf16 pairs [BK / 2]
  static constexpr uint32_t kColS = 0, kColO = NT * BK, kColP = NT * BK + NT * FA_HD;
  static constexpr int kSoftmaxRegs = NT == 2 ? 232 : 152;   // NT x 128 x regs + 128 x 40 <= 65 536
  // NT = 1 (256 threads): every thread may keep up to 255 registers from launch, so no setmaxnreg hand-over is needed
  static constexpr bool kMoveRegs = NT >= 2;
  static_assert(kColP + NT * BK / 2 <= 512, "TMEM");
  static_assert(kSmem <= 232448, "shared memory");
  static_assert((NT * 128 * kSoftmaxRegs + 128 * 40) <= 65536, "registers");
};
// VAR bits (template parameter; A/B knobs, see launch_attention_tcgen05):
//   16  free-running tiles: no ping-pong turns between the softmax warpgroups (the default with 3 x 64)
//   64  split phases: a row's MUFU pairs first, under the tile's turn (one warp saturates the XU pipe), the turn is
//       handed over, then the polynomial pairs (FMA pipe) run under the next tile's MUFU phase; without it POLY of
//       every 16 pairs are interleaved and the turn is handed over one chunk before the end of the row
//   128 speculative blocks: no row maximum after an item's first block (see the softmax loop)
//   256 eight row-maximum chains instead of four
constexpr int FA_VAR_INTERLEAVED = 0;
constexpr int FA_VAR_SPLIT = 64;
constexpr int FA_VAR_SPEC = 64 | 128;
constexpr int FA_VAR_FREE = 16;
struct FaParams {
  int q_len, kv_len;
  int heads, q_blocks, n_items;      // work item w = (b * heads + head) * q_blocks + qb
  // Ragged queries (the aggregator's cross-attention, CW:361-366 with JES:377-388's dict K/V): utterance b owns the packed
  // query / output rows cu_q[b] .. cu_q[b + 1]; keys and values stay [batch, kv_len].  Q tiles are fetched at their packed
  // row (rows past the utterance belong to its neighbour or are out-of-bounds zeros: computed, never stored) and the
  // output pass writes each thread's own row with plain stores, masked by the utterance's row count.
  const int* cu_q;     // null = fixed q_len rows per utterance
  uint16_t* o;         // ragged mode only
  int ldo;
};

// Work item w = (b * heads + head) * q_blocks + qb, walked with stride gridDim.x.  The stride is decomposed once; each
// step is three adds with carries (integer division runs on the XU pipe, behind the softmax warps' exponentials).
struct FaItemWalk {
  int qb, head, b;
  int d_qb, d_head, d_b;
  __device__ FaItemWalk(int first, int stride, int q_blocks, int heads) {
    qb = first % q_blocks;
    const int bh = first / q_blocks;
    head = bh % heads;
    b = bh / heads;
    d_qb = stride % q_blocks;
    const int dbh = stride / q_blocks;
    d_head = dbh % heads;
    d_b = dbh / heads;
  }
  __device__ void next(int q_blocks, int heads) {
    qb += d_qb;
    head += d_head;
    b += d_b;
    if (qb >= q_blocks) {
      qb -= q_blocks;
      ++head;
    }
    if (head >= heads) {
      head -= heads;
      ++b;
    }
  }
};

// Persistent: one CTA per SM loops over work items (256 queries of one (utterance, head)); TMEM, barriers and the
// K/V ring live across items, the next item's Q / K / V loads and first QK^T run under the current item's tail, so
// the ~4 us of per-CTA set-up and drain measured on the non-persistent version is paid once per launch.
template <int VAR, int POLY, int NT, int BK>
__global__ void __launch_bounds__((FaCfg<NT, BK>::kThreads), 1)
attention_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                         const __grid_constant__ CUtensorMap tma_v, const __grid_constant__ CUtensorMap tma_o,
                         const FaParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using Cfg = FaCfg<NT, BK>;
  constexpr int FA_STAGES = Cfg::kStages;
  constexpr uint32_t FA_TILE_BYTES = Cfg::kQTile, FA_KV_BYTES = Cfg::kKvTile, FA_OUT_BYTES = Cfg::kOutBytes;
  constexpr uint32_t FA_COL_S = Cfg::kColS, FA_COL_O = Cfg::kColO, FA_COL_P = Cfg::kColP;
  constexpr int FA_BK = BK;
  constexpr int kTmaWarp = 4 * NT, kMmaWarp0 = 4 * NT + 1;      // the last warpgroup: TMA, NT MMA issuers, (idle)
  constexpr int NC = BK / 32;                                   // 32-score chunks of a row
  uint8_t* sQ = smem;                                   // 2 buffers x NT tiles
  uint8_t* sKV = smem + size_t(2 * NT) * FA_TILE_BYTES; // FA_STAGES x {K, V}
  uint8_t* sOut = sKV + size_t(2 * FA_STAGES) * FA_KV_BYTES;          // 4 NT warps x 4 KB (1024-byte aligned)
  uint64_t* bars = reinterpret_cast<uint64_t*>(sOut + size_t(4 * NT) * FA_OUT_BYTES);
  uint64_t* q_full = bars;                              // [2]
  uint64_t* q_empty = bars + 2;                         // [2]  all QK^T of the item retired (MMA -> producer)
  uint64_t* kv_full = bars + 4;                         // [FA_STAGES]
  uint64_t* kv_empty = kv_full + FA_STAGES;             // [FA_STAGES]
  uint64_t* s_full = kv_empty + FA_STAGES;              // [NT]  S_i ready               (MMA -> softmax)
  uint64_t* s_free = s_full + NT;                       // [NT]  S_i read from TMEM      (softmax -> MMA), 4 warp arrivals
  uint64_t* p_full = s_free + NT;                       // [NT]  P_i written             (softmax -> MMA), 4 warp arrivals
  uint64_t* o_full = p_full + NT;                       // [NT]  P_i V done              (MMA -> softmax)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + NT);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_blocks = (p.kv_len + FA_BK - 1) / FA_BK;
  const int my_items = (p.n_items - int(blockIdx.x) + int(gridDim.x) - 1) / int(gridDim.x);
  const int total_g = my_items * n_blocks;              // key blocks this CTA walks through, over all its items

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    tma_prefetch_desc(&tma_o);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&q_full[s], 1);
      mbar_init(&q_empty[s], NT);           // one tcgen05.commit per MMA issuer
    }
    for (int s = 0; s < FA_STAGES; ++s) {
      mbar_init(&kv_full[s], 1);
      mbar_init(&kv_empty[s], NT);
    }
    for (int i = 0; i < NT; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&p_full[i], 4);
      mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // Producer and MMA warps run warp-uniform loops and elect one lane around the TMA / tcgen05 instructions (inside an
  // `if (lane == 0)` region every UTCHMMA is wrapped in an ELECT loop: ~90 cycles per MMA, measured with an in-kernel clock64 trace in round 1).
  // Register reallocation between warpgroups (setmaxnreg): the softmax threads hold a whole score row and schedule
  // their exp2 phase far better with 232 registers (2 x 128; 152 for 3 x 64); the TMA / MMA warps need almost nothing.
  // 256 x 232 + 128 x 40 <= 384 x 168.  (The instruction sits at the head of each role's branch: ptxas budgets
  // registers per region it dominates.)
#define FA_REGS_SMALL()                                                      \
  do {                                                                       \
    if constexpr (Cfg::kMoveRegs) asm volatile("setmaxnreg.dec.sync.aligned.u32 40;"); \
  } while (0)
#define FA_REGS_LARGE()                                                      \
  do {                                                                       \
    if constexpr (Cfg::kMoveRegs) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::kSoftmaxRegs)); \
  } while (0)
  if (warp > kMmaWarp0 + NT - 1) {
    FA_REGS_SMALL();        // idle: only completes the last warpgroup
  } else if (warp == kTmaWarp) {
    // ===================== TMA producer =====================
    FA_REGS_SMALL();
    int stage = 0;
    uint32_t phase = 0;
    FaItemWalk walk(int(blockIdx.x), int(gridDim.x), p.q_blocks, p.heads);
    for (int it = 0; it < my_items; ++it, walk.next(p.q_blocks, p.heads)) {
      const int head = walk.head;
      const int b = walk.b;
      const int q0 = walk.qb * (NT * FA_BQ);
      const int qbuf = it & 1;
      const int q_row0 = p.cu_q ? __ldg(p.cu_q + b) : 0;      // ragged: packed rows, one "batch" entry
      const int q_b = p.cu_q ? 0 : b;
      mbar_wait_relaxed(&q_empty[qbuf], uint32_t(((it >> 1) & 1) ^ 1));
      if (elect_one()) {
        uint8_t* sq = sQ + size_t(NT * qbuf) * FA_TILE_BYTES;
        mbar_expect_tx(&q_full[qbuf], NT * FA_TILE_BYTES);
#pragma unroll
        for (int t = 0; t < NT; ++t)
          tma_load_3d(sq + size_t(t) * FA_TILE_BYTES, &tma_q, &q_full[qbuf], head * FA_HD, q_row0 + q0 + t * FA_BQ, q_b);
      }
      __syncwarp();
      for (int j = 0; j < n_blocks; ++j) {
        mbar_wait_relaxed(&kv_empty[stage], phase ^ 1);
        uint8_t* sk = sKV + size_t(2 * stage) * FA_KV_BYTES;
        if (elect_one()) {
          mbar_expect_tx(&kv_full[stage], 2 * FA_KV_BYTES);
          tma_load_3d(sk, &tma_k, &kv_full[stage], head * FA_HD, j * FA_BK, b);
          tma_load_3d(sk + FA_KV_BYTES, &tma_v, &kv_full[stage], head * FA_HD, j * FA_BK, b);
        }
        __syncwarp();
        if (++stage == FA_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp >= kMmaWarp0) {
    // ===================== MMA issuers: one warp per query tile =====================
    // One issuing thread per tile, so neither tile's GEMMs ever queue behind a barrier that only the other tile's
    // softmax warps can satisfy (with a single issuer and a fixed wait order the two warpgroups throttled each other).
    FA_REGS_SMALL();
    const int i = warp - kMmaWarp0;
    constexpr uint32_t idesc_qk = umma_idesc(FA_BQ, FA_BK, kActBf16, 0, 0);        // A, B K-major
    constexpr uint32_t idesc_pv = umma_idesc(FA_BQ, FA_HD, kActBf16, 0, 1);        // A from TMEM, B (V) MN-major
    // No integer division in this loop: I2F / MUFU.RCP / F2I queue behind the softmax warps' exponentials on the XU
    // pipe (in-kernel timeline: 600-700 cycles per loop step, which left the issuers with no slack at all), so the
    // position of the block whose QK^T is issued next (one ahead of the block whose P V is issued) is carried along.
    constexpr uint64_t kQBufStep = (NT * FA_TILE_BYTES) >> 4;    // descriptor address units between the Q buffers
    constexpr uint64_t kStageStep = (2 * FA_KV_BYTES) >> 4;      // ... and between K/V ring stages
    const uint64_t da_base = umma_desc_k_sw128(smem_u32(sQ + size_t(i) * FA_TILE_BYTES));
    const uint64_t db_base = umma_desc_k_sw128(smem_u32(sKV));
    const uint64_t dv_base = umma_desc_mn_sw128(smem_u32(sKV + FA_KV_BYTES), 0);
    int nj = 0, nit = 0, nstage = 0;            // next QK^T: key block within its item, item, K/V ring stage
    uint32_t nphase = 0;
    // everything that block needs from the producer: its K/V stage and, on the first block of an item, the item's Q
    // (also waits for `also`, polled together with the K/V stage: a successful try_wait takes ~150 cycles to return)
    auto wait_inputs_next = [&](uint64_t* also, uint32_t also_parity) {
      const bool kv_ok = mbar_try_wait(&kv_full[nstage], nphase);
      const bool also_ok = also == nullptr || mbar_try_wait(also, also_parity);
      if (nj == 0) mbar_wait(&q_full[nit & 1], uint32_t((nit >> 1) & 1));
      if (!kv_ok) mbar_wait(&kv_full[nstage], nphase);
      if (!also_ok) mbar_wait(also, also_parity);
      tc_fence_after();
    };
    auto issue_qk_next = [&]() {
      const uint64_t da = da_base + uint64_t(nit & 1) * kQBufStep;
      const uint64_t db = db_base + uint64_t(nstage) * kStageStep;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FA_HD / 16; ++k)
          umma_ss(tmem_base + FA_COL_S + uint32_t(i * FA_BK), da + uint64_t(k * 2), db + uint64_t(k * 2), idesc_qk,
                  k != 0 ? 1u : 0u);
        umma_commit(&s_full[i]);
      }
      __syncwarp();
      if (++nj == n_blocks) {
        nj = 0;
        ++nit;
      }
      if (++nstage == FA_STAGES) {
        nstage = 0;
        nphase ^= 1;
      }
    };
    if (total_g > 0) {
      wait_inputs_next(nullptr, 0);
      // tile i starts when tile i - 1 has read its first S tile, so that one warpgroup is in its exp2-heavy pass
      // while the next reads / reduces scores instead of all hitting the MUFU together
      if (i > 0) mbar_wait(&s_free[i - 1], 0);
      issue_qk_next();
    }
    int j = 0, stage = 0, qbuf = 0;             // the block whose P V is issued
    for (int g = 0; g < total_g; ++g) {
      const bool more = g + 1 < total_g;
      if (more) wait_inputs_next(&s_free[i], uint32_t(g & 1));
      else mbar_wait(&s_free[i], uint32_t(g & 1));
      tc_fence_after();
      if (more) issue_qk_next();          // next block's scores first: the softmax warps wait on these
      mbar_wait(&p_full[i], uint32_t(g & 1));
      tc_fence_after();
      const uint64_t dv = dv_base + uint64_t(stage) * kStageStep;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < FA_BK / 16; ++k)     // 16 keys per MMA: 8 TMEM columns of P, 2048 B of V
          umma_ts(tmem_base + FA_COL_O + uint32_t(i * FA_HD), tmem_base + FA_COL_P + uint32_t(i * (BK / 2) + k * 8),
                  dv + uint64_t(k * (2048 >> 4)), idesc_pv, (j | k) != 0 ? 1u : 0u);
        umma_commit(&o_full[i]);
        umma_commit(&kv_empty[stage]);                                   // needs both tiles' commits
        if (j == n_blocks - 1) umma_commit(&q_empty[qbuf]);              // likewise
      }
      __syncwarp();
      if (++j == n_blocks) {
        j = 0;
        qbuf ^= 1;
      }
      if (++stage == FA_STAGES) stage = 0;
    }
  } else {
    // ===================== softmax + output (warps 0 .. 4 NT - 1) =====================
    FA_REGS_LARGE();
    const int i = warp >> 2;                        // query tile
    const int q = warp & 3;                         // TMEM lane quarter
    const uint32_t lane_base = tmem_base + (uint32_t(q * 32) << 16);
    const uint32_t t_s = lane_base + FA_COL_S + uint32_t(i * FA_BK);
    const uint32_t t_o = lane_base + FA_COL_O + uint32_t(i * FA_HD);
    const uint32_t t_p = lane_base + FA_COL_P + uint32_t(i * (BK / 2));
    const float kLog2e = 1.4426950408889634f;
    constexpr bool kSplit = (VAR & 64) != 0;
    constexpr bool kSpec = (VAR & 128) != 0;
    constexpr int kMuPairs = NC * (16 - POLY);      // split phases: pairs [0, kMuPairs) of a row on the MUFU
    // Ping-pong turn (named barrier 1 + i belongs to tile i; 128 arriving + 128 waiting threads): the tiles take the
    // exp2-heavy part of their blocks in rotation 0, 1, .., NT - 1, 0, ..
    // (VAR bit 4: no turns, the tiles run free)
    constexpr bool kTurns = (VAR & 16) == 0;
    auto turn_wait = [&]() {
      if constexpr (kTurns) asm volatile("bar.sync %0, 256;" ::"r"(1 + i) : "memory");
    };
    auto turn_pass = [&](int g_now) {
      if constexpr (kTurns) {
        if (i + 1 < NT) asm volatile("bar.arrive %0, 256;" ::"r"(2 + i) : "memory");
        else if (g_now + 1 < total_g) asm volatile("bar.arrive 1, 256;" ::: "memory");
      }
    };
    if (kTurns && i == NT - 1 && total_g > 0) asm volatile("bar.arrive 1, 256;" ::: "memory");      // tile 0 takes the first turn
    int g = 0;
    bool s_ready = false;            // early poll of s_full for the block about to start
    // ---- output of a finished item: O_i / l (called once the item's last P V may be waited for; g = blocks done) ----
    float out_inv = 0.f;
    int out_head = 0, out_b = 0, out_row = 0, out_it = 0;
    int out_row0 = 0, out_rows_valid = 0;      // ragged mode: first packed row and row count of the utterance
    auto write_output = [&]() {
      // Thread-per-row global stores (32 rows x 16 B per instruction) cost ~2 cycles per 16-byte request: 1900 cycles
      // per item on the in-kernel timeline, 12 % of the kernel.  Each warp now stages its 32 x 64 bf16 tile in shared
      // memory (128-byte swizzle: chunk ^ (row & 7), conflict-free 16-byte stores) and lane 0 hands it to the TMA,
      // which also clips the rows past the utterance's last query.
      mbar_wait(&o_full[i], uint32_t((g - 1) & 1));
      tc_fence_after();
      const float inv = out_inv;
      if (p.cu_q) {
        // ragged: this thread's row straight from TMEM to global memory (128 B per row), masked by the utterance's rows
#pragma unroll
        for (int hc = 0; hc < 2; ++hc) {
          uint32_t o[32];
          tmem_ld_32x32b_x32(t_o + uint32_t(hc * 32), o);
          tmem_ld_wait();
          if (out_row < out_rows_valid) {
            uint4* dst = reinterpret_cast<uint4*>(p.o + (int64_t(out_row0) + out_row) * p.ldo + out_head * FA_HD + hc * 32);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              uint4 u;
              u.x = pack_act2(__uint_as_float(o[8 * e + 0]) * inv, __uint_as_float(o[8 * e + 1]) * inv);
              u.y = pack_act2(__uint_as_float(o[8 * e + 2]) * inv, __uint_as_float(o[8 * e + 3]) * inv);
              u.z = pack_act2(__uint_as_float(o[8 * e + 4]) * inv, __uint_as_float(o[8 * e + 5]) * inv);
              u.w = pack_act2(__uint_as_float(o[8 * e + 6]) * inv, __uint_as_float(o[8 * e + 7]) * inv);
              dst[e] = u;
            }
          }
        }
        tc_fence_before();
        return;
      }
      uint8_t* stage_out = sOut + size_t(warp) * FA_OUT_BYTES;
      if (lane == 0) tma_store_wait_read();          // the previous item's store has read the staging tile
      __syncwarp();
      uint4* srow = reinterpret_cast<uint4*>(stage_out + lane * 128);
      // (32 columns at a time: in the deferred call the next block's 128 scores are live as well)
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {
        uint32_t o[32];
        tmem_ld_32x32b_x32(t_o + uint32_t(hc * 32), o);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          uint4 u;
          u.x = pack_act2(__uint_as_float(o[8 * e + 0]) * inv, __uint_as_float(o[8 * e + 1]) * inv);
          u.y = pack_act2(__uint_as_float(o[8 * e + 2]) * inv, __uint_as_float(o[8 * e + 3]) * inv);
          u.z = pack_act2(__uint_as_float(o[8 * e + 4]) * inv, __uint_as_float(o[8 * e + 5]) * inv);
          u.w = pack_act2(__uint_as_float(o[8 * e + 6]) * inv, __uint_as_float(o[8 * e + 7]) * inv);
          srow[(hc * 4 + e) ^ (lane & 7)] = u;
        }
      }
      tc_fence_before();      // O_i has been read: the next item's first P V (accumulate = 0) may overwrite it; that MMA
                              // is issued only after this warpgroup's next p_full arrive, which follows in program order
      fence_proxy_async();
      __syncwarp();
      if (lane == 0 && out_row < p.q_len) {
        tma_store_3d(&tma_o, stage_out, out_head * FA_HD, out_row, out_b);      // lane 0: `row` is the warp's first query
        tma_store_commit();
      }
    };
    FaItemWalk walk(int(blockIdx.x), int(gridDim.x), p.q_blocks, p.heads);
    for (int it = 0; it < my_items; ++it, walk.next(p.q_blocks, p.heads)) {
      const int head = walk.head;
      const int b = walk.b;
      const int row = walk.qb * (NT * FA_BQ) + i * FA_BQ + q * 32 + lane;   // query index within the utterance
      float m_used = -INFINITY;      // stale running maximum (raw score units)
      float l_run = 0.f;
      bool pending = false;          // P of the previous block written but not yet signalled

      for (int j = 0; j < n_blocks; ++j, ++g) {
        // (a successful mbarrier.try_wait still takes ~150 cycles to return - in-kernel timeline - so the polls of
        // barriers that are normally complete by the time they are needed are issued early and consumed here)
        if (!s_ready) mbar_wait(&s_full[i], uint32_t(g & 1));
        tc_fence_after();
        // The whole S row (128 fp32) is read into registers ONCE: one exposed TMEM latency per block, and the S tile
        // can be handed back to the MMA warp immediately, so the next block's QK^T runs under this block's entire
        // softmax.  P goes back to TMEM in four 16-column stores as soon as each 32-score chunk is exponentiated.
        // (Tried and rejected, with measurements in DESIGN.md: two passes over TMEM with 64 live scores; 16 softmax
        // warps with two threads per row; refilling each chunk's registers with the next block's scores under the
        // exponentials.)
        const int valid = p.kv_len - j * FA_BK;       // keys of this block that exist
        uint32_t r[NC][32];
#pragma unroll
        for (int c = 0; c < NC / 2; ++c) tmem_ld_32x32b_x32(t_s + uint32_t(c * 32), r[c]);
        if (pending) {
          // P of the previous block: its stores are waited for only now, under the latency of the loads above
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[i]);
          pending = false;
        }
#pragma unroll
        for (int c = NC / 2; c < NC; ++c) tmem_ld_32x32b_x32(t_s + uint32_t(c * 32), r[c]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[i]);       // every read of S has landed: the next QK^T may overwrite it
        if (valid < FA_BK) {
#pragma unroll
          for (int c = 0; c < NC; ++c)
            if (valid < (c + 1) * 32) {               // warp-uniform: only the chunks that reach past the last key
#pragma unroll
              for (int e = 0; e < 32; ++e)
                if (c * 32 + e >= valid) r[c][e] = 0xff800000u;      // -inf
            }
        }
        // Row maximum over the block: independent chains (a single chain of 64 dependent FMNMX3 costs ~400 cycles).
        auto row_max = [&]() {
          constexpr int kChains = (VAR & 256) ? 8 : 4;
          float mxc[kChains];
#pragma unroll
          for (int c = 0; c < kChains; ++c) mxc[c] = -INFINITY;
#pragma unroll
          for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              constexpr int kPer = kChains / NC;      // chains per chunk
              const int ch = c * kPer + e / (32 / kPer);
              mxc[ch] = fmaxf(mxc[ch], __uint_as_float(r[c][e]));
            }
          float mx = fmaxf(fmaxf(mxc[0], mxc[1]), fmaxf(mxc[2], mxc[3]));
          if constexpr (kChains == 8) mx = fmaxf(mx, fmaxf(fmaxf(mxc[4], mxc[5]), fmaxf(mxc[6], mxc[7])));
          return mx;
        };
        // Lazy rescale: the running reference m_used moves only when the block maximum exceeds it by more than 2^8
        // (returns the factor for l_run and O; 1 for rows that keep their reference).
        auto adopt_max = [&](float mx, bool& any_grow) {
          const bool grow = mx > m_used + 5.545177f;     // 8 in log2 units; first block: m_used = -inf
          any_grow = __any_sync(0xffffffffu, grow);
          float alpha = 1.0f;
          if (any_grow) {
            const float m_new = grow ? mx : m_used;
            alpha = (m_used == -INFINITY) ? 0.f : fast_exp2((m_used - m_new) * kLog2e);
            l_run *= alpha;
            m_used = m_new;
          }
          return alpha;
        };
        auto rescale_o = [&](float alpha) {
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
        };
        // One pass over the row: p = 2^(s * log2e - m_used * log2e), P (bf16) to TMEM chunk by chunk, returns the row
        // sum.  Exponentials in pairs (packed FFMA2 / FADD2).  16/clk/SM of MUFU.EX2 alone would cap the tensor pipe
        // at 50 %, so POLY of every 16 pairs are evaluated on the FMA pipe instead (degree-3 polynomial).
        // kTurn: this is the tile's regular pass, taken under the ping-pong turn (see below); the redo pass of the
        // speculative mode runs outside the protocol.
        auto exp_pass = [&](auto turn_tag, bool rescale, float alpha) {
          constexpr bool kTurn = decltype(turn_tag)::value;
          bool o_ready = false;
          if (kTurn && j > 0) o_ready = mbar_try_wait(&o_full[i], uint32_t((g - 1) & 1));      // consumed after chunk 0
          const float neg_m = -m_used * kLog2e;
          const uint64_t negm2 = f2_pack(neg_m, neg_m);
          const uint64_t log2e2 = f2_pack(kLog2e, kLog2e);
          const float c1 = kLog2e * (1.0f / 252.0f);                 // exp2_poly2_sat: t' = (t + 126) / 252
          const float c0 = (126.0f + neg_m) * (1.0f / 252.0f);
          uint64_t sum2 = f2_pack(0.f, 0.f);
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            uint32_t pk[16];
            if (kTurn && c == NC - 1) s_ready = (g + 1 < total_g) && mbar_try_wait(&s_full[i], uint32_t((g + 1) & 1));
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              if constexpr (kTurn && kSplit && POLY > 0) {
                if (c * 16 + e == kMuPairs) turn_pass(g);       // MUFU phase over: the next tile's turn
              }
              // split phases: the last 4 * POLY pairs of the row; otherwise POLY of every 16, evenly spread
              const bool use_poly = POLY > 0 && (kSplit ? (c * 16 + e >= kMuPairs) : (((e * POLY) & 15) < POLY));
              float p0, p1;
              if (use_poly && kSpec) {
                exp2_poly2_sat(__uint_as_float(r[c][2 * e]), __uint_as_float(r[c][2 * e + 1]), c1, c0, p0, p1);
              } else {
                const uint64_t t2 = f2_fma(f2_pack(__uint_as_float(r[c][2 * e]), __uint_as_float(r[c][2 * e + 1])), log2e2, negm2);
                if (use_poly) {
                  exp2_poly2(t2, p0, p1);
                } else {
                  f2_unpack(t2, p0, p1);
                  p0 = fast_exp2(p0);
                  p1 = fast_exp2(p1);
                }
              }
              sum2 = f2_add(sum2, f2_pack(p0, p1));
              pk[e] = pack_act2(p0, p1);
            }
            if (kTurn && c == 0 && j > 0) {
              // Only now is the previous block's P V needed: P_i has been consumed (it may be overwritten) and O_i is
              // stable (it may be rescaled).  On the first block of an item the output pass below already waited.
              if (!o_ready) mbar_wait(&o_full[i], uint32_t((g - 1) & 1));
              tc_fence_after();
              if (rescale) rescale_o(alpha);
            }
            tmem_st_32x32b_x16(t_p + uint32_t(c * 16), pk);
            if constexpr (kTurn && !(kSplit && POLY > 0)) {
              if (c == NC - 2) turn_pass(g);                    // hand the turn over one chunk early
            }
          }
          float sum0, sum1;
          f2_unpack(sum2, sum0, sum1);
          return sum0 + sum1;
        };

        // Speculative mode (VAR bit 7): only the first block of an item computes its row maximum.  Later blocks
        // exponentiate against the stale reference straight away - fp32 / bf16 carry 2^(+-126), so a reference that
        // is too low costs no accuracy - and a row whose block sum shows that an argument came near the fp32 range
        // (any p >= 2^100; the polynomial path clamps at 2^126 instead of wrapping) takes the exact path afterwards:
        // maximum, rescale, second pass.  This removes ~80 of ~800 issue slots per block and the maximum -> exp2
        // dependency from the per-tile serial loop.
        const bool exact = !kSpec || j == 0;
        bool any_grow = false;
        float alpha = 1.0f;
        if (exact) {
          const float mx = row_max();
          alpha = adopt_max(mx, any_grow);
        }
        // Ping-pong: the two tiles' exp2 phases alternate (named barriers 1 / 2, FA3-style) instead of drifting into
        // lock-step, where both warps of a scheduler fight for the MUFU queue and neither feeds the FMA pipe; the other
        // tile's TMEM loads, row maximum and waits run under this tile's exponentials.
        turn_wait();
        float bsum = exp_pass(std::true_type{}, any_grow, alpha);
        if constexpr (kSpec) {
          if (!exact) {
            const bool bad = !(bsum < 1.0e30f);            // also catches inf and NaN
            if (__any_sync(0xffffffffu, bad)) {
              alpha = adopt_max(row_max(), any_grow);
              rescale_o(alpha);                            // o_full of the previous block was waited for in the pass
              bsum = exp_pass(std::false_type{}, false, 1.0f);
            }
          }
        }
        l_run += bsum;
        if (j + 1 < n_blocks) {
          pending = true;              // the P stores are waited for (and p_full signalled) under the next block's loads
        } else {
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[i]);
        }
      }

      // (Deferring this pass into the next item's first block, under its loads and row maximum, hides the wait for the
      // last P V but costs more than it saves: with the extra live ranges in the block loop ptxas schedules the whole
      // loop worse - measured 1.17 ms against 1.01 ms per layer.)
      out_inv = l_run > 0.f ? __fdividef(1.0f, l_run) : 0.f;
      out_head = head;
      out_b = b;
      out_row = row;
      out_it = it;
      if (p.cu_q) {
        out_row0 = __ldg(p.cu_q + b);
        out_rows_valid = __ldg(p.cu_q + b + 1) - out_row0;
      }
      write_output();
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// {64 * heads columns, rows, batch} view of a [batch * rows, ld] bf16 activation; box = 64 x 128 x 1
static int make_map(EncodeTiledFn enc, CUtensorMap* m, const void* base, int heads, int rows, int batch, int ld,
                    int box_rows = FA_BQ) {
  cuuint64_t dims[3] = {(cuuint64_t)heads * FA_HD, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rows};
  cuuint32_t box[3] = {FA_HD, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, kActBf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error((int)r, "attention: tensor map encode failed (%d)", (int)r);
  return 0;
}

// Per-device launch state, resolved once (cudaFuncSetAttribute and the SM count are per device).
struct FaDevice {
  bool configured = false;
  int n_sm = 0;
};
static FaDevice g_fa_dev[kMaxDevices];

// Which instantiation serves a problem.  The geometry was chosen by measurement and is resolved here
// from the problem alone: no environment lookups on the launch path.
//   encoder self-attention, >= 384 queries : 3 tiles x 64 keys, free-running, 4 of 16 exp2 pairs on the FMA pipe
//   fixed-length, 256 .. 383 queries       : 2 tiles x 128 keys, split phases under ping-pong turns, 5 of 16
//   ragged queries (aggregator cross-attn) : 1 tile x 128 keys (utterances have <= 448 query rows, typically ~70)
#define FA_VARIANTS(X) X(FA_VAR_FREE, 4, 3, 64) X(FA_VAR_SPLIT, 5, 2, 128) X(FA_VAR_FREE, 4, 1, 128)

bool attention_tcgen05_eligible(const AttnDesc& d) {
  if (d.cu_kv || d.causal) return false;                                // causal / ragged keys: the mma.sync kernel
  if (d.kv_len < 128) return false;
  if (!d.cu_q && d.q_len < 2 * FA_BQ) return false;                     // small fixed-length problems: the mma.sync kernel
  if ((reinterpret_cast<uintptr_t>(d.q) | reinterpret_cast<uintptr_t>(d.k) | reinterpret_cast<uintptr_t>(d.v) |
       reinterpret_cast<uintptr_t>(d.o)) & 15)
    return false;
  if ((d.ldq | d.ldk | d.ldv | d.ldo) % 8 != 0) return false;
  if (d.cu_q && d.total_q <= 0) return false;                           // ragged mode needs the packed row count
  return true;
}

int launch_attention_tcgen05(const AttnDesc& d, cudaStream_t stream) {
  EncodeTiledFn enc = get_tensor_map_encoder();
  if (!enc) return set_error(TASTE_E_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");
  const int dev = current_device();
  FaDevice& fd = g_fa_dev[dev];
  if (!fd.configured) {
#define FA_CFG(V, P, T, K)                                                                                         \
  TASTE_CUDA_OK(cudaFuncSetAttribute(attention_tcgen05_kernel<(V), (P), T, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                     (int)FaCfg<T, K>::kSmem));
    FA_VARIANTS(FA_CFG)
#undef FA_CFG
    cudaDeviceGetAttribute(&fd.n_sm, cudaDevAttrMultiProcessorCount, dev);
    if (fd.n_sm <= 0) fd.n_sm = 148;
    fd.configured = true;
  }
  const bool ragged = d.cu_q != nullptr;
  int want_var = FA_VAR_FREE, want_poly = 4, want_nt = 3, want_bk = 64;
  if (ragged) want_nt = 1, want_bk = 128;
  else if (d.q_len < 3 * FA_BQ) want_var = FA_VAR_SPLIT, want_poly = 5, want_nt = 2, want_bk = 128;
  CUtensorMap mq, mk, mv, mo;
  int rc;
  // ragged queries / outputs: one packed [total_q, heads * 64] matrix (a single "batch" entry)
  const int q_rows = ragged ? d.total_q : d.q_len, q_batch = ragged ? 1 : d.batch;
  if ((rc = make_map(enc, &mq, d.q, d.heads, q_rows, q_batch, d.ldq))) return rc;
  if ((rc = make_map(enc, &mk, d.k, d.heads, d.kv_len, d.batch, d.ldk, want_bk))) return rc;
  if ((rc = make_map(enc, &mv, d.v, d.heads, d.kv_len, d.batch, d.ldv, want_bk))) return rc;
  if ((rc = make_map(enc, &mo, d.o, d.heads, q_rows, q_batch, d.ldo, 32))) return rc;      // one warp's rows per store
  FaParams p;
  p.q_len = d.q_len;
  p.kv_len = d.kv_len;
  p.heads = d.heads;
  p.q_blocks = (d.q_len + want_nt * FA_BQ - 1) / (want_nt * FA_BQ);
  p.n_items = p.q_blocks * d.heads * d.batch;
  p.cu_q = d.cu_q;
  p.o = static_cast<uint16_t*>(d.o);
  p.ldo = d.ldo;
  const int n_cta = p.n_items < fd.n_sm ? p.n_items : fd.n_sm;
  dim3 grid(n_cta);
  const double tq = ragged ? double(d.total_q) : double(d.batch) * d.q_len;
  ProfScope ps(stream, d.kclass == KC_ATTN_ENC ? KC_ATTN_ENC : (ragged ? KC_ATTN_AGG : KC_ATTN_TC),
               4.0 * tq * d.kv_len * FA_HD * d.heads,
               2.0 * FA_HD * d.heads * (2.0 * tq + 2.0 * double(d.batch) * d.kv_len));
  bool launched = false;
#define FA_GO(V, P, T, K)                                                                                           \
  if (!launched && want_var == (V) && want_poly == (P) && want_nt == T && want_bk == K) {                            \
    attention_tcgen05_kernel<(V), (P), T, K><<<grid, FaCfg<T, K>::kThreads, FaCfg<T, K>::kSmem, stream>>>(mq, mk, mv, mo, p); \
    launched = true;                                                                                                \
  }
  FA_VARIANTS(FA_GO)
#undef FA_GO
#undef FA_VARIANTS
  if (!launched)
    return set_error(TASTE_E_ARG, "attention: variant %d / poly %d / %dx%d is not compiled", want_var, want_poly, want_nt, want_bk);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace taste
