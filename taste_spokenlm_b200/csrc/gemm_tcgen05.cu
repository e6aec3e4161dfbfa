// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> smem ring -> tcgen05.mma (fp32 accumulators in TMEM,
// double buffered) -> tcgen05.ld epilogue with fused bias / GELU(erf) / residual / positional add.
//
// Serves every linear layer of the Whisper encoder and the aggregator (CW:342,365-366,407,699-701) and,
// through the multi-tap A addressing, the two stem convolutions as implicit GEMMs (JES:174-175): tap t of the k=3
// kernel is a GEMM over the same activations shifted by one row, and the zero padding is TMA out-of-bounds fill.
//
// Two kernels share this file: the CTA-pair kernel (cta_group::2, 256 x 256 tiles, 8 epilogue warps, coalesced
// epilogues, optional LayerNorm folding) that carries the encoder, and the single-CTA kernel below it in the file for
// small problems.  Roles of the single-CTA kernel (256 threads, 1 CTA / SM):  warp 0 = TMA producer, warp 1 = MMA
// issuer, warp 2 = TMEM allocator, warps 4-7 = epilogue (warp w owns TMEM lanes 32*(w%4)..+31 = rows of the tile).
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "internal.h"

namespace taste {

constexpr int BM = 128;
constexpr int BK = 64;            // 64 bf16 = 128 B = one swizzle row
constexpr int kGemmThreads = 256;

struct GemmParams {
  int rows_out;            // valid output rows per batch entry
  int m_tiles_per_batch;
  int n_tiles;
  int total_tiles;
  int kb_per_tap;
  int taps;
  int tap_s[3];
  int tap_dr[3];
  const float* bias;
  void* out;
  int ldc;
  const float* pos;
  // LayerNorm folding: a producer epilogue (LN == 2) also writes a bf16 copy of its fp32 output
  // rows and, per row and 128-column segment, the partial (sum, sum of squares); a consumer epilogue (LN == 1) turns
  // those into the row's mean / rstd and applies  rstd * (acc - mean * colsum[n]) + bias[n]  to the GEMM of the RAW
  // bf16 rows against gamma-folded weights, which equals Linear(LayerNorm(row)).
  const float* ln_stats;      // [rows][ln_nseg][2]
  int ln_nseg;
  float ln_inv_k;             // 1 / (128 * ln_nseg)
  const float* ln_colsum;     // [n]
  float* stats_out;           // [rows][n / 128][2]
  __nv_bfloat16* out_bf16;    // [rows, ldc]
  int ab_bf16;                // tcgen05 operand format: 1 = bf16, 0 = fp16
  int kclass;                 // host-side accounting only
  double alg_bytes;
};

template <int BN>
struct GemmCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr uint32_t kABytes = BM * BK * 2;
  static constexpr uint32_t kBBytes = BN * BK * 2;
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BN;       // two accumulator stages
  static constexpr size_t kSmemBytes = 1024 /*align slack*/ + size_t(kStages) * kStageBytes + 256 /*barriers*/;
};

template <int EPI>
TASTE_DEVINL void epilogue_chunk(const uint32_t (&acc)[32], const GemmParams& p, int64_t grow, int row_in_batch, int n0) {
  float v[32];
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
      v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
      v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
      v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  }
  if (EPI == EPI_GELU_BF16 || EPI == EPI_GELU_POS_F32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf<true>(v[j]);
  }
  if (EPI == EPI_BF16 || EPI == EPI_GELU_BF16) {
    uint4* o = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + grow * p.ldc + n0);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u;
      u.x = pack_act2(v[8 * j + 0], v[8 * j + 1]);
      u.y = pack_act2(v[8 * j + 2], v[8 * j + 3]);
      u.z = pack_act2(v[8 * j + 4], v[8 * j + 5]);
      u.w = pack_act2(v[8 * j + 6], v[8 * j + 7]);
      o[j] = u;
    }
  } else {
    float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + grow * p.ldc + n0);
    if (EPI == EPI_RESID_F32) {
      float4 r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = o[j];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r[j].x += v[4 * j + 0];
        r[j].y += v[4 * j + 1];
        r[j].z += v[4 * j + 2];
        r[j].w += v[4 * j + 3];
        o[j] = r[j];
      }
    } else if (EPI == EPI_GELU_POS_F32) {
      const float4* pp = reinterpret_cast<const float4*>(p.pos + int64_t(row_in_batch) * p.ldc + n0);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 q = __ldg(pp + j);
        o[j] = make_float4(v[4 * j + 0] + q.x, v[4 * j + 1] + q.y, v[4 * j + 2] + q.z, v[4 * j + 3] + q.w);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
}

// fp32 outputs (residual add in place, plain, GELU + positions) with coalesced global access: the warp's 32 rows x 32
// columns go through a swizzled 4 KB smem tile so that every load / store instruction covers 4 rows x 128 contiguous
// bytes instead of 32 rows x 16 bytes (thread-per-row).  The residual read-modify-write of out_proj / fc2 is what
// bounds those GEMMs (K = 1280: 1.2 GB of HBM traffic against 0.3 TFLOP).
template <int EPI, int LN>
TASTE_DEVINL void epilogue_chunk_f32_coalesced(const uint32_t (&acc)[32], const GemmParams& p, int64_t grow0,
                                               int row0_in_batch, int n0, float4* scratch, int lane, float (&psum)[8],
                                               float (&psq)[8]) {
  float v[32];
  if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
      v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
      v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
      v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  }
  if (EPI == EPI_GELU_POS_F32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = gelu_erf<true>(v[j]);
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)            // own row, 16-byte chunk j -> slot j ^ (row & 7): conflict-free both ways
    scratch[lane * 8 + (j ^ (lane & 7))] = make_float4(v[4 * j + 0], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int j = lane & 7;
  const int r_base = lane >> 3;
  float* out_base = reinterpret_cast<float*>(p.out) + grow0 * p.ldc + n0 + j * 4;
  float4 res[8];
  if (EPI == EPI_RESID_F32) {            // all eight residual loads in flight before the first store
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int R = i * 4 + r_base;
      res[i] = (row0_in_batch + R < p.rows_out) ? *reinterpret_cast<const float4*>(out_base + int64_t(R) * p.ldc)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else if (EPI == EPI_GELU_POS_F32) {
    const float* pos_base = p.pos + int64_t(row0_in_batch) * p.ldc + n0 + j * 4;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int R = i * 4 + r_base;
      res[i] = (row0_in_batch + R < p.rows_out) ? __ldg(reinterpret_cast<const float4*>(pos_base + int64_t(R) * p.ldc))
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int R = i * 4 + r_base;
    float4 x = scratch[R * 8 + (j ^ (R & 7))];
    if (EPI == EPI_RESID_F32 || EPI == EPI_GELU_POS_F32) {
      x.x += res[i].x; x.y += res[i].y; x.z += res[i].z; x.w += res[i].w;
    }
    const bool row_ok = row0_in_batch + R < p.rows_out;
    if (row_ok) *reinterpret_cast<float4*>(out_base + int64_t(R) * p.ldc) = x;
    if (LN == 2) {
      // bf16 copy of the new fp32 row (the next GEMM's A operand) and this lane's share of the row statistics
      if (row_ok) {
        uint2 u;
        u.x = pack_act2(x.x, x.y);
        u.y = pack_act2(x.z, x.w);
        *reinterpret_cast<uint2*>(p.out_bf16 + (grow0 + R) * p.ldc + n0 + j * 4) = u;
      }
      psum[i] += (x.x + x.y) + (x.z + x.w);          // this lane's 4 columns; lanes are combined once per tile
      psq[i] = fmaf(x.x, x.x, fmaf(x.y, x.y, fmaf(x.z, x.z, fmaf(x.w, x.w, psq[i]))));
    }
  }
  __syncwarp();
}

// bf16 outputs, 64 columns (two TMEM chunks) at a time: bias (+ GELU) on the thread's own row, pack to bf16, then the
// same swizzled 32 x 128-byte smem transpose so that stores cover 4 rows x 128 contiguous bytes per instruction.
template <int EPI, int LN>
TASTE_DEVINL void epilogue_pack_bf16(const uint32_t (&acc)[32], const GemmParams& p, int n0, uint32_t (&out16)[16],
                                     float mu, float rstd) {
  float v[32];
  if (LN == 1) {
    // Linear(LayerNorm(row)) from the GEMM of the raw row against gamma-folded weights (p.bias holds b + W beta)
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
    const float4* c4 = reinterpret_cast<const float4*>(p.ln_colsum + n0);
    const float nmr = -mu * rstd;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      const float4 c = __ldg(c4 + j);
      v[4 * j + 0] = fmaf(__uint_as_float(acc[4 * j + 0]), rstd, fmaf(nmr, c.x, b.x));
      v[4 * j + 1] = fmaf(__uint_as_float(acc[4 * j + 1]), rstd, fmaf(nmr, c.y, b.y));
      v[4 * j + 2] = fmaf(__uint_as_float(acc[4 * j + 2]), rstd, fmaf(nmr, c.z, b.z));
      v[4 * j + 3] = fmaf(__uint_as_float(acc[4 * j + 3]), rstd, fmaf(nmr, c.w, b.w));
    }
  } else if (p.bias != nullptr) {
    const float4* b4 = reinterpret_cast<const float4*>(p.bias + n0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 b = __ldg(b4 + j);
      v[4 * j + 0] = __uint_as_float(acc[4 * j + 0]) + b.x;
      v[4 * j + 1] = __uint_as_float(acc[4 * j + 1]) + b.y;
      v[4 * j + 2] = __uint_as_float(acc[4 * j + 2]) + b.z;
      v[4 * j + 3] = __uint_as_float(acc[4 * j + 3]) + b.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
  }
  if (EPI == EPI_GELU_BF16) {
#pragma unroll
    for (int j = 0; j < 16; ++j) gelu_erf_pair(v[2 * j], v[2 * j + 1], v[2 * j], v[2 * j + 1]);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) out16[j] = pack_act2(v[2 * j], v[2 * j + 1]);
}

TASTE_DEVINL void epilogue_store_bf16_coalesced(const uint32_t (&lo)[16], const uint32_t (&hi)[16], const GemmParams& p,
                                                int64_t grow0, int row0_in_batch, int n0, float4* scratch, int lane) {
  uint4* s4 = reinterpret_cast<uint4*>(scratch);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    s4[lane * 8 + (j ^ (lane & 7))] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
    s4[lane * 8 + ((j + 4) ^ (lane & 7))] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
  }
  __syncwarp();
  const int j = lane & 7;
  __nv_bfloat16* out_base = reinterpret_cast<__nv_bfloat16*>(p.out) + grow0 * p.ldc + n0 + j * 8;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int R = i * 4 + (lane >> 3);
    const uint4 x = s4[R * 8 + (j ^ (R & 7))];
    if (row0_in_batch + R < p.rows_out) *reinterpret_cast<uint4*>(out_base + int64_t(R) * p.ldc) = x;
  }
  __syncwarp();
}

template <int BN, int EPI>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + size_t(kStages) * Cfg::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2] accumulator ready   (MMA -> epilogue)
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);          // one arrive per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kb_total = p.taps * p.kb_per_tap;

  // Producer and MMA warps run their loops with all 32 lanes (warp-uniform control flow) and elect one lane around
  // the TMA / tcgen05 instructions: inside an `if (lane == 0)` region the compiler wraps every UTCHMMA / UTMALDG in
  // an ELECT loop (~90 cycles per MMA measured), which made the issuing thread the bottleneck.
  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mg = tile / p.n_tiles;
      const int b = mg / p.m_tiles_per_batch;
      const int mt = mg - b * p.m_tiles_per_batch;
      for (int kb = 0; kb < kb_total; ++kb) {
        const int tap = kb / p.kb_per_tap;
        const int kc = kb - tap * p.kb_per_tap;
        mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + size_t(stage) * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        const int ts = tap == 0 ? p.tap_s[0] : (tap == 1 ? p.tap_s[1] : p.tap_s[2]);
        const int tr = tap == 0 ? p.tap_dr[0] : (tap == 1 ? p.tap_dr[1] : p.tap_dr[2]);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_4d(sa, &tma_a, &full_bar[stage], kc * BK, ts, mt * BM + tr, b);
          tma_load_2d(sb, &tma_b, &full_bar[stage], kb * BK, nt * BN);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const uint32_t idesc = umma_idesc(BM, BN, p.ab_bf16, 0, 0);            // operand format: runtime (the DFT is always bf16)
    int stage = 0;
    uint32_t phase = 0;
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(&tempty_bar[as], aphase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + uint32_t(as * BN);
      for (int kb = 0; kb < kb_total; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + size_t(stage) * Cfg::kStageBytes);
        const uint64_t da = umma_desc_k_sw128(sa);
        const uint64_t db = umma_desc_k_sw128(sa + Cfg::kABytes);
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-B swizzle row: +2 in the (addr >> 4) field
            umma_ss(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);      // frees the smem slot when these MMAs retire
          if (kb == kb_total - 1) umma_commit(&tfull_bar[as]);   // accumulator complete
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                    // TMEM lane quarter this warp may access
    int as = 0;
    uint32_t aphase = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int nt = tile % p.n_tiles;
      const int mg = tile / p.n_tiles;
      const int b = mg / p.m_tiles_per_batch;
      const int mt = mg - b * p.m_tiles_per_batch;
      const int row_in_batch = mt * BM + q * 32 + lane;
      const bool valid = row_in_batch < p.rows_out;
      const int64_t grow = int64_t(b) * p.rows_out + row_in_batch;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN);
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        uint32_t acc[32];
        tmem_ld_32x32b_x32(taddr + uint32_t(c * 32), acc);
        tmem_ld_wait();
        if (valid) epilogue_chunk<EPI>(acc, p, grow, row_in_batch, nt * BN + c * 32);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[as]);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): one 256 x 256 output tile per pair of SMs.  Each CTA stages its own 128 rows of A
// and 128 of the 256 B rows per k-block (32 KB instead of the 48 KB a single-CTA 128 x 256 tile needs), which is
// what lifts the kernel off the L2 -> SM bandwidth ceiling.  The leader CTA issues the MMAs for both; accumulators
// (128 lanes x 256 columns per CTA, double buffered) live in each CTA's own TMEM.
// Roles per CTA (384 threads): warp 0 = TMA producer, warp 1 = MMA issuer (leader only), warp 2 = TMEM allocator,
// warps 4-11 = epilogue (warp w: TMEM lanes 32*(w%4)..+31, columns 128*((w-4)/4)..+127).
// ------------------------------------------------------------------------------------------------
constexpr int kGemm2Threads = 384;
constexpr int BN2 = 256;
struct Gemm2Cfg {
  static constexpr int kStages = 6;
  static constexpr uint32_t kABytes = BM * BK * 2;           // this CTA's 128 rows of A
  static constexpr uint32_t kBBytes = (BN2 / 2) * BK * 2;    // this CTA's 128 of the 256 B rows
  static constexpr uint32_t kStageBytes = kABytes + kBBytes;
  static constexpr uint32_t kTmemCols = 2 * BN2;
  static constexpr size_t kScratchBytes = 8 * 4096;         // one 32 x 32 fp32 transpose tile per epilogue warp
  static constexpr size_t kSmemBytes = 1024 + size_t(kStages) * kStageBytes + 256 + kScratchBytes;
};

template <int EPI, int LN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kGemm2Threads, 1)
gemm_bf16_tcgen05_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                              const GemmParams p) {
  using Cfg = Gemm2Cfg;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + size_t(kStages) * Cfg::kStageBytes);   // used in the leader
  uint64_t* empty_bar = full_bar + kStages;    // per CTA: smem slot free (MMA commit, multicast to both CTAs)
  uint64_t* tfull_bar = empty_bar + kStages;   // per CTA: accumulator ready (MMA commit, multicast)
  uint64_t* tempty_bar = tfull_bar + 2;        // leader: accumulator drained by the 16 epilogue warps of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);            // the leader's arrive.expect_tx covers the bytes of both CTAs
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 16);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int kb_total = p.taps * p.kb_per_tap;

  if (warp == 0) {
    // ===================== TMA producer (both CTAs; warp-uniform loop, one elected lane issues) =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
      const int nt = tile % p.n_tiles;
      const int mg = tile / p.n_tiles;
      const int b = mg / p.m_tiles_per_batch;
      const int mt = mg - b * p.m_tiles_per_batch;
      const int row0 = mt * (2 * BM) + int(rank) * BM;
      const int col0 = nt * BN2 + int(rank) * (BN2 / 2);
      for (int kb = 0; kb < kb_total; ++kb) {
        const int tap = kb / p.kb_per_tap;
        const int kc = kb - tap * p.kb_per_tap;
        mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + size_t(stage) * Cfg::kStageBytes;
        uint8_t* sb = sa + Cfg::kABytes;
        // Both CTAs' TMA bytes are counted on the LEADER's barrier; only the leader arrives (expecting both halves).
        // The peer cannot run a phase ahead: its slot is released by the same multicast commit as the leader's.
        const uint32_t lead_full = mapa_shared(smem_u32(&full_bar[stage]), 0);
        const int ts = tap == 0 ? p.tap_s[0] : (tap == 1 ? p.tap_s[1] : p.tap_s[2]);
        const int tr = tap == 0 ? p.tap_dr[0] : (tap == 1 ? p.tap_dr[1] : p.tap_dr[2]);
        if (elect_one()) {
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::kStageBytes);
          tma_load_4d_2sm(sa, &tma_a, lead_full, kc * BK, ts, row0 + tr, b);
          tma_load_2d_2sm(sb, &tma_b, lead_full, kb * BK, col0);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA only) =====================
      const uint32_t idesc = umma_idesc(2 * BM, BN2, p.ab_bf16, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int as = 0;
      uint32_t aphase = 0;
      for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
        mbar_wait(&tempty_bar[as], aphase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + uint32_t(as * BN2);
        for (int kb = 0; kb < kb_total; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + size_t(stage) * Cfg::kStageBytes);
          const uint64_t da = umma_desc_k_sw128(sa);
          const uint64_t db = umma_desc_k_sw128(sa + Cfg::kABytes);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)
              umma_ss_2sm(d_tmem, da + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            umma_commit_2sm(&empty_bar[stage], 0x3);    // both CTAs may refill this slot once the MMAs retire
            if (kb == kb_total - 1) umma_commit_2sm(&tfull_bar[as], 0x3);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (++as == 2) {
          as = 0;
          aphase ^= 1;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (both CTAs) =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    float4* scratch = reinterpret_cast<float4*>(smem + size_t(kStages) * Cfg::kStageBytes + 256) + (warp - 4) * 256;
    int as = 0;
    uint32_t aphase = 0;
    const uint32_t lead_tempty0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    const uint32_t lead_tempty1 = mapa_shared(smem_u32(&tempty_bar[1]), 0);
    for (int tile = pair; tile < p.total_tiles; tile += n_pairs) {
      const int nt = tile % p.n_tiles;
      const int mg = tile / p.n_tiles;
      const int b = mg / p.m_tiles_per_batch;
      const int mt = mg - b * p.m_tiles_per_batch;
      const int row0_in_batch = mt * (2 * BM) + int(rank) * BM + q * 32;
      const int row_in_batch = row0_in_batch + lane;
      const bool valid = row_in_batch < p.rows_out;
      const int64_t grow0 = int64_t(b) * p.rows_out + row0_in_batch;
      const int64_t grow = grow0 + lane;
      mbar_wait(&tfull_bar[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(as * BN2 + half * (BN2 / 2));
      if (EPI == EPI_RESID_F32 || EPI == EPI_F32 || EPI == EPI_GELU_POS_F32) {
        float psum[8], psq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) psum[i] = psq[i] = 0.f;
#pragma unroll 1
        for (int c = 0; c < BN2 / 64; ++c) {
          uint32_t acc[32];
          tmem_ld_32x32b_x32(taddr + uint32_t(c * 32), acc);
          tmem_ld_wait();
          epilogue_chunk_f32_coalesced<EPI, LN>(acc, p, grow0, row0_in_batch, nt * BN2 + half * (BN2 / 2) + c * 32,
                                                scratch, lane, psum, psq);
        }
        if (LN == 2) {
          // this warp's 128 columns are one statistics segment of its 32 rows; the 8 lanes that share a row combine
          // their partials with three shuffles, lane j == 0 writes the (sum, sum of squares) pair
          const int seg = nt * (BN2 / 128) + half;
          const int nseg = p.n_tiles * (BN2 / 128);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float s1 = psum[i], s2 = psq[i];
#pragma unroll
            for (int o = 1; o < 8; o <<= 1) {
              s1 += __shfl_xor_sync(0xffffffffu, s1, o);
              s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            }
            const int R = i * 4 + (lane >> 3);
            if ((lane & 7) == 0 && row0_in_batch + R < p.rows_out)
              *reinterpret_cast<float2*>(p.stats_out + ((grow0 + R) * nseg + seg) * 2) = make_float2(s1, s2);
          }
        }
      } else {
        float mu = 0.f, rstd = 0.f;
        if (LN == 1 && valid) {
          float s1 = 0.f, s2 = 0.f;
          const float2* st = reinterpret_cast<const float2*>(p.ln_stats) + grow * p.ln_nseg;
          for (int sgm = 0; sgm < p.ln_nseg; ++sgm) {        // fixed order: deterministic
            const float2 t = st[sgm];
            s1 += t.x;
            s2 += t.y;
          }
          mu = s1 * p.ln_inv_k;
          rstd = rsqrtf(fmaxf(fmaf(-mu, mu, s2 * p.ln_inv_k), 0.f) + 1e-5f);
        }
#pragma unroll 1
        for (int c = 0; c < BN2 / 128; ++c) {        // 64 columns per step
          const int n0 = nt * BN2 + half * (BN2 / 2) + c * 64;
          uint32_t acc0[32], acc1[32], lo[16], hi[16];
          tmem_ld_32x32b_x32(taddr + uint32_t(c * 64), acc0);
          tmem_ld_32x32b_x32(taddr + uint32_t(c * 64 + 32), acc1);
          tmem_ld_wait();
          epilogue_pack_bf16<EPI, LN>(acc0, p, n0, lo, mu, rstd);
          epilogue_pack_bf16<EPI, LN>(acc1, p, n0 + 32, hi, mu, rstd);
          epilogue_store_bf16_coalesced(lo, hi, p, grow0, row0_in_batch, n0, scratch, lane);
        }
      }
      (void)valid;
      (void)grow;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(as == 0 ? lead_tempty0 : lead_tempty1);
      if (++as == 2) {
        as = 0;
        aphase ^= 1;
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();          // the peer's remote arrives and smem reads are complete before either CTA exits
  if (warp == 2) tmem_dealloc_2sm(tmem_base, Cfg::kTmemCols);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
EncodeTiledFn get_tensor_map_encoder() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// Launch state is cached per device ordinal: shared-memory attributes and occupancy belong to the CURRENT device.
static int num_sms() {
  static int n[kMaxDevices] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

template <int BN, int EPI>
static int launch_cfg(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05_kernel<BN, EPI>;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    TASTE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)GemmCfg<BN>::kSmemBytes));
    configured = true;
  }
  const int grid = p.total_tiles < num_sms() ? p.total_tiles : num_sms();
  const double m = double(p.rows_out) * (p.total_tiles / (p.n_tiles * p.m_tiles_per_batch));
  const double n = double(p.n_tiles) * BN, k = double(p.taps) * p.kb_per_tap * BK;
  const double out_b = (EPI == EPI_BF16 || EPI == EPI_GELU_BF16) ? 2.0 : (EPI == EPI_RESID_F32 ? 8.0 : 4.0);
  ProfScope ps(stream, p.kclass, 2.0 * m * n * k, p.alg_bytes >= 0 ? p.alg_bytes : 2.0 * (m * k / p.taps + n * k) + out_b * m * n);
  kern<<<grid, kGemmThreads, GemmCfg<BN>::kSmemBytes, stream>>>(ta, tb, p);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

template <int EPI, int LN>
static int launch_cfg2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05_2cta_kernel<EPI, LN>;
  static int max_pairs_dev[kMaxDevices] = {};
  int& max_pairs = max_pairs_dev[current_device()];
  if (max_pairs == 0) {
    TASTE_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Gemm2Cfg::kSmemBytes));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * num_sms());
    cfg.blockDim = dim3(kGemm2Threads);
    cfg.dynamicSmemBytes = Gemm2Cfg::kSmemBytes;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2;
    attr.val.clusterDim.y = 1;
    attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    int n = 0;
    TASTE_CUDA_OK(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
    if (n <= 0) return set_error(TASTE_E_NO_DEVICE, "gemm: no CTA pair fits on this device");
    max_pairs = n;
    if (getenv("TASTE_DEBUG")) fprintf(stderr, "[taste] gemm pair kernel: %d co-resident CTA pairs\n", n);
  }
  int pairs = p.total_tiles < max_pairs ? p.total_tiles : max_pairs;
  const double m = double(p.rows_out) * (p.total_tiles / (p.n_tiles * p.m_tiles_per_batch));
  const double n = double(p.n_tiles) * BN2, k = double(p.taps) * p.kb_per_tap * BK;
  const double out_b = (EPI == EPI_BF16 || EPI == EPI_GELU_BF16) ? 2.0 : (EPI == EPI_RESID_F32 ? 8.0 : 4.0);
  ProfScope ps(stream, p.kclass, 2.0 * m * n * k, p.alg_bytes >= 0 ? p.alg_bytes : 2.0 * (m * k / p.taps + n * k) + out_b * m * n);
  kern<<<2 * pairs, kGemm2Threads, Gemm2Cfg::kSmemBytes, stream>>>(ta, tb, p);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

static int launch_epi2(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int epi, int ln, cudaStream_t s) {
  if (ln == 1) {
    if (epi == EPI_BF16) return launch_cfg2<EPI_BF16, 1>(ta, tb, p, s);
    if (epi == EPI_GELU_BF16) return launch_cfg2<EPI_GELU_BF16, 1>(ta, tb, p, s);
    return set_error(TASTE_E_ARG, "gemm: LayerNorm-in folding needs a bf16 epilogue");
  }
  if (ln == 2) {
    if (epi == EPI_RESID_F32) return launch_cfg2<EPI_RESID_F32, 2>(ta, tb, p, s);
    if (epi == EPI_GELU_POS_F32) return launch_cfg2<EPI_GELU_POS_F32, 2>(ta, tb, p, s);
    return set_error(TASTE_E_ARG, "gemm: LayerNorm-out statistics need an fp32 epilogue");
  }
  switch (epi) {
    case EPI_BF16: return launch_cfg2<EPI_BF16, 0>(ta, tb, p, s);
    case EPI_GELU_BF16: return launch_cfg2<EPI_GELU_BF16, 0>(ta, tb, p, s);
    case EPI_RESID_F32: return launch_cfg2<EPI_RESID_F32, 0>(ta, tb, p, s);
    case EPI_F32: return launch_cfg2<EPI_F32, 0>(ta, tb, p, s);
    case EPI_GELU_POS_F32: return launch_cfg2<EPI_GELU_POS_F32, 0>(ta, tb, p, s);
  }
  return set_error(TASTE_E_ARG, "gemm: unknown epilogue %d", epi);
}

template <int BN>
static int launch_epi(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int epi, cudaStream_t s) {
  switch (epi) {
    case EPI_BF16: return launch_cfg<BN, EPI_BF16>(ta, tb, p, s);
    case EPI_GELU_BF16: return launch_cfg<BN, EPI_GELU_BF16>(ta, tb, p, s);
    case EPI_RESID_F32: return launch_cfg<BN, EPI_RESID_F32>(ta, tb, p, s);
    case EPI_F32: return launch_cfg<BN, EPI_F32>(ta, tb, p, s);
    case EPI_GELU_POS_F32: return launch_cfg<BN, EPI_GELU_POS_F32>(ta, tb, p, s);
  }
  return set_error(TASTE_E_ARG, "gemm: unknown epilogue %d", epi);
}

// 0 = automatic, 1 = force the single-CTA kernel (tests / A-B timing)
static int g_gemm_mode = 0;
void set_gemm_mode(int mode) { g_gemm_mode = mode; }

int launch_gemm(const GemmDesc& d, cudaStream_t stream) {
  if (!d.a || !d.w || !d.out) return set_error(TASTE_E_ARG, "gemm: null pointer");
  if (d.k_inner % BK != 0 || d.n % 128 != 0 || d.taps < 1 || d.taps > 3)
    return set_error(TASTE_E_SHAPE, "gemm: need K %% 64 == 0 and N %% 128 == 0 (k_inner=%d n=%d taps=%d)", d.k_inner,
                     d.n, d.taps);
  if (d.rows_out <= 0 || d.batches <= 0) return 0;
  if (d.epilogue == EPI_GELU_POS_F32 && !d.pos) return set_error(TASTE_E_ARG, "gemm: pos table missing");
  EncodeTiledFn enc = get_tensor_map_encoder();
  if (!enc) return set_error(TASTE_E_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");

  // tile shape: CTA pairs (256 x 256) when there is at least one wave of pair tiles, else single-CTA 128 x {256,128}
  const int64_t pair_tiles = (d.n % 256 == 0) ? int64_t((d.rows_out + 2 * BM - 1) / (2 * BM)) * d.batches * (d.n / 256) : 0;
  const int ln = d.ln_stats ? 1 : (d.stats_out ? 2 : 0);
  if (ln && (d.n % 256 != 0))
    return set_error(TASTE_E_SHAPE, "gemm: LayerNorm folding needs N %% 256 == 0 (CTA-pair kernel)");
  if (ln == 1 && (!d.ln_colsum || !d.bias || d.ln_nseg <= 0)) return set_error(TASTE_E_ARG, "gemm: LayerNorm-in folding needs stats, colsum and bias");
  if (ln == 2 && !d.out_bf16) return set_error(TASTE_E_ARG, "gemm: LayerNorm-out needs the bf16 copy buffer");
  const bool use_pair = ln != 0 || d.force_pair || (g_gemm_mode != 1 && pair_tiles >= num_sms() / 2);
  const int m_tiles = use_pair ? (d.rows_out + 2 * BM - 1) / (2 * BM) : (d.rows_out + BM - 1) / BM;
  const int64_t tiles256 = (d.n % 256 == 0) ? int64_t(m_tiles) * d.batches * (d.n / 256) : 0;
  const int bn = use_pair ? 128 /* B box: this CTA's half of the 256 columns */ : ((tiles256 >= num_sms()) ? 256 : 128);

  const int ab_bf16 = d.ab_bf16 < 0 ? kActBf16 : d.ab_bf16;
  CUtensorMap ta, tb;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d.k_inner, (cuuint64_t)d.s_count, (cuuint64_t)d.rows_in, (cuuint64_t)d.batches};
    cuuint64_t strides[3] = {(cuuint64_t)d.s_stride, (cuuint64_t)d.r_stride, (cuuint64_t)d.b_stride};
    cuuint32_t box[4] = {BK, 1, BM, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(&ta, ab_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(d.a), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error((int)r, "gemm: tensor map A encode failed (%d)", (int)r);
  }
  {
    const int64_t ktot = int64_t(d.taps) * d.k_inner;
    cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)d.n};
    cuuint64_t strides[1] = {(cuuint64_t)(ktot * 2)};
    cuuint32_t box[2] = {BK, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tb, ab_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(d.w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error((int)r, "gemm: tensor map B encode failed (%d)", (int)r);
  }
  GemmParams p;
  p.rows_out = d.rows_out;
  p.m_tiles_per_batch = m_tiles;
  p.n_tiles = use_pair ? d.n / 256 : d.n / bn;
  p.total_tiles = m_tiles * d.batches * p.n_tiles;
  p.kb_per_tap = d.k_inner / BK;
  p.taps = d.taps;
  for (int i = 0; i < 3; ++i) {
    p.tap_s[i] = d.tap_s[i];
    p.tap_dr[i] = d.tap_dr[i];
  }
  p.bias = d.bias;
  p.out = d.out;
  p.ldc = d.ldc;
  p.pos = d.pos;
  p.ln_stats = d.ln_stats;
  p.ln_nseg = d.ln_nseg;
  p.ln_inv_k = d.ln_nseg > 0 ? 1.0f / float(128 * d.ln_nseg) : 0.f;
  p.ln_colsum = d.ln_colsum;
  p.stats_out = d.stats_out;
  p.out_bf16 = static_cast<__nv_bfloat16*>(d.out_bf16);
  p.ab_bf16 = ab_bf16;
  p.kclass = d.kclass;
  p.alg_bytes = d.alg_bytes;
  if (use_pair) return launch_epi2(ta, tb, p, d.epilogue, ln, stream);
  return bn == 256 ? launch_epi<256>(ta, tb, p, d.epilogue, stream) : launch_epi<128>(ta, tb, p, d.epilogue, stream);
}

int gemm_plain(const void* a, const void* w, const float* bias, void* out, int m, int n, int k, int epilogue,
               cudaStream_t stream) {
  GemmDesc d;
  d.a = a;
  d.k_inner = k;
  d.s_count = 1;
  d.rows_in = m;
  d.rows_out = m;
  d.batches = 1;
  d.s_stride = int64_t(k) * 2;
  d.r_stride = int64_t(k) * 2;
  d.b_stride = int64_t(m) * k * 2;
  d.taps = 1;
  d.w = w;
  d.n = n;
  d.bias = bias;
  d.out = out;
  d.ldc = n;
  d.epilogue = epilogue;
  return launch_gemm(d, stream);
}

}  // namespace taste
