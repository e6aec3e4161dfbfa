// Residual vector quantiser (AQ:109-124 -> RVQ:359-490 -> VQ:955-1217 -> VQ:462-566; Appendix A6).
//
// Contract: indices bit-identical to the reference's fp32 formula (VQ:44-48, VQ:102)
//     d_j = sqrt(max((|r|^2 + |e_j|^2) + (-2 <r, e_j>), 0)),   index = first minimum of d_j
// given the same input.  Three kernels:
//   1. rvq_sgemm_kernel      x = project_in(z)  (RVQ:371), fp32 FFMA: every output is one k-ascending fma chain
//   2. rvq_search_kernel     128 tokens per CTA, all levels fused, residuals never leave the SM:
//        * the 512 dot products <r, e_j> of a level run on the tensor cores (tcgen05.mma, fp32 accumulators in TMEM) as a
//          split-bf16 three-product GEMM: r = r_hi + r_lo, e = e_hi + e_lo, <r,e> ~ r_hi.e_hi + r_hi.e_lo + r_lo.e_hi
//          (worst-case error 1.1e-5 |r||e|); the codebook planes arrive by TMA through a 3-stage ring, the residual planes
//          are written by the row threads straight into the 128-byte-swizzled K-major layout the MMA reads;
//        * the fp32 residual of every token lives in TMEM (columns 256..511, one lane per token) next to the 256
//          distance columns of the current half of the codebook;
//        * one thread per token scans its distances with a fused arg-min that keeps every code whose lower error bound
//          lies below the smallest upper bound (error bound 3e-5 (|r|^2 + |e_j|^2) per score).  A single survivor
//          (99.7 % of the tokens) IS the fp32 arg-min; otherwise the survivors are re-evaluated with the exact fp32
//          formula, operation order and first-minimum tie-break of the reference, so indices stay bit-exact;
//        * gather, residual -= code, re-split into bf16 planes for the next level.
//   3. rvq_sgemm_kernel      quantized = project_out(sum of codes)  (RVQ:470), fp32 FFMA as above
#include "common.cuh"
#include "internal.h"

namespace taste {

constexpr int RVQ_ROWS = 16;         // decode kernel: tokens per CTA
constexpr int RVQ_THREADS = 256;
constexpr int RVQ_DC = 256;          // codebook dim
constexpr int RVQ_K = 512;           // codes per level
constexpr int RVQ_MAXQ = 8;

// ------------------------------------------------------------------------------------------------
// fp32 projection GEMM: C[m, n] = bias[n] + sum_k A[m, k] * Bt[k, n]; each output is ONE fma chain over ascending k
// (the reference is fp32; a tensor-core product would change the indices it feeds).  64 x 64 tile, 4 x 4 per thread.
// ------------------------------------------------------------------------------------------------
constexpr int SG_T = 64, SG_K = 16;
__global__ void __launch_bounds__(256)
rvq_sgemm_kernel(const float* __restrict__ A, int lda, const float* __restrict__ Bt, int ldb,
                 const float* __restrict__ bias, float* __restrict__ C, int ldc, int M, int N, int K) {
  __shared__ __align__(16) float sA[SG_K][SG_T + 4];      // [k][m]  (transposed on the way in)
  __shared__ __align__(16) float sB[SG_K][SG_T];          // [k][n]
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;                 // 16 x 16 threads, 4 x 4 outputs each
  const int m0 = blockIdx.y * SG_T, n0 = blockIdx.x * SG_T;
  uint64_t acc2[4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc2[i][0] = acc2[i][1] = f2_pack(0.f, 0.f);
  const int a_row = tid >> 2, a_k4 = (tid & 3) * 4;       // A tile: 64 rows x 16 k, one float4 per thread
  const int b_k = tid >> 4, b_n4 = (tid & 15) * 4;        // B tile: 16 k x 64 n, one float4 per thread
  for (int k0 = 0; k0 < K; k0 += SG_K) {
    float4 av = make_float4(0.f, 0.f, 0.f, 0.f), bv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (m0 + a_row < M) av = *reinterpret_cast<const float4*>(A + int64_t(m0 + a_row) * lda + k0 + a_k4);
    if (n0 + b_n4 < N) bv = __ldg(reinterpret_cast<const float4*>(Bt + int64_t(k0 + b_k) * ldb + n0 + b_n4));
    __syncthreads();
    sA[a_k4 + 0][a_row] = av.x;
    sA[a_k4 + 1][a_row] = av.y;
    sA[a_k4 + 2][a_row] = av.z;
    sA[a_k4 + 3][a_row] = av.w;
    *reinterpret_cast<float4*>(&sB[b_k][b_n4]) = bv;
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < SG_K; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&sA[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&sB[kk][tx * 4]);
      const float ar[4] = {a.x, a.y, a.z, a.w};
      const uint64_t b01 = f2_pack(b.x, b.y), b23 = f2_pack(b.z, b.w);
#pragma unroll
      for (int i = 0; i < 4; ++i) {                // packed FFMA2: two independent fma.rn per instruction, same rounding
        const uint64_t aa = f2_splat(ar[i]);
        acc2[i][0] = f2_fma(aa, b01, acc2[i][0]);
        acc2[i][1] = f2_fma(aa, b23, acc2[i][1]);
      }
    }
  }
  if (n0 + tx * 4 < N) {
    const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + n0 + tx * 4));
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f2_unpack(acc2[i][0], acc[i][0], acc[i][1]);
      f2_unpack(acc2[i][1], acc[i][2], acc[i][3]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int m = m0 + ty * 4 + i;
      if (m < M)
        *reinterpret_cast<float4*>(C + int64_t(m) * ldc + n0 + tx * 4) =
            make_float4(acc[i][0] + bb.x, acc[i][1] + bb.y, acc[i][2] + bb.z, acc[i][3] + bb.w);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// the search: tcgen05 / TMEM / TMA
// ------------------------------------------------------------------------------------------------
constexpr int RS_ROWS = 128;                       // tokens per CTA = TMEM lanes
constexpr int RS_ROW_WARPS = 8;                    // two threads per token: warp w owns lanes 32 (w % 4).., column half w / 4
constexpr int RS_THREADS = (RS_ROW_WARPS + 2) * 32;   // + TMA producer warp + MMA issuer warp
constexpr int RS_STAGES = 5;
constexpr int RS_QN = 128;                         // codes per distance block (a quarter of the codebook)
constexpr uint32_t RS_A_ATOM = 128 * 128;          // [128 tokens x 64 dims] bf16, 128-byte rows, 8-row swizzle atoms
constexpr uint32_t RS_A_PLANE = 4 * RS_A_ATOM;     // 256 dims
constexpr uint32_t RS_B_TILE = RS_QN * 128;        // [128 codes x 64 dims] bf16
constexpr uint32_t RS_XCH = RS_ROWS * 32;          // per-token exchange between its two threads (8 floats)
constexpr uint32_t RS_TAB = 2 * RVQ_K * 4;         // |e_j|^2 and |e_j| of the current level
constexpr uint32_t RS_FALL = 4 * RVQ_DC * 4;       // one staged residual row per deciding warp (cooperative re-evaluation)
constexpr size_t RS_SMEM = 1024 + 2 * RS_A_PLANE + RS_STAGES * RS_B_TILE + RS_XCH + RS_TAB + RS_FALL + 256;
constexpr uint32_t RS_COL_D = 0, RS_COL_R = 256;   // TMEM columns: two distance blocks of 128 / the fp32 residual
// |score error| <= RS_KAPPA * |r| |e_j|: the three dropped / rounded terms of the split are each <= 2^-18 |r_c||e_c| per
// dimension, i.e. <= 1.15e-5 |r||e| on the dot product (Cauchy-Schwarz), x 2 in the score, + the fp32 accumulation of 768
// products in the tensor core; 3e-5 leaves a third of margin.
constexpr float RS_KAPPA = 3.0e-5f;
static_assert(RS_SMEM <= 232448, "shared memory");

struct RsParams {
  const float* x;            // [n_rows, ldx] fp32: project_in output (or the codes themselves, RVQ:258-357)
  int ldx;
  const int32_t* lengths;    // [batch] or null
  int tmax, n_rows, n_q;
  const float* code;         // [Q][512][256] fp32
  const float* code_sq;      // [Q][512]
  int64_t* indices;          // [n_rows, n_q]
  float* code_sum;           // [n_rows, 256] fp32 or null: sum over levels of the selected codes (RVQ:455-456)
};

// Smallest three lower bounds (with the codes of the first two) and the smallest upper bound seen so far, branch-free.
struct RsTop {
  float lo1, lo2, lo3, up;
  int j1, j2;
  __device__ void init() {
    lo1 = lo2 = lo3 = up = INFINITY;
    j1 = j2 = 0;
  }
  __device__ __forceinline__ void push(float lo, float hi, int j) {
    up = fminf(up, hi);
    const bool a = lo < lo1, b = lo < lo2;
    lo3 = fminf(lo3, fmaxf(lo, lo2));               // third smallest of {lo1, lo2, lo3, lo}
    lo3 = b ? lo2 : lo3;
    const float n2 = a ? lo1 : lo;
    const int m2 = a ? j1 : j;
    lo2 = b ? n2 : lo2;
    j2 = b ? m2 : j2;
    lo1 = a ? lo : lo1;
    j1 = a ? j : j1;
  }
  __device__ void merge(const RsTop& o) {           // the other thread's list: at most its first three matter
    const float olo3 = o.lo3;
    push(o.lo1, o.up, o.j1);
    push(o.lo2, INFINITY, o.j2);
    lo3 = fminf(lo3, olo3);
  }
};

__global__ void __launch_bounds__(RS_THREADS, 1)
rvq_search_kernel(const __grid_constant__ CUtensorMap tma_code, const RsParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sAhi = smem;
  uint8_t* sAlo = smem + RS_A_PLANE;
  uint8_t* sB = smem + 2 * RS_A_PLANE;
  float* sX = reinterpret_cast<float*>(sB + size_t(RS_STAGES) * RS_B_TILE);          // [128 tokens][8]
  float* sE2 = sX + RS_XCH / 4;                                                      // [512] |e_j|^2 of the level
  float* sEn = sE2 + RVQ_K;                                                          // [512] |e_j|
  float* sFall = sEn + RVQ_K;                                                        // [4 warps][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sFall + 4 * RVQ_DC);
  uint64_t* b_full = bars;                      // [RS_STAGES]
  uint64_t* b_empty = bars + RS_STAGES;         // [RS_STAGES]
  uint64_t* a_full = bars + 2 * RS_STAGES;      // residual planes of a level written      (rows -> MMA), 8 warp arrivals
  uint64_t* d_full = a_full + 1;                // [2] distance block complete               (MMA -> rows)
  uint64_t* d_free = a_full + 3;                // [2] distance block read                   (rows -> MMA), 8 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 5);
  constexpr int kTmaWarp = RS_ROW_WARPS, kMmaWarp = RS_ROW_WARPS + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tma_code);
    for (int s = 0; s < RS_STAGES; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    mbar_init(a_full, RS_ROW_WARPS);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&d_full[i], 1);
      mbar_init(&d_free[i], RS_ROW_WARPS);
    }
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == kTmaWarp) {
    // ===================== TMA producer: codebook planes, in the order the MMA warp consumes them =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int q = 0; q < p.n_q; ++q)
      for (int blk = 0; blk < RVQ_K / RS_QN; ++blk)
        for (int kc = 0; kc < 4; ++kc)
          for (int plane = 0; plane < 2; ++plane) {
            mbar_wait_relaxed(&b_empty[stage], phase ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&b_full[stage], RS_B_TILE);
              tma_load_2d(sB + size_t(stage) * RS_B_TILE, &tma_code, &b_full[stage], kc * 64,
                          (q * 2 + plane) * RVQ_K + blk * RS_QN);
            }
            __syncwarp();
            if (++stage == RS_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc(RS_ROWS, RS_QN, /*bf16*/ 1, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int g = 0;                                   // distance block counter over the whole kernel; buffer g & 1
    for (int q = 0; q < p.n_q; ++q) {
      mbar_wait(a_full, uint32_t(q & 1));
      tc_fence_after();
      for (int blk = 0; blk < RVQ_K / RS_QN; ++blk, ++g) {
        const int buf = g & 1;
        if (g >= 2) {
          mbar_wait(&d_free[buf], uint32_t(((g >> 1) - 1) & 1));
          tc_fence_after();
        }
        const uint32_t d_tmem = tmem_base + RS_COL_D + uint32_t(buf * RS_QN);
        for (int kc = 0; kc < 4; ++kc)
          for (int plane = 0; plane < 2; ++plane) {
            mbar_wait(&b_full[stage], phase);
            tc_fence_after();
            const uint64_t db = umma_desc_k_sw128(smem_u32(sB + size_t(stage) * RS_B_TILE));
            const uint64_t da_hi = umma_desc_k_sw128(smem_u32(sAhi + size_t(kc) * RS_A_ATOM));
            const uint64_t da_lo = umma_desc_k_sw128(smem_u32(sAlo + size_t(kc) * RS_A_ATOM));
            if (elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)          // r_hi . e_{hi|lo}
                umma_ss(d_tmem, da_hi + uint64_t(k * 2), db + uint64_t(k * 2), idesc, (kc | plane | k) != 0 ? 1u : 0u);
              if (plane == 0) {
#pragma unroll
                for (int k = 0; k < 4; ++k)        // r_lo . e_hi
                  umma_ss(d_tmem, da_lo + uint64_t(k * 2), db + uint64_t(k * 2), idesc, 1u);
              }
              umma_commit(&b_empty[stage]);
              if (kc == 3 && plane == 1) umma_commit(&d_full[buf]);
            }
            __syncwarp();
            if (++stage == RS_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
      }
    }
  } else {
    // ===================== two threads per token =====================
    const int hcol = warp >> 2;                     // which half of every column range this thread owns
    const int row_in = (warp & 3) * 32 + lane;
    const int row = blockIdx.x * RS_ROWS + row_in;
    const bool in_range = row < p.n_rows;
    bool valid = in_range;
    if (in_range && p.lengths) {
      const int b = row / p.tmax;
      valid = (row - b * p.tmax) < __ldg(p.lengths + b);
    }
    const uint32_t lane_base = tmem_base + (uint32_t((warp & 3) * 32) << 16);
    const uint32_t t_d = lane_base + RS_COL_D, t_r = lane_base + RS_COL_R;
    float* xch = sX + row_in * 8;
    // bf16 split of 32 residual values -> the swizzled K-major planes (chunk c of the row: dims 32 c .. 32 c + 31)
    auto write_planes = [&](int c, const float (&v)[32]) {
      uint8_t* hi = sAhi + size_t(c >> 1) * RS_A_ATOM + size_t(row_in) * 128;
      uint8_t* lo = sAlo + size_t(c >> 1) * RS_A_ATOM + size_t(row_in) * 128;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        uint32_t hw[4], lw[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float a = v[8 * u + 2 * e], b = v[8 * u + 2 * e + 1];
          const float ah = __bfloat162float(__float2bfloat16_rn(a)), bh = __bfloat162float(__float2bfloat16_rn(b));
          hw[e] = pack_true_bf16x2(ah, bh);
          lw[e] = pack_true_bf16x2(a - ah, b - bh);
        }
        const int chunk = ((c & 1) * 4 + u) ^ (row_in & 7);
        *reinterpret_cast<uint4*>(hi + chunk * 16) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
        *reinterpret_cast<uint4*>(lo + chunk * 16) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
      }
    };
    // |r|^2 of this thread's 128 dims: 32 strided partial sums (chunk-ascending fma chains), then a fixed tree
    float x2p[32];
    auto x2_half = [&]() {
      float a[16], b[8], c4[4];
#pragma unroll
      for (int l = 0; l < 16; ++l) a[l] = x2p[l] + x2p[l + 16];
#pragma unroll
      for (int l = 0; l < 8; ++l) b[l] = a[l] + a[l + 8];
#pragma unroll
      for (int l = 0; l < 4; ++l) c4[l] = b[l] + b[l + 4];
      return (c4[0] + c4[2]) + (c4[1] + c4[3]);
    };
    // named barrier among the 256 row threads (the two threads of a token sit in different warps)
    auto rows_sync = [&]() { asm volatile("bar.sync 1, 256;" ::: "memory"); };

    // ---- level 0 residual = x : this thread's dims 128 hcol .. 128 hcol + 127 ----
#pragma unroll
    for (int l = 0; l < 32; ++l) x2p[l] = 0.f;
#pragma unroll 1
    for (int cc = 0; cc < 4; ++cc) {
      const int c = hcol * 4 + cc;
      float v[32];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (in_range) t = *reinterpret_cast<const float4*>(p.x + int64_t(row) * p.ldx + c * 32 + u * 4);
        v[4 * u + 0] = t.x; v[4 * u + 1] = t.y; v[4 * u + 2] = t.z; v[4 * u + 3] = t.w;
      }
      uint32_t w[32];
#pragma unroll
      for (int l = 0; l < 32; ++l) {
        w[l] = __float_as_uint(v[l]);
        x2p[l] = fmaf(v[l], v[l], x2p[l]);
      }
      tmem_st_32x32b_x32(t_r + uint32_t(c * 32), w);
      write_planes(c, v);
    }
    float x2_mine = x2_half();
    tmem_st_wait();
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(a_full);

    int chosen[RVQ_MAXQ];
    int g = 0;
#pragma unroll 1
    for (int q = 0; q < p.n_q; ++q) {
      // |r|^2 = (dims 0..127) + (dims 128..255): exchange the halves; code norms of the level into shared memory
      xch[hcol] = x2_mine;
      const float* e2 = p.code_sq + q * RVQ_K;
      {
        const int t = (warp * 32 + lane) * 2;       // 256 threads x 2 codes
        const float2 y = __ldg(reinterpret_cast<const float2*>(e2 + t));
        *reinterpret_cast<float2*>(sE2 + t) = y;
        *reinterpret_cast<float2*>(sEn + t) = make_float2(sqrtf(y.x), sqrtf(y.y));
      }
      rows_sync();
      const float x2 = xch[0] + xch[1];
      const float kr = RS_KAPPA * sqrtf(x2);        // error bound of a score: kr * |e_j|
      RsTop top[2];                                 // two independent chains (even / odd codes): shorter dependency chains
      top[0].init();
      top[1].init();
#pragma unroll 1
      for (int blk = 0; blk < RVQ_K / RS_QN; ++blk, ++g) {
        const int buf = g & 1;
        mbar_wait(&d_full[buf], uint32_t((g >> 1) & 1));
        tc_fence_after();
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {            // this thread's 64 of the block's 128 codes
          const int col = hcol * 64 + cc * 32;
          uint32_t dv[32];
          tmem_ld_32x32b_x32(t_d + uint32_t(buf * RS_QN + col), dv);
          tmem_ld_wait();
          const int j0 = blk * RS_QN + col;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const float4 y = *reinterpret_cast<const float4*>(sE2 + j0 + 4 * u);       // broadcast reads
            const float4 n = *reinterpret_cast<const float4*>(sEn + j0 + 4 * u);
            const float yy[4] = {y.x, y.y, y.z, y.w}, nn[4] = {n.x, n.y, n.z, n.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              // score = |e|^2 - 2 <r,e>  (+ |r|^2, constant); bounds: score -+ kappa |r| |e|
              const float sc = fmaf(-2.0f, __uint_as_float(dv[4 * u + e]), yy[e]);
              top[e & 1].push(fmaf(-kr, nn[e], sc), fmaf(kr, nn[e], sc), j0 + 4 * u + e);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&d_free[buf]);
      }
      top[0].merge(top[1]);
      // ---- merge with the token's other thread ----
      rows_sync();                                  // everyone has read x2 from the exchange area
      if (hcol == 1) {
        xch[0] = top[0].lo1; xch[1] = top[0].lo2; xch[2] = top[0].lo3; xch[3] = top[0].up;
        xch[4] = __int_as_float(top[0].j1); xch[5] = __int_as_float(top[0].j2);
      }
      rows_sync();
      if (hcol == 0) {
        RsTop o;
        o.lo1 = xch[0]; o.lo2 = xch[1]; o.lo3 = xch[2]; o.up = xch[3];
        o.j1 = __float_as_int(xch[4]); o.j2 = __float_as_int(xch[5]);
        top[0].merge(o);
      }
      // ---- decide (hcol 0 threads; warps 0..3 are uniformly hcol 0) ----
      int best = top[0].j1;
      if (hcol == 0) {
        // survivors: codes whose lower bound does not exceed the smallest upper bound.  One survivor IS the fp32 arg-min.
        const bool two = valid && top[0].lo2 <= top[0].up;
        const bool many = valid && top[0].lo3 <= top[0].up;          // three or more: identities beyond two are not kept
        const float* cb = p.code + int64_t(q) * RVQ_K * RVQ_DC;
        if (__any_sync(0xffffffffu, two && !many)) {
          // exact fp32 re-evaluation (VQ:44-48 operation order, first minimum) of the two survivors.  tcgen05.ld is
          // warp-collective, so the loop is warp-uniform and lanes that need nothing ride along.
          float bd = INFINITY;
          int bi = 0;
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            const int j = t == 0 ? top[0].j1 : top[0].j2;
            float dot = 0.f;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              uint32_t rv[32];
              tmem_ld_32x32b_x32(t_r + uint32_t(c * 32), rv);
              tmem_ld_wait();
#pragma unroll
              for (int u = 0; u < 8; ++u) {
                const float4 ev = __ldg(reinterpret_cast<const float4*>(cb + int64_t(j) * RVQ_DC + c * 32 + 4 * u));
                dot = fmaf(__uint_as_float(rv[4 * u + 3]), ev.w,
                           fmaf(__uint_as_float(rv[4 * u + 2]), ev.z,
                                fmaf(__uint_as_float(rv[4 * u + 1]), ev.y, fmaf(__uint_as_float(rv[4 * u + 0]), ev.x, dot))));
              }
            }
            const float sx = __fadd_rn(__fadd_rn(x2, sE2[j]), __fmul_rn(dot, -2.0f));
            const float dj = __fsqrt_rn(fmaxf(sx, 0.f));
            if (dj < bd || (dj == bd && j < bi)) {
              bd = dj;
              bi = j;
            }
          }
          if (two && !many) best = bi;
        }
        // Three or more survivors (identities beyond two are not kept; ~1e-5 of the tokens on well-scaled inputs): the
        // whole warp re-evaluates that token exactly over all 512 codes, 16 codes per lane, the token's residual staged
        // in shared memory.  Same formula, operation order and tie-break.
        unsigned todo = __ballot_sync(0xffffffffu, many);
        if (todo) {
          float* fr = sFall + (warp & 3) * RVQ_DC;
          while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
#pragma unroll 1
            for (int c = 0; c < 8; ++c) {
              uint32_t rv[32];
              tmem_ld_32x32b_x32(t_r + uint32_t(c * 32), rv);
              tmem_ld_wait();
              if (lane == src) {
#pragma unroll
                for (int l = 0; l < 32; ++l) fr[c * 32 + l] = __uint_as_float(rv[l]);
              }
            }
            __syncwarp();
            const float x2s = __shfl_sync(0xffffffffu, x2, src);
            float bd = INFINITY;
            int bi = 0;
#pragma unroll 1
            for (int jj = 0; jj < RVQ_K / 32; ++jj) {
              const int j = jj * 32 + lane;
              const float4* er = reinterpret_cast<const float4*>(cb + int64_t(j) * RVQ_DC);
              float dot = 0.f;
#pragma unroll 8
              for (int c4 = 0; c4 < RVQ_DC / 4; ++c4) {
                const float4 ev = __ldg(er + c4);
                const float4 rr = *reinterpret_cast<const float4*>(fr + 4 * c4);
                dot = fmaf(rr.w, ev.w, fmaf(rr.z, ev.z, fmaf(rr.y, ev.y, fmaf(rr.x, ev.x, dot))));
              }
              const float sx = __fadd_rn(__fadd_rn(x2s, sE2[j]), __fmul_rn(dot, -2.0f));
              const float dj = __fsqrt_rn(fmaxf(sx, 0.f));
              if (dj < bd) {                         // j ascends within a lane: strict < keeps the first minimum
                bd = dj;
                bi = j;
              }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const float od = __shfl_xor_sync(0xffffffffu, bd, o);
              const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
              if (od < bd || (od == bd && oi < bi)) {
                bd = od;
                bi = oi;
              }
            }
            if (lane == src) best = bi;
            __syncwarp();
          }
        }
        __syncwarp();
        xch[7] = __int_as_float(best);
        if (in_range) p.indices[int64_t(row) * p.n_q + q] = valid ? int64_t(best) : int64_t(-1);          // VQ:1205-1210
      }
      rows_sync();
      best = __float_as_int(xch[7]);
      chosen[q] = best;
      // ---- gather, residual -= code (masked rows: zero code, VQ:1192-1203), planes of the next level ----
      if (q + 1 < p.n_q) {
        const float* cv = p.code + (int64_t(q) * RVQ_K + best) * RVQ_DC;
#pragma unroll
        for (int l = 0; l < 32; ++l) x2p[l] = 0.f;
#pragma unroll 1
        for (int cc = 0; cc < 4; ++cc) {
          const int c = hcol * 4 + cc;
          uint32_t rv[32];
          tmem_ld_32x32b_x32(t_r + uint32_t(c * 32), rv);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            float4 ev = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) ev = __ldg(reinterpret_cast<const float4*>(cv + c * 32 + 4 * u));
            v[4 * u + 0] = __uint_as_float(rv[4 * u + 0]) - ev.x;
            v[4 * u + 1] = __uint_as_float(rv[4 * u + 1]) - ev.y;
            v[4 * u + 2] = __uint_as_float(rv[4 * u + 2]) - ev.z;
            v[4 * u + 3] = __uint_as_float(rv[4 * u + 3]) - ev.w;
          }
#pragma unroll
          for (int l = 0; l < 32; ++l) {
            rv[l] = __float_as_uint(v[l]);
            x2p[l] = fmaf(v[l], v[l], x2p[l]);
          }
          tmem_st_32x32b_x32(t_r + uint32_t(c * 32), rv);
          write_planes(c, v);
        }
        x2_mine = x2_half();
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_full);
      }
    }
    // ---- sum of the selected codes, level order (RVQ:455-456): the input of project_out; dims split as above ----
    if (p.code_sum && in_range) {
      float* dst = p.code_sum + int64_t(row) * RVQ_DC + hcol * 128;
#pragma unroll 1
      for (int c4 = 0; c4 < 32; ++c4) {
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
          for (int q = 0; q < p.n_q; ++q) {
            int j = 0;
#pragma unroll
            for (int t = 0; t < RVQ_MAXQ; ++t)
              if (t == q) j = chosen[t];
            const float4 ev =
                __ldg(reinterpret_cast<const float4*>(p.code + (int64_t(q) * RVQ_K + j) * RVQ_DC + hcol * 128) + c4);
            a.x += ev.x; a.y += ev.y; a.z += ev.z; a.w += ev.w;
          }
        }
        reinterpret_cast<float4*>(dst)[c4] = a;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) tmem_dealloc(tmem_base, 512);
}

// get_output_from_indices / get_code_from_indices (RVQ:183-242): one CTA per 16 rows.
__global__ void __launch_bounds__(RVQ_THREADS)
rvq_decode_kernel(const int64_t* __restrict__ indices, int n_rows, int n_q, int d_model, const float* __restrict__ code,
                  const float* __restrict__ wout_t, const float* __restrict__ bout, int project_out,
                  float* __restrict__ out) {
  __shared__ float s_acc[RVQ_ROWS][RVQ_DC];
  const int tid = threadIdx.x;
  const int row0 = blockIdx.x * RVQ_ROWS;
#pragma unroll
  for (int r = 0; r < RVQ_ROWS; ++r) {
    float a = 0.f;
    if (row0 + r < n_rows) {
      for (int q = 0; q < n_q; ++q) {
        const int64_t i = indices[int64_t(row0 + r) * n_q + q];
        if (i >= 0 && i < RVQ_K) a += __ldg(code + (int64_t(q) * RVQ_K + i) * RVQ_DC + tid);   // -1 -> zero code
      }
    }
    s_acc[r][tid] = a;
  }
  __syncthreads();
  if (!project_out) {
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r)
      if (row0 + r < n_rows) out[int64_t(row0 + r) * RVQ_DC + tid] = s_acc[r][tid];
    return;
  }
  for (int c0 = 0; c0 < d_model; c0 += RVQ_THREADS) {
    const int col = c0 + tid;
    if (col >= d_model) break;
    float acc[RVQ_ROWS];
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r) acc[r] = 0.f;
#pragma unroll 4
    for (int k = 0; k < RVQ_DC; ++k) {
      const float wv = __ldg(wout_t + int64_t(k) * d_model + col);
#pragma unroll
      for (int r = 0; r < RVQ_ROWS; ++r) acc[r] = fmaf(s_acc[r][k], wv, acc[r]);
    }
    const float bv = __ldg(bout + col);
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r)
      if (row0 + r < n_rows) out[int64_t(row0 + r) * d_model + col] = acc[r] + bv;
  }
}

static int check_rvq(const taste_weights_t& w) {
  if (w.dims.codebook_dim != RVQ_DC || w.dims.codebook_size != RVQ_K || w.dims.num_quantizers > RVQ_MAXQ ||
      w.dims.num_quantizers < 1)
    return set_error(TASTE_E_SHAPE, "rvq: kernel is built for codebook_dim 256, codebook_size 512, <= 8 levels");
  if (!w.rvq_win_t || !w.rvq_bin || !w.rvq_code || !w.rvq_code_sq || !w.rvq_wout_t || !w.rvq_bout)
    return set_error(TASTE_E_ARG, "rvq: weights missing from the handle");
  return 0;
}

size_t rvq_ws_bytes(int n_rows) { return size_t(n_rows > 0 ? n_rows : 0) * RVQ_DC * sizeof(float) * 2 + 256; }

static int launch_sgemm(const float* A, int lda, const float* Bt, int ldb, const float* bias, float* C, int ldc, int M,
                        int N, int K, cudaStream_t stream) {
  if (K % SG_K != 0 || N % 4 != 0 || lda % 4 != 0 || ldb % 4 != 0 || ldc % 4 != 0)
    return set_error(TASTE_E_SHAPE, "rvq projection: K must be a multiple of 16 and N / strides multiples of 4");
  dim3 grid((N + SG_T - 1) / SG_T, (M + SG_T - 1) / SG_T);
  ProfScope ps(stream, KC_RVQ_PROJ, 2.0 * M * double(N) * K, 4.0 * (double(M) * K + double(K) * N + double(M) * N));
  rvq_sgemm_kernel<<<grid, 256, 0, stream>>>(A, lda, Bt, ldb, bias, C, ldc, M, N, K);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_rvq_encode(const taste_weights_t& w, const float* z, const int32_t* lengths, int batch, int tmax, int in_dim,
                      int64_t* indices, float* quantized, void* ws, size_t ws_bytes, cudaStream_t stream) {
  if (int rc = check_rvq(w)) return rc;
  if (!z || !indices) return set_error(TASTE_E_ARG, "rvq_encode: null pointer");
  if (!w.rvq_code_split) return set_error(TASTE_E_ARG, "rvq_encode: split codebook planes missing from the handle");
  if (in_dim != w.dims.d_model && in_dim != RVQ_DC) return set_error(TASTE_E_SHAPE, "rvq_encode: in_dim must be d_model or 256");
  const int n_rows = batch * tmax;
  if (n_rows <= 0) return 0;
  if (!ws || ws_bytes < rvq_ws_bytes(n_rows))
    return set_error(TASTE_E_WORKSPACE, "rvq_encode: workspace %zu < %zu", ws_bytes, rvq_ws_bytes(n_rows));
  float* x = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~uintptr_t(255));
  float* code_sum = x + size_t(n_rows) * RVQ_DC;
  const int n_q = w.dims.num_quantizers;
  int rc;
  const float* x_in = z;
  int ldx = in_dim;
  if (in_dim != RVQ_DC) {                                                                               // RVQ:371
    if ((rc = launch_sgemm(z, in_dim, w.rvq_win_t, RVQ_DC, w.rvq_bin, x, RVQ_DC, n_rows, RVQ_DC, in_dim, stream))) return rc;
    x_in = x;
    ldx = RVQ_DC;
  }
  EncodeTiledFn enc = get_tensor_map_encoder();
  if (!enc) return set_error(TASTE_E_NO_DEVICE, "cuTensorMapEncodeTiled entry point unavailable");
  CUtensorMap tm;
  {
    cuuint64_t dims[2] = {(cuuint64_t)RVQ_DC, (cuuint64_t)n_q * 2 * RVQ_K};
    cuuint64_t strides[1] = {(cuuint64_t)RVQ_DC * 2};
    cuuint32_t box[2] = {64, RS_QN};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w.rvq_code_split), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error((int)r, "rvq: tensor map encode failed (%d)", (int)r);
  }
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    TASTE_CUDA_OK(cudaFuncSetAttribute(rvq_search_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_SMEM));
    configured = true;
  }
  RsParams p;
  p.x = x_in;
  p.ldx = ldx;
  p.lengths = lengths;
  p.tmax = tmax;
  p.n_rows = n_rows;
  p.n_q = n_q;
  p.code = w.rvq_code;
  p.code_sq = w.rvq_code_sq;
  p.indices = indices;
  p.code_sum = quantized ? code_sum : nullptr;
  {
    // algorithmic work: the three bf16 products of every distance GEMM; bytes: residual in, indices (+ code sums) out
    ProfScope ps(stream, KC_RVQ_ENCODE, 3.0 * 2.0 * RVQ_DC * RVQ_K * double(n_q) * n_rows,
                 double(n_rows) * (4.0 * RVQ_DC + 8.0 * n_q + (quantized ? 4.0 * RVQ_DC : 0.0)));
    rvq_search_kernel<<<(n_rows + RS_ROWS - 1) / RS_ROWS, RS_THREADS, RS_SMEM, stream>>>(tm, p);
    TASTE_CUDA_OK(cudaGetLastError());
  }
  if (quantized)                                                                                         // RVQ:470
    if ((rc = launch_sgemm(code_sum, RVQ_DC, w.rvq_wout_t, w.dims.d_model, w.rvq_bout, quantized, w.dims.d_model, n_rows,
                           w.dims.d_model, RVQ_DC, stream)))
      return rc;
  return 0;
}

int launch_rvq_decode(const taste_weights_t& w, const int64_t* indices, int n, bool project_out, float* out,
                      cudaStream_t stream) {
  if (int rc = check_rvq(w)) return rc;
  if (!indices || !out) return set_error(TASTE_E_ARG, "rvq_decode: null pointer");
  if (n <= 0) return 0;
  const int blocks = (n + RVQ_ROWS - 1) / RVQ_ROWS;
  ProfScope ps(stream, KC_RVQ_DECODE, project_out ? 2.0 * RVQ_DC * w.dims.d_model * n : 0.0,
               double(n) * (8.0 * w.dims.num_quantizers + 4.0 * (project_out ? w.dims.d_model : RVQ_DC)));
  rvq_decode_kernel<<<blocks, RVQ_THREADS, 0, stream>>>(indices, n, w.dims.num_quantizers, w.dims.d_model, w.rvq_code,
                                                        w.rvq_wout_t, w.rvq_bout, project_out ? 1 : 0, out);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace taste
