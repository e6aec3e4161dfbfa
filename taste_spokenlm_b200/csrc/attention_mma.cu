// Segmented (ragged) flash attention, head_dim 64, bf16 in / fp32 softmax / bf16 out.
//
// One kernel serves the three attention shapes of the path (CW:377-394 eager semantics, no dropout):
//   - encoder self-attention: 1500 x 1500, no mask (JES:198-203)                      [fixed lengths]
//   - aggregator causal self-attention over T'_b <= 448 assembled tokens (CW:786-797)  [ragged, causal]
//   - aggregator cross-attention: T'_b queries x 1500 frames, keys from the final encoder state and values from the
//     layer-6 input (CW:361-366, JES:377-388), no encoder mask                         [ragged queries]
// q is pre-scaled by head_dim^-0.5 (folded into W_q at pack time, CW:342).  Scores never touch HBM: 64 x 64 tiles,
// warp-level online softmax in registers (each warp owns 16 query rows), K/V tiles double-buffered with cp.async.
#include "common.cuh"
#include "internal.h"

namespace taste {

constexpr int ATT_BQ = 64;       // queries per CTA (4 warps x 16)
constexpr int ATT_BK = 64;       // keys per tile
constexpr int ATT_HD = 64;
constexpr int ATT_THREADS = 128;

struct AttnParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* o;
  int ldq, ldk, ldv, ldo;
  const int32_t* cu_q;
  const int32_t* cu_kv;
  int q_len, kv_len;
  int causal;
};

TASTE_DEVINL void cp_async_16(uint32_t smem, const void* gmem, bool valid) {
  const int sz = valid ? 16 : 0;   // src-size 0 => 16 zero bytes
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem), "l"(gmem), "r"(sz) : "memory");
}
TASTE_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
TASTE_DEVINL void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

TASTE_DEVINL void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
TASTE_DEVINL void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
TASTE_DEVINL void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      #if TASTE_F16
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
#else
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
#endif
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// tile = 64 rows x 64 bf16 (128 B per row); 16-byte chunk c of row r lives at r*128 + ((c ^ (r & 7)) * 16)
TASTE_DEVINL uint32_t tile_off(int row, int chunk) { return uint32_t(row * 128 + ((chunk ^ (row & 7)) << 4)); }

TASTE_DEVINL void load_tile(uint32_t smem_base, const __nv_bfloat16* g, int ld, int row0, int rows_valid, int tid) {
  // 64 rows x 8 chunks = 512 chunks, 4 per thread
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * ATT_THREADS;
    const int r = idx >> 3;
    const int c = idx & 7;
    const bool ok = r < rows_valid;
    const __nv_bfloat16* src = g + int64_t(row0 + (ok ? r : 0)) * ld + c * 8;
    cp_async_16(smem_base + tile_off(r, c), src, ok);
  }
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_mma_kernel(const AttnParams p) {
  __shared__ __align__(128) uint8_t smem[ATT_BQ * 128 + 2 * 2 * ATT_BK * 128];   // Q, K[2], V[2] = 40 KB
  const int tid = threadIdx.x;
  const int warp = tid >> 5;
  const int lane = tid & 31;
  const int b = blockIdx.z;
  const int head = blockIdx.y;

  const int q_start = p.cu_q ? p.cu_q[b] : b * p.q_len;
  const int q_len = p.cu_q ? (p.cu_q[b + 1] - q_start) : p.q_len;
  const int kv_start = p.cu_kv ? p.cu_kv[b] : b * p.kv_len;
  const int kv_len = p.cu_kv ? (p.cu_kv[b + 1] - kv_start) : p.kv_len;
  const int q0 = blockIdx.x * ATT_BQ;
  if (q0 >= q_len) return;
  const int q_rows = min(ATT_BQ, q_len - q0);

  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK0 = sQ + ATT_BQ * 128;
  const uint32_t sV0 = sK0 + 2 * ATT_BK * 128;

  const __nv_bfloat16* gq = p.q + head * ATT_HD;
  const __nv_bfloat16* gk = p.k + head * ATT_HD;
  const __nv_bfloat16* gv = p.v + head * ATT_HD;

  // keys this CTA needs: all kv (non-causal) or up to the last query of the tile (causal)
  const int kv_end = p.causal ? min(kv_len, q0 + q_rows) : kv_len;
  const int n_tiles = (kv_end + ATT_BK - 1) / ATT_BK;

  load_tile(sQ, gq, p.ldq, q_start + q0, q_rows, tid);
  load_tile(sK0, gk, p.ldk, kv_start, min(ATT_BK, kv_end), tid);
  load_tile(sV0, gv, p.ldv, kv_start, min(ATT_BK, kv_end), tid);
  cp_async_commit();

  float o_acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  uint32_t qf[4][4];
  const float kLog2e = 1.4426950408889634f;

  for (int t = 0; t < n_tiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < n_tiles) {
      const int k0n = (t + 1) * ATT_BK;
      const int rows = min(ATT_BK, kv_end - k0n);
      load_tile(sK0 + (buf ^ 1) * ATT_BK * 128, gk, p.ldk, kv_start + k0n, rows, tid);
      load_tile(sV0 + (buf ^ 1) * ATT_BK * 128, gv, p.ldv, kv_start + k0n, rows, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (t == 0) {
      // Q fragments: warp rows w*16..+15, 4 k-steps of 16
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const int r = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = ks * 2 + (lane >> 4);
        ldsm_x4(sQ + tile_off(r, c), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
      }
    }
    const uint32_t sK = sK0 + buf * ATT_BK * 128;
    const uint32_t sV = sV0 + buf * ATT_BK * 128;
    const int k0 = t * ATT_BK;

    // ---- S = Q K^T : 16 x 64 per warp = 8 n-tiles ----
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {       // pairs of n-tiles (16 keys)
        uint32_t b0, b1, b2, b3;
        const int key = np * 16 + (lane & 7) + (lane >> 4) * 8;
        const int c = ks * 2 + ((lane >> 3) & 1);
        ldsm_x4(sK + tile_off(key, c), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
        mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
      }
    }

    // ---- masking ----
    const int row_a = q0 + warp * 16 + (lane >> 2);     // local query index (within the utterance) of c0/c1
    const bool need_mask = (k0 + ATT_BK > kv_end) || (p.causal && (k0 + ATT_BK > q0 + warp * 16));
    if (need_mask) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int col = k0 + i * 8 + (lane & 3) * 2 + (e & 1);
          const int row = row_a + (e >> 1) * 8;
          const bool dead = (col >= kv_end) || (p.causal && col > row);
          if (dead) s[i][e] = -INFINITY;
        }
      }
    }

    // ---- online softmax (rows row_a and row_a + 8) ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float scale[2], mnew_l2[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float mnew = fmaxf(m_run[r], mx[r]);
      const float msafe = (mnew == -INFINITY) ? 0.f : mnew;
      scale[r] = fast_exp2((m_run[r] - msafe) * kLog2e);      // m_run = -inf -> 0
      m_run[r] = mnew;
      mnew_l2[r] = msafe * kLog2e;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];       // P as A fragments for 4 k-steps of 16 keys
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = fast_exp2(fmaf(s[i][0], kLog2e, -mnew_l2[0]));
      const float p1 = fast_exp2(fmaf(s[i][1], kLog2e, -mnew_l2[0]));
      const float p2 = fast_exp2(fmaf(s[i][2], kLog2e, -mnew_l2[1]));
      const float p3 = fast_exp2(fmaf(s[i][3], kLog2e, -mnew_l2[1]));
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      const int ks = i >> 1;
      if ((i & 1) == 0) {
        pf[ks][0] = pack_act2(p0, p1);
        pf[ks][1] = pack_act2(p2, p3);
      } else {
        pf[ks][2] = pack_act2(p0, p1);
        pf[ks][3] = pack_act2(p2, p3);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * scale[r] + rs[r];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o_acc[i][0] *= scale[0];
      o_acc[i][1] *= scale[0];
      o_acc[i][2] *= scale[1];
      o_acc[i][3] *= scale[1];
    }

    // ---- O += P V ----
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {       // pairs of d n-tiles (16 dims)
        uint32_t b0, b1, b2, b3;
        const int key = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int c = np * 2 + (lane >> 4);
        ldsm_x4_trans(sV + tile_off(key, c), b0, b1, b2, b3);
        mma_bf16_16816(o_acc[2 * np], pf[ks], b0, b1);
        mma_bf16_16816(o_acc[2 * np + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();     // everyone done with buf before it is refilled two iterations later
  }

  // ---- finalise ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  const int r0 = warp * 16 + (lane >> 2);
  __nv_bfloat16* go = p.o + head * ATT_HD;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int col = i * 8 + (lane & 3) * 2;
    if (r0 < q_rows) {
      *reinterpret_cast<uint32_t*>(go + int64_t(q_start + q0 + r0) * p.ldo + col) =
          pack_act2(o_acc[i][0] * inv0, o_acc[i][1] * inv0);
    }
    if (r0 + 8 < q_rows) {
      *reinterpret_cast<uint32_t*>(go + int64_t(q_start + q0 + r0 + 8) * p.ldo + col) =
          pack_act2(o_acc[i][2] * inv1, o_acc[i][3] * inv1);
    }
  }
}

static int g_attn_mode = 0;
void set_attention_mode(int mode) { g_attn_mode = mode; }

int launch_attention(const AttnDesc& d, cudaStream_t stream) {
  if (!d.q || !d.k || !d.v || !d.o) return set_error(TASTE_E_ARG, "attention: null pointer");
  if (d.batch <= 0 || d.heads <= 0 || d.q_len <= 0 || d.kv_len <= 0) return 0;
  if (g_attn_mode == 0 && attention_tcgen05_eligible(d)) return launch_attention_tcgen05(d, stream);
  if ((d.ldq | d.ldk | d.ldv) % 8 != 0 || d.ldo % 2 != 0)
    return set_error(TASTE_E_SHAPE, "attention: row strides must be multiples of 8 elements");
  AttnParams p;
  p.q = static_cast<const __nv_bfloat16*>(d.q);
  p.k = static_cast<const __nv_bfloat16*>(d.k);
  p.v = static_cast<const __nv_bfloat16*>(d.v);
  p.o = static_cast<__nv_bfloat16*>(d.o);
  p.ldq = d.ldq; p.ldk = d.ldk; p.ldv = d.ldv; p.ldo = d.ldo;
  p.cu_q = d.cu_q; p.cu_kv = d.cu_kv;
  p.q_len = d.q_len; p.kv_len = d.kv_len;
  p.causal = d.causal;
  dim3 grid((d.q_len + ATT_BQ - 1) / ATT_BQ, d.heads, d.batch);
  const double tq = d.total_q > 0 ? double(d.total_q) : double(d.batch) * d.q_len;
  const double pairs = tq * d.kv_len * (d.causal ? 0.5 : 1.0);                  // causal: about half the tile
  ProfScope ps(stream, d.kclass, 4.0 * pairs * ATT_HD * d.heads,
               2.0 * ATT_HD * d.heads * (2.0 * tq + 2.0 * double(d.batch) * d.kv_len));
  attention_mma_kernel<<<grid, ATT_THREADS, 0, stream>>>(p);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace taste
