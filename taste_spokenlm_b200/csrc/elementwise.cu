// HBM-bound helpers of the path: LayerNorm (CW:660,666,1034,1188), bf16 cast, decoder embedding (CW:1300,1328-1341),
// segmented word-mean pooling (JES:393-458 + MT:170-172) and the extract_vq word-start mapping (MT:1438-1450).
#include "common.cuh"
#include "internal.h"

namespace taste {

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, row held in registers, two-pass statistics in fp32 (eps = 1e-5).
// ------------------------------------------------------------------------------------------------
template <int VEC /* float4 per lane */, bool OUT_BF16>
__global__ void __launch_bounds__(256)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ b,
                 void* __restrict__ y, int rows, int d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + int64_t(warp) * d);
  float4 v[VEC];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx * 4 < d) {
      v[i] = xr[idx];
      sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    } else {
      v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  const float mean = warp_sum(sum) / float(d);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx * 4 < d) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
      sq += (a * a + bb * bb) + (c * c + e * e);
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / float(d) + 1e-5f);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(b);
#pragma unroll
  for (int i = 0; i < VEC; ++i) {
    const int idx = lane + i * 32;
    if (idx * 4 < d) {
      const float4 g = __ldg(w4 + idx);
      const float4 be = __ldg(b4 + idx);
      float4 o;
      o.x = (v[i].x - mean) * rstd * g.x + be.x;
      o.y = (v[i].y - mean) * rstd * g.y + be.y;
      o.z = (v[i].z - mean) * rstd * g.z + be.z;
      o.w = (v[i].w - mean) * rstd * g.w + be.w;
      if (OUT_BF16) {
        uint2 u;
        u.x = pack_act2(o.x, o.y);
        u.y = pack_act2(o.z, o.w);
        reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(y) + int64_t(warp) * d)[idx] = u;
      } else {
        reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + int64_t(warp) * d)[idx] = o;
      }
    }
  }
}

int launch_layernorm(const float* x, const float* w, const float* b, void* y, int rows, int d, bool out_bf16,
                     cudaStream_t stream) {
  if (!x || !w || !b || !y) return set_error(TASTE_E_ARG, "layernorm: null pointer");
  if (d % 4 != 0 || d > 4 * 32 * 16) return set_error(TASTE_E_SHAPE, "layernorm: d=%d unsupported", d);
  if (rows <= 0) return 0;
  const int threads = 256;
  const int blocks = (rows + 7) / 8;
  const int vec = (d / 4 + 31) / 32;
  ProfScope ps(stream, KC_LAYERNORM, 8.0 * rows * d, double(rows) * d * (4.0 + (out_bf16 ? 2.0 : 4.0)));
#define TASTE_LN(V)                                                                           \
  do {                                                                                        \
    if (out_bf16) layernorm_kernel<V, true><<<blocks, threads, 0, stream>>>(x, w, b, y, rows, d);  \
    else layernorm_kernel<V, false><<<blocks, threads, 0, stream>>>(x, w, b, y, rows, d);          \
  } while (0)
  if (vec <= 1) TASTE_LN(1);
  else if (vec <= 2) TASTE_LN(2);
  else if (vec <= 4) TASTE_LN(4);
  else if (vec <= 8) TASTE_LN(8);
  else if (vec <= 10) TASTE_LN(10);
  else TASTE_LN(16);
#undef TASTE_LN
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32 -> bf16 cast (feats for the stem, captured layer-6 input)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float4* __restrict__ x, uint2* __restrict__ y, int64_t n4) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = x[i];
    uint2 u;
    u.x = pack_act2(v.x, v.y);
    u.y = pack_act2(v.z, v.w);
    y[i] = u;
  }
}

int launch_cast_bf16(const float* x, void* y, int64_t n, cudaStream_t stream) {
  if (n % 4 != 0) return set_error(TASTE_E_SHAPE, "cast: n %% 4 != 0");
  if (n == 0) return 0;
  const int64_t n4 = n / 4;
  int blocks = int((n4 + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  ProfScope ps(stream, KC_CAST, 0.0, 6.0 * double(n));
  cast_bf16_kernel<<<blocks, 256, 0, stream>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<uint2*>(y), n4);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// token assembly (MT:144-152), packed: row b = prefix ++ ids[b, :T_b] ++ (ids[b, T_b] if T_b < Tmax else EOS).
// The reference appends EOS after the PADDED width; by causality only the first T_b + 5 entries of a row can reach
// the states the path consumes, and entry T_b + 4 is whatever sits there in the padded row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
assemble_tokens_kernel(const int64_t* __restrict__ ids, const int32_t* __restrict__ lens, const int32_t* __restrict__ cu,
                       int tmax, int32_t* __restrict__ tokens) {
  const int b = blockIdx.x;
  const int T = min(max(lens[b], 0), tmax);
  const int base = cu[b];
  for (int p = threadIdx.x; p < T + 5; p += blockDim.x) {
    int32_t v;
    if (p == 0) v = 50258;
    else if (p == 1) v = 50259;
    else if (p == 2) v = 50360;
    else if (p == 3) v = 50364;
    else if (p - 4 < tmax) v = int32_t(ids[int64_t(b) * tmax + p - 4]);
    else v = 50257;
    tokens[base + p] = v;
  }
}

int launch_assemble_tokens(const int64_t* ids, const int32_t* lens, const int32_t* cu, int batch, int tmax,
                           int32_t* tokens, cudaStream_t stream) {
  if (!ids || !lens || !cu || !tokens) return set_error(TASTE_E_ARG, "assemble_tokens: null pointer");
  if (batch <= 0) return 0;
  ProfScope ps(stream, KC_EMBED, 0.0, 12.0 * double(batch) * (tmax + 5));
  assemble_tokens_kernel<<<batch, 256, 0, stream>>>(ids, lens, cu, tmax, tokens);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// decoder input: E_tok[token] + P_dec[position]  (one warp per assembled row)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
embed_kernel(const int32_t* __restrict__ tokens, const int32_t* __restrict__ cu, int batch, int sum_tokens,
             const float* __restrict__ tok_emb, const float* __restrict__ pos_emb, int d, int vocab, int max_pos,
             float* __restrict__ out) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= sum_tokens) return;
  // utterance of this row: binary search in cu
  int lo = 0, hi = batch;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (cu[mid] <= row) lo = mid; else hi = mid;
  }
  int pos = row - cu[lo];
  if (pos >= max_pos) pos = max_pos - 1;     // guarded by the host (T' <= max_target_positions)
  int tok = tokens[row];
  if (tok < 0 || tok >= vocab) tok = 0;
  const float4* e = reinterpret_cast<const float4*>(tok_emb + int64_t(tok) * d);
  const float4* pe = reinterpret_cast<const float4*>(pos_emb + int64_t(pos) * d);
  float4* o = reinterpret_cast<float4*>(out + int64_t(row) * d);
  for (int i = lane; i < d / 4; i += 32) {
    const float4 a = __ldg(e + i), c = __ldg(pe + i);
    o[i] = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
  }
}

int launch_embed(const int32_t* tokens, const int32_t* cu_tokens, int batch, int sum_tokens, const float* tok_emb,
                 const float* pos_emb, int d, int vocab, int max_pos, float* out, cudaStream_t stream) {
  if (sum_tokens <= 0) return 0;
  const int blocks = (sum_tokens + 7) / 8;
  ProfScope ps(stream, KC_EMBED, double(sum_tokens) * d, 12.0 * double(sum_tokens) * d);
  embed_kernel<<<blocks, 256, 0, stream>>>(tokens, cu_tokens, batch, sum_tokens, tok_emb, pos_emb, d, vocab, max_pos, out);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// word pooling.  For output position (b, t): the run of equal word ids containing t is found on the PADDED row
// (JES:437-458); it is pooled iff run length > 1 and run end <= T_b + 1 (the "+1" is the EOS slot that the
// reference's length still counts when the runs are formed, MT:152 / JES:396).  Source rows are the decoder states
// at assembled positions 4 + t (prefix skipped, JES:393-396).  Rows t >= T_b are zeroed (they are padding; the
// reference leaves decoder garbage there, which no caller reads).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
word_pool_kernel(const float* __restrict__ dec, const int32_t* __restrict__ cu, const int32_t* __restrict__ word_ids,
                 const int32_t* __restrict__ lens, int tmax, int d, float* __restrict__ z) {
  const int b = blockIdx.y;
  const int t = blockIdx.x;
  const int T = lens[b];
  float4* o = reinterpret_cast<float4*>(z + (int64_t(b) * tmax + t) * d);
  if (t >= T) {
    for (int i = threadIdx.x; i < d / 4; i += blockDim.x) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    return;
  }
  const int32_t* wid = word_ids + int64_t(b) * tmax;
  const int w = wid[t];
  int s = t, e = t + 1;
  while (s > 0 && wid[s - 1] == w) --s;
  while (e < tmax && wid[e] == w) ++e;
  const bool pooled = (e - s > 1) && (e <= T + 1);
  const int64_t base = int64_t(cu[b]) + 4;
  if (!pooled) {
    const float4* src = reinterpret_cast<const float4*>(dec + (base + t) * d);
    for (int i = threadIdx.x; i < d / 4; i += blockDim.x) o[i] = src[i];
    return;
  }
  // the run may include position T (the first pad / EOS slot) when e == T + 1: that row exists in the packed
  // decoder output (each utterance carries T + 5 assembled rows)
  const float inv = 1.0f / float(e - s);
  for (int i = threadIdx.x; i < d / 4; i += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = s; r < e; ++r) {
      const float4 v = reinterpret_cast<const float4*>(dec + (base + r) * d)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    o[i] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
  }
}

int launch_word_pool(const float* dec_out, const int32_t* cu_tokens, const int32_t* word_ids,
                     const int32_t* token_lengths, int batch, int tmax, int d, float* z, cudaStream_t stream) {
  if (!dec_out || !cu_tokens || !word_ids || !token_lengths || !z) return set_error(TASTE_E_ARG, "word_pool: null pointer");
  if (d % 4 != 0) return set_error(TASTE_E_SHAPE, "word_pool: d %% 4 != 0");
  if (batch <= 0 || tmax <= 0) return 0;
  dim3 grid(tmax, batch);
  ProfScope ps(stream, KC_WORD_POOL, double(batch) * tmax * d, 8.0 * double(batch) * tmax * d);
  word_pool_kernel<<<grid, 256, 0, stream>>>(dec_out, cu_tokens, word_ids, token_lengths, tmax, d, z);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// extract_vq epilogue (MT:1438-1450, 1877-1881).  One thread per (b, l).
// M[l,t] = same word & both valid.  W1 keeps, per l, the first matching t.  W2[l,t] = M[l,t] & (cumsum_l W1[:,t] == 1).
// llm[l] = sum_t W2[l,t] * asr[t]  - [no t].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
map_llm_kernel(const int64_t* __restrict__ asr_idx, const int32_t* __restrict__ asr_wid, const int32_t* __restrict__ asr_len,
               const int32_t* __restrict__ llm_wid, const int32_t* __restrict__ llm_len, int tmax, int lmax, int nq,
               int64_t* __restrict__ out) {
  const int b = blockIdx.y;
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= lmax) return;
  const int T = min(asr_len[b], tmax), L = min(llm_len[b], lmax);
  const int32_t* aw = asr_wid + int64_t(b) * tmax;
  const int32_t* lw = llm_wid + int64_t(b) * lmax;
  int64_t acc[8];
  for (int q = 0; q < nq; ++q) acc[q] = 0;
  int hits = 0;
  if (l < L) {
    const int w = lw[l];
    for (int t = 0; t < T; ++t) {
      if (aw[t] != w) continue;                    // M[l,t]
      // cumsum over l' <= l of W1[l',t]: W1[l',t] = 1 iff t is the first asr position whose word == lw[l'] (l' < L)
      const int wt = aw[t];
      bool first_t = true;
      for (int t2 = 0; t2 < t; ++t2) if (aw[t2] == wt) { first_t = false; break; }
      if (!first_t) continue;                      // column t of W1 is all zero
      int cnt = 0;
      for (int l2 = 0; l2 <= l; ++l2) cnt += (lw[l2] == wt) ? 1 : 0;
      if (cnt == 1) {
        ++hits;
        for (int q = 0; q < nq; ++q) acc[q] += asr_idx[(int64_t(b) * tmax + t) * nq + q];
      }
    }
  }
  for (int q = 0; q < nq; ++q) out[(int64_t(b) * lmax + l) * nq + q] = hits ? acc[q] : -1;
}

int launch_map_llm(const int64_t* asr_indices, const int32_t* asr_wid, const int32_t* asr_len, const int32_t* llm_wid,
                   const int32_t* llm_len, int batch, int tmax, int lmax, int nq, int64_t* out, cudaStream_t stream) {
  if (!asr_indices || !asr_wid || !asr_len || !llm_wid || !llm_len || !out) return set_error(TASTE_E_ARG, "map_llm: null pointer");
  if (nq > 8) return set_error(TASTE_E_SHAPE, "map_llm: num_q > 8");
  if (batch <= 0 || lmax <= 0) return 0;
  dim3 grid((lmax + 127) / 128, batch);
  ProfScope ps(stream, KC_MAP_LLM, 0.0, 8.0 * double(batch) * (double(tmax) + lmax) * nq);
  map_llm_kernel<<<grid, 128, 0, stream>>>(asr_indices, asr_wid, asr_len, llm_wid, llm_len, tmax, lmax, nq, out);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace taste
