// Fused residual vector quantiser, fp32 (AQ:109-124 -> RVQ:359-490 -> VQ:955-1217 -> VQ:462-566; Appendix A6).
//
// One CTA owns 16 tokens end to end: project_in -> 4 x { distances to 512 codes, first-arg-min, gather,
// residual -= code, acc += code } -> project_out.  Residuals and the running sum live in shared memory: nothing
// round-trips HBM between levels.  Arithmetic is deliberately fp32 FFMA, not tensor cores: the contract is
// bit-identical indices against the reference's fp32 path (SURVEY §0 finding 8), and the stage is <0.01 % of the
// path's FLOPs.  The distance follows the reference's operation order exactly (VQ:44-48):
//   d = sqrt(max((|r|^2 + |e|^2) + (-2 * <r,e>), 0)),  index = first minimum of d (argmax of -d, VQ:102).
#include "common.cuh"
#include "internal.h"

namespace taste {

constexpr int RVQ_ROWS = 16;
constexpr int RVQ_THREADS = 256;
constexpr int RVQ_DC = 256;      // codebook dim
constexpr int RVQ_K = 512;       // codes per level
constexpr int RVQ_MAXQ = 8;

// x rows are staged in chunks of 256 input dims
__global__ void __launch_bounds__(RVQ_THREADS)
rvq_encode_kernel(const float* __restrict__ z, const int32_t* __restrict__ lengths, int tmax, int n_rows, int in_dim,
                  int d_model, int n_q, const float* __restrict__ win_t, const float* __restrict__ bin,
                  const float* __restrict__ code_t, const float* __restrict__ code, const float* __restrict__ code_sq,
                  const float* __restrict__ wout_t, const float* __restrict__ bout, int64_t* __restrict__ indices,
                  float* __restrict__ quantized) {
  __shared__ __align__(16) float s_res[RVQ_ROWS][RVQ_DC];       // residual
  __shared__ __align__(16) float s_acc[RVQ_ROWS][RVQ_DC];       // sum of selected codes
  float (*s_in)[RVQ_DC] = s_acc;                  // input staging chunk; dead before s_acc is first used
  __shared__ float s_x2[RVQ_ROWS];
  __shared__ float s_bd[RVQ_THREADS / 32][RVQ_ROWS];
  __shared__ int s_bi[RVQ_THREADS / 32][RVQ_ROWS];
  __shared__ int s_idx[RVQ_ROWS];
  __shared__ int s_valid[RVQ_ROWS];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int row0 = blockIdx.x * RVQ_ROWS;

  if (tid < RVQ_ROWS) {
    const int r = row0 + tid;
    int ok = 0;
    if (r < n_rows) {
      const int b = r / tmax, t = r - b * tmax;
      ok = lengths ? (t < lengths[b]) : 1;
    }
    s_valid[tid] = ok;
  }

  // ---- project_in (RVQ:371) or pass-through when the input is already a code (RVQ:258-357) ----
  if (in_dim == RVQ_DC) {
    for (int i = tid; i < RVQ_ROWS * RVQ_DC; i += RVQ_THREADS) {
      const int r = i / RVQ_DC, c = i - r * RVQ_DC;
      s_res[r][c] = (row0 + r < n_rows) ? z[int64_t(row0 + r) * in_dim + c] : 0.f;
    }
  } else {
    float acc[RVQ_ROWS];
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r) acc[r] = 0.f;
    for (int k0 = 0; k0 < in_dim; k0 += RVQ_DC) {
      __syncthreads();
      for (int i = tid; i < RVQ_ROWS * RVQ_DC; i += RVQ_THREADS) {
        const int r = i / RVQ_DC, c = i - r * RVQ_DC;
        s_in[r][c] = (row0 + r < n_rows && k0 + c < in_dim) ? z[int64_t(row0 + r) * in_dim + k0 + c] : 0.f;
      }
      __syncthreads();
      // four k at a time: one 16-byte broadcast read of the staged row per 4 multiply-adds (the loop is LDS-bound)
      const int kmax = min(RVQ_DC, in_dim - k0);           // in_dim is a multiple of 4 (checked by the launcher)
#pragma unroll 2
      for (int k = 0; k < kmax; k += 4) {
        const float w0 = __ldg(win_t + int64_t(k0 + k + 0) * RVQ_DC + tid);
        const float w1 = __ldg(win_t + int64_t(k0 + k + 1) * RVQ_DC + tid);
        const float w2 = __ldg(win_t + int64_t(k0 + k + 2) * RVQ_DC + tid);
        const float w3 = __ldg(win_t + int64_t(k0 + k + 3) * RVQ_DC + tid);
#pragma unroll
        for (int r = 0; r < RVQ_ROWS; ++r) {
          const float4 x = *reinterpret_cast<const float4*>(&s_in[r][k]);
          acc[r] = fmaf(x.w, w3, fmaf(x.z, w2, fmaf(x.y, w1, fmaf(x.x, w0, acc[r]))));
        }
      }
    }
    const float bv = __ldg(bin + tid);
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r) s_res[r][tid] = acc[r] + bv;
  }
  __syncthreads();       // all reads of the staging chunk (aliases s_acc) are done
#pragma unroll
  for (int r = 0; r < RVQ_ROWS; ++r) s_acc[r][tid] = 0.f;
  __syncthreads();

  for (int q = 0; q < n_q; ++q) {
    // |r|^2 per row: warp w handles rows 2w, 2w+1
    for (int r = warp * 2; r < warp * 2 + 2; ++r) {
      float p = 0.f;
      for (int c = lane; c < RVQ_DC; c += 32) p = fmaf(s_res[r][c], s_res[r][c], p);
      p = warp_sum(p);
      if (lane == 0) s_x2[r] = p;
    }
    // <r, e_j> for codes j = tid and tid + 256
    float d0[RVQ_ROWS], d1[RVQ_ROWS];
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r) d0[r] = d1[r] = 0.f;
    const float* ct = code_t + int64_t(q) * RVQ_DC * RVQ_K;
#pragma unroll 1
    for (int c = 0; c < RVQ_DC; c += 4) {
      float e0[4], e1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        e0[u] = __ldg(ct + (c + u) * RVQ_K + tid);
        e1[u] = __ldg(ct + (c + u) * RVQ_K + tid + RVQ_THREADS);
      }
#pragma unroll
      for (int r = 0; r < RVQ_ROWS; ++r) {
        const float4 rv = *reinterpret_cast<const float4*>(&s_res[r][c]);
        // same summation order as one element at a time (c ascending), so the distances are bit-identical
        d0[r] = fmaf(rv.w, e0[3], fmaf(rv.z, e0[2], fmaf(rv.y, e0[1], fmaf(rv.x, e0[0], d0[r]))));
        d1[r] = fmaf(rv.w, e1[3], fmaf(rv.z, e1[2], fmaf(rv.y, e1[1], fmaf(rv.x, e1[0], d1[r]))));
      }
    }
    __syncthreads();     // s_x2 visible
    const float y0 = __ldg(code_sq + q * RVQ_K + tid);
    const float y1 = __ldg(code_sq + q * RVQ_K + tid + RVQ_THREADS);
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r) {
      const float x2 = s_x2[r];
      const float s0 = __fadd_rn(__fadd_rn(x2, y0), __fmul_rn(d0[r], -2.0f));
      const float s1 = __fadd_rn(__fadd_rn(x2, y1), __fmul_rn(d1[r], -2.0f));
      float bd = __fsqrt_rn(fmaxf(s0, 0.f));
      int bi = tid;
      const float dd1 = __fsqrt_rn(fmaxf(s1, 0.f));
      if (dd1 < bd) { bd = dd1; bi = tid + RVQ_THREADS; }
      // warp arg-min, ties -> smaller index (first maximum of -d)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float od = __shfl_xor_sync(0xffffffffu, bd, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
      }
      if (lane == 0) { s_bd[warp][r] = bd; s_bi[warp][r] = bi; }
    }
    __syncthreads();
    if (tid < RVQ_ROWS) {
      float bd = s_bd[0][tid];
      int bi = s_bi[0][tid];
      for (int w = 1; w < RVQ_THREADS / 32; ++w) {
        const float od = s_bd[w][tid];
        const int oi = s_bi[w][tid];
        if (od < bd || (od == bd && oi < bi)) { bd = od; bi = oi; }
      }
      s_idx[tid] = bi;
      const int r = row0 + tid;
      if (r < n_rows) indices[int64_t(r) * n_q + q] = s_valid[tid] ? int64_t(bi) : int64_t(-1);   // VQ:1205-1210
    }
    __syncthreads();
    // gather + residual update; masked rows contribute a zero code (VQ:1192-1203, RVQ:455-456)
    const float* cb = code + int64_t(q) * RVQ_K * RVQ_DC;
#pragma unroll
    for (int r = 0; r < RVQ_ROWS; ++r) {
      if (s_valid[r]) {
        const float cv = __ldg(cb + int64_t(s_idx[r]) * RVQ_DC + tid);
        s_res[r][tid] = s_res[r][tid] - cv;
        s_acc[r][tid] = s_acc[r][tid] + cv;
      }
    }
    __syncthreads();
  }

  // ---- project_out (RVQ:470) ----
  if (quantized != nullptr) {
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
        if (row0 + r < n_rows) quantized[int64_t(row0 + r) * d_model + col] = acc[r] + bv;
    }
  }
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
  if (!w.rvq_win_t || !w.rvq_bin || !w.rvq_code_t || !w.rvq_code || !w.rvq_code_sq || !w.rvq_wout_t || !w.rvq_bout)
    return set_error(TASTE_E_ARG, "rvq: weights missing from the handle");
  return 0;
}

int launch_rvq_encode(const taste_weights_t& w, const float* z, const int32_t* lengths, int batch, int tmax, int in_dim,
                      int64_t* indices, float* quantized, cudaStream_t stream) {
  if (int rc = check_rvq(w)) return rc;
  if (!z || !indices) return set_error(TASTE_E_ARG, "rvq_encode: null pointer");
  if (in_dim != w.dims.d_model && in_dim != RVQ_DC) return set_error(TASTE_E_SHAPE, "rvq_encode: in_dim must be d_model or 256");
  const int n_rows = batch * tmax;
  if (n_rows <= 0) return 0;
  const int blocks = (n_rows + RVQ_ROWS - 1) / RVQ_ROWS;
  const double per_row = (in_dim == RVQ_DC ? 0.0 : 2.0 * in_dim * RVQ_DC) + w.dims.num_quantizers * 2.0 * RVQ_DC * RVQ_K +
                         (quantized ? 2.0 * RVQ_DC * w.dims.d_model : 0.0);
  ProfScope ps(stream, KC_RVQ_ENCODE, per_row * n_rows,
               double(n_rows) * (4.0 * in_dim + 8.0 * w.dims.num_quantizers + (quantized ? 4.0 * w.dims.d_model : 0.0)));
  rvq_encode_kernel<<<blocks, RVQ_THREADS, 0, stream>>>(z, lengths, tmax, n_rows, in_dim, w.dims.d_model,
                                                        w.dims.num_quantizers, w.rvq_win_t, w.rvq_bin, w.rvq_code_t,
                                                        w.rvq_code, w.rvq_code_sq, w.rvq_wout_t, w.rvq_bout, indices,
                                                        quantized);
  TASTE_CUDA_OK(cudaGetLastError());
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
