// C ABI of libtaste_b200.so (include/taste_b200.h): handle management, workspace carving and the launch sequences of
// the encoder (JES:133-223) and the aggregator (CW:1200-1437).  All device memory belongs to the caller.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <vector>

#include "internal.h"

namespace taste {

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

struct Carver {
  uint8_t* base;
  size_t off = 0;
  explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
  void* take(size_t bytes) {
    void* p = base ? base + off : nullptr;
    off += (bytes + 255) & ~size_t(255);
    return p;
  }
};

}  // namespace taste

struct taste_handle_s {
  taste_weights_t w;
  std::vector<taste_enc_layer_t> enc;
  std::vector<taste_dec_layer_t> dec;
};

using namespace taste;

namespace {

struct EncWs {
  void* feats_bf16;
  float* h;
  void* a;
  void* big;
  void* hb;        // bf16 copy of h (LayerNorm folding)
  float* stats;    // per row and 128-column segment: (sum, sum of squares) of h
  size_t bytes;
};
int g_encoder_mode = 0;
EncWs carve_encoder(const taste_dims_t& d, int batch, void* ws) {
  Carver c(ws);
  EncWs e;
  const size_t rows = size_t(batch) * TASTE_ENC_FRAMES;
  const size_t wide = size_t(d.ffn) > size_t(3) * d.d_model ? size_t(d.ffn) : size_t(3) * d.d_model;
  e.feats_bf16 = c.take(size_t(batch) * TASTE_N_FRAMES * TASTE_N_MELS * 2);
  e.h = static_cast<float*>(c.take(rows * d.d_model * 4));
  e.a = c.take(rows * d.d_model * 2);
  e.big = c.take(rows * wide * 2);       // conv1 output [B*3000, D] == [B*1500, 2D] also fits (wide >= 3D)
  e.hb = c.take(rows * d.d_model * 2);
  e.stats = static_cast<float*>(c.take(rows * size_t(d.d_model / 128 + 1) * 2 * 4));
  e.bytes = c.off;
  return e;
}

struct AggWs {
  void *kx, *vx;
  float* d;
  void *a, *qkv, *att, *qx, *mid;
  size_t bytes;
};
AggWs carve_aggregator(const taste_dims_t& d, int batch, int sum_tokens, void* ws) {
  Carver c(ws);
  AggWs g;
  const size_t frames = size_t(batch) * TASTE_ENC_FRAMES;
  const size_t rows = size_t(sum_tokens > 0 ? sum_tokens : 1);
  g.kx = c.take(frames * d.d_model * 2);
  g.vx = c.take(frames * d.d_model * 2);
  g.d = static_cast<float*>(c.take(rows * d.d_model * 4));
  g.a = c.take(rows * d.d_model * 2);
  g.qkv = c.take(rows * 3 * d.d_model * 2);
  g.att = c.take(rows * d.d_model * 2);
  g.qx = c.take(rows * d.d_model * 2);
  g.mid = c.take(rows * d.ffn * 2);
  g.bytes = c.off;
  return g;
}

struct MelWs {
  float* logspec;
  unsigned int* umax;
  void* planes;        // tensor-core DFT: split waveform
  float* spectrum;     // ... and its fp32 spectrum
  size_t bytes;
};
MelWs carve_logmel(int batch, void* ws) {
  Carver c(ws);
  MelWs m;
  m.logspec = static_cast<float*>(c.take(size_t(batch) * TASTE_N_FRAMES * TASTE_N_MELS * 4));
  m.umax = static_cast<unsigned int*>(c.take(size_t(batch) * 4));
  m.planes = c.take(size_t(batch) * 2 * size_t(LOGMEL_PLANE) * 2);
  m.spectrum = static_cast<float*>(c.take(size_t(batch) * TASTE_N_FRAMES * TASTE_DFT_N * 4));
  m.bytes = c.off;
  return m;
}

}  // namespace

extern "C" {

int taste_operand_dtype(void) { return TASTE_F16 ? 1 : 0; }

int taste_abi_version(void) { return TASTE_ABI_VERSION; }
const char* taste_last_error(void) { return taste::g_err; }

int taste_handle_create(const taste_weights_t* w, taste_handle_t* out) {
  if (!w || !out) return set_error(TASTE_E_ARG, "handle_create: null pointer");
  const taste_dims_t& d = w->dims;
  if (d.d_model <= 0 || d.d_model % 128 != 0 || d.heads * 64 != d.d_model || d.ffn <= 0 || d.ffn % 128 != 0)
    return set_error(TASTE_E_SHAPE, "handle_create: need d_model %% 128 == 0, head_dim == 64, ffn %% 128 == 0 (got d=%d heads=%d ffn=%d)",
                     d.d_model, d.heads, d.ffn);
  if (d.enc_layers < 0 || d.dec_layers < 0 || d.target_layer < 0 || (d.enc_layers > 0 && d.target_layer >= d.enc_layers))
    return set_error(TASTE_E_SHAPE, "handle_create: bad layer counts (enc=%d dec=%d target=%d)", d.enc_layers, d.dec_layers,
                     d.target_layer);
  if ((d.enc_layers > 0 && !w->enc) || (d.dec_layers > 0 && !w->dec))
    return set_error(TASTE_E_ARG, "handle_create: layer arrays missing");
  taste_handle_s* h = new (std::nothrow) taste_handle_s();
  if (!h) return set_error(TASTE_E_ARG, "handle_create: out of host memory");
  h->w = *w;
  h->enc.assign(w->enc, w->enc + d.enc_layers);
  h->dec.assign(w->dec, w->dec + d.dec_layers);
  h->w.enc = h->enc.data();
  h->w.dec = h->dec.data();
  *out = h;
  return 0;
}

int taste_handle_destroy(taste_handle_t h) {
  delete h;
  return 0;
}

size_t taste_ws_bytes(taste_handle_t h, int batch, int sum_tokens) {
  if (!h || batch <= 0) return 0;
  const size_t e = carve_encoder(h->w.dims, batch, nullptr).bytes;
  const size_t a = carve_aggregator(h->w.dims, batch, sum_tokens, nullptr).bytes;
  const size_t m = carve_logmel(batch, nullptr).bytes;
  size_t mx = e > a ? e : a;
  if (m > mx) mx = m;
  return mx + 256;
}

int taste_logmel_f32(taste_handle_t h, const float* wav, const int32_t* n_samples, int batch, int64_t wav_stride,
                     float* feats_f32, void* feats_bf16, void* ws, size_t ws_bytes, void* stream) {
  if (!h) return set_error(TASTE_E_ARG, "logmel: null handle");
  if (batch <= 0) return 0;
  if (!ws) return set_error(TASTE_E_ARG, "logmel: null workspace");
  MelWs m = carve_logmel(batch, ws);
  if (m.bytes > ws_bytes) return set_error(TASTE_E_WORKSPACE, "logmel: workspace %zu < %zu", ws_bytes, m.bytes);
  return launch_logmel(h->w, wav, n_samples, batch, wav_stride, feats_f32, feats_bf16, m.logspec, m.umax, m.planes,
                       m.spectrum, static_cast<cudaStream_t>(stream));
}

int taste_encoder_fwd(taste_handle_t h, const float* feats_f32, const void* feats_bf16, int batch, void* h_last_bf16,
                      void* h_target_bf16, void* ws, size_t ws_bytes, void* stream_) {
  if (!h || (!feats_f32 && !feats_bf16) || !h_last_bf16 || !h_target_bf16 || !ws)
    return set_error(TASTE_E_ARG, "encoder_fwd: null pointer");
  if (batch <= 0) return 0;
  const taste_weights_t& w = h->w;
  const taste_dims_t& d = w.dims;
  if (!w.conv1_w || !w.conv2_w || !w.conv1_b || !w.conv2_b || !w.enc_pos || !w.enc_ln_w || !w.enc_ln_b)
    return set_error(TASTE_E_ARG, "encoder_fwd: encoder weights missing from the handle");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  EncWs e = carve_encoder(d, batch, ws);
  if (e.bytes > ws_bytes) return set_error(TASTE_E_WORKSPACE, "encoder_fwd: workspace %zu < %zu", ws_bytes, e.bytes);
  const int D = d.d_model;
  const int rows = batch * TASTE_ENC_FRAMES;
  int rc;

  const void* fb = feats_bf16;
  if (feats_f32) {
    if ((rc = launch_cast_bf16(feats_f32, e.feats_bf16, int64_t(batch) * TASTE_N_FRAMES * TASTE_N_MELS, stream))) return rc;
    fb = e.feats_bf16;
  }
  {   // conv1 (k3, p1) + GELU as a 3-tap implicit GEMM over [B, 3000, 128]            JES:174
    GemmDesc g;
    g.a = fb;
    g.k_inner = TASTE_N_MELS;
    g.s_count = 1;
    g.rows_in = TASTE_N_FRAMES;
    g.rows_out = TASTE_N_FRAMES;
    g.batches = batch;
    g.s_stride = TASTE_N_MELS * 2;
    g.r_stride = TASTE_N_MELS * 2;
    g.b_stride = int64_t(TASTE_N_FRAMES) * TASTE_N_MELS * 2;
    g.taps = 3;
    g.tap_dr[0] = -1; g.tap_dr[1] = 0; g.tap_dr[2] = 1;
    g.w = w.conv1_w;
    g.n = D;
    g.bias = w.conv1_b;
    g.out = e.big;
    g.ldc = D;
    g.epilogue = EPI_GELU_BF16;
    if ((rc = launch_gemm(g, stream))) return rc;
  }
  // LayerNorm folding: available when every layer carries the folded weights, the shapes fit the CTA-pair kernel and
  // the batch is large enough for it (>= 2048 rows); otherwise the separate LayerNorm kernel runs (small batches).
  bool fold = g_encoder_mode != 1 && rows >= 2048 && D % 256 == 0 && d.ffn % 256 == 0 && d.enc_layers > 0;
  for (int l = 0; l < d.enc_layers && fold; ++l) {
    const taste_enc_layer_t& L = h->enc[l];
    fold = L.wqkv_ln && L.bqkv_ln && L.cqkv_ln && L.w1_ln && L.b1_ln && L.c1_ln;
  }
  const int nseg = D / 128;
  {   // conv2 (k3, s2, p1) + GELU + positions: input viewed as [B, 1500, 2, D]; output frame j reads
      // frames 2j-1, 2j, 2j+1 = (j-1, phase 1), (j, phase 0), (j, phase 1)            JES:175-180
    GemmDesc g;
    g.a = e.big;
    g.k_inner = D;
    g.s_count = 2;
    g.rows_in = TASTE_ENC_FRAMES;
    g.rows_out = TASTE_ENC_FRAMES;
    g.batches = batch;
    g.s_stride = int64_t(D) * 2;
    g.r_stride = int64_t(D) * 4;
    g.b_stride = int64_t(TASTE_N_FRAMES) * D * 2;
    g.taps = 3;
    g.tap_s[0] = 1; g.tap_s[1] = 0; g.tap_s[2] = 1;
    g.tap_dr[0] = -1; g.tap_dr[1] = 0; g.tap_dr[2] = 0;
    g.w = w.conv2_w;
    g.n = D;
    g.bias = w.conv2_b;
    g.out = e.h;
    g.ldc = D;
    g.epilogue = EPI_GELU_POS_F32;
    g.pos = w.enc_pos;
    if (fold) {
      g.stats_out = e.stats;
      g.out_bf16 = e.hb;
    }
    if ((rc = launch_gemm(g, stream))) return rc;
  }
  auto gemm_ex = [&](const void* a, const void* wt, const float* bias, void* out, int n, int k, int epi,
                     const float* ln_stats, const float* colsum, float* stats_out, void* out_bf16) -> int {
    GemmDesc g;
    g.a = a;
    g.k_inner = k;
    g.s_count = 1;
    g.rows_in = rows;
    g.rows_out = rows;
    g.batches = 1;
    g.s_stride = int64_t(k) * 2;
    g.r_stride = int64_t(k) * 2;
    g.b_stride = int64_t(rows) * k * 2;
    g.taps = 1;
    g.w = wt;
    g.n = n;
    g.bias = bias;
    g.out = out;
    g.ldc = n;
    g.epilogue = epi;
    g.ln_stats = ln_stats;
    g.ln_nseg = ln_stats ? nseg : 0;
    g.ln_colsum = colsum;
    g.stats_out = stats_out;
    g.out_bf16 = out_bf16;
    return launch_gemm(g, stream);
  };
  for (int l = 0; l < d.enc_layers; ++l) {
    const taste_enc_layer_t& L = h->enc[l];
    if (l == d.target_layer) {                                                       // JES:192-193
      if (fold) {
        TASTE_CUDA_OK(cudaMemcpyAsync(h_target_bf16, e.hb, size_t(rows) * D * 2, cudaMemcpyDeviceToDevice, stream));
      } else if ((rc = launch_cast_bf16(e.h, h_target_bf16, int64_t(rows) * D, stream))) {
        return rc;
      }
    }
    if (fold) {
      if ((rc = gemm_ex(e.hb, L.wqkv_ln, L.bqkv_ln, e.big, 3 * D, D, EPI_BF16, e.stats, L.cqkv_ln, nullptr, nullptr))) return rc;   // CW:690 + 342,365
    } else {
      if ((rc = launch_layernorm(e.h, L.ln1_w, L.ln1_b, e.a, rows, D, true, stream))) return rc;              // CW:690
      if ((rc = gemm_plain(e.a, L.wqkv, L.bqkv, e.big, rows, 3 * D, D, EPI_BF16, stream))) return rc;         // CW:342,365
    }
    AttnDesc at;
    at.q = e.big;
    at.k = static_cast<const uint16_t*>(e.big) + D;
    at.v = static_cast<const uint16_t*>(e.big) + 2 * D;
    at.o = e.a;
    at.ldq = at.ldk = at.ldv = 3 * D;
    at.ldo = D;
    at.cu_q = at.cu_kv = nullptr;
    at.q_len = at.kv_len = TASTE_ENC_FRAMES;
    at.batch = batch;
    at.heads = d.heads;
    at.causal = 0;
    at.kclass = KC_ATTN_ENC;
    if ((rc = launch_attention(at, stream))) return rc;                                                      // CW:377-394
    if (fold && g_encoder_mode == 2) {
      // both LayerNorms folded (measured slower than the hybrid below: the GELU epilogue of fc1 and the HBM-bound
      // out_proj epilogue have no slack for the extra work; kept for A/B runs)
      if ((rc = gemm_ex(e.a, L.wo, L.bo, e.h, D, D, EPI_RESID_F32, nullptr, nullptr, e.stats, e.hb))) return rc;            // CW:407,692
      if ((rc = gemm_ex(e.hb, L.w1_ln, L.b1_ln, e.big, d.ffn, D, EPI_GELU_BF16, e.stats, L.c1_ln, nullptr, nullptr))) return rc;   // CW:698-699
      if ((rc = gemm_ex(e.big, L.w2, L.b2, e.h, D, d.ffn, EPI_RESID_F32, nullptr, nullptr, e.stats, e.hb))) return rc;      // CW:701-703
    } else if (fold) {
      // hybrid: self_attn_layer_norm is folded (producer = fc2 with K = 5120, whose epilogue has slack; consumer = the
      // QKV GEMM, whose bias-only epilogue has slack); final_layer_norm stays a kernel
      if ((rc = gemm_plain(e.a, L.wo, L.bo, e.h, rows, D, D, EPI_RESID_F32, stream))) return rc;                            // CW:407,692
      if ((rc = launch_layernorm(e.h, L.ln2_w, L.ln2_b, e.a, rows, D, true, stream))) return rc;                            // CW:698
      if ((rc = gemm_plain(e.a, L.w1, L.b1, e.big, rows, d.ffn, D, EPI_GELU_BF16, stream))) return rc;                      // CW:699
      if ((rc = gemm_ex(e.big, L.w2, L.b2, e.h, D, d.ffn, EPI_RESID_F32, nullptr, nullptr, e.stats, e.hb))) return rc;      // CW:701-703
    } else {
      if ((rc = gemm_plain(e.a, L.wo, L.bo, e.h, rows, D, D, EPI_RESID_F32, stream))) return rc;              // CW:407,692
      if ((rc = launch_layernorm(e.h, L.ln2_w, L.ln2_b, e.a, rows, D, true, stream))) return rc;              // CW:698
      if ((rc = gemm_plain(e.a, L.w1, L.b1, e.big, rows, d.ffn, D, EPI_GELU_BF16, stream))) return rc;        // CW:699
      if ((rc = gemm_plain(e.big, L.w2, L.b2, e.h, rows, D, d.ffn, EPI_RESID_F32, stream))) return rc;        // CW:701-703
    }
  }
  if (d.target_layer >= d.enc_layers) return set_error(TASTE_E_SHAPE, "encoder_fwd: target layer not reached");
  return launch_layernorm(e.h, w.enc_ln_w, w.enc_ln_b, h_last_bf16, rows, D, true, stream);                 // JES:211
}

int taste_aggregator_fwd(taste_handle_t h, const void* h_last_bf16, const void* h_target_bf16, const int32_t* tokens,
                         const int32_t* cu_tokens, int batch, int sum_tokens, int max_tokens, float* dec_out, void* ws,
                         size_t ws_bytes, void* stream_) {
  if (!h || !h_last_bf16 || !h_target_bf16 || !tokens || !cu_tokens || !dec_out || !ws)
    return set_error(TASTE_E_ARG, "aggregator_fwd: null pointer");
  if (batch <= 0 || sum_tokens <= 0) return 0;
  const taste_weights_t& w = h->w;
  const taste_dims_t& d = w.dims;
  if (!w.tok_emb || !w.dec_pos || !w.dec_ln_w || !w.dec_ln_b)
    return set_error(TASTE_E_ARG, "aggregator_fwd: decoder weights missing from the handle");
  if (max_tokens > d.max_target_pos)
    return set_error(TASTE_E_SHAPE, "aggregator_fwd: %d assembled tokens exceed max_target_positions %d", max_tokens,
                     d.max_target_pos);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  AggWs g = carve_aggregator(d, batch, sum_tokens, ws);
  if (g.bytes > ws_bytes) return set_error(TASTE_E_WORKSPACE, "aggregator_fwd: workspace %zu < %zu", ws_bytes, g.bytes);
  const int D = d.d_model;
  const int frames = batch * TASTE_ENC_FRAMES;
  const int rows = sum_tokens;
  int rc;
  if ((rc = launch_embed(tokens, cu_tokens, batch, sum_tokens, w.tok_emb, w.dec_pos, D, d.vocab, d.max_target_pos, g.d,
                         stream)))
    return rc;                                                                                             // CW:1300-1341
  for (int l = 0; l < d.dec_layers; ++l) {
    const taste_dec_layer_t& L = h->dec[l];
    // causal self-attention                                                                                  CW:784-797
    if ((rc = launch_layernorm(g.d, L.ln1_w, L.ln1_b, g.a, rows, D, true, stream))) return rc;
    if ((rc = gemm_plain(g.a, L.wqkv, L.bqkv, g.qkv, rows, 3 * D, D, EPI_BF16, stream))) return rc;
    AttnDesc at;
    at.q = g.qkv;
    at.k = static_cast<const uint16_t*>(g.qkv) + D;
    at.v = static_cast<const uint16_t*>(g.qkv) + 2 * D;
    at.o = g.att;
    at.ldq = at.ldk = at.ldv = 3 * D;
    at.ldo = D;
    at.cu_q = at.cu_kv = cu_tokens;
    at.q_len = at.kv_len = max_tokens;
    at.batch = batch;
    at.heads = d.heads;
    at.causal = 1;
    at.total_q = sum_tokens;
    at.kclass = KC_ATTN_AGG;
    if ((rc = launch_attention(at, stream))) return rc;
    if ((rc = gemm_plain(g.att, L.wo, L.bo, g.d, rows, D, D, EPI_RESID_F32, stream))) return rc;
    // cross-attention: keys from the final encoder state, values from the layer-6 input, no mask        CW:361-366, 801-813
    if ((rc = launch_layernorm(g.d, L.lnx_w, L.lnx_b, g.a, rows, D, true, stream))) return rc;
    if ((rc = gemm_plain(g.a, L.wq_x, L.bq_x, g.qx, rows, D, D, EPI_BF16, stream))) return rc;
    if ((rc = gemm_plain(h_last_bf16, L.wk_x, nullptr, g.kx, frames, D, D, EPI_BF16, stream))) return rc;
    if ((rc = gemm_plain(h_target_bf16, L.wv_x, L.bv_x, g.vx, frames, D, D, EPI_BF16, stream))) return rc;
    at.q = g.qx;
    at.k = g.kx;
    at.v = g.vx;
    at.o = g.att;
    at.ldq = at.ldk = at.ldv = at.ldo = D;
    at.cu_q = cu_tokens;
    at.cu_kv = nullptr;
    at.q_len = max_tokens;
    at.kv_len = TASTE_ENC_FRAMES;
    at.causal = 0;
    if ((rc = launch_attention(at, stream))) return rc;
    if ((rc = gemm_plain(g.att, L.wo_x, L.bo_x, g.d, rows, D, D, EPI_RESID_F32, stream))) return rc;
    // MLP                                                                                                    CW:816-824
    if ((rc = launch_layernorm(g.d, L.ln2_w, L.ln2_b, g.a, rows, D, true, stream))) return rc;
    if ((rc = gemm_plain(g.a, L.w1, L.b1, g.mid, rows, d.ffn, D, EPI_GELU_BF16, stream))) return rc;
    if ((rc = gemm_plain(g.mid, L.w2, L.b2, g.d, rows, D, d.ffn, EPI_RESID_F32, stream))) return rc;
  }
  return launch_layernorm(g.d, w.dec_ln_w, w.dec_ln_b, dec_out, rows, D, false, stream);                    // CW:1415
}

int taste_assemble_tokens(const int64_t* asr_token_ids, const int32_t* token_lengths, const int32_t* cu_tokens, int batch,
                          int tmax, int32_t* tokens, void* stream) {
  return launch_assemble_tokens(asr_token_ids, token_lengths, cu_tokens, batch, tmax, tokens,
                                static_cast<cudaStream_t>(stream));
}

int taste_word_pool_f32(const float* dec_out, const int32_t* cu_tokens, const int32_t* word_ids,
                        const int32_t* token_lengths, int batch, int tmax, int d_model, float* z, void* stream) {
  return launch_word_pool(dec_out, cu_tokens, word_ids, token_lengths, batch, tmax, d_model, z,
                          static_cast<cudaStream_t>(stream));
}

size_t taste_rvq_ws_bytes(int n_rows) { return rvq_ws_bytes(n_rows); }

int taste_rvq_encode_f32(taste_handle_t h, const float* z, const int32_t* lengths, int batch, int tmax, int in_dim,
                         int64_t* indices, float* quantized, void* ws, size_t ws_bytes, void* stream) {
  if (!h) return set_error(TASTE_E_ARG, "rvq_encode: null handle");
  return launch_rvq_encode(h->w, z, lengths, batch, tmax, in_dim, indices, quantized, ws, ws_bytes,
                           static_cast<cudaStream_t>(stream));
}

int taste_rvq_decode_f32(taste_handle_t h, const int64_t* indices, int n, int project_out, float* out, void* stream) {
  if (!h) return set_error(TASTE_E_ARG, "rvq_decode: null handle");
  return launch_rvq_decode(h->w, indices, n, project_out != 0, out, static_cast<cudaStream_t>(stream));
}

int taste_map_to_llm_tokens(const int64_t* asr_indices, const int32_t* asr_word_ids, const int32_t* asr_lengths,
                            const int32_t* llm_word_ids, const int32_t* llm_lengths, int batch, int tmax, int lmax,
                            int num_q, int64_t* llm_indices, void* stream) {
  return launch_map_llm(asr_indices, asr_word_ids, asr_lengths, llm_word_ids, llm_lengths, batch, tmax, lmax, num_q,
                        llm_indices, static_cast<cudaStream_t>(stream));
}

int taste_gemm_bf16(const void* a, const void* w, const float* bias, void* out, int m, int n, int k, int epilogue,
                    void* stream) {
  if (epilogue < 0 || epilogue > 3) return set_error(TASTE_E_ARG, "gemm: epilogue must be 0..3");
  return gemm_plain(a, w, bias, out, m, n, k, epilogue, static_cast<cudaStream_t>(stream));
}

int taste_logmel_set_mode(int mode) {
  if (mode < 0 || mode > 1) return set_error(TASTE_E_ARG, "logmel_set_mode: mode must be 0 or 1");
  set_logmel_mode(mode);
  return 0;
}

int taste_encoder_set_mode(int mode) {
  if (mode < 0 || mode > 2) return set_error(TASTE_E_ARG, "encoder_set_mode: mode must be 0, 1 or 2");
  g_encoder_mode = mode;
  return 0;
}

int taste_gemm_ex(const taste_gemm_ex_t* x, void* stream) {
  if (!x) return set_error(TASTE_E_ARG, "gemm_ex: null descriptor");
  if (x->epilogue < 0 || x->epilogue > 3) return set_error(TASTE_E_ARG, "gemm_ex: epilogue must be 0..3");
  GemmDesc g;
  g.a = x->a;
  g.k_inner = x->k;
  g.s_count = 1;
  g.rows_in = x->m;
  g.rows_out = x->m;
  g.batches = 1;
  g.s_stride = int64_t(x->k) * 2;
  g.r_stride = int64_t(x->k) * 2;
  g.b_stride = int64_t(x->m) * x->k * 2;
  g.taps = 1;
  g.w = x->w;
  g.n = x->n;
  g.bias = x->bias;
  g.out = x->out;
  g.ldc = x->n;
  g.epilogue = x->epilogue;
  g.ln_stats = x->ln_stats;
  g.ln_nseg = x->ln_nseg;
  g.ln_colsum = x->ln_colsum;
  g.stats_out = x->stats_out;
  g.out_bf16 = x->out_bf16;
  return launch_gemm(g, static_cast<cudaStream_t>(stream));
}

int taste_attention_set_mode(int mode) {
  if (mode < 0 || mode > 1) return set_error(TASTE_E_ARG, "attention_set_mode: mode must be 0 or 1");
  set_attention_mode(mode);
  return 0;
}

int taste_gemm_set_mode(int mode) {
  if (mode < 0 || mode > 1) return set_error(TASTE_E_ARG, "gemm_set_mode: mode must be 0 or 1");
  set_gemm_mode(mode);
  return 0;
}

int taste_layernorm_f32(const float* x, const float* w, const float* b, void* y, int rows, int d, int out_bf16,
                        void* stream) {
  return launch_layernorm(x, w, b, y, rows, d, out_bf16 != 0, static_cast<cudaStream_t>(stream));
}

int taste_attention_ragged_bf16(const void* q, const void* k, const void* v, void* o, int ldq, int ldk, int ldv, int ldo,
                                const int32_t* cu_q, int total_q, int max_q_len, int kv_len, int batch, int heads,
                                void* stream) {
  if (!cu_q || total_q <= 0) return set_error(TASTE_E_ARG, "attention_ragged: cu_q and total_q are required");
  AttnDesc at;
  at.q = q; at.k = k; at.v = v; at.o = o;
  at.ldq = ldq; at.ldk = ldk; at.ldv = ldv; at.ldo = ldo;
  at.cu_q = cu_q; at.cu_kv = nullptr;
  at.q_len = max_q_len; at.kv_len = kv_len;
  at.batch = batch; at.heads = heads; at.causal = 0;
  at.total_q = total_q;
  return launch_attention(at, static_cast<cudaStream_t>(stream));
}

int taste_attention_bf16(const void* q, const void* k, const void* v, void* o, int ldq, int ldk, int ldv, int ldo,
                         const int32_t* cu_q, const int32_t* cu_kv, int q_len, int kv_len, int batch, int heads,
                         int causal, void* stream) {
  AttnDesc at;
  at.q = q; at.k = k; at.v = v; at.o = o;
  at.ldq = ldq; at.ldk = ldk; at.ldv = ldv; at.ldo = ldo;
  at.cu_q = cu_q; at.cu_kv = cu_kv;
  at.q_len = q_len; at.kv_len = kv_len;
  at.batch = batch; at.heads = heads; at.causal = causal;
  return launch_attention(at, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
