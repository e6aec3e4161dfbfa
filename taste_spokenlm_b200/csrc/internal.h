// Host-side internal interfaces between the translation units of libtaste_b200.so.
#pragma once
#ifndef TASTE_F16
#define TASTE_F16 0          /* library flavour: see common.cuh */
#endif
#include <cuda_runtime.h>
#include <cuda.h>
#include <stddef.h>
#include <stdint.h>

#include "../../include/taste_b200.h"

namespace taste {

// Per-device caches (cudaFuncSetAttribute, occupancy, SM count) are indexed by the current device ordinal.
constexpr int kMaxDevices = 64;
inline int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < kMaxDevices) ? d : 0;
}

// ---- error plumbing (api.cu) ----
int set_error(int code, const char* fmt, ...);
#define TASTE_CUDA_OK(expr)                                                                        \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return ::taste::set_error((int)_e, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

// ---- launch accounting / device timing (prof.cu) ----
enum KernelClass {
  KC_GEMM = 0, KC_ATTN_ENC, KC_ATTN_AGG, KC_LAYERNORM, KC_CAST, KC_LOGMEL_TILE, KC_LOGMEL_FINISH, KC_EMBED,
  KC_WORD_POOL, KC_RVQ_ENCODE, KC_RVQ_DECODE, KC_MAP_LLM, KC_ATTN_TC, KC_RESAMPLE, KC_LOGMEL_SPLIT, KC_LOGMEL_DFT,
  KC_LOGMEL_MEL, KC_RVQ_PROJ, KC_COUNT
};
// Wrap a kernel launch: counts it and, when profiling is on, brackets it with CUDA events on `stream`.
// flops / bytes are the ALGORITHMIC work of the launch (DESIGN.md "Kernels").
struct ProfScope {
  ProfScope(cudaStream_t stream, int kc, double flops, double bytes);
  ~ProfScope();
  cudaStream_t stream_;
  int slot_;
};

// ---- TMA descriptors: cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda needed) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_tensor_map_encoder();   // gemm_tcgen05.cu; nullptr when the entry point is unavailable

// ---- GEMM (gemm_tcgen05.cu) ----
enum GemmEpilogue {
  EPI_BF16 = 0,          // out bf16 = acc + bias
  EPI_GELU_BF16 = 1,     // out bf16 = gelu(acc + bias)
  EPI_RESID_F32 = 2,     // out fp32 += acc + bias
  EPI_F32 = 3,           // out fp32 = acc + bias
  EPI_GELU_POS_F32 = 4,  // out fp32 = gelu(acc + bias) + pos[row_in_batch][col]
};

// out[b*R + r, :] = epi( sum_tap A[b, r + tap_dr[tap], tap_s[tap], :] @ W[:, tap*k_inner : (tap+1)*k_inner]^T + bias )
// A is addressed as a 4-D tensor {k_inner, s_count, rows_in, batches} with byte strides; rows outside
// [0, rows_in) read as zero (TMA out-of-bounds fill) which implements the convolution's zero padding.
struct GemmDesc {
  const void* a = nullptr;
  int k_inner = 0;
  int s_count = 1;
  int rows_in = 0;            // addressable rows per batch entry in A
  int rows_out = 0;           // output rows per batch entry (R)
  int batches = 1;
  int64_t s_stride = 0, r_stride = 0, b_stride = 0;   // bytes
  int taps = 1;
  int tap_s[3] = {0, 0, 0};
  int tap_dr[3] = {0, 0, 0};
  const void* w = nullptr;    // bf16 [n, taps*k_inner]
  int n = 0;
  const float* bias = nullptr;
  void* out = nullptr;
  int ldc = 0;
  int epilogue = EPI_BF16;
  const float* pos = nullptr;
  // LayerNorm folding (CTA-pair kernel only; see GemmParams in gemm_tcgen05.cu)
  const float* ln_stats = nullptr;    // consumer: [rows][ln_nseg][2]; requires ln_colsum and bias (= b + W beta)
  int ln_nseg = 0;
  const float* ln_colsum = nullptr;
  float* stats_out = nullptr;         // producer: [rows][n/128][2]; requires out_bf16
  void* out_bf16 = nullptr;
  bool force_pair = false;
  int ab_bf16 = -1;                   // operand format: -1 = the library flavour (kActBf16), 1 = bf16 (the log-mel DFT)
  // accounting (prof.cu): kernel class and, when >= 0, the algorithmic bytes booked on this launch
  int kclass = KC_GEMM;
  double alg_bytes = -1.0;
};
int launch_gemm(const GemmDesc& d, cudaStream_t stream);
void set_gemm_mode(int mode);
int gemm_plain(const void* a, const void* w, const float* bias, void* out, int m, int n, int k, int epilogue,
               cudaStream_t stream);

// ---- attention (attention_mma.cu) ----
struct AttnDesc {
  const void *q, *k, *v;
  void* o;
  int ldq, ldk, ldv, ldo;
  const int32_t* cu_q;
  const int32_t* cu_kv;
  int q_len, kv_len;     // fixed lengths when cu_* is null; otherwise upper bounds used for the grid
  int batch, heads, causal;
  int total_q = 0;       // sum of query rows (accounting only; 0 = batch * q_len)
  int kclass = KC_ATTN_AGG;
};
int launch_attention(const AttnDesc& d, cudaStream_t stream);      // dispatches between the two kernels below
bool attention_tcgen05_eligible(const AttnDesc& d);
int launch_attention_tcgen05(const AttnDesc& d, cudaStream_t stream);
void set_attention_mode(int mode);                                  // 0 = automatic, 1 = always the mma.sync kernel

// ---- elementwise (elementwise.cu) ----
int launch_layernorm(const float* x, const float* w, const float* b, void* y, int rows, int d, bool out_bf16,
                     cudaStream_t stream);
int launch_cast_bf16(const float* x, void* y, int64_t n, cudaStream_t stream);
int launch_assemble_tokens(const int64_t* ids, const int32_t* lens, const int32_t* cu, int batch, int tmax,
                           int32_t* tokens, cudaStream_t stream);
int launch_embed(const int32_t* tokens, const int32_t* cu_tokens, int batch, int sum_tokens, const float* tok_emb,
                 const float* pos_emb, int d, int vocab, int max_pos, float* out, cudaStream_t stream);
int launch_word_pool(const float* dec_out, const int32_t* cu_tokens, const int32_t* word_ids,
                     const int32_t* token_lengths, int batch, int tmax, int d, float* z, cudaStream_t stream);
int launch_map_llm(const int64_t* asr_indices, const int32_t* asr_wid, const int32_t* asr_len, const int32_t* llm_wid,
                   const int32_t* llm_len, int batch, int tmax, int lmax, int nq, int64_t* out, cudaStream_t stream);

// ---- log-mel (logmel.cu) ----
// planes: bf16 [batch][2][LOGMEL_PLANE] (hi / lo halves of the reflect-padded waveform), spectrum: fp32
// [batch * 3000][TASTE_DFT_N]; both only used by the tensor-core formulation (may be null for the FMA kernel)
constexpr int64_t LOGMEL_PLANE = int64_t(TASTE_HOP) * TASTE_N_FRAMES + TASTE_DFT_K;      // 480 448 samples
int launch_logmel(const taste_weights_t& w, const float* wav, const int32_t* n_samples, int batch, int64_t wav_stride,
                  float* feats_f32, void* feats_bf16, float* scratch_logspec, unsigned int* scratch_max,
                  void* scratch_planes, float* scratch_spectrum, cudaStream_t stream);
void set_logmel_mode(int mode);

// ---- RVQ (rvq.cu) ----
int launch_rvq_encode(const taste_weights_t& w, const float* z, const int32_t* lengths, int batch, int tmax, int in_dim,
                      int64_t* indices, float* quantized, void* ws, size_t ws_bytes, cudaStream_t stream);
size_t rvq_ws_bytes(int n_rows);
int launch_rvq_decode(const taste_weights_t& w, const int64_t* indices, int n, bool project_out, float* out,
                      cudaStream_t stream);

// ---- ingest (resample.cu) ----
int launch_resample_mean(const float* in, const int64_t* in_off, const int32_t* channels, const int32_t* n_in, int batch,
                         int orig, int nw, int width, const float* taps, const int32_t* kstart, int knz, int knz_ld,
                         int max_out, double total_in_elems, double total_out_elems, float* wav, int64_t wav_stride,
                         int32_t* n_out, cudaStream_t stream);

}  // namespace taste
