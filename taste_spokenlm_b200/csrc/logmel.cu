// Whisper log-mel front-end (WF:56-113):
//   pad/trim to 30 s + reflect padding -> periodic-Hann window -> 400-point real DFT -> |X|^2 -> sparse Slaney mel
//   projection (394 non-zeros) -> log10(clamp 1e-10) -> per-utterance max; a last small kernel applies the `max - 8`
//   floor and the (x+4)/4 scaling and emits fp32 and/or bf16 features.
// Two formulations of the DFT (taste_logmel_set_mode; ncu: profiles/r1v12_ncu_misc_summary.json and r1v13):
//  0 (default) DFT-as-GEMM on the tcgen05 GEMM kernel.  bf16 alone cannot carry the 80 dB dynamic range kept by the
//    `max - 8` floor (SURVEY 7 hard part 6), so samples and twiddles are split into bf16 halves, x = hi + lo (|lo| <=
//    2^-9 |x|), and  x t ~= hi(x) hi(t) + hi(x) lo(t) + lo(x) hi(t)  (relative error ~2^-17, fp32 accumulation) is
//    ONE GEMM with three K slabs.  logmel_split_kernel writes the reflect-padded waveform once as two bf16 planes;
//    the frames are never materialised: the A operand is a 4-D tensor map over the planes whose row stride is the hop
//    (160 samples = 320 bytes), so TMA reads the overlapping 448-sample slabs (400 + zero-weighted tail) directly;
//    the Hann window is folded into the twiddle matrix; logmel_mel_kernel turns the fp32 spectrum into log-mel.
//  1 fp32 FMA kernel, fused per 64-frame tile: the DFT folded over the n <-> 400-n symmetry so only 199 x 201 twiddle
//    products are needed per frame (packed FFMA2).  Exact fp32 arithmetic, but latency-bound on its twiddle loads.
#include "common.cuh"
#include "internal.h"

namespace taste {

constexpr int LM_FRAMES = 64;          // frames per CTA
constexpr int LM_THREADS = 256;        // 8 frame groups x 32 bin lanes
constexpr int LM_FPT = 8;              // frames per thread (4 packed pairs)
constexpr int LM_KPT = 4;              // half-spectrum bins per thread: k' = lane + 32*j  (128 >= 101)
constexpr int LM_NH = 199;             // folded terms n = 1..199
constexpr int LM_LDF = 66;             // padded frame stride of the folded arrays (even: 8-byte aligned frame pairs)
constexpr int LM_LDP = 209;            // padded bin stride of the power tile
constexpr int LM_SMEM = (2 * 200 * LM_LDF + LM_FRAMES) * 4;   // E, O, y200

static_assert(LM_FRAMES * LM_LDP * 4 <= 2 * 200 * LM_LDF * 4, "power tile must fit in the folded arrays it aliases");

__device__ __forceinline__ float padded_sample(const float* __restrict__ wav, int n_valid, int i) {
  int j = i - TASTE_N_FFT / 2;                                   // torch.stft(center=True): reflect pad 200
  if (j < 0) j = -j;
  else if (j >= TASTE_N_SAMPLES) j = 2 * (TASTE_N_SAMPLES - 1) - j;
  return j < n_valid ? __ldg(wav + j) : 0.f;                     // whisper.pad_or_trim zero padding (WF:98-99)
}

__device__ __forceinline__ unsigned int float_to_ordered(float f) {
  const unsigned int b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(unsigned int k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// The 400-point real DFT of a windowed frame y uses two symmetries:
//   n <-> 400-n :  Re X[k] = y[0] + (-1)^k y[200] + sum_{n=1..199} E[n] cos(2 pi k n / 400),  E[n] = y[n] + y[400-n]
//                  Im X[k] =                      - sum_{n=1..199} O[n] sin(2 pi k n / 400),  O[n] = y[n] - y[400-n]
//   k <-> 200-k :  cos(2 pi (200-k) n / 400) = (-1)^n cos(2 pi k n / 400),  sin(...) = -(-1)^n sin(2 pi k n / 400)
// so with the sums split by the parity of n (Ce/Co over E, Se/So over O) only k' = 0..100 is accumulated:
//   Re X[k'] = Ce + Co + s,  Re X[200-k'] = Ce - Co + s,  |Im X[k']| = |Se + So|,  |Im X[200-k']| = |Se - So|.
// That is a quarter of the multiply-adds of the plain 400 x 201 DFT; they run as packed FFMA2 over frame pairs.
__global__ void __launch_bounds__(LM_THREADS, 1)
logmel_tile_kernel(const float* __restrict__ wav, const int32_t* __restrict__ n_samples, int64_t wav_stride,
                   const float* __restrict__ dft_cos, const float* __restrict__ dft_sin, const float* __restrict__ hann,
                   const int32_t* __restrict__ mel_start, const int32_t* __restrict__ mel_count,
                   const float* __restrict__ mel_weight, float* __restrict__ logspec, unsigned int* __restrict__ umax) {
  extern __shared__ __align__(16) float lm_smem[];
  float* sE = lm_smem;                       // [200][66]  w[n] * (x[n] + x[400-n]),  row n-1
  float* sO = sE + 200 * LM_LDF;             // [200][66]  w[n] * (x[n] - x[400-n])
  float* sY200 = sO + 200 * LM_LDF;          // [64]
  float* sP = lm_smem;                       // aliases sE/sO after the DFT: [64][209]
  __shared__ float s_red[LM_THREADS / 32];

  const int b = blockIdx.y;
  const int f0 = blockIdx.x * LM_FRAMES;
  const int tid = threadIdx.x;
  const float* w = wav + int64_t(b) * wav_stride;
  int n_valid = n_samples ? n_samples[b] : TASTE_N_SAMPLES;
  n_valid = max(0, min(n_valid, TASTE_N_SAMPLES));

  // ---- stage 1: folded, windowed frames ----
  for (int idx = tid; idx < LM_FRAMES * 200; idx += LM_THREADS) {
    const int f = idx / 200;
    const int n = idx - f * 200 + 1;                  // 1..200
    const int base = (f0 + f) * TASTE_HOP;
    if (n <= LM_NH) {
      const float a = padded_sample(w, n_valid, base + n);
      const float c = padded_sample(w, n_valid, base + TASTE_N_FFT - n);
      const float hw = __ldg(hann + n);
      sE[(n - 1) * LM_LDF + f] = hw * a + hw * c;
      sO[(n - 1) * LM_LDF + f] = hw * a - hw * c;
    } else {
      sY200[f] = __ldg(hann + 200) * padded_sample(w, n_valid, base + 200);
      sE[199 * LM_LDF + f] = 0.f;
      sO[199 * LM_LDF + f] = 0.f;
    }
  }
  __syncthreads();

  // ---- stage 2: half-spectrum DFT, register tile of 8 frames (4 packed pairs) x 4 bins x {Ce, Co, Se, So} ----
  const int fg = tid >> 5;                    // frame group: frames fg*8 .. fg*8+7
  const int lane = tid & 31;
  uint64_t ce[LM_FPT / 2][LM_KPT], co[LM_FPT / 2][LM_KPT], se[LM_FPT / 2][LM_KPT], so[LM_FPT / 2][LM_KPT];
#pragma unroll
  for (int i = 0; i < LM_FPT / 2; ++i)
#pragma unroll
    for (int j = 0; j < LM_KPT; ++j) ce[i][j] = co[i][j] = se[i][j] = so[i][j] = f2_pack(0.f, 0.f);

  auto accumulate = [&](int row, uint64_t (&accc)[LM_FPT / 2][LM_KPT], uint64_t (&accs)[LM_FPT / 2][LM_KPT]) {
    float c[LM_KPT], sn[LM_KPT];
#pragma unroll
    for (int j = 0; j < LM_KPT; ++j) {
      c[j] = __ldg(dft_cos + row * TASTE_DFT_LD + lane + 32 * j);
      sn[j] = __ldg(dft_sin + row * TASTE_DFT_LD + lane + 32 * j);
    }
    const uint64_t* e2 = reinterpret_cast<const uint64_t*>(sE + row * LM_LDF + fg * LM_FPT);
    const uint64_t* o2 = reinterpret_cast<const uint64_t*>(sO + row * LM_LDF + fg * LM_FPT);
#pragma unroll
    for (int i = 0; i < LM_FPT / 2; ++i) {
      const uint64_t e = e2[i];               // frames (2i, 2i+1) of this group: warp-wide broadcast
      const uint64_t o = o2[i];
#pragma unroll
      for (int j = 0; j < LM_KPT; ++j) {
        accc[i][j] = f2_fma(e, f2_pack(c[j], c[j]), accc[i][j]);
        accs[i][j] = f2_fma(o, f2_pack(sn[j], sn[j]), accs[i][j]);
      }
    }
  };
#pragma unroll 3
  for (int row = 0; row < LM_NH - 1; row += 2) {      // n = row + 1: odd first, then even (99 iterations = 3 x 33)
    accumulate(row, co, so);
    accumulate(row + 1, ce, se);
  }
  accumulate(LM_NH - 1, co, so);                      // n = 199 (odd)

  float y200[LM_FPT];
#pragma unroll
  for (int i = 0; i < LM_FPT; ++i) y200[i] = sY200[fg * LM_FPT + i];
  __syncthreads();                            // everyone finished reading sE/sO: reuse as the power tile
#pragma unroll
  for (int j = 0; j < LM_KPT; ++j) {
    const int k = lane + 32 * j;
    if (k <= 100) {
      const float sgn = (k & 1) ? -1.f : 1.f;          // y[0] = 0 (periodic Hann); y[200] enters with (-1)^k
#pragma unroll
      for (int i = 0; i < LM_FPT / 2; ++i) {
        float cev[2], cov[2], sev[2], sov[2];
        f2_unpack(ce[i][j], cev[0], cev[1]);
        f2_unpack(co[i][j], cov[0], cov[1]);
        f2_unpack(se[i][j], sev[0], sev[1]);
        f2_unpack(so[i][j], sov[0], sov[1]);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int f = fg * LM_FPT + 2 * i + h;
          const float s = sgn * y200[2 * i + h];
          const float re_lo = (cev[h] + cov[h]) + s, im_lo = sev[h] + sov[h];
          const float re_hi = (cev[h] - cov[h]) + s, im_hi = sev[h] - sov[h];
          sP[f * LM_LDP + k] = re_lo * re_lo + im_lo * im_lo;
          if (k < 100) sP[f * LM_LDP + 200 - k] = re_hi * re_hi + im_hi * im_hi;
        }
      }
    }
  }
  __syncthreads();

  // ---- stage 3: mel projection, log10, running max ----
  float lmax = -INFINITY;
  const int m = tid & 127;
  const int ms = __ldg(mel_start + m);
  const int mc = __ldg(mel_count + m);
  float wgt[TASTE_MEL_MAXW];
#pragma unroll
  for (int j = 0; j < TASTE_MEL_MAXW; ++j) wgt[j] = __ldg(mel_weight + m * TASTE_MEL_MAXW + j);
  for (int f = tid >> 7; f < LM_FRAMES; f += 2) {
    if (f0 + f >= TASTE_N_FRAMES) break;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < TASTE_MEL_MAXW; ++j)
      if (j < mc) acc = fmaf(wgt[j], sP[f * LM_LDP + ms + j], acc);
    const float lg = log10f(fmaxf(acc, 1e-10f));
    logspec[(int64_t(b) * TASTE_N_FRAMES + f0 + f) * TASTE_N_MELS + m] = lg;
    lmax = fmaxf(lmax, lg);
  }
  lmax = warp_max(lmax);
  if (lane == 0) s_red[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float v = s_red[0];
    for (int i = 1; i < LM_THREADS / 32; ++i) v = fmaxf(v, s_red[i]);
    if (v > -INFINITY) atomicMax(umax + b, float_to_ordered(v));
  }
}

// ---- tensor-core formulation: split planes -> (GEMM) -> mel ----
// planes[b][0 | 1][i], i in [0, LOGMEL_PLANE): hi / lo bf16 halves of the padded waveform sample i (reflect padding of
// 200, zeros past the utterance and past the 30 s window).  8 samples per thread, 16-byte stores.
__global__ void __launch_bounds__(256)
logmel_split_kernel(const float* __restrict__ wav, const int32_t* __restrict__ n_samples, int64_t wav_stride,
                    __nv_bfloat16* __restrict__ planes) {
  const int b = blockIdx.y;
  const int64_t i0 = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) * 8;
  if (i0 >= LOGMEL_PLANE) return;
  const float* w = wav + int64_t(b) * wav_stride;
  int n_valid = n_samples ? n_samples[b] : TASTE_N_SAMPLES;
  n_valid = max(0, min(n_valid, TASTE_N_SAMPLES));
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float x[2], h[2];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      const int i = int(i0) + 2 * e + t;
      x[t] = i < TASTE_N_SAMPLES + TASTE_N_FFT ? padded_sample(w, n_valid, i) : 0.f;
      h[t] = __bfloat162float(__float2bfloat16_rn(x[t]));
    }
    hi[e] = pack_true_bf16x2(h[0], h[1]);
    lo[e] = pack_true_bf16x2(x[0] - h[0], x[1] - h[1]);
  }
  __nv_bfloat16* p0 = planes + int64_t(b) * 2 * LOGMEL_PLANE + i0;
  *reinterpret_cast<uint4*>(p0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
  *reinterpret_cast<uint4*>(p0 + LOGMEL_PLANE) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

// spectrum [b * 3000 + f][TASTE_DFT_N] (Re at column k, Im at 256 + k) -> power -> mel -> log10 -> logspec, running max.
constexpr int MM_FRAMES = 32;
constexpr int MM_LDL = TASTE_N_MELS + 1;      // log tile stride (odd)
__global__ void __launch_bounds__(256)
logmel_mel_kernel(const float* __restrict__ spectrum, const int32_t* __restrict__ mel_start,
                  const int32_t* __restrict__ mel_count, const float* __restrict__ mel_weight,
                  float* __restrict__ logspec, unsigned int* __restrict__ umax) {
  __shared__ float sP[MM_FRAMES * LM_LDP];
  __shared__ float sL[MM_FRAMES * MM_LDL];
  __shared__ float s_red[8];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * MM_FRAMES;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  // 16-byte loads, 51 per half row (bins 0..203; 201..203 are zero columns of the GEMM), four rows in flight per thread
  constexpr int kQuads = 51;
#pragma unroll 4
  for (int idx = tid; idx < MM_FRAMES * 64; idx += 256) {
    const int f = idx >> 6;
    const int k4 = idx & 63;
    if (k4 < kQuads && f0 + f < TASTE_N_FRAMES) {
      const float4* row = reinterpret_cast<const float4*>(spectrum + (int64_t(b) * TASTE_N_FRAMES + f0 + f) * TASTE_DFT_N);
      const float4 re = __ldg(row + k4), im = __ldg(row + 64 + k4);
      float* dst = sP + f * LM_LDP + 4 * k4;
      dst[0] = re.x * re.x + im.x * im.x;
      dst[1] = re.y * re.y + im.y * im.y;
      dst[2] = re.z * re.z + im.z * im.z;
      dst[3] = re.w * re.w + im.w * im.w;
    }
  }
  __syncthreads();
  // mel projection with lane = frame (power-tile stride 209 is odd: conflict-free) and 16 filters per warp (start,
  // count and weights are warp-uniform: broadcast loads, no divergence); the log values go through shared memory so
  // that the global stores are coalesced along the mel axis.  (With lane = filter the scattered power-tile reads
  // made this phase shared-memory-bound.)
  float lmax = -INFINITY;
  const int warp = tid >> 5;
  const bool frame_ok = f0 + lane < TASTE_N_FRAMES;
#pragma unroll 4
  for (int mm = 0; mm < 16; ++mm) {
    const int m = warp * 16 + mm;
    const int ms = __ldg(mel_start + m);
    const int mc = __ldg(mel_count + m);
    const float* pw = sP + lane * LM_LDP + ms;
    float acc = 0.f;
    for (int j = 0; j < mc; ++j) acc = fmaf(__ldg(mel_weight + m * TASTE_MEL_MAXW + j), pw[j], acc);
    const float lg = log10f(fmaxf(acc, 1e-10f));
    sL[lane * MM_LDL + m] = lg;
    if (frame_ok) lmax = fmaxf(lmax, lg);
  }
  __syncthreads();
  for (int idx = tid; idx < MM_FRAMES * TASTE_N_MELS; idx += 256) {
    const int f = idx >> 7, m = idx & 127;
    if (f0 + f < TASTE_N_FRAMES) logspec[(int64_t(b) * TASTE_N_FRAMES + f0 + f) * TASTE_N_MELS + m] = sL[f * MM_LDL + m];
  }
  lmax = warp_max(lmax);
  if (lane == 0) s_red[tid >> 5] = lmax;
  __syncthreads();
  if (tid == 0) {
    float v = s_red[0];
    for (int i = 1; i < 8; ++i) v = fmaxf(v, s_red[i]);
    if (v > -INFINITY) atomicMax(umax + b, float_to_ordered(v));
  }
}

__global__ void __launch_bounds__(256)
logmel_finish_kernel(const float* __restrict__ logspec, const unsigned int* __restrict__ umax, float* __restrict__ out_f32,
                     __nv_bfloat16* __restrict__ out_bf16) {
  const int b = blockIdx.y;
  const float floor_v = ordered_to_float(umax[b]) - 8.0f;                 // WF:79-82
  const int64_t base = int64_t(b) * TASTE_N_FRAMES * TASTE_N_MELS;
  const int n4 = TASTE_N_FRAMES * TASTE_N_MELS / 4;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 v = reinterpret_cast<const float4*>(logspec + base)[i];
    v.x = (fmaxf(v.x, floor_v) + 4.0f) / 4.0f;                            // WF:83
    v.y = (fmaxf(v.y, floor_v) + 4.0f) / 4.0f;
    v.z = (fmaxf(v.z, floor_v) + 4.0f) / 4.0f;
    v.w = (fmaxf(v.w, floor_v) + 4.0f) / 4.0f;
    if (out_f32) reinterpret_cast<float4*>(out_f32 + base)[i] = v;
    if (out_bf16) {
      uint2 u;
      u.x = pack_act2(v.x, v.y);
      u.y = pack_act2(v.z, v.w);
      reinterpret_cast<uint2*>(out_bf16 + base)[i] = u;
    }
  }
}

static int g_logmel_mode = 0;
void set_logmel_mode(int mode) { g_logmel_mode = mode; }

int launch_logmel(const taste_weights_t& w, const float* wav, const int32_t* n_samples, int batch, int64_t wav_stride,
                  float* feats_f32, void* feats_bf16, float* scratch_logspec, unsigned int* scratch_max,
                  void* scratch_planes, float* scratch_spectrum, cudaStream_t stream) {
  if (!wav || (!feats_f32 && !feats_bf16)) return set_error(TASTE_E_ARG, "logmel: null pointer");
  if (!w.dft_cos || !w.dft_sin || !w.hann || !w.mel_start || !w.mel_count || !w.mel_weight)
    return set_error(TASTE_E_ARG, "logmel: tables missing from the handle");
  if (batch <= 0) return 0;
  static bool configured_dev[kMaxDevices] = {};
  bool& configured = configured_dev[current_device()];
  if (!configured) {
    TASTE_CUDA_OK(cudaFuncSetAttribute(logmel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LM_SMEM));
    configured = true;
  }
  // un-normalised log spectrum goes to the fp32 output when present (normalised in place), else to scratch
  float* logspec = feats_f32 ? feats_f32 : scratch_logspec;
  TASTE_CUDA_OK(cudaMemsetAsync(scratch_max, 0, sizeof(unsigned int) * batch, stream));
  dim3 grid((TASTE_N_FRAMES + LM_FRAMES - 1) / LM_FRAMES, batch);
  const double feat_elems = double(batch) * TASTE_N_FRAMES * TASTE_N_MELS;
  const bool tensor_dft = g_logmel_mode == 0 && w.dft_w_bf16 && scratch_planes && scratch_spectrum;
  if (tensor_dft) {
    const double plane_bytes = double(batch) * 2 * double(LOGMEL_PLANE) * 2;
    const double spec_bytes = double(batch) * TASTE_N_FRAMES * TASTE_DFT_N * 4.0;
    {
      ProfScope ps(stream, KC_LOGMEL_SPLIT, double(batch) * double(LOGMEL_PLANE) * 4.0,
                   double(batch) * TASTE_N_SAMPLES * 4.0 + plane_bytes);
      dim3 g((unsigned)((LOGMEL_PLANE / 8 + 255) / 256), batch);
      logmel_split_kernel<<<g, 256, 0, stream>>>(wav, n_samples, wav_stride, static_cast<__nv_bfloat16*>(scratch_planes));
      TASTE_CUDA_OK(cudaGetLastError());
    }
    GemmDesc d;
    d.a = scratch_planes;
    d.k_inner = TASTE_DFT_K;
    d.s_count = 2;                                     // s = 0: hi plane, 1: lo plane
    d.rows_in = TASTE_N_FRAMES;
    d.rows_out = TASTE_N_FRAMES;
    d.batches = batch;
    d.s_stride = LOGMEL_PLANE * 2;
    d.r_stride = TASTE_HOP * 2;                        // overlapping frames: one hop per row
    d.b_stride = 2 * LOGMEL_PLANE * 2;
    d.taps = 3;                                        // K slabs hi(x) | hi(x) | lo(x)  against  hi(t) | lo(t) | hi(t)
    d.tap_s[0] = 0, d.tap_s[1] = 0, d.tap_s[2] = 1;
    d.w = w.dft_w_bf16;
    d.ab_bf16 = 1;                                     // split-bf16 planes and twiddles in both library flavours
    d.n = TASTE_DFT_N;
    d.out = scratch_spectrum;
    d.ldc = TASTE_DFT_N;
    d.epilogue = EPI_F32;
    d.kclass = KC_LOGMEL_DFT;
    // algorithmic bytes of the whole front-end: waveform in + features out (3.456 MB per utterance),
    // booked on its dominant kernel; the helper passes are booked with their own reads + writes
    d.alg_bytes = double(batch) * TASTE_N_SAMPLES * 4.0 + feat_elems * 4.0;
    if (int rc = launch_gemm(d, stream)) return rc;
    {
      ProfScope ps(stream, KC_LOGMEL_MEL, double(batch) * TASTE_N_FRAMES * (3.0 * 201 + 2.0 * 394), spec_bytes + feat_elems * 4.0);
      dim3 g((TASTE_N_FRAMES + MM_FRAMES - 1) / MM_FRAMES, batch);
      logmel_mel_kernel<<<g, 256, 0, stream>>>(scratch_spectrum, w.mel_start, w.mel_count, w.mel_weight, logspec, scratch_max);
      TASTE_CUDA_OK(cudaGetLastError());
    }
  } else {
    // algorithmic bytes of the whole front-end: waveform in + features out (3.456 MB per utterance),
    // booked on the tile kernel; the finish pass is booked with its own read + write
    ProfScope ps(stream, KC_LOGMEL_TILE, double(batch) * TASTE_N_FRAMES * (2.0 * 2 * 199 * 101 + 2.0 * 394),
                 double(batch) * TASTE_N_SAMPLES * 4.0 + feat_elems * 4.0);
    logmel_tile_kernel<<<grid, LM_THREADS, LM_SMEM, stream>>>(wav, n_samples, wav_stride, w.dft_cos, w.dft_sin, w.hann,
                                                              w.mel_start, w.mel_count, w.mel_weight, logspec, scratch_max);
    TASTE_CUDA_OK(cudaGetLastError());
  }
  {
    dim3 grid2(48, batch);
    ProfScope ps(stream, KC_LOGMEL_FINISH, feat_elems * 3.0,
                 feat_elems * (4.0 + (feats_f32 ? 4.0 : 0.0) + (feats_bf16 ? 2.0 : 0.0)));
    logmel_finish_kernel<<<grid2, 256, 0, stream>>>(logspec, scratch_max, feats_f32,
                                                    static_cast<__nv_bfloat16*>(feats_bf16));
    TASTE_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

}  // namespace taste
