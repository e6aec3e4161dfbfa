// Corpus ingest: polyphase windowed-sinc resampling to 16 kHz fused with the channel mean, i.e.
//   waveform = torchaudio.transforms.Resample(orig_sr, 16000)(speech_pt).mean(0)            (DS:52-60)
// for a whole batch of decoded PCM arrays, written straight into the [batch, wav_stride] waveform buffer that the
// log-mel kernel reads (no per-utterance tensors, no padded copy).
//
// torchaudio's algorithm (functional._apply_sinc_resample_kernel): with orig / new reduced by their gcd, pad the signal
// with `width` zeros on the left and `width + orig` on the right, then
//   y[f * new + p] = sum_{k < 2*width+orig} kernel[p][k] * xpad[f * orig + k],     keep the first ceil(new * n / orig).
// Most of each phase's 2*width+orig taps are the clamped window's tail (|tap| ~ 1e-35, "almost zero" in torchaudio's
// own words), so the host passes, per phase, the first significant tap and a dense run of `knz` taps.
//
// HBM-bound by contract: 4 * (channels * n_in + n_out) bytes per utterance.  One CTA produces a tile of `frames`
// input strides x `new` phases; each channel's input span is staged through shared memory with coalesced loads
// (every input sample is read from HBM once), the phase taps live in shared memory (row stride padded to an odd
// number of words), accumulation is fp32 in tap order, and the channel mean is applied at the store.
#include "common.cuh"
#include "internal.h"

namespace taste {

constexpr int RS_THREADS = 256;
constexpr int RS_MAX_PER_THREAD = 8;                 // outputs per thread per tile
constexpr int RS_TILE_OUT = RS_THREADS * RS_MAX_PER_THREAD;
constexpr int RS_MAX_SPAN = 8192;                    // floats of one channel's input span held in shared memory

struct ResampleParams {
  const float* in;
  const int64_t* in_off;       // [batch + 1] element offsets: utterance b is [channels[b], n_in[b]] row-major at in + in_off[b]
  const int32_t* channels;     // [batch]
  const int32_t* n_in;         // [batch] samples per channel
  const float* taps;           // [nw][knz_ld]
  const int32_t* kstart;       // [nw] first tap (index into the full 2*width+orig kernel) of each phase's run
  int orig, nw, width, knz, knz_ld;
  int frames;                  // input strides per tile
  int span;                    // floats staged per channel per tile
  float* wav;
  int64_t wav_stride;
  int32_t* n_out;              // [batch] ceil(nw * n_in / orig), NOT clipped to wav_stride (nullable)
};

__global__ void __launch_bounds__(RS_THREADS)
resample_mean_kernel(const ResampleParams p) {
  extern __shared__ float rs_smem[];
  float* taps_s = rs_smem;                                   // nw * knz_ld
  float* xs = rs_smem + ((p.nw * p.knz_ld + 3) & ~3);       // span
  int* ks = reinterpret_cast<int*>(xs + p.span);             // nw
  const int b = blockIdx.y;
  const int n = p.n_in[b];
  const int C = p.channels[b];
  const int64_t target = (int64_t(p.nw) * n + p.orig - 1) / p.orig;
  if (blockIdx.x == 0 && threadIdx.x == 0 && p.n_out) p.n_out[b] = int32_t(target < 2147483647LL ? target : 2147483647LL);
  const int64_t limit = target < p.wav_stride ? target : p.wav_stride;       // outputs this row stores
  const int tile_out = p.frames * p.nw;
  const int64_t o0 = int64_t(blockIdx.x) * tile_out;
  if (o0 >= limit) return;
  for (int i = threadIdx.x; i < p.nw * p.knz_ld; i += RS_THREADS) taps_s[i] = p.taps[i];
  for (int i = threadIdx.x; i < p.nw; i += RS_THREADS) ks[i] = p.kstart[i];
  const int64_t f0 = int64_t(blockIdx.x) * p.frames;
  const int64_t x0 = f0 * p.orig - p.width;                  // signal index of xs[0]
  float acc[RS_MAX_PER_THREAD];
#pragma unroll
  for (int u = 0; u < RS_MAX_PER_THREAD; ++u) acc[u] = 0.f;
  const float* xin = p.in + p.in_off[b];
  for (int c = 0; c < C; ++c) {
    __syncthreads();                                         // previous channel's reads done (and taps visible)
    const float* xc = xin + int64_t(c) * n;
    for (int j = threadIdx.x; j < p.span; j += RS_THREADS) {
      const int64_t g = x0 + j;
      xs[j] = (g >= 0 && g < n) ? __ldg(xc + g) : 0.f;       // the zero padding of _apply_sinc_resample_kernel
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < RS_MAX_PER_THREAD; ++u) {
      const int ol = threadIdx.x + u * RS_THREADS;
      if (ol < tile_out) {
        const int f = ol / p.nw;
        const int ph = ol - f * p.nw;
        const float* t = taps_s + ph * p.knz_ld;
        const float* x = xs + f * p.orig + ks[ph];
        float a = 0.f;
        for (int k = 0; k < p.knz; ++k) a = fmaf(t[k], x[k], a);
        acc[u] += a;
      }
    }
  }
  const float inv_c = 1.0f / float(C);
  float* out = p.wav + int64_t(b) * p.wav_stride;
#pragma unroll
  for (int u = 0; u < RS_MAX_PER_THREAD; ++u) {
    const int ol = threadIdx.x + u * RS_THREADS;
    const int64_t o = o0 + ol;
    if (ol < tile_out && o < limit) out[o] = C == 1 ? acc[u] : acc[u] * inv_c;
  }
}

int launch_resample_mean(const float* in, const int64_t* in_off, const int32_t* channels, const int32_t* n_in, int batch,
                         int orig, int nw, int width, const float* taps, const int32_t* kstart, int knz, int knz_ld,
                         int max_out, double total_in_elems, double total_out_elems, float* wav, int64_t wav_stride, int32_t* n_out,
                         cudaStream_t stream) {
  if (!in || !in_off || !channels || !n_in || !taps || !kstart || !wav) return set_error(TASTE_E_ARG, "resample: null pointer");
  if (batch <= 0 || max_out <= 0) return 0;
  if (orig <= 0 || nw <= 0 || width < 0 || knz <= 0 || knz_ld < knz || wav_stride <= 0)
    return set_error(TASTE_E_ARG, "resample: bad geometry (orig %d new %d width %d knz %d ld %d)", orig, nw, width, knz, knz_ld);
  if (nw > RS_TILE_OUT) return set_error(TASTE_E_SHAPE, "resample: reduced new_freq %d > %d phases", nw, RS_TILE_OUT);
  const int kfull = 2 * width + orig;
  if (orig + kfull > RS_MAX_SPAN) return set_error(TASTE_E_SHAPE, "resample: reduced orig_freq %d too large", orig);
  ResampleParams p;
  p.in = in; p.in_off = in_off; p.channels = channels; p.n_in = n_in; p.taps = taps; p.kstart = kstart;
  p.orig = orig; p.nw = nw; p.width = width; p.knz = knz; p.knz_ld = knz_ld;
  int frames = RS_TILE_OUT / nw;
  const int by_span = (RS_MAX_SPAN - kfull) / orig;
  if (frames > by_span) frames = by_span;
  if (frames < 1) frames = 1;
  p.frames = frames;
  p.span = frames * orig + kfull;
  p.wav = wav; p.wav_stride = wav_stride; p.n_out = n_out;
  const size_t smem = (size_t((nw * knz_ld + 3) & ~3) + p.span + nw) * sizeof(float);
  if (smem > 200 * 1024) return set_error(TASTE_E_SHAPE, "resample: tap table needs %zu bytes of shared memory", smem);
  static size_t configured_dev[kMaxDevices] = {};
  size_t& configured = configured_dev[current_device()];
  if (configured == 0) configured = 48 * 1024;
  if (smem > configured) {
    TASTE_CUDA_OK(cudaFuncSetAttribute(resample_mean_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int64_t lim = max_out < wav_stride ? max_out : wav_stride;
  const int tiles = int((lim + int64_t(frames) * nw - 1) / (int64_t(frames) * nw));
  dim3 grid(tiles, batch);
  // algorithmic bytes: every input sample read once, every output sample written once
  ProfScope ps(stream, KC_RESAMPLE, 2.0 * knz * total_in_elems * nw / orig,
               4.0 * (total_in_elems + total_out_elems));
  resample_mean_kernel<<<grid, RS_THREADS, smem, stream>>>(p);
  TASTE_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace taste

extern "C" int taste_resample_mean_f32(const float* in, const int64_t* in_offsets, const int32_t* channels,
                                       const int32_t* n_in, int batch, int orig_reduced, int new_reduced, int width,
                                       const float* taps, const int32_t* tap_start, int taps_per_phase, int taps_ld,
                                       int max_out, int64_t total_in_elems, int64_t total_out_elems, float* wav,
                                       int64_t wav_stride,
                                       int32_t* n_out, void* stream) {
  return taste::launch_resample_mean(in, in_offsets, channels, n_in, batch, orig_reduced, new_reduced, width, taps,
                                     tap_start, taps_per_phase, taps_ld, max_out, double(total_in_elems),
                                     double(total_out_elems), wav,
                                     wav_stride, n_out, static_cast<cudaStream_t>(stream));
}
