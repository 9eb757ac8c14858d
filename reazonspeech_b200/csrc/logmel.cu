// Fused log-mel frontend (N1), ONE kernel: replaces NeMo's AudioToMelSpectrogramPreprocessor / FilterbankFeatures.forward
// reached through model.transcribe (pkg/nemo-asr/src/transcribe.py:48-53):
//   pre-emphasis -> framing (center=True, zero pad) -> Hann(400) in a 512 frame -> 512-point real FFT -> |X|^2
//   -> slaney mel (80 x 257, sparse triangular) -> log(x + 2^-24)  -> per-feature (mean, 1 / (unbiased std + 1e-5)) over
//   the utterance's valid frames.
// The features leave the kernel UN-normalised together with the statistics; the only consumer on the transcribe path
// (sub_conv0_dw1_kernel, subsample.cu) applies (x - mean) * rstd and the zero tail while it stages its mel patch, so the
// 31.7 MB feature tensor is written once and read once.  rs_logmel (the stage entry point of the C ABI) runs
// mel_apply_norm_kernel afterwards to hand out NeMo's normalised, zero-tailed tensor.
//
// Memory-bound by traffic (2.98 MB per 30 s clip; 15 us at the measured copy bandwidth for 32 clips), instruction-issue
// bound in practice: the design minimises warp instructions per frame.
//   * CTA = kFrames consecutive frames of one utterance, 256 threads = 16 half-warps, a half-warp per frame in flight.
//   * Samples are staged once with 16-byte loads and PRE-EMPHASISED while staged (each sample serves 3.2 frames).
//   * 512-point real FFT = 256-point complex FFT as 16 x 16 held in registers (fft16.cuh), one transpose through shared
//     memory, then the paired real-FFT split (logmel_frame.cuh): E and W.O are formed once per bin pair.
//   * Mel: 80 filters sorted by width and dealt 16 at a time to the 16 lanes ("slots"), zero-padded to the slot's widest
//     filter: 38 taps per lane in straight-line, divergence-free loops (31.25 is the unpadded mean).
//   * Statistics: per CTA and feature (sum x, sum x^2) in registers across the CTA's frames, one partial row per CTA; the
//     LAST CTA of an utterance to finish (atomic ticket) adds the partial rows in tile order -- a fixed order whichever CTA
//     is last, so the result is bit-reproducible -- in double precision and writes (mean, rstd).
#include "common.cuh"
#include "logmel.h"
#include "logmel_frame.cuh"

namespace rs {

namespace {

constexpr int kLmThreads = 256;
constexpr int kLmHalfWarps = kLmThreads / lm::kLanes;      // 16 frames in flight per CTA
constexpr int kScrUsed = 2 * lm::kLanes * lm::kTrPitch;    // per half-warp scratch: transpose buffer (544 floats), later power spectrum + mel row
constexpr int kScrFloats = kScrUsed + 16;                  // pitch = 16 mod 32 banks: the two half-warps of a warp run the same access pattern on
                                                           // their own scratch, and with a pitch of 0 mod 32 every 32-bit access of theirs collided 2-way
constexpr int kPwOff = 0;                                  // power spectrum: 257 floats
constexpr int kMelRowOff = 264;                            // raw mel sums: up to 16 * kMaxSlots floats, then 16 parking rows
static_assert(kMelRowOff + lm::kLanes * lm::kMaxSlots + lm::kLanes <= kScrUsed && kScrFloats % 32 == 16, "scratch layout");   // + 16 parking rows of empty slots

struct LmSmem {          // float offsets into dynamic shared memory
  int y, win, twb, twx, mw, meta, scr, total;
};

__host__ __device__ inline LmSmem lm_layout(int frames, int hop, int n_taps) {
  LmSmem s;
  int o = 0;
  s.y = o; o += ((frames - 1) * hop + lm::kNfft + 3) & ~3;
  s.win = o; o += lm::kNfft;
  s.twb = o; o += 2 * 256;
  s.twx = o; o += 2 * lm::kPairs * lm::kLanes;
  s.mw = o; o += (n_taps * lm::kLanes + 3) & ~3;
  s.meta = o; o += kLmMetaInts;
  s.scr = o; o += kLmHalfWarps * kScrFloats;
  s.total = o;
  return s;
}

template <int K2>
__device__ __forceinline__ void split_step(const float2 (&v)[16], int t, int partner, const float2* __restrict__ s_twx, float* __restrict__ s_pw) {
  const float2 mine = lm::provided<K2>(v, t);
  float2 zc;
  zc.x = __shfl_sync(0xffffffffu, mine.x, partner, lm::kLanes);
  zc.y = __shfl_sync(0xffffffffu, mine.y, partner, lm::kLanes);
  float pp, pm;
  lm::split_pair(v[K2], zc, s_twx[K2 * lm::kLanes + t], pp, pm);
  s_pw[lm::bin_plus(t, K2)] = pp;
  s_pw[lm::bin_minus(t, K2)] = pm;
}

// PCM16 ingest: the staging loop reads int16 samples (half the HBM and PCIe bytes) and scales by 2^-15, which is exactly
// what decoding a 16-bit WAV to float32 does (soundfile / librosa: sample / 32768), so both input types give identical features.
template <bool kI16> struct LmSample { using type = float; };
template <> struct LmSample<true> { using type = int16_t; };
template <bool kI16>
__device__ __forceinline__ float lm_load(const typename LmSample<kI16>::type* p) {
  if constexpr (kI16) return static_cast<float>(__ldg(p)) * (1.0f / 32768.0f);
  else return __ldg(p);
}

template <int kFrames, int kSlots, bool kI16>
__global__ void __launch_bounds__(kLmThreads, kFrames <= 32 ? 3 : 2)
logmel_fused_kernel(const typename LmSample<kI16>::type* __restrict__ wav, const int32_t* __restrict__ len, int L_max, float* __restrict__ mel,
                    int32_t* __restrict__ mel_len, LmTables tb, float* __restrict__ partials, float* __restrict__ stats,
                    unsigned int* __restrict__ tickets, int F_max, int tiles_max, int n_mels, int hop, float preemph,
                    float guard, float eps) {
  extern __shared__ __align__(16) float lm_smem[];
  __shared__ int s_last;
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kFrames;
  const int n = len[b];
  // valid frames = FilterbankFeatures.get_seq_len = n / hop (n_fft even): one less than the centred STFT yields; NeMo masks
  // that final frame to zero and keeps it out of the statistics (config.py::mel_valid)
  const int n_frames = n / hop;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    mel_len[b] = n_frames;
    if (n_frames == 0) tickets[b] = 0;                     // no CTA of this utterance takes a ticket: nothing to normalise
  }
  if (f0 >= n_frames) return;
  const int tiles_b = (n_frames + kFrames - 1) / kFrames;

  const LmSmem L = lm_layout(kFrames, hop, tb.n_taps);
  float* s_y = lm_smem + L.y;
  const float* s_win = lm_smem + L.win;
  const float2* s_twb = reinterpret_cast<const float2*>(lm_smem + L.twb);
  const float2* s_twx = reinterpret_cast<const float2*>(lm_smem + L.twx);
  const float* s_mw = lm_smem + L.mw;
  const int* s_meta = reinterpret_cast<const int*>(lm_smem + L.meta);

  // ---- stage the tile's samples, pre-emphasised: y[i] = x[i] - preemph * x[i-1] for 0 <= i < n (x[-1] = 0), else 0
  {
    const int n_stage = (kFrames - 1) * hop + lm::kNfft;
    const int start = f0 * hop - lm::kHalf;                // global sample index of s_y[0]
    using Sample = typename LmSample<kI16>::type;
    const Sample* xw = wav + static_cast<size_t>(b) * L_max;
    const bool vec = ((reinterpret_cast<uintptr_t>(xw) | static_cast<uintptr_t>(start * sizeof(Sample))) & (4 * sizeof(Sample) - 1)) == 0;
    // all of the thread's loads are issued before the first value is used (the tile is (kFrames - 1) * hop + 512 samples: at most
    // kStageIters groups of four per thread), so one trip to L2 / HBM covers the whole staging instead of one per group
    constexpr int kStageIters = ((kFrames - 1) * 160 + lm::kNfft + 4 * kLmThreads - 1) / (4 * kLmThreads);
    if (n_stage <= kStageIters * 4 * kLmThreads) {
      float x[kStageIters][5];
#pragma unroll
      for (int it = 0; it < kStageIters; ++it) {
        const int i = (threadIdx.x + it * kLmThreads) * 4, idx = start + i;
        if (vec && idx >= 4 && idx + 3 < n && i + 3 < n_stage) {
          x[it][0] = lm_load<kI16>(xw + idx - 1);
          if constexpr (kI16) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(xw + idx));       // four int16 samples
            constexpr float k = 1.0f / 32768.0f;
            x[it][1] = static_cast<float>(static_cast<int16_t>(q.x & 0xffffu)) * k; x[it][2] = static_cast<float>(static_cast<int16_t>(q.x >> 16)) * k;
            x[it][3] = static_cast<float>(static_cast<int16_t>(q.y & 0xffffu)) * k; x[it][4] = static_cast<float>(static_cast<int16_t>(q.y >> 16)) * k;
          } else {
            const float4 q = __ldg(reinterpret_cast<const float4*>(xw + idx));
            x[it][1] = q.x; x[it][2] = q.y; x[it][3] = q.z; x[it][4] = q.w;
          }
        } else if (i < n_stage) {
#pragma unroll
          for (int k = 0; k < 5; ++k) { const int g = idx - 1 + k; x[it][k] = (g >= 0 && g < n) ? lm_load<kI16>(xw + g) : 0.0f; }
        }
      }
#pragma unroll
      for (int it = 0; it < kStageIters; ++it) {
        const int i = (threadIdx.x + it * kLmThreads) * 4, idx = start + i;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int g = idx + k;
          if (i + k < n_stage) s_y[i + k] = (g >= 0 && g < n) ? fmaf(-preemph, x[it][k], x[it][k + 1]) : 0.0f;
        }
      }
    } else
    for (int i = threadIdx.x * 4; i < n_stage; i += kLmThreads * 4) {
      const int idx = start + i;
      float x[5];                                          // x[idx-1 .. idx+3]
      if (vec && idx >= 4 && idx + 3 < n && i + 3 < n_stage) {
        x[0] = lm_load<kI16>(xw + idx - 1);
        if constexpr (kI16) {
          const uint2 q = __ldg(reinterpret_cast<const uint2*>(xw + idx));       // four int16 samples
          constexpr float k = 1.0f / 32768.0f;
          x[1] = static_cast<float>(static_cast<int16_t>(q.x & 0xffffu)) * k; x[2] = static_cast<float>(static_cast<int16_t>(q.x >> 16)) * k;
          x[3] = static_cast<float>(static_cast<int16_t>(q.y & 0xffffu)) * k; x[4] = static_cast<float>(static_cast<int16_t>(q.y >> 16)) * k;
        } else {
          const float4 q = __ldg(reinterpret_cast<const float4*>(xw + idx));
          x[1] = q.x; x[2] = q.y; x[3] = q.z; x[4] = q.w;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 5; ++k) { const int g = idx - 1 + k; x[k] = (g >= 0 && g < n) ? lm_load<kI16>(xw + g) : 0.0f; }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int g = idx + k;
        if (i + k < n_stage) s_y[i + k] = (g >= 0 && g < n) ? fmaf(-preemph, x[k], x[k + 1]) : 0.0f;
      }
    }
    float* dst = lm_smem + L.win;
    for (int i = threadIdx.x; i < lm::kNfft; i += kLmThreads) dst[i] = tb.window[i];
    dst = lm_smem + L.twb;
    for (int i = threadIdx.x; i < 512; i += kLmThreads) dst[i] = tb.tw_b[i];
    dst = lm_smem + L.twx;
    for (int i = threadIdx.x; i < 2 * lm::kPairs * lm::kLanes; i += kLmThreads) dst[i] = tb.tw_x[i];
    dst = lm_smem + L.mw;
    for (int i = threadIdx.x; i < tb.n_taps * lm::kLanes; i += kLmThreads) dst[i] = tb.mel_w[i];
    int* dm = reinterpret_cast<int*>(lm_smem + L.meta);
    for (int i = threadIdx.x; i < kLmMetaInts; i += kLmThreads) dm[i] = tb.mel_meta[i];
  }
  __syncthreads();

  const int hw = threadIdx.x >> 4, t = threadIdx.x & 15;   // half-warp = frame slot, lane within the frame
  float* s_scr = lm_smem + L.scr + hw * kScrFloats;
  float2* s_tr = reinterpret_cast<float2*>(s_scr);
  float* s_pw = s_scr + kPwOff;
  float* s_o = s_scr + kMelRowOff;
  const int partner = lm::partner_lane(t);
  const int n_slots = s_meta[0];
  float sum[kSlots], sq[kSlots];                           // kSlots = ceil(n_mels / 16): features t, t + 16, ... of this lane
#pragma unroll
  for (int j = 0; j < kSlots; ++j) { sum[j] = 0.f; sq[j] = 0.f; }

  for (int it = 0; it < kFrames / kLmHalfWarps; ++it) {
    const int fi = it * kLmHalfWarps + hw;
    const int f = f0 + fi;
    if (f0 + (fi & ~1) >= n_frames) break;                 // warp-uniform: both frames of this warp are beyond the utterance
    const bool live = f < n_frames;
    // ---- windowed samples of this lane: complex z[16 n1 + t] = (y[32 n1 + 2t], y[32 n1 + 2t + 1]) * window
    float2 v[16];
    {
      const float2* yp = reinterpret_cast<const float2*>(s_y + fi * hop) + t;
      const float2* wp = reinterpret_cast<const float2*>(s_win) + t;
#pragma unroll
      for (int n1 = 0; n1 < 16; ++n1) {
        const float2 y = yp[16 * n1], w = wp[16 * n1];
        v[n1] = make_float2(y.x * w.x, y.y * w.y);
      }
    }
    fft16(v);                                              // over n1: v[k1] = A[t][k1]
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float2 w = s_twb[k1 * lm::kLanes + t];
      s_tr[k1 * lm::kTrPitch + t] = make_float2(v[k1].x * w.x - v[k1].y * w.y, v[k1].x * w.y + v[k1].y * w.x);
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = s_tr[t * lm::kTrPitch + n2];
    __syncwarp();                                          // the power spectrum reuses the transpose buffer
    fft16(v);                                              // over n2: v[k2] = Z[t + 16 k2]
    // ---- paired real-FFT split -> 4 |X|^2 for bins 0..256
    split_step<0>(v, t, partner, s_twx, s_pw); split_step<1>(v, t, partner, s_twx, s_pw);
    split_step<2>(v, t, partner, s_twx, s_pw); split_step<3>(v, t, partner, s_twx, s_pw);
    split_step<4>(v, t, partner, s_twx, s_pw); split_step<5>(v, t, partner, s_twx, s_pw);
    split_step<6>(v, t, partner, s_twx, s_pw); split_step<7>(v, t, partner, s_twx, s_pw);
    if (t == 0) s_pw[128] = 4.0f * (v[8].x * v[8].x + v[8].y * v[8].y);
    __syncwarp();
    // ---- this lane's mel filters, one per slot; every lane runs the slot's (padded) tap count
    {
      const float* wp = s_mw + t;
      for (int s = 0; s < n_slots; ++s) {
        const int c = s_meta[8 + s];
        const float* pp = s_pw + s_meta[kLmMetaStart + s * lm::kLanes + t];
        float acc = 0.f;
#pragma unroll 2
        for (int j = 0; j < c; ++j) acc = fmaf(wp[j * lm::kLanes], pp[j], acc);
        wp += c * lm::kLanes;
        s_o[s_meta[kLmMetaOut + s * lm::kLanes + t]] = acc;
      }
    }
    __syncwarp();
    // ---- log, store, statistics: lane t owns features t, t + 16, ...
    {
      float* orow = mel + (static_cast<size_t>(b) * F_max + f) * n_mels;
#pragma unroll
      for (int j = 0; j < kSlots; ++j) {
        const int m = t + lm::kLanes * j;
        if (m < n_mels && live) {
          const float x = __logf(s_o[m] + guard);
          orow[m] = x;
          sum[j] += x;
          sq[j] = fmaf(x, x, sq[j]);
        }
      }
    }
    __syncwarp();
  }

  // ---- CTA partial sums in a fixed order (half-warp 0..15), then the utterance's last CTA finalises
#pragma unroll
  for (int j = 0; j < kSlots; ++j) {
    const int m = t + lm::kLanes * j;
    if (m < n_mels) { s_scr[2 * m] = sum[j]; s_scr[2 * m + 1] = sq[j]; }
  }
  __syncthreads();
  float* prow = partials + (static_cast<size_t>(b) * tiles_max + blockIdx.x) * 2 * n_mels;
  for (int i = threadIdx.x; i < 2 * n_mels; i += kLmThreads) {
    float a = 0.f;
#pragma unroll
    for (int h = 0; h < kLmHalfWarps; ++h) a += lm_smem[L.scr + h * kScrFloats + i];
    prow[i] = a;
    __threadfence();                                       // by the writers only: the row is visible before this CTA's ticket is
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(tickets + b, 1u) == static_cast<unsigned>(tiles_b - 1));
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int m = threadIdx.x; m < n_mels; m += kLmThreads) {
    double sx = 0.0, sxx = 0.0;
    const float* p = partials + static_cast<size_t>(b) * tiles_max * 2 * n_mels + 2 * m;
    for (int tile = 0; tile < tiles_b; ++tile) {
      sx += static_cast<double>(__ldcg(p + static_cast<size_t>(tile) * 2 * n_mels));
      sxx += static_cast<double>(__ldcg(p + static_cast<size_t>(tile) * 2 * n_mels + 1));
    }
    const double mean = sx / n_frames;
    double var = (sxx - sx * mean) / (n_frames > 1 ? n_frames - 1 : 1);     // unbiased, as torch.std
    var = var > 0.0 ? var : 0.0;
    stats[(static_cast<size_t>(b) * n_mels + m) * 2] = static_cast<float>(mean);
    stats[(static_cast<size_t>(b) * n_mels + m) * 2 + 1] = static_cast<float>(1.0 / (sqrt(var) + static_cast<double>(eps)));
  }
  if (threadIdx.x == 0) tickets[b] = 0;                    // ready for the next launch
}

// rs_logmel only: (x - mean) * rstd over the valid frames, zero tail -- NeMo's normalised feature tensor
__global__ void __launch_bounds__(256)
mel_apply_norm_kernel(float* __restrict__ mel, const int32_t* __restrict__ mel_len, const float* __restrict__ stats, int F_max, int n_mels) {
  const int b = blockIdx.y;
  const int nf = mel_len[b];
  const size_t total = static_cast<size_t>(F_max) * n_mels;
  float* base = mel + static_cast<size_t>(b) * total;
  const float* st = stats + static_cast<size_t>(b) * n_mels * 2;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int f = static_cast<int>(i / n_mels), m = static_cast<int>(i % n_mels);
    base[i] = f < nf ? (base[i] - st[2 * m]) * st[2 * m + 1] : 0.0f;
  }
}

template <int kFrames, int kSlots, bool kI16>
cudaError_t launch_fused(const LogmelArgs& a, cudaStream_t stream) {
  const LmSmem L = lm_layout(kFrames, a.hop, a.tb.n_taps);
  const size_t smem = static_cast<size_t>(L.total) * 4;
  if (smem > 200 * 1024) return cudaErrorInvalidValue;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(logmel_fused_kernel<kFrames, kSlots, kI16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_once.set();
  }
  const int F_max = a.L_max / a.hop + 1;
  const int tiles = logmel_tiles(a.L_max, a.hop);
  if (tiles * kLmTileFrames < F_max - 1) return cudaErrorInvalidValue;
  const dim3 grid((F_max - 1 + kFrames - 1) / kFrames > 0 ? (F_max - 1 + kFrames - 1) / kFrames : 1, a.B);
  logmel_fused_kernel<kFrames, kSlots, kI16><<<grid, kLmThreads, smem, stream>>>(static_cast<const typename LmSample<kI16>::type*>(a.wav), a.len, a.L_max, a.mel, a.mel_len, a.tb, a.partials, a.stats,
                                                                   a.tickets, F_max, tiles, a.n_mels, a.hop, a.preemph, a.guard, a.eps);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_logmel_fused(const LogmelArgs& a, cudaStream_t stream) {
  if (a.n_fft != lm::kNfft || a.win > lm::kNfft || a.n_mels > lm::kLanes * lm::kMaxSlots || a.hop <= 0 || (a.hop & 1) ||
      a.tb.n_taps <= 0 || a.tb.n_taps > 128)
    return cudaErrorInvalidValue;
  cudaError_t e;
  if (a.n_mels <= 5 * lm::kLanes) e = a.wav_i16 ? launch_fused<kLmTileFrames, 5, true>(a, stream) : launch_fused<kLmTileFrames, 5, false>(a, stream);
  else e = a.wav_i16 ? launch_fused<kLmTileFrames, lm::kMaxSlots, true>(a, stream) : launch_fused<kLmTileFrames, lm::kMaxSlots, false>(a, stream);
  if (e != cudaSuccess || !a.normalise_in_place) return e;
  const int F_max = a.L_max / a.hop + 1;
  mel_apply_norm_kernel<<<dim3(64, a.B), 256, 0, stream>>>(a.mel, a.mel_len, a.stats, F_max, a.n_mels);
  return cudaGetLastError();
}

}  // namespace rs
