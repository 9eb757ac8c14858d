// Fused log-mel frontend (N1): replaces NeMo's AudioToMelSpectrogramPreprocessor /
// FilterbankFeatures.forward reached through model.transcribe (pkg/nemo-asr/src/transcribe.py:48-53):
//   pre-emphasis -> framing (center=True, zero pad) -> Hann(400) in a 512 frame -> 512-pt real FFT
//   -> |X|^2 -> slaney mel (80 x 257, sparse triangular) -> log(x + 2^-24)            [kernel 1]
//   -> per-feature mean / unbiased std over the valid frames -> normalise -> zero tail  [kernel 2]
//
// Kernel 1: one CTA per (32-frame tile, utterance).  The tile's sample range is staged once in
// shared memory with coalesced loads (frames overlap 2.5x), together with the window, the FFT
// twiddles and the sparse mel table.  Each warp owns frames: a 512-pt real FFT is computed as a
// 256-pt complex radix-4 Stockham FFT in shared memory plus the even/odd split.
// Output is time-major [B, F_max, n_mels] so the subsampling convs read channels-last.
#include "common.cuh"
#include "fft16.cuh"
#include "kernels.h"

namespace rs {

constexpr int kNfft = 512;
constexpr int kHalf = 256;
constexpr int kFramesPerBlock = 32;
constexpr int kFeWarps = 8;
constexpr int kMelMaxW = 40;   // widest triangular filter in FFT bins (checked at pack time)

struct FeTables {             // device pointers, filled by the engine from the packed weights
  const float* window;        // [512]  Hann(400) centred in the FFT frame
  const float* tw256;         // [256][2]  exp(-2*pi*i*k/256)
  const float* tw512;         // [257][2]  exp(-2*pi*i*k/512)
  const int32_t* mel_start;   // [n_mels]
  const int32_t* mel_count;   // [n_mels]
  const float* mel_w;         // [n_mels][kMelMaxW]
};

__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

__global__ void __launch_bounds__(32 * kFeWarps)
logmel_kernel(const float* __restrict__ wav, const int32_t* __restrict__ len, int L_max, float* __restrict__ mel,
              int32_t* __restrict__ mel_len, FeTables tb, int F_max, int n_mels, int hop, float preemph, float guard) {
  extern __shared__ __align__(16) uint8_t fe_smem[];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kFramesPerBlock;
  const int n = len[b];
  // valid frames = FilterbankFeatures.get_seq_len = n / hop (n_fft even): one less than the centred STFT yields;
  // NeMo masks that final frame to zero and keeps it out of the statistics (config.py::mel_valid)
  const int n_frames = n / hop;
  if (blockIdx.x == 0 && threadIdx.x == 0) mel_len[b] = n_frames;
  if (f0 >= n_frames) return;

  const int n_stage = (kFramesPerBlock - 1) * hop + kNfft + 1;      // +1: x[n-1] of the first sample
  float* s_x = reinterpret_cast<float*>(fe_smem);                    // [n_stage]
  float* s_win = s_x + ((n_stage + 3) & ~3);                         // [512]
  float2* s_tw256 = reinterpret_cast<float2*>(s_win + kNfft);        // [256]
  float2* s_tw512 = s_tw256 + kHalf;                                 // [257] (+1 pad)
  float* s_melw = reinterpret_cast<float*>(s_tw512 + kHalf + 2);     // [n_mels][kMelMaxW]
  int* s_mels = reinterpret_cast<int*>(s_melw + n_mels * kMelMaxW);  // [n_mels] start
  int* s_melc = s_mels + n_mels;                                     // [n_mels] count
  float2* s_fft = reinterpret_cast<float2*>(s_melc + n_mels + ((2 * n_mels) & 1)); // [warps][2][256]

  const int start = f0 * hop - kHalf - 1;                            // global index of s_x[0]
  const float* xw = wav + static_cast<size_t>(b) * L_max;
  for (int i = threadIdx.x; i < n_stage; i += blockDim.x) {
    const int idx = start + i;
    s_x[i] = (idx >= 0 && idx < n) ? __ldg(xw + idx) : 0.0f;
  }
  for (int i = threadIdx.x; i < kNfft; i += blockDim.x) s_win[i] = tb.window[i];
  for (int i = threadIdx.x; i < kHalf; i += blockDim.x) s_tw256[i] = reinterpret_cast<const float2*>(tb.tw256)[i];
  for (int i = threadIdx.x; i < kHalf + 1; i += blockDim.x) s_tw512[i] = reinterpret_cast<const float2*>(tb.tw512)[i];
  for (int i = threadIdx.x; i < n_mels * kMelMaxW; i += blockDim.x) s_melw[i] = tb.mel_w[i];
  for (int i = threadIdx.x; i < n_mels; i += blockDim.x) { s_mels[i] = tb.mel_start[i]; s_melc[i] = tb.mel_count[i]; }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float2* bufA = s_fft + warp * 2 * kHalf;
  float2* bufB = bufA + kHalf;

  for (int fi = warp; fi < kFramesPerBlock; fi += kFeWarps) {
    const int f = f0 + fi;
    if (f >= n_frames) break;                                        // warp-uniform
    // ---- windowed, pre-emphasised frame packed as 256 complex values
    const int off = fi * hop + 1;                                    // s_x index of frame sample 0
    const int g0 = f * hop - kHalf;                                  // global index of frame sample 0
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int k = lane + 32 * r;                                   // complex index: samples 2k, 2k+1
      float2 z;
      {
        const int j = 2 * k, gi = g0 + j;
        const float y = (gi >= 0 && gi < n) ? s_x[off + j] - preemph * s_x[off + j - 1] : 0.0f;
        z.x = y * s_win[j];
      }
      {
        const int j = 2 * k + 1, gi = g0 + j;
        const float y = (gi >= 0 && gi < n) ? s_x[off + j] - preemph * s_x[off + j - 1] : 0.0f;
        z.y = y * s_win[j];
      }
      bufA[k] = z;
    }
    __syncwarp();
    // ---- 256-pt complex FFT: 4 radix-4 Stockham passes (Ns = 1, 4, 16, 64), 64 butterflies each
    float2* src = bufA;
    float2* dst = bufB;
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int Ns = 1 << (2 * pass);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int j = lane + 32 * h;
        const int jm = j & (Ns - 1);
        const int tw = jm * (64 / Ns);                               // W256 exponent for r = 1
        float2 v0 = src[j], v1 = src[j + 64], v2 = src[j + 128], v3 = src[j + 192];
        if (pass > 0) {
          v1 = cmul(v1, s_tw256[tw]);
          v2 = cmul(v2, s_tw256[2 * tw]);
          v3 = cmul(v3, s_tw256[3 * tw]);
        }
        const float2 a0 = make_float2(v0.x + v2.x, v0.y + v2.y);
        const float2 a1 = make_float2(v0.x - v2.x, v0.y - v2.y);
        const float2 a2 = make_float2(v1.x + v3.x, v1.y + v3.y);
        const float2 a3 = make_float2(v1.y - v3.y, v3.x - v1.x);     // (v1 - v3) * (-i)
        const int j0 = ((j - jm) << 2) + jm;
        dst[j0] = make_float2(a0.x + a2.x, a0.y + a2.y);
        dst[j0 + Ns] = make_float2(a1.x + a3.x, a1.y + a3.y);
        dst[j0 + 2 * Ns] = make_float2(a0.x - a2.x, a0.y - a2.y);
        dst[j0 + 3 * Ns] = make_float2(a1.x - a3.x, a1.y - a3.y);
      }
      __syncwarp();
      float2* t = src; src = dst; dst = t;
    }
    // result is in `src` (== bufA after four swaps); power spectrum goes to `dst` as floats
    float* pw = reinterpret_cast<float*>(dst);
#pragma unroll
    for (int r = 0; r < 9; ++r) {
      const int k = lane + 32 * r;
      if (k <= kHalf) {
        const float2 zk = src[k & (kHalf - 1)];
        const float2 zc = src[(kHalf - k) & (kHalf - 1)];
        const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));       // (Zk + conj Zc)/2
        const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));      // (Zk - conj Zc)/(2i)
        const float2 wo = cmul(s_tw512[k], o);
        const float re = e.x + wo.x, im = e.y + wo.y;
        pw[k] = re * re + im * im;
      }
    }
    __syncwarp();
    // ---- sparse mel + log
    float* orow = mel + (static_cast<size_t>(b) * F_max + f) * n_mels;
    for (int m = lane; m < n_mels; m += 32) {
      const int s = s_mels[m], c = s_melc[m];
      const float* wr = s_melw + m * kMelMaxW;
      float acc = 0.f;
      for (int j = 0; j < c; ++j) acc = fmaf(wr[j], pw[s + j], acc);
      orow[m] = logf(acc + guard);
    }
    __syncwarp();
  }
}

// Kernel 2: per-feature normalisation over valid frames, two-pass variance (N-1), tail zeroed.
// grid (B, n_mels/16); block 512 = 16 features x 32 time lanes.
__global__ void __launch_bounds__(512)
mel_normalize_kernel(float* __restrict__ mel, const int32_t* __restrict__ mel_len, int F_max, int n_mels, float eps) {
  __shared__ float s_red[32][17];
  const int b = blockIdx.x;
  const int fm = threadIdx.x & 15, tl = threadIdx.x >> 4;
  const int m = blockIdx.y * 16 + fm;
  const int nf = mel_len[b];
  float* base = mel + static_cast<size_t>(b) * F_max * n_mels + m;
  const bool ok = m < n_mels;
  auto block_sum = [&](float v) -> float {
    s_red[tl][fm] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < 32; ++i) t += s_red[i][fm];
    __syncthreads();
    return t;
  };
  float s = 0.f;
  if (ok) for (int f = tl; f < nf; f += 32) s += base[static_cast<size_t>(f) * n_mels];
  const float mean = block_sum(s) / static_cast<float>(nf);
  float ss = 0.f;
  if (ok) for (int f = tl; f < nf; f += 32) { const float d = base[static_cast<size_t>(f) * n_mels] - mean; ss = fmaf(d, d, ss); }
  const float var = block_sum(ss) / static_cast<float>(nf > 1 ? nf - 1 : 1);
  const float inv = 1.0f / (sqrtf(var) + eps);
  if (ok) {
    for (int f = tl; f < nf; f += 32) { float* p = base + static_cast<size_t>(f) * n_mels; *p = (*p - mean) * inv; }
    for (int f = nf + tl; f < F_max; f += 32) base[static_cast<size_t>(f) * n_mels] = 0.0f;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// EXPERIMENT (RS_LOGMEL_VARIANT=B, unmeasured).  Same arithmetic as logmel_kernel up to the order of the FFT's additions,
// organised to issue about half the instructions per frame and to keep two independent frames in flight per warp:
//   * sixteen lanes per frame; the 256-point complex FFT is 16 x 16: lane t transforms the stride-16 samples t + 16 n1 in
//     registers (fft16.cuh), multiplies by W256^(t k1), the sixteen lanes transpose through shared memory (row pitch 17),
//     and a second register FFT over n2 leaves lane t with Z[t + 16 k2];
//   * the real-FFT split needs Z[256 - k]: that is lane (16 - t) & 15, register 15 - k2 (lane 0: itself, (16 - k2) & 15)
//     -- one pair of width-16 shuffles per bin;
//   * every lane owns about 31 mel taps (filters dealt longest first at pack time, never split across lanes);
//   * 64 frames per CTA, so the tables are staged once per 4x more frames than in logmel_kernel.
// The per-frame arithmetic is replayed on the CPU from this description by tests/test_logmel_b_host.py.
constexpr int kBFrames = 64;
constexpr int kBLaneBins = 8;
constexpr int kBLaneTaps = 47;
constexpr int kBTrPitch = 17;

__global__ void __launch_bounds__(32 * kFeWarps)
logmel_b_kernel(const float* __restrict__ wav, const int32_t* __restrict__ len, int L_max, float* __restrict__ mel,
                int32_t* __restrict__ mel_len, FeTables tb, FeTablesB tbb, int F_max, int n_mels, int hop, float preemph, float guard) {
  extern __shared__ __align__(16) uint8_t fe_smem[];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kBFrames;
  const int n = len[b];
  const int n_frames = n / hop;                                      // FilterbankFeatures.get_seq_len, as in logmel_kernel
  if (blockIdx.x == 0 && threadIdx.x == 0) mel_len[b] = n_frames;
  if (f0 >= n_frames) return;

  const int n_stage = (kBFrames - 1) * hop + kNfft + 1;
  float* s_x = reinterpret_cast<float*>(fe_smem);                    // [n_stage]
  float* s_win = s_x + ((n_stage + 3) & ~3);                         // [512]
  float2* s_twb = reinterpret_cast<float2*>(s_win + kNfft);          // [16][16]  W256^(t k1) at [k1][t]
  float2* s_twx = s_twb + 256;                                       // [16][16]  W512^(t + 16 k2) at [k2][t]
  float* s_lw = reinterpret_cast<float*>(s_twx + 256);               // [16][kBLaneTaps]
  int* s_lb = reinterpret_cast<int*>(s_lw + 16 * kBLaneTaps);        // [16][kBLaneBins]
  int* s_lnb = s_lb + 16 * kBLaneBins;                               // [16]
  float2* s_tr_all = reinterpret_cast<float2*>(s_lnb + 16);          // [16 half-warps][16 * 17]; later the frame's mel row
  float* s_pw_all = reinterpret_cast<float*>(s_tr_all + 16 * 16 * kBTrPitch);   // [16 half-warps][260]

  const int start = f0 * hop - kHalf - 1;
  const float* xw = wav + static_cast<size_t>(b) * L_max;
  for (int i = threadIdx.x; i < n_stage; i += blockDim.x) {
    const int idx = start + i;
    s_x[i] = (idx >= 0 && idx < n) ? __ldg(xw + idx) : 0.0f;
  }
  for (int i = threadIdx.x; i < kNfft; i += blockDim.x) s_win[i] = tb.window[i];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    s_twb[i] = reinterpret_cast<const float2*>(tbb.tw_b)[i];
    s_twx[i] = reinterpret_cast<const float2*>(tbb.tw_x)[i];
  }
  for (int i = threadIdx.x; i < 16 * kBLaneTaps; i += blockDim.x) s_lw[i] = tbb.lane_w[i];
  for (int i = threadIdx.x; i < 16 * kBLaneBins; i += blockDim.x) s_lb[i] = tbb.lane_bins[i];
  if (threadIdx.x < 16) s_lnb[threadIdx.x] = tbb.lane_nb[threadIdx.x];
  __syncthreads();

  const int hw = threadIdx.x >> 4, t = threadIdx.x & 15;            // half-warp = frame slot, lane within the frame
  float2* s_tr = s_tr_all + hw * 16 * kBTrPitch;
  float* s_pw = s_pw_all + hw * 260;
  float* s_o = reinterpret_cast<float*>(s_tr);                       // the transpose buffer is free again when the mel row is built
  const int partner = (16 - t) & 15;

  for (int it = 0; it < kBFrames / 16; ++it) {                       // both half-warps of a warp always run the whole body
    const int fi = it * 16 + hw;
    const int f = f0 + fi;
    const bool live = f < n_frames;
    // ---- windowed, pre-emphasised samples of this lane: complex z[16 n1 + t] = (s[32 n1 + 2t], s[32 n1 + 2t + 1])
    const int off = fi * hop + 1;
    const int g0 = f * hop - kHalf;
    float2 v[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
      const int j = 32 * n1 + 2 * t;
      const int gi = g0 + j;
      const float xm = s_x[off + j - 1], x0 = s_x[off + j], x1 = s_x[off + j + 1];
      const float y0 = (gi >= 0 && gi < n) ? x0 - preemph * xm : 0.0f;
      const float y1 = (gi + 1 >= 0 && gi + 1 < n) ? x1 - preemph * x0 : 0.0f;
      v[n1] = make_float2(y0 * s_win[j], y1 * s_win[j + 1]);
    }
    fft16(v);                                                        // over n1: v[k1] = A[t][k1]
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) s_tr[k1 * kBTrPitch + t] = cmul(v[k1], s_twb[k1 * 16 + t]);
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = s_tr[t * kBTrPitch + n2];
    __syncwarp();
    fft16(v);                                                        // over n2: v[k2] = Z[t + 16 k2]
    // ---- real-FFT split and power spectrum
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) {
      float2 zc;
      zc.x = __shfl_sync(0xffffffffu, v[15 - k2].x, partner, 16);
      zc.y = __shfl_sync(0xffffffffu, v[15 - k2].y, partner, 16);
      if (t == 0) zc = v[(16 - k2) & 15];
      const float2 zk = v[k2];
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
      const float2 wo = cmul(s_twx[k2 * 16 + t], o);
      const float re = e.x + wo.x, im = e.y + wo.y;
      s_pw[t + 16 * k2] = re * re + im * im;
    }
    if (t == 0) { const float d = v[0].x - v[0].y; s_pw[256] = d * d; }
    __syncwarp();
    // ---- this lane's mel filters
    {
      const int nb = s_lnb[t];
      const float* lw = s_lw + t * kBLaneTaps;
      int pos = 0;
      for (int bi = 0; bi < nb; ++bi) {
        const int e = s_lb[t * kBLaneBins + bi];
        const int m = e & 255, s0 = (e >> 8) & 1023, c = e >> 18;
        float acc = 0.f;
        for (int j = 0; j < c; ++j) acc = fmaf(lw[pos + j], s_pw[s0 + j], acc);
        pos += c;
        s_o[m] = logf(acc + guard);
      }
    }
    __syncwarp();
    if (live) {
      float* orow = mel + (static_cast<size_t>(b) * F_max + f) * n_mels;
      for (int m = t; m < n_mels; m += 16) orow[m] = s_o[m];
    }
    __syncwarp();
  }
}

static size_t logmel_b_smem_bytes(int hop) {
  const int n_stage = (kBFrames - 1) * hop + kNfft + 1;
  size_t bytes = static_cast<size_t>(((n_stage + 3) & ~3) + kNfft) * 4 + 2 * 256 * 8 + 16 * kBLaneTaps * 4 + 16 * kBLaneBins * 4 + 16 * 4;
  bytes += static_cast<size_t>(16) * 16 * kBTrPitch * 8 + static_cast<size_t>(16) * 260 * 4;
  return bytes + 16;
}

cudaError_t launch_logmel_b(const float* wav, const int32_t* len, int B, int L_max, float* mel, int32_t* mel_len,
                            const void* tables, const FeTablesB& tbb, int n_mels, int hop, int n_fft, int win,
                            float preemph, float guard, float eps, cudaStream_t stream) {
  if (n_fft != kNfft || win > kNfft || n_mels > 128 || n_mels * 4 > 16 * kBTrPitch * 8 || hop <= 0) return cudaErrorInvalidValue;
  const FeTables tb = *static_cast<const FeTables*>(tables);
  const int F_max = L_max / hop + 1;
  const size_t smem = logmel_b_smem_bytes(hop);
  if (smem > 112 * 1024) return cudaErrorInvalidValue;               // two CTAs per SM
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(logmel_b_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024);
    if (e != cudaSuccess) return e;
    attr_once.set();
  }
  const dim3 grid((F_max + kBFrames - 1) / kBFrames, B);
  logmel_b_kernel<<<grid, 32 * kFeWarps, smem, stream>>>(wav, len, L_max, mel, mel_len, tb, tbb, F_max, n_mels, hop, preemph, guard);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  mel_normalize_kernel<<<dim3(B, (n_mels + 15) / 16), 512, 0, stream>>>(mel, mel_len, F_max, n_mels, eps);
  return cudaGetLastError();
}

static size_t logmel_smem_bytes(int n_mels, int hop) {
  const int n_stage = (kFramesPerBlock - 1) * hop + kNfft + 1;
  size_t fl = ((n_stage + 3) & ~3) + kNfft + 2 * kHalf + 2 * (kHalf + 2) + static_cast<size_t>(n_mels) * kMelMaxW + 2 * n_mels + ((2 * n_mels) & 1);
  return fl * 4 + static_cast<size_t>(kFeWarps) * 2 * kHalf * 8 + 16;
}

cudaError_t launch_logmel(const float* wav, const int32_t* len, int B, int L_max, float* mel, int32_t* mel_len,
                          float* /*unused*/, const void* tables, int n_mels, int hop, int n_fft, int win,
                          float preemph, float guard, float eps, cudaStream_t stream) {
  if (n_fft != kNfft || win > kNfft || n_mels > 128 || hop <= 0) return cudaErrorInvalidValue;
  const FeTables tb = *static_cast<const FeTables*>(tables);
  const int F_max = L_max / hop + 1;
  const size_t smem = logmel_smem_bytes(n_mels, hop);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(logmel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    if (e != cudaSuccess) return e;
    attr_once.set();
  }
  if (smem > 160 * 1024) return cudaErrorInvalidValue;
  const dim3 grid((F_max + kFramesPerBlock - 1) / kFramesPerBlock, B);
  logmel_kernel<<<grid, 32 * kFeWarps, smem, stream>>>(wav, len, L_max, mel, mel_len, tb, F_max, n_mels, hop, preemph, guard);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  mel_normalize_kernel<<<dim3(B, (n_mels + 15) / 16), 512, 0, stream>>>(mel, mel_len, F_max, n_mels, eps);
  return cudaGetLastError();
}

}  // namespace rs
