// Device-side norm_audio: resample to 16 kHz and average the channels (pkg/nemo-asr/src/audio.py:54-68: librosa.resample, then
// librosa.to_mono), with the transcribe() padding (audio.py:70-83, 0.5 s of zeros on both sides) written in the same pass.
// The host path (nemo/asr/audio.py) uses scipy.signal.resample_poly; this kernel evaluates the same polyphase FIR
//     out[m] = sum_n x[n] h[(m + n_pre_remove) * down - n * up]
// with the SAME filter (firwin(20 * max(up, down) + 1, 1 / max(up, down), kaiser 5.0) * up, designed on the host by
// engine.py::resample_taps and handed over in polyphase order taps[phase][j] = h_padded[phase + j * up]), so device and host
// agree to fp32 rounding.  A 30 s 48 kHz clip costs the host ~10 ms in scipy -- more than the whole engine spends on it.
// Averaging the channels first and resampling once is the same linear map as the reference's resample-then-average.
#include "common.cuh"
#include "kernels.h"

namespace rs {

namespace {

template <bool kI16>
__global__ void __launch_bounds__(256)
resample_mono_kernel(const void* __restrict__ in_raw, const int32_t* __restrict__ len_in, int C, int L_in_max,
                     const float* __restrict__ taps, int taps_per_phase, int up, int down, int n_pre_remove, int pad,
                     float* __restrict__ out, int L_out_row, int32_t* __restrict__ len_out) {
  const int b = blockIdx.y;
  const int n_in = len_in[b];
  const long long scaled = static_cast<long long>(n_in) * up;
  const int n_out = static_cast<int>((scaled + down - 1) / down);            // scipy: n_in * up // down + bool(n_in * up % down)
  if (blockIdx.x == 0 && threadIdx.x == 0) len_out[b] = n_out + 2 * pad;
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= L_out_row) return;
  float acc = 0.f;
  const int mo = m - pad;
  if (mo >= 0 && mo < n_out) {
    const long long t = static_cast<long long>(mo + n_pre_remove) * down;
    const int n_hi = static_cast<int>(t / up), phase = static_cast<int>(t - static_cast<long long>(n_hi) * up);
    const float* h = taps + static_cast<size_t>(phase) * taps_per_phase;
    const float inv_c = 1.0f / static_cast<float>(C);
    const size_t base = static_cast<size_t>(b) * C * L_in_max;
    for (int j = 0; j < taps_per_phase; ++j) {
      const int n = n_hi - j;
      if (n < 0) break;
      if (n >= n_in) continue;
      float x = 0.f;
      for (int c = 0; c < C; ++c) {
        const size_t at = base + static_cast<size_t>(c) * L_in_max + n;
        if constexpr (kI16) x += static_cast<float>(__ldg(static_cast<const int16_t*>(in_raw) + at)) * (1.0f / 32768.0f);
        else x += __ldg(static_cast<const float*>(in_raw) + at);
      }
      acc = fmaf(__ldg(h + j), C > 1 ? x * inv_c : x, acc);
    }
  }
  out[static_cast<size_t>(b) * L_out_row + m] = acc;
}

}  // namespace

cudaError_t launch_resample_mono(const ResampleArgs& a, cudaStream_t stream) {
  if (a.B <= 0 || a.C <= 0 || a.L_in_max <= 0 || a.up <= 0 || a.down <= 0 || a.taps_per_phase <= 0 || a.L_out_row <= 0 || a.pad < 0)
    return cudaErrorInvalidValue;
  const dim3 grid((a.L_out_row + 255) / 256, a.B);
  if (a.in_i16)
    resample_mono_kernel<true><<<grid, 256, 0, stream>>>(a.in, a.len_in, a.C, a.L_in_max, a.taps, a.taps_per_phase, a.up, a.down,
                                                         a.n_pre_remove, a.pad, a.out, a.L_out_row, a.len_out);
  else
    resample_mono_kernel<false><<<grid, 256, 0, stream>>>(a.in, a.len_in, a.C, a.L_in_max, a.taps, a.taps_per_phase, a.up, a.down,
                                                          a.n_pre_remove, a.pad, a.out, a.L_out_row, a.len_out);
  return cudaGetLastError();
}

}  // namespace rs
