// CPU replay of logmel_fused_kernel's per-frame arithmetic (reazonspeech_b200/csrc/logmel.cu), lane by lane and phase by
// phase, from the SAME tables the kernel stages, with the kernel's own register FFT (csrc/fft16.cuh) and its own split /
// index helpers (csrc/logmel_frame.cuh).  Shared memory is an array, a __syncwarp is the end of a loop over the sixteen
// lanes, a width-16 shuffle is an index into the partner lane's provided register.  Built and driven by
// tests/test_logmel_host.py; test infrastructure, not product code.
#include <cmath>
#include <cstdint>

#include "logmel_frame.cuh"

using namespace rs;

// x: the utterance (n samples), frame index f; tables as packed by reazonspeech_b200/logmel_tables.py.
// out_pw[257] = 4 |X|^2, out_mel[n_mels] = log(mel + guard), NOT normalised.
extern "C" void replay_frame(const float* x, int n, int f, int hop, float preemph, float guard, const float* window /*[512]*/,
                             const float* tw_b /*[16][16][2]*/, const float* tw_x /*[8][16][2]*/, const float* mel_w /*[n_taps][16]*/,
                             const int32_t* meta, int n_mels, float* out_pw, float* out_mel) {
  float2 v[16][16];                       // v[t][.]: lane t's registers
  float2 tr[16 * lm::kTrPitch];
  float s_o[16 * lm::kMaxSlots + 16];
  const int g0 = f * hop - lm::kHalf;
  // what the staging loop leaves in s_y: y[g] = x[g] - preemph * x[g-1] inside [0, n), 0 outside
  auto y = [&](int g) { return (g >= 0 && g < n) ? std::fmaf(-preemph, g >= 1 ? x[g - 1] : 0.0f, x[g]) : 0.0f; };
  for (int t = 0; t < 16; ++t) {
    for (int n1 = 0; n1 < 16; ++n1) {
      const int j = 32 * n1 + 2 * t;
      v[t][n1] = make_float2(y(g0 + j) * window[j], y(g0 + j + 1) * window[j + 1]);
    }
    fft16(v[t]);
    for (int k1 = 0; k1 < 16; ++k1) {
      const float wx = tw_b[2 * (k1 * 16 + t)], wy = tw_b[2 * (k1 * 16 + t) + 1];
      tr[k1 * lm::kTrPitch + t] = make_float2(v[t][k1].x * wx - v[t][k1].y * wy, v[t][k1].x * wy + v[t][k1].y * wx);
    }
  }
  for (int t = 0; t < 16; ++t) {          // after the __syncwarp
    for (int n2 = 0; n2 < 16; ++n2) v[t][n2] = tr[t * lm::kTrPitch + n2];
    fft16(v[t]);
  }
  for (int i = 0; i <= 256; ++i) out_pw[i] = NAN;          // every bin must be written exactly once
  int writes[257] = {0};
  float2 given[16][lm::kPairs];           // what each lane hands over in exchange K2
  for (int t = 0; t < 16; ++t) {
    given[t][0] = lm::provided<0>(v[t], t); given[t][1] = lm::provided<1>(v[t], t);
    given[t][2] = lm::provided<2>(v[t], t); given[t][3] = lm::provided<3>(v[t], t);
    given[t][4] = lm::provided<4>(v[t], t); given[t][5] = lm::provided<5>(v[t], t);
    given[t][6] = lm::provided<6>(v[t], t); given[t][7] = lm::provided<7>(v[t], t);
  }
  for (int t = 0; t < 16; ++t) {
    for (int k2 = 0; k2 < lm::kPairs; ++k2) {
      const float2 zc = given[lm::partner_lane(t)][k2];    // the pair of width-16 shuffles
      float pp, pm;
      lm::split_pair(v[t][k2], zc, make_float2(tw_x[2 * (k2 * 16 + t)], tw_x[2 * (k2 * 16 + t) + 1]), pp, pm);
      out_pw[lm::bin_plus(t, k2)] = pp; writes[lm::bin_plus(t, k2)]++;
      out_pw[lm::bin_minus(t, k2)] = pm; writes[lm::bin_minus(t, k2)]++;
    }
    if (t == 0) { out_pw[128] = 4.0f * (v[0][8].x * v[0][8].x + v[0][8].y * v[0][8].y); writes[128]++; }
  }
  for (int i = 0; i <= 256; ++i) if (writes[i] != 1) out_pw[i] = NAN;
  for (int i = 0; i < 16 * lm::kMaxSlots + 16; ++i) s_o[i] = NAN;
  const int n_slots = meta[0];
  for (int t = 0; t < 16; ++t) {
    const float* wp = mel_w + t;
    for (int s = 0; s < n_slots; ++s) {
      const int c = meta[8 + s];
      const float* pp = out_pw + meta[16 + s * 16 + t];
      float acc = 0.f;
      for (int j = 0; j < c; ++j) acc = std::fmaf(wp[j * 16], pp[j], acc);
      wp += c * 16;
      s_o[meta[16 + 128 + s * 16 + t]] = acc;
    }
  }
  for (int m = 0; m < n_mels; ++m) out_mel[m] = std::log(s_o[m] + guard);
}
