// CPU replay of logmel_b_kernel's per-frame arithmetic (reazonspeech_b200/csrc/frontend.cu), lane by lane and phase by
// phase, from the SAME tables the kernel stages and with the kernel's own register FFT (csrc/fft16.cuh).  Shared memory is
// an array, a __syncwarp is the end of a loop over the sixteen lanes, a width-16 shuffle is an index into the other
// lane's registers.  Built and driven by tests/test_logmel_b_host.py; test infrastructure, not product code.
#include <cmath>
#include <cstdint>

#include "fft16.cuh"

namespace {
inline float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
constexpr int kBLaneBins = 8, kBLaneTaps = 47, kBTrPitch = 17;
}  // namespace

// x: the utterance (n samples), frame index f; tables as packed by engine.py::frontend_tables / frontend_tables_b.
// out_pw[257], out_mel[n_mels] (log(mel + guard), NOT normalised).
extern "C" void replay_frame(const float* x, int n, int f, int hop, float preemph, float guard, const float* window /*[512]*/,
                             const float* tw_b /*[16][16][2]*/, const float* tw_x /*[16][16][2]*/, const float* lane_w,
                             const int32_t* lane_bins, const int32_t* lane_nb, int n_mels, float* out_pw, float* out_mel) {
  float2 v[16][16];                       // v[t][.]: lane t's registers
  float2 tr[16 * kBTrPitch];
  const int g0 = f * hop - 256;
  auto sample = [&](int gi) { return (gi >= 0 && gi < n) ? x[gi] : 0.0f; };   // what the staging loop puts into s_x
  for (int t = 0; t < 16; ++t) {
    for (int n1 = 0; n1 < 16; ++n1) {
      const int j = 32 * n1 + 2 * t, gi = g0 + j;
      const float xm = sample(gi - 1), x0 = sample(gi), x1 = sample(gi + 1);
      const float y0 = (gi >= 0 && gi < n) ? x0 - preemph * xm : 0.0f;
      const float y1 = (gi + 1 >= 0 && gi + 1 < n) ? x1 - preemph * x0 : 0.0f;
      v[t][n1] = make_float2(y0 * window[j], y1 * window[j + 1]);
    }
    rs::fft16(v[t]);
    for (int k1 = 0; k1 < 16; ++k1)
      tr[k1 * kBTrPitch + t] = cmul(v[t][k1], make_float2(tw_b[2 * (k1 * 16 + t)], tw_b[2 * (k1 * 16 + t) + 1]));
  }
  for (int t = 0; t < 16; ++t) {          // after the __syncwarp
    for (int n2 = 0; n2 < 16; ++n2) v[t][n2] = tr[t * kBTrPitch + n2];
    rs::fft16(v[t]);
  }
  for (int t = 0; t < 16; ++t) {
    const int partner = (16 - t) & 15;
    for (int k2 = 0; k2 < 16; ++k2) {
      float2 zc = v[partner][15 - k2];    // the pair of width-16 shuffles
      if (t == 0) zc = v[t][(16 - k2) & 15];
      const float2 zk = v[t][k2];
      const float2 e = make_float2(0.5f * (zk.x + zc.x), 0.5f * (zk.y - zc.y));
      const float2 o = make_float2(0.5f * (zk.y + zc.y), -0.5f * (zk.x - zc.x));
      const float2 wo = cmul(make_float2(tw_x[2 * (k2 * 16 + t)], tw_x[2 * (k2 * 16 + t) + 1]), o);
      const float re = e.x + wo.x, im = e.y + wo.y;
      out_pw[t + 16 * k2] = re * re + im * im;
    }
    if (t == 0) { const float d = v[0][0].x - v[0][0].y; out_pw[256] = d * d; }
  }
  for (int m = 0; m < n_mels; ++m) out_mel[m] = NAN;     // every filter must be written by exactly one lane
  for (int t = 0; t < 16; ++t) {
    const float* lw = lane_w + t * kBLaneTaps;
    int pos = 0;
    for (int bi = 0; bi < lane_nb[t]; ++bi) {
      const int e = lane_bins[t * kBLaneBins + bi];
      const int m = e & 255, s0 = (e >> 8) & 1023, c = e >> 18;
      float acc = 0.f;
      for (int j = 0; j < c; ++j) acc = std::fmaf(lw[pos + j], out_pw[s0 + j], acc);
      pos += c;
      out_mel[m] = std::log(acc + guard);
    }
  }
}
