// Per-frame arithmetic of the fused log-mel kernel (logmel.cu) that does not depend on CUDA: index maps of the 16 x 16
// FFT decomposition and the paired real-FFT split.  Host + device code: tests/host/logmel_replay.cpp compiles this header
// (and fft16.cuh) with g++ and replays a frame lane by lane against numpy, so the index algebra is checked without a GPU.
//
// One frame = 512 real samples = 256 complex z[n] = (s[2n], s[2n+1]).  Sixteen lanes per frame:
//   lane t holds z[16 n1 + t], n1 = 0..15        -> fft16 over n1            -> A[t][k1]
//   times W256^(t k1), transposed through shared memory: lane k1 holds the sixteen t -> fft16 over t -> Z[k1 + 16 k2]
// so after the second pass lane t register k2 is Z[t + 16 k2].
// Real-FFT split, two bins at a time: with E = (Z[k] + conj Z[256-k]) / 2, O = (Z[k] - conj Z[256-k]) / 2i,
//   X[k] = E + W512^k O      and      X[256-k] = conj(E - W512^k O),
// so one lane forms E, W O once and takes |.|^2 of the sum and of the difference.  Lane t owns the pairs k = t + 16 k2,
// k2 = 0..7, i.e. its own bins t + 16 k2 and the bins 256 - t - 16 k2 of lane (16 - t) & 15, which in turn covers this
// lane's upper eight bins; Z[256 - k] comes from that lane by one pair of width-16 shuffles.  Lanes 0 and 8 are their own
// partners; lane 0 additionally owns the self-paired bin 128 and, through k = 0 (Z[256] = Z[0]), bins 0 and 256.
// The factors 1/2 are dropped: every power is 4 |X|^2 and the mel weights are packed with a factor 1/4.
#pragma once
#include "fft16.cuh"

namespace rs {
namespace lm {

constexpr int kNfft = 512;
constexpr int kHalf = 256;
constexpr int kLanes = 16;       // lanes per frame
constexpr int kTrPitch = 17;     // float2 pitch of the transpose buffer (conflict-free both ways)
constexpr int kPairs = 8;        // bin pairs per lane
constexpr int kMaxSlots = 8;     // mel filters per lane (n_mels <= 128)

RS_FFT_HD int partner_lane(int t) { return (16 - t) & 15; }
// register a lane hands to its partner for pair K2 (the partner's Z[256 - k]); lane 0 is its own partner with a shifted map
template <int K2>
RS_FFT_HD float2 provided(const float2 (&v)[16], int lane) { return lane == 0 ? v[(16 - K2) & 15] : v[15 - K2]; }
RS_FFT_HD int bin_plus(int t, int k2) { return t + 16 * k2; }
RS_FFT_HD int bin_minus(int t, int k2) { return kHalf - t - 16 * k2; }

// zk = Z[k], zc = Z[256 - k], w = W512^k = (cos, -sin)(2 pi k / 512):  p_plus = 4 |X[k]|^2,  p_minus = 4 |X[256 - k]|^2
RS_FFT_HD void split_pair(float2 zk, float2 zc, float2 w, float& p_plus, float& p_minus) {
  const float ex = zk.x + zc.x, ey = zk.y - zc.y;            // 2 E
  const float ox = zk.y + zc.y, oy = zc.x - zk.x;            // 2 O
  const float wx = w.x * ox - w.y * oy, wy = w.x * oy + w.y * ox;
  const float ax = ex + wx, ay = ey + wy, bx = ex - wx, by = ey - wy;
  p_plus = ax * ax + ay * ay;
  p_minus = bx * bx + by * by;
}

}  // namespace lm
}  // namespace rs
