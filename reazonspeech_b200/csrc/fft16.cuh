// 16-point complex FFT held entirely in one thread's registers, the building block of the fused log-mel kernel
// (logmel.cu: a 256-point complex FFT as 16 x 16, sixteen lanes per frame, two frames per warp).
// Host + device code: tests/test_logmel_host.py compiles this header with g++ and runs the kernel's whole per-frame
// arithmetic (both FFT16 passes, inter-pass twiddles, the paired real-FFT split with its partner-lane mapping) on the CPU
// against numpy, so the index algebra is checked without a GPU.
#pragma once
#if defined(__CUDACC__)
#include <cuda_runtime.h>
#define RS_FFT_HD __host__ __device__ __forceinline__
#else
#include <vector_types.h>
#define RS_FFT_HD inline
static inline float2 make_float2(float x, float y) { float2 r; r.x = x; r.y = y; return r; }
#endif

namespace rs {

// y_k = sum_n x_n (-i)^(n k), k = 0..3, in place
RS_FFT_HD void dft4(float2& x0, float2& x1, float2& x2, float2& x3) {
  const float2 s02 = make_float2(x0.x + x2.x, x0.y + x2.y), d02 = make_float2(x0.x - x2.x, x0.y - x2.y);
  const float2 s13 = make_float2(x1.x + x3.x, x1.y + x3.y), d13 = make_float2(x1.x - x3.x, x1.y - x3.y);
  x0 = make_float2(s02.x + s13.x, s02.y + s13.y);
  x2 = make_float2(s02.x - s13.x, s02.y - s13.y);
  x1 = make_float2(d02.x + d13.y, d02.y - d13.x);     // d02 - i d13
  x3 = make_float2(d02.x - d13.y, d02.y + d13.x);     // d02 + i d13
}

RS_FFT_HD float2 cmulf(float2 a, float cr, float ci) { return make_float2(a.x * cr - a.y * ci, a.x * ci + a.y * cr); }

// In place, natural order in and out:  v[k] <- sum_n v[n] exp(-2 pi i n k / 16).
// n = 4 n1 + n2, k = k1 + 4 k2:  W16^(nk) = W4^(n1 k1) W16^(n2 k1) W4^(n2 k2).
RS_FFT_HD void fft16(float2 (&v)[16]) {
  constexpr float c1 = 0.92387953251128674f, s1 = 0.38268343236508977f, r2 = 0.70710678118654752f;
  // a[n2][k1]: radix-4 over n1 for each n2
#pragma unroll
  for (int n2 = 0; n2 < 4; ++n2) dft4(v[n2], v[n2 + 4], v[n2 + 8], v[n2 + 12]);      // v[n2 + 4 k1] = a[n2][k1]
  // twiddles W16^(n2 k1) = (cos, -sin)(2 pi n2 k1 / 16); n2 k1 in {1, 2, 3, 2, 4, 6, 3, 6, 9}
  v[1 + 4] = cmulf(v[1 + 4], c1, -s1);        // n2 = 1, k1 = 1: m = 1
  v[1 + 8] = cmulf(v[1 + 8], r2, -r2);        //          k1 = 2: m = 2
  v[1 + 12] = cmulf(v[1 + 12], s1, -c1);      //          k1 = 3: m = 3
  v[2 + 4] = cmulf(v[2 + 4], r2, -r2);        // n2 = 2, k1 = 1: m = 2
  v[2 + 8] = make_float2(v[2 + 8].y, -v[2 + 8].x);                                    // m = 4: times -i
  v[2 + 12] = cmulf(v[2 + 12], -r2, -r2);     //          k1 = 3: m = 6
  v[3 + 4] = cmulf(v[3 + 4], s1, -c1);        // n2 = 3, k1 = 1: m = 3
  v[3 + 8] = cmulf(v[3 + 8], -r2, -r2);       //          k1 = 2: m = 6
  v[3 + 12] = cmulf(v[3 + 12], -c1, s1);      //          k1 = 3: m = 9
  // radix-4 over n2 for each k1: inputs a[0..3][k1] = v[4 k1 + 0..3]; outputs Y[k1 + 4 k2]
  float2 y[16];
#pragma unroll
  for (int k1 = 0; k1 < 4; ++k1) {
    float2 x0 = v[4 * k1], x1 = v[4 * k1 + 1], x2 = v[4 * k1 + 2], x3 = v[4 * k1 + 3];
    dft4(x0, x1, x2, x3);
    y[k1] = x0; y[k1 + 4] = x1; y[k1 + 8] = x2; y[k1 + 12] = x3;
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = y[k];
}

}  // namespace rs
