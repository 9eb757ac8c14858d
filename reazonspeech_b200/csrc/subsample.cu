// ConvSubsampling (dw_striding, x8) direct-convolution parts (N2).  NeMo reference module:
// parts/submodules/subsampling.py ConvSubsampling (reached via model.transcribe,
// pkg/nemo-asr/src/transcribe.py:48-53).  The 1x1 convs and the output Linear run on the
// tcgen05 GEMM over channels-last activations; this file holds the HBM-bound 3x3 convs.
//
// Kernel A fuses conv.0 (1->C, 3x3, s2, p1) + ReLU + conv.2 (depthwise 3x3, s2, p1): the
// [B, T1, 40, C] intermediate (1 GB at 32 x 30 s in bf16) never reaches HBM.  One thread per
// channel; the mel patch of the tile sits in shared memory and is read as a warp broadcast.
// Every stage treats frames at or beyond the utterance's length at that stage as zero, which is
// what batch=1 execution sees as conv zero-padding (padding invariance, SURVEY.md finding 4).
#include "common.cuh"
#include "kernels.h"

namespace rs {

__host__ __device__ __forceinline__ int conv_len(int n) { return (n - 1) / 2 + 1; }

constexpr int kSubTT = 4;     // t2 rows per CTA

__global__ void __launch_bounds__(256)
sub_conv0_dw1_kernel(const float* __restrict__ mel, const int32_t* __restrict__ mel_len, int F_max, int n_mels, int C,
                     const float* __restrict__ w0, const float* __restrict__ b0, const float* __restrict__ wd,
                     const float* __restrict__ bd, __nv_bfloat16* __restrict__ out, int T2, int F1, int F2) {
  extern __shared__ float s_mel[];                       // [(4*TT+3)][n_mels + 2], column 0 == mel bin -1
  const int b = blockIdx.y;
  const int t2_0 = blockIdx.x * kSubTT;
  const int len0 = mel_len[b];
  const int len1 = conv_len(len0);
  const int len2 = conv_len(len1);
  const int rows = 4 * kSubTT + 3;
  const int ld = n_mels + 2;
  const int t0_base = 4 * t2_0 - 3;                      // first mel row needed: 2*(2*t2_0-1)-1
  for (int i = threadIdx.x; i < rows * ld; i += blockDim.x) {
    const int r = i / ld, cidx = i % ld - 1;
    const int t0 = t0_base + r;
    float v = 0.f;
    if (t0 >= 0 && t0 < len0 && t0 < F_max && cidx >= 0 && cidx < n_mels)
      v = mel[(static_cast<size_t>(b) * F_max + t0) * n_mels + cidx];
    s_mel[i] = v;
  }
  __syncthreads();

  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float k0[9], kd[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) { k0[j] = __ldg(w0 + c * 9 + j); kd[j] = __ldg(wd + c * 9 + j); }
    const float bias0 = __ldg(b0 + c), biasd = __ldg(bd + c);
    for (int tt = 0; tt < kSubTT; ++tt) {
      const int t2 = t2_0 + tt;
      if (t2 >= T2) break;
      __nv_bfloat16* orow = out + ((static_cast<size_t>(b) * T2 + t2) * F2) * C + c;
      if (t2 >= len2) {                                   // padded frame: defined value, never read as valid
        for (int f2 = 0; f2 < F2; ++f2) orow[static_cast<size_t>(f2) * C] = __float2bfloat16_rn(0.f);
        continue;
      }
      // conv.0 outputs at (t1 = 2*t2-1+dt, f1) for dt = 0..2; recomputed per f2 with a sliding 3x3 window
      float win[3][3];                                    // [dt][df] conv0+ReLU values, df <-> f1 = 2*f2-1+df
      auto conv0_at = [&](int dt, int f1) -> float {
        const int t1 = 2 * t2 - 1 + dt;
        if (t1 < 0 || t1 >= len1 || f1 < 0 || f1 >= F1) return 0.f;      // dw zero padding / masked frame
        // mel rows 2*t1-1 .. 2*t1+1  ->  smem rows (2*t1-1) - t0_base; mel cols 2*f1-1 .. 2*f1+1 -> +1 offset
        const float* p = s_mel + (2 * t1 - 1 - t0_base) * ld + 2 * f1;
        float a = bias0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
          for (int j = 0; j < 3; ++j) a = fmaf(p[i * ld + j], k0[i * 3 + j], a);
        return fmaxf(a, 0.f);
      };
#pragma unroll
      for (int dt = 0; dt < 3; ++dt) { win[dt][1] = 0.f; win[dt][2] = conv0_at(dt, -1); }
      for (int f2 = 0; f2 < F2; ++f2) {
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
          win[dt][0] = win[dt][2];
          win[dt][1] = conv0_at(dt, 2 * f2);
          win[dt][2] = conv0_at(dt, 2 * f2 + 1);
        }
        float a = biasd;
#pragma unroll
        for (int dt = 0; dt < 3; ++dt)
#pragma unroll
          for (int df = 0; df < 3; ++df) a = fmaf(win[dt][df], kd[dt * 3 + df], a);
        orow[static_cast<size_t>(f2) * C] = __float2bfloat16_rn(a);
      }
    }
  }
}

cudaError_t launch_sub_conv0_dw1(const SubsampleArgs& a, cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(4 * kSubTT + 3) * (a.n_mels + 2) * sizeof(float);
  const dim3 grid((a.T2 + kSubTT - 1) / kSubTT, a.B);
  sub_conv0_dw1_kernel<<<grid, 256, smem, stream>>>(a.mel, a.mel_len, a.F_max, a.n_mels, a.C, a.w0, a.b0, a.wd1, a.bd1,
                                                   static_cast<__nv_bfloat16*>(a.out1), a.T2, a.F1, a.F2);
  return cudaGetLastError();
}

// Depthwise 3x3 s2 p1 on channels-last bf16 [B, Tin, Fin, C] -> [B, Tout, Fout, C].
// The valid input length of utterance b is conv_len applied `len_shift` times to mel_len[b].
__global__ void __launch_bounds__(128)
sub_dw_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, const float* __restrict__ w,
              const float* __restrict__ bias, const int32_t* __restrict__ mel_len, int len_shift, int Tin, int Fin,
              int Tout, int Fout, int C) {
  const int b = blockIdx.z, t = blockIdx.y;
  int lin = mel_len[b];
  for (int i = 0; i < len_shift; ++i) lin = conv_len(lin);
  const int lout = conv_len(lin);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;        // over Fout * C/2
  const int c2 = C / 2;
  if (idx >= Fout * c2) return;
  const int f = idx / c2, c = (idx % c2) * 2;
  uint32_t* o = reinterpret_cast<uint32_t*>(out + ((static_cast<size_t>(b) * Tout + t) * Fout + f) * C + c);
  if (t >= lout) { *o = 0u; return; }
  float ax = __ldg(bias + c), ay = __ldg(bias + c + 1);
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int ti = 2 * t - 1 + i;
    if (ti < 0 || ti >= lin || ti >= Tin) continue;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int fi = 2 * f - 1 + j;
      if (fi < 0 || fi >= Fin) continue;
      const float2 v = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(in + ((static_cast<size_t>(b) * Tin + ti) * Fin + fi) * C + c)));
      ax = fmaf(v.x, __ldg(w + c * 9 + i * 3 + j), ax);
      ay = fmaf(v.y, __ldg(w + (c + 1) * 9 + i * 3 + j), ay);
    }
  }
  *o = pack_bf16x2(ax, ay);
}

cudaError_t launch_sub_dw(const void* in, void* out, const float* w, const float* b, const int32_t* mel_len, int len_shift,
                          int B, int Tin, int Fin, int Tout, int Fout, int C, cudaStream_t stream) {
  const int work = Fout * (C / 2);
  const dim3 grid((work + 127) / 128, Tout, B);
  sub_dw_kernel<<<grid, 128, 0, stream>>>(static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(out), w, b,
                                          mel_len, len_shift, Tin, Fin, Tout, Fout, C);
  return cudaGetLastError();
}

}  // namespace rs
