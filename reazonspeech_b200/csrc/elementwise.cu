// HBM-bound row kernels of the Conformer layer: LayerNorm (N7 / pre-norms) and the middle of
// ConformerConvolution (mask -> depthwise k=9 -> BatchNorm(eval, folded) -> Swish) (N6).
// Warp-shuffle reductions, 128-bit / 32-bit coalesced accesses; no tensor cores on purpose.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace rs {

// ------------------------------------------------------------------------------ LayerNorm
// One warp per row of d = 128*NV fp32 values (two-pass mean / variance in registers).  Optional second LayerNorm chained
// on the result (norm_out of layer i feeding norm_feed_forward1 of layer i+1) so the residual stream is read once.
//
// Rows reach the warp through its own ring of kLnStages row buffers in shared memory, filled by 1-D bulk copies
// (cp.async.bulk, one 4 KB request per row, completion on an mbarrier): the warp always has kLnStages rows in flight
// without holding them in registers.  The version this replaces prefetched ONE row ahead with eight LDG.128 per lane and
// was bound by that round trip (~2.7 us per row and warp under load: 3.1 TB/s, 0.47 of the HBM peak, ncu r01_v4); more
// resident warps did not help, the register file was full.
constexpr int kLnWarps = 8;
constexpr int kLnStages = 3;

__device__ __forceinline__ void bulk_load_row(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

template <int NV>
__global__ void __launch_bounds__(32 * kLnWarps)
layernorm_kernel(const float* __restrict__ x, const float* __restrict__ g1, const float* __restrict__ b1,
                 float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
                 const float* __restrict__ g2, const float* __restrict__ b2, int rows, float eps) {
  constexpr int D = 128 * NV;
  constexpr uint32_t kRowBytes = D * 4;
  extern __shared__ __align__(128) uint8_t ln_smem[];
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const int wstride = gridDim.x * kLnWarps;
  const int row0 = blockIdx.x * kLnWarps + warp;
  const uint32_t ring = smem_u32(ln_smem) + static_cast<uint32_t>(warp) * kLnStages * kRowBytes;
  const uint32_t bars = smem_u32(ln_smem) + kLnWarps * kLnStages * kRowBytes + static_cast<uint32_t>(warp) * kLnStages * 8;
  if (lane == 0) {
    for (int s = 0; s < kLnStages; ++s) mbar_init(bars + 8 * s, 1);
    fence_barrier_init();
  }
  __syncwarp();
  if (row0 >= rows) return;
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < kLnStages; ++s) {
      const int r = row0 + s * wstride;
      if (r < rows) {
        mbar_arrive_expect_tx(bars + 8 * s, kRowBytes);
        bulk_load_row(ring + s * kRowBytes, x + static_cast<size_t>(r) * D, kRowBytes, bars + 8 * s);
      }
    }
  }
  float4 v[NV];
  auto normalize = [&](const float* __restrict__ g, const float* __restrict__ b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
    const float mean = warp_sum(s) * (1.0f / D);
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, dd = v[i].w - mean;
      ss += (a * a + bb * bb) + (c * c + dd * dd);
    }
    const float rstd = rsqrtf(warp_sum(ss) * (1.0f / D) + eps);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 gg = __ldg(reinterpret_cast<const float4*>(g) + lane + 32 * i);
      const float4 bb = __ldg(reinterpret_cast<const float4*>(b) + lane + 32 * i);
      v[i].x = (v[i].x - mean) * rstd * gg.x + bb.x;
      v[i].y = (v[i].y - mean) * rstd * gg.y + bb.y;
      v[i].z = (v[i].z - mean) * rstd * gg.z + bb.z;
      v[i].w = (v[i].w - mean) * rstd * gg.w + bb.w;
    }
  };

  int stage = 0; uint32_t phase = 0;
  for (int row = row0; row < rows; row += wstride) {
    mbar_wait(bars + 8 * stage, phase);
    {
      const uint32_t src = ring + stage * kRowBytes + lane * 16;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(src + 512 * i));
    }
    __syncwarp();                                              // every lane has its part of the row: the buffer may be refilled
    const int nrow = row + kLnStages * wstride;
    if (lane == 0 && nrow < rows) {
      mbar_arrive_expect_tx(bars + 8 * stage, kRowBytes);
      bulk_load_row(ring + stage * kRowBytes, x + static_cast<size_t>(nrow) * D, kRowBytes, bars + 8 * stage);
    }
    if (++stage == kLnStages) { stage = 0; phase ^= 1u; }
    normalize(g1, b1);
    if (out_f32 != nullptr) {
      float4* o = reinterpret_cast<float4*>(out_f32 + static_cast<size_t>(row) * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) o[lane + 32 * i] = v[i];
    }
    if (g2 != nullptr) normalize(g2, b2);
    if (out_bf16 != nullptr) {
      uint2* o = reinterpret_cast<uint2*>(out_bf16 + static_cast<size_t>(row) * D);
#pragma unroll
      for (int i = 0; i < NV; ++i) o[lane + 32 * i] = make_uint2(pack_bf16x2(v[i].x, v[i].y), pack_bf16x2(v[i].z, v[i].w));
    }
  }
}

template <int NV>
static cudaError_t launch_ln(const float* x, const float* g1, const float* b1, float* out_f32, __nv_bfloat16* out_bf16,
                             const float* g2, const float* b2, int rows, float eps, int num_sms, cudaStream_t stream) {
  constexpr int kSmem = kLnWarps * kLnStages * (128 * NV * 4 + 8);
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(layernorm_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return e;
    attr_once.set();
  }
  const int need = (rows + kLnWarps - 1) / kLnWarps;
  const int cap = num_sms * 2;                                 // two resident CTAs per SM at d = 1024 (2 x 96 KB of row buffers)
  layernorm_kernel<NV><<<need < cap ? need : cap, 32 * kLnWarps, kSmem, stream>>>(x, g1, b1, out_f32, out_bf16, g2, b2, rows, eps);
  return cudaGetLastError();
}

cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float* out_f32, void* out_bf16,
                             const float* gamma2, const float* beta2, int rows, int d, float eps, cudaStream_t stream) {
  if (rows <= 0) return cudaSuccess;
  if (reinterpret_cast<uintptr_t>(x) & 15u) return cudaErrorInvalidValue;      // bulk copies want 16-byte aligned rows
  int num_sms = 0, dev = 0;                                    // per call: engines on different devices share this code
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  auto* ob = static_cast<__nv_bfloat16*>(out_bf16);
  switch (d) {
    case 256: return launch_ln<2>(x, gamma, beta, out_f32, ob, gamma2, beta2, rows, eps, num_sms, stream);
    case 512: return launch_ln<4>(x, gamma, beta, out_f32, ob, gamma2, beta2, rows, eps, num_sms, stream);
    case 1024: return launch_ln<8>(x, gamma, beta, out_f32, ob, gamma2, beta2, rows, eps, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

// ------------------------------------------------------------------------------ conv module middle
// u: GLU output, bf16 [B, T_max, d].  Frames t >= len[b] read as zero (masked_fill before the
// depthwise conv); t < 0 is the conv's own zero padding.  BatchNorm(eval) is folded at pack time:
// w'[j][c] = w[c][j] * gamma/sqrt(var+eps),  shift[c] = (bias - mean) * gamma/sqrt(var+eps) + beta.
//
// Persistent CTAs of d/2 threads (two channels each) walk work items (utterance, TT output frames).  The TT + KW - 1 input
// rows of an item are ONE contiguous block of [B, T_max, d], fetched with a single bulk copy into one of two shared-memory
// buffers while the previous item is computed: the thread slides a KW-row register window down its two channels (one LDS.32
// per output frame) and stores bf16x2, a warp covering 128 contiguous bytes.  The version this replaces had every thread
// request its own TT + KW - 1 rows from L2 (2x read amplification, 1.6 TB/s: 0.31 of the HBM peak).
constexpr int kDwTT = 16;

template <int KW, int TT>
__global__ void __launch_bounds__(512)
conv_dw_kernel(const __nv_bfloat16* __restrict__ u, __nv_bfloat16* __restrict__ out, const float* __restrict__ w,
               const float* __restrict__ shift, const int32_t* __restrict__ len, int B, int T_max, int d) {
  constexpr int PAD = (KW - 1) / 2;
  constexpr int ROWS = TT + KW - 1;
  extern __shared__ __align__(128) uint8_t dw_smem[];
  const uint32_t row_bytes = static_cast<uint32_t>(d) * 2u;
  const uint32_t buf_bytes = ROWS * row_bytes;
  const uint32_t bars = smem_u32(dw_smem) + 2 * buf_bytes;
  const int chunks = (T_max + TT - 1) / TT;
  const int items = B * chunks;
  const int c = threadIdx.x * 2;
  if (threadIdx.x == 0) { mbar_init(bars, 1); mbar_init(bars + 8, 1); fence_barrier_init(); }
  __syncthreads();
  auto fetch = [&](int item, int buf) {                         // one thread: the item's rows that exist, at their place in the buffer
    const int b = item / chunks, t0 = (item % chunks) * TT;
    const int lo = max(t0 - PAD, 0), hi = min(t0 + TT + PAD, T_max);
    const uint32_t bytes = static_cast<uint32_t>(hi - lo) * row_bytes;
    mbar_arrive_expect_tx(bars + 8 * buf, bytes);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dw_smem) + buf * buf_bytes + static_cast<uint32_t>(lo - (t0 - PAD)) * row_bytes),
                   "l"(u + (static_cast<size_t>(b) * T_max + lo) * d), "r"(bytes), "r"(bars + 8 * buf) : "memory");
  };
  if (threadIdx.x == 0) {
    if (static_cast<int>(blockIdx.x) < items) fetch(blockIdx.x, 0);
    if (static_cast<int>(blockIdx.x + gridDim.x) < items) fetch(blockIdx.x + gridDim.x, 1);
  }
  float2 wt[KW];
#pragma unroll
  for (int j = 0; j < KW; ++j) wt[j] = __ldg(reinterpret_cast<const float2*>(w + static_cast<size_t>(j) * d + c));
  const float2 sh = __ldg(reinterpret_cast<const float2*>(shift + c));
  int k = 0;
  for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
    const int buf = k & 1;
    const int b = item / chunks, t0 = (item % chunks) * TT;
    const int n = len[b];
    mbar_wait(bars + 8 * buf, (k >> 1) & 1u);
    const uint32_t base = smem_u32(dw_smem) + buf * buf_bytes + static_cast<uint32_t>(c) * 2u;
    auto row = [&](int r) -> float2 {                           // buffer row r holds frame t0 - PAD + r; masked / padded frames read as 0
      const int t = t0 - PAD + r;
      uint32_t v = 0;
      if (t >= 0 && t < n) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + static_cast<uint32_t>(r) * row_bytes));
      return make_float2(bf16_lo(v), bf16_hi(v));
    };
    float2 win[KW];
#pragma unroll
    for (int j = 0; j < KW - 1; ++j) win[j + 1] = row(j);
    __nv_bfloat16* o = out + (static_cast<size_t>(b) * T_max + t0) * d + c;
#pragma unroll
    for (int q = 0; q < TT; ++q) {
#pragma unroll
      for (int j = 0; j < KW - 1; ++j) win[j] = win[j + 1];
      win[KW - 1] = row(q + KW - 1);
      float2 a = sh;
#pragma unroll
      for (int j = 0; j < KW; ++j) { a.x = fmaf(win[j].x, wt[j].x, a.x); a.y = fmaf(win[j].y, wt[j].y, a.y); }
      if (t0 + q < T_max) *reinterpret_cast<uint32_t*>(o + static_cast<size_t>(q) * d) = pack_bf16x2(swishf_fast(a.x), swishf_fast(a.y));
    }
    __syncthreads();                                            // everyone is done with this buffer
    const int nxt = item + 2 * gridDim.x;
    if (threadIdx.x == 0 && nxt < items) fetch(nxt, buf);
  }
}

cudaError_t launch_conv_dw(const void* u, void* out, const float* w, const float* shift, const int32_t* enc_len,
                           int B, int T_max, int d, int k, cudaStream_t stream) {
  if (k != 9 || (d & 63) || d > 1024 || (reinterpret_cast<uintptr_t>(u) & 15u)) return cudaErrorInvalidValue;
  if (B <= 0 || T_max <= 0) return cudaSuccess;
  constexpr int TT = kDwTT;
  const int smem = 2 * (TT + 8) * d * 2 + 16;
  int num_sms = 0, dev = 0;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0) num_sms = 148;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(conv_dw_kernel<9, TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * (TT + 8) * 1024 * 2 + 16);
    if (e != cudaSuccess) return e;
    attr_once.set();
  }
  const int items = B * ((T_max + TT - 1) / TT);
  const int per_sm = smem <= 100 * 1024 ? 2 : 1;
  const int cap = num_sms * per_sm;
  conv_dw_kernel<9, TT><<<items < cap ? items : cap, d / 2, smem, stream>>>(static_cast<const __nv_bfloat16*>(u), static_cast<__nv_bfloat16*>(out),
            w, shift, enc_len, B, T_max, d);
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------ utilities
__global__ void f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n) {
  int64_t i = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(in + i);
    *reinterpret_cast<uint2*>(out + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
  } else {
    for (; i < n; ++i) out[i] = __float2bfloat16_rn(in[i]);
  }
}
cudaError_t launch_f32_to_bf16(const float* in, void* out, int64_t n, cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  const int threads = 256;
  const int64_t blocks = (n / 4 + threads) / threads;
  f32_to_bf16_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(in, static_cast<__nv_bfloat16*>(out), n);
  return cudaGetLastError();
}

// Zero rows t >= len[b] of a padded fp32 [B, T_max, d] tensor (final encoder output hygiene).
__global__ void zero_pad_rows_kernel(float* __restrict__ x, const int32_t* __restrict__ len, int T_max, int d) {
  const int b = blockIdx.y, t = blockIdx.x;
  if (t < len[b]) return;
  float4* r = reinterpret_cast<float4*>(x + (static_cast<size_t>(b) * T_max + t) * d);
  for (int i = threadIdx.x; i < d / 4; i += blockDim.x) r[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}
cudaError_t launch_zero_pad_rows(float* x, const int32_t* len, int B, int T_max, int d, cudaStream_t stream) {
  zero_pad_rows_kernel<<<dim3(T_max, B), 128, 0, stream>>>(x, len, T_max, d);
  return cudaGetLastError();
}

}  // namespace rs
