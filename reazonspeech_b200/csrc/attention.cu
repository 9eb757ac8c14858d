// Relative-position local attention with a global token (N5): the core of NeMo's
// RelPositionMultiHeadAttentionLongformer (parts/submodules/multi_head_attention.py), reached
// through model.transcribe (pkg/nemo-asr/src/transcribe.py:48-53).  Semantics restated in
// oracle/nemo_restated.py::local_attention_core.
//
// One CTA = 64 query rows of one (utterance, head); 4 warps x 16 rows, flash-style online
// softmax in fp32, bf16 mma.sync m16n8k16 for QK^T and PV.  The positional term
// BD[i][c] = (q_i + v).p[c] for all 2w+1 relative offsets c is produced beforehand by ONE batched
// tcgen05 GEMM over all heads (A = q columns of the QKV buffer, W = linear_pos table, bias =
// v.p[c]); this kernel adds it to each key tile through the skewed index c = j - i + w_left
// (no rel-shift pass, no [T,T] score tensor in HBM).  With BD out of the kernel the CTA needs
// 53 KB of shared memory, so 3-4 CTAs share an SM and hide each other's load / mma latency.
// Only key tiles intersecting [i-w_left, i+w_right] are visited, so cost is linear in T
// (long-form audio).  The global token enters as the initial state of the online softmax; its
// own row (full attention) is a separate small kernel.
#include "common.cuh"
#include "kernels.h"

namespace rs {

constexpr int DK = 128;
constexpr int QT = 64;            // query rows per CTA
constexpr int KT = 64;            // keys per tile
constexpr int LDS = DK + 8;       // padded smem row (bf16 elements): conflict-free ldmatrix

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct AttnDev {
  const __nv_bfloat16* qkv; const __half* bd; const float* bias_u;
  __nv_bfloat16* out; const int32_t* enc_len;
  int T_max, H, w_left, w_right, n_global, n_rel_pad;
};

// Asynchronous copy (cp.async, 16 B, L2-only) of 64 rows x 128 bf16 into padded smem; rows >= valid are zero-filled.
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* s, const __nv_bfloat16* g, size_t ld, int valid_rows) {
  for (int id = threadIdx.x; id < 64 * 16; id += blockDim.x) {
    const int r = id >> 4, c = (id & 15) * 8;
    const bool ok = r < valid_rows;
    const __nv_bfloat16* src = g + static_cast<size_t>(ok ? r : 0) * ld + c;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(s + r * LDS + c)), "l"(src), "r"(ok ? 16 : 0) : "memory");
  }
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Copy 64 rows x 128 bf16 (row stride `ld` elements, rows >= valid read as zero) into padded smem.
__device__ __forceinline__ void load_tile(__nv_bfloat16* s, const __nv_bfloat16* g, size_t ld, int valid_rows) {
  for (int id = threadIdx.x; id < 64 * 16; id += blockDim.x) {
    const int r = id >> 4, c = (id & 15) * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (r < valid_rows) v = *reinterpret_cast<const uint4*>(g + static_cast<size_t>(r) * ld + c);
    *reinterpret_cast<uint4*>(s + r * LDS + c) = v;
  }
}

constexpr int kKvStages = 2;      // K/V tile ring (cp.async prefetch of the next key tile during the current one)

__global__ void __launch_bounds__(128, 2)
local_attention_kernel(const AttnDev p) {
  extern __shared__ __align__(16) uint8_t at_smem[];
  __nv_bfloat16* sQU = reinterpret_cast<__nv_bfloat16*>(at_smem);
  __nv_bfloat16* sKV = sQU + QT * LDS;                    // [stages][K tile | V tile]
  float* sG = reinterpret_cast<float*>(sKV + kKvStages * 2 * KT * LDS);   // [QT] global-key score (already scaled)

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
  const int len = p.enc_len[b];
  const int d = p.H * DK;
  const size_t ld = static_cast<size_t>(3) * d;
  const __nv_bfloat16* qbase = p.qkv + (static_cast<size_t>(b) * p.T_max) * ld + h * DK;
  const __nv_bfloat16* kbase = qbase + d;
  const __nv_bfloat16* vbase = qbase + 2 * d;
  __nv_bfloat16* obase = p.out + (static_cast<size_t>(b) * p.T_max) * d + h * DK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const float scale = rsqrtf(static_cast<float>(DK));

  if (q0 >= len) {                                   // fully padded tile: defined zeros
    for (int id = threadIdx.x; id < QT * 16; id += blockDim.x) {
      const int r = id >> 4, c = (id & 15) * 8;
      if (q0 + r < p.T_max) *reinterpret_cast<uint4*>(obase + static_cast<size_t>(q0 + r) * d + c) = make_uint4(0, 0, 0, 0);
    }
    return;
  }

  // ---- phase 0: Q tile (already q + pos_bias_u: folded into the QKV projection's bias at pack time); global-key scores
  for (int id = threadIdx.x; id < QT * 16; id += blockDim.x) {
    const int r = id >> 4, c = (id & 15) * 8;
    uint4 raw = make_uint4(0, 0, 0, 0);
    if (q0 + r < p.T_max) raw = *reinterpret_cast<const uint4*>(qbase + static_cast<size_t>(q0 + r) * ld + c);
    *reinterpret_cast<uint4*>(sQU + r * LDS + c) = raw;
  }
  if (p.n_global > 0) {
    const int r = threadIdx.x >> 1, hf = threadIdx.x & 1;     // 2 threads per row, 64 dims each
    float acc = 0.f;
    if (q0 + r < p.T_max) {
      const uint4* qr = reinterpret_cast<const uint4*>(qbase + static_cast<size_t>(q0 + r) * ld + hf * 64);
      const uint4* k0 = reinterpret_cast<const uint4*>(kbase + hf * 64);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 a = qr[i], bb = __ldg(k0 + i);
        const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 x = unpack_bf16x2(aw[j]), y = unpack_bf16x2(bw[j]);
          const float2 u2 = __ldg(reinterpret_cast<const float2*>(p.bias_u + h * DK + hf * 64 + i * 8 + 2 * j));   // the global key is scored against q, not q + u
          acc = fmaf(x.x - u2.x, y.x, acc); acc = fmaf(x.y - u2.y, y.y, acc);
        }
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (hf == 0) sG[r] = acc * scale;
  }
  __syncthreads();

  const int r0 = warp * 16;
  const uint32_t sQU_a = smem_u32(sQU), sKV_a = smem_u32(sKV);
  // per-lane ldmatrix offsets (bytes)
  const uint32_t a_off = ((r0 + (lane & 15)) * LDS + (lane >> 4) * 8) * 2;             // A: rows r0.., k-halves
  const uint32_t b_off = ((((lane >> 4) & 1) * 8 + (lane & 7)) * LDS + ((lane >> 3) & 1) * 8) * 2;   // B (K-major rows)
  const uint32_t v_off = ((((lane >> 3) & 1) * 8 + (lane & 7)) * LDS + ((lane >> 4) & 1) * 8) * 2;   // B via .trans (V)

  // ---- phase 2: online softmax over the key tiles that intersect the band
  float o[16][4];
  float m_run[2], l_run[2];
  const int i_lo = q0 + r0 + g, i_hi = i_lo + 8;             // the two query rows this thread owns
  if (p.n_global > 0) {
    m_run[0] = sG[r0 + g]; m_run[1] = sG[r0 + g + 8];
    l_run[0] = l_run[1] = 1.f;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float2 v0 = unpack_bf16x2(__ldg(reinterpret_cast<const uint32_t*>(vbase + nt * 8 + 2 * t)));
      o[nt][0] = v0.x; o[nt][1] = v0.y; o[nt][2] = v0.x; o[nt][3] = v0.y;
    }
  } else {
    m_run[0] = m_run[1] = -INFINITY;
    l_run[0] = l_run[1] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
  }
  // BD rows of the two query rows this thread owns (skewed read: column j - i + w_left)
  const __half* bd_lo = p.bd + (static_cast<size_t>(b) * p.T_max + min(i_lo, p.T_max - 1)) * (static_cast<size_t>(p.H) * p.n_rel_pad) + static_cast<size_t>(h) * p.n_rel_pad;
  const __half* bd_hi = p.bd + (static_cast<size_t>(b) * p.T_max + min(i_hi, p.T_max - 1)) * (static_cast<size_t>(p.H) * p.n_rel_pad) + static_cast<size_t>(h) * p.n_rel_pad;
  const int j_first = max(0, q0 - p.w_left);
  const int j_last = min(len - 1, q0 + QT - 1 + p.w_right);
  const int kt_first = j_first / KT, kt_last = j_last / KT;
  auto prefetch = [&](int kt, int st) {
    const int jj = kt * KT;
    __nv_bfloat16* dst = sKV + st * 2 * KT * LDS;
    load_tile_async(dst, kbase + static_cast<size_t>(jj) * ld, ld, min(KT, p.T_max - jj));
    load_tile_async(dst + KT * LDS, vbase + static_cast<size_t>(jj) * ld, ld, min(KT, p.T_max - jj));
    cp_async_commit();
  };
  prefetch(kt_first, 0);
  for (int kt = kt_first; kt <= kt_last; ++kt) {
    const int j0 = kt * KT;
    const int st = (kt - kt_first) & 1;
    if (kt < kt_last) { prefetch(kt + 1, st ^ 1); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
    __syncthreads();
    const uint32_t sK_a = sKV_a + st * 2 * KT * LDS * 2, sV_a = sK_a + KT * LDS * 2;
    // positional terms for this key tile: issued before the QK^T mma chain so the L2 latency hides behind it
    float bdv[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + nt * 8 + 2 * t + (e & 1);
        const int rel = j - ((e < 2) ? i_lo : i_hi);
        const bool ok = (j < len) && (rel >= -p.w_left) && (rel <= p.w_right);
        bdv[nt][e] = ok ? __half2float(__ldg(((e < 2) ? bd_lo : bd_hi) + rel + p.w_left)) : 0.f;
      }
    }
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; }
#pragma unroll
    for (int ks = 0; ks < DK / 16; ++ks) {
      uint32_t a[4];
      ldsm_x4(sQU_a + a_off + ks * 32, a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int np = 0; np < 4; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(sK_a + b_off + (np * 16 * LDS) * 2 + ks * 32, b0, b1, b2, b3);
        mma_bf16(s[2 * np], a, b0, b1);
        mma_bf16(s[2 * np + 1], a, b2, b3);
      }
    }
    // positional term, band / padding mask, scale
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int j = j0 + nt * 8 + 2 * t + (e & 1);
        const int i = (e < 2) ? i_lo : i_hi;
        const int rel = j - i;
        const bool ok = (j < len) && (rel >= -p.w_left) && (rel <= p.w_right);
        float v = -INFINITY;
        if (ok) v = (s[nt][e] + bdv[nt][e]) * scale;
        s[nt][e] = v;
        mx[e >> 1] = fmaxf(mx[e >> 1], v);
      }
    }
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      mx[hrow] = fmaxf(mx[hrow], __shfl_xor_sync(0xffffffffu, mx[hrow], 1));
      mx[hrow] = fmaxf(mx[hrow], __shfl_xor_sync(0xffffffffu, mx[hrow], 2));
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      const float m_new = fmaxf(m_run[hrow], mx[hrow]);
      msafe[hrow] = (m_new == -INFINITY) ? 0.f : m_new;
      corr[hrow] = __expf(m_run[hrow] - msafe[hrow]);           // exp(-inf) == 0 when nothing seen yet
      m_run[hrow] = m_new;
      l_run[hrow] *= corr[hrow];
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = __expf(s[nt][e] - msafe[e >> 1]);
        s[nt][e] = pv;
        rs[e >> 1] += pv;
      }
    }
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
      rs[hrow] += __shfl_xor_sync(0xffffffffu, rs[hrow], 1);
      rs[hrow] += __shfl_xor_sync(0xffffffffu, rs[hrow], 2);
      l_run[hrow] += rs[hrow];
    }
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) { o[nt][0] *= corr[0]; o[nt][1] *= corr[0]; o[nt][2] *= corr[1]; o[nt][3] *= corr[1]; }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < KT / 16; ++kk) {
      uint32_t a[4];
      a[0] = pack_bf16x2(s[2 * kk][0], s[2 * kk][1]);
      a[1] = pack_bf16x2(s[2 * kk][2], s[2 * kk][3]);
      a[2] = pack_bf16x2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      a[3] = pack_bf16x2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 8; ++np) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(sV_a + v_off + (kk * 16 * LDS) * 2 + np * 32, b0, b1, b2, b3);
        mma_bf16(o[2 * np], a, b0, b1);
        mma_bf16(o[2 * np + 1], a, b2, b3);
      }
    }
    __syncthreads();                                   // all warps done with this stage before it is refilled
  }
  // ---- write back
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    const int i = hrow == 0 ? i_lo : i_hi;
    if (i >= p.T_max) continue;
    const bool valid = i < len;
    const float inv = valid ? 1.0f / l_run[hrow] : 0.f;
    __nv_bfloat16* orow = obase + static_cast<size_t>(i) * d;
#pragma unroll
    for (int nt = 0; nt < 16; ++nt) {
      const float x = valid ? o[nt][2 * hrow] * inv : 0.f, y = valid ? o[nt][2 * hrow + 1] * inv : 0.f;
      *reinterpret_cast<uint32_t*>(orow + nt * 8 + 2 * t) = pack_bf16x2(x, y);
    }
  }
}

// Row(s) of the global token(s): full attention softmax_j((q_g / sqrt(dk)) . k_j) v_j, no positional
// terms.  grid (H, B), 256 threads.  Scores: one key per thread (16 independent 16-byte loads in flight
// per thread); weighted sum of V: 4 key groups x 64 dim pairs, unrolled by 8.
__global__ void __launch_bounds__(256)
global_row_attention_kernel(const AttnDev p) {
  extern __shared__ __align__(16) float gs[];        // [T_max] scores | [8] reduction scratch | [128] q
  const int h = blockIdx.x, b = blockIdx.y;
  const int len = p.enc_len[b];
  if (len <= 0) return;
  const int d = p.H * DK;
  const size_t ld = static_cast<size_t>(3) * d;
  const __nv_bfloat16* qrow = p.qkv + (static_cast<size_t>(b) * p.T_max) * ld + h * DK;   // query = frame 0
  const __nv_bfloat16* kbase = qrow + d;
  const __nv_bfloat16* vbase = qrow + 2 * d;
  const float scale = rsqrtf(static_cast<float>(DK));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int t_pad = ((p.T_max > 1024 ? p.T_max : 1024) + 3) & ~3;
  float* red = gs + t_pad;
  float* sq = red + 8;
  if (tid < DK) sq[tid] = (__bfloat162float(qrow[tid]) - p.bias_u[h * DK + tid]) * scale;   // q columns hold q + pos_bias_u
  __syncthreads();
  float mx = -INFINITY;
  for (int j = tid; j < len; j += 256) {
    const uint4* kr = reinterpret_cast<const uint4*>(kbase + static_cast<size_t>(j) * ld);
    uint4 kk[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) kk[i] = __ldg(kr + i);
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float4 qa = *reinterpret_cast<const float4*>(sq + 8 * i), qb = *reinterpret_cast<const float4*>(sq + 8 * i + 4);
      d0 = fmaf(bf16_lo(kk[i].x), qa.x, d0); d1 = fmaf(bf16_hi(kk[i].x), qa.y, d1);
      d2 = fmaf(bf16_lo(kk[i].y), qa.z, d2); d3 = fmaf(bf16_hi(kk[i].y), qa.w, d3);
      d0 = fmaf(bf16_lo(kk[i].z), qb.x, d0); d1 = fmaf(bf16_hi(kk[i].z), qb.y, d1);
      d2 = fmaf(bf16_lo(kk[i].w), qb.z, d2); d3 = fmaf(bf16_hi(kk[i].w), qb.w, d3);
    }
    const float v = (d0 + d1) + (d2 + d3);
    gs[j] = v;
    mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3])), fmaxf(fmaxf(red[4], red[5]), fmaxf(red[6], red[7])));
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < len; j += 256) { const float e = __expf(gs[j] - mx); gs[j] = e; sum += e; }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  const float inv = 1.0f / (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7])));
  const int kg = tid >> 6, cp = tid & 63;            // key group (4), dim pair (64)
  float a0 = 0.f, a1 = 0.f;
  for (int j0 = kg; j0 < len; j0 += 32) {
    uint32_t vv[8]; float pj[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = j0 + 4 * u;
      const int jc = min(j, len - 1);
      vv[u] = __ldg(reinterpret_cast<const uint32_t*>(vbase + static_cast<size_t>(jc) * ld) + cp);
      pj[u] = j < len ? gs[jc] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) { a0 = fmaf(pj[u], bf16_lo(vv[u]), a0); a1 = fmaf(pj[u], bf16_hi(vv[u]), a1); }
  }
  __syncthreads();                                   // everyone is done reading gs
  float2* part = reinterpret_cast<float2*>(gs);      // [4 key groups][64 dim pairs]
  part[kg * 64 + cp] = make_float2(a0, a1);
  __syncthreads();
  if (tid < 64) {
    const float2 x0 = part[tid], x1 = part[64 + tid], x2 = part[128 + tid], x3 = part[192 + tid];
    const float r0 = ((x0.x + x1.x) + (x2.x + x3.x)) * inv, r1 = ((x0.y + x1.y) + (x2.y + x3.y)) * inv;
    *reinterpret_cast<uint32_t*>(p.out + (static_cast<size_t>(b) * p.T_max) * d + h * DK + 2 * tid) = pack_bf16x2(r0, r1);
  }
}

cudaError_t launch_attention(const AttnArgs& a, cudaStream_t stream) {
  if (a.dk != DK || a.n_global < 0 || a.n_global > 1) return cudaErrorInvalidValue;
  AttnDev p;
  p.qkv = static_cast<const __nv_bfloat16*>(a.qkv); p.bd = static_cast<const __half*>(a.bd);
  p.bias_u = a.bias_u; p.out = static_cast<__nv_bfloat16*>(a.out); p.enc_len = a.enc_len;
  p.T_max = a.T_max; p.H = a.H; p.w_left = a.w_left; p.w_right = a.w_right; p.n_global = a.n_global;
  p.n_rel_pad = a.n_rel_pad;
  if (a.n_rel_pad < a.w_left + a.w_right + 1) return cudaErrorInvalidValue;
  const size_t smem = static_cast<size_t>(1 + 2 * kKvStages) * QT * LDS * 2 + QT * 4;
  static DeviceOnce attr_once;
  if (attr_once.pending()) {
    cudaError_t e = cudaFuncSetAttribute(local_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(global_row_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    attr_once.set();
  }
  const dim3 grid((a.T_max + QT - 1) / QT, a.H, a.B);
  local_attention_kernel<<<grid, 128, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (a.n_global > 0) {
    const size_t gsmem = (static_cast<size_t>(((a.T_max > 1024 ? a.T_max : 1024) + 3) & ~3) + 8 + DK) * sizeof(float);
    if (gsmem > 200 * 1024) return cudaErrorInvalidValue;
    global_row_attention_kernel<<<dim3(a.H, a.B), 256, gsmem, stream>>>(p);
    e = cudaGetLastError();
  }
  return e;
}

}  // namespace rs
