// Persistent RNN-T greedy decode (N8 + N9): RNNTDecoder.predict + RNNTJoint.joint +
// GreedyRNNTInfer._greedy_decode of NeMo (modules/rnnt.py, rnnt_greedy_decoding.py), reached
// through model.transcribe (pkg/nemo-asr/src/transcribe.py:48-53).  In NeMo the (t,u) loop is
// host-driven with a device->host sync per step; here one thread-block CLUSTER owns one
// utterance for its whole lifetime and the token loop never returns to the host.
//
// Per joint evaluation:  logits = W_out relu(enc_proj[t] + pred_proj) + b  (W_out bf16, fp32
// activations and accumulation), argmax; blank -> next frame; otherwise emit, one LSTM step on
// (embed[k], h, c) and pred_proj = W_pred h + b.  The GEMV rows (vocabulary, LSTM units, joint
// rows) are split across the CTAs of the cluster; partial argmax / new h / new pred_proj are
// exchanged through distributed shared memory with one cluster barrier each.  Weights stream
// from L2 (11 MB bf16 total, resident in the 126 MB L2); a half-warp owns a weight row so every
// load is a coalesced 256 B segment and 5-10 independent 16 B loads are in flight per lane.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace cg = cooperative_groups;

namespace rs {

constexpr int kDecThreads = 512;
constexpr int kDecUnits = kDecThreads / 16;      // half-warps
constexpr int kMaxCluster = 8;

struct DecodeDev {
  const float* enc_proj; const int32_t* enc_len;
  const __nv_bfloat16* w_out; const float* b_out; const float* embed;
  const __nv_bfloat16* w_lstm; const float* b_lstm; const __nv_bfloat16* w_pred; const float* b_pred;
  int32_t* tokens; int32_t* frames; int32_t* n_tok;
  int T_max, Hj, Hp, V, U_max, max_symbols;
};

// dot(W[row, :], x) for one row handled by a half-warp: lane hl owns chunks hl, hl+16, ...
template <int NCH>
__device__ __forceinline__ float row_dot(const __nv_bfloat16* __restrict__ wrow, const float* __restrict__ x, int hl) {
  uint4 w[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) w[i] = ldg_nc_v4(wrow + (hl + 16 * i) * 8);
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const float4 x0 = *reinterpret_cast<const float4*>(x + (hl + 16 * i) * 8);
    const float4 x1 = *reinterpret_cast<const float4*>(x + (hl + 16 * i) * 8 + 4);
    acc = fmaf(bf16_lo(w[i].x), x0.x, acc); acc = fmaf(bf16_hi(w[i].x), x0.y, acc);
    acc = fmaf(bf16_lo(w[i].y), x0.z, acc); acc = fmaf(bf16_hi(w[i].y), x0.w, acc);
    acc = fmaf(bf16_lo(w[i].z), x1.x, acc); acc = fmaf(bf16_hi(w[i].z), x1.y, acc);
    acc = fmaf(bf16_lo(w[i].w), x1.z, acc); acc = fmaf(bf16_hi(w[i].w), x1.w, acc);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// Two rows at once: 2*NCH independent 16 B loads in flight per lane (the GEMV is bound by
// memory-level parallelism against ~1 us of L2 latency, not by bandwidth or FMAs).
template <int NCH>
__device__ __forceinline__ void row_dot2(const __nv_bfloat16* __restrict__ w0, const __nv_bfloat16* __restrict__ w1,
                                         const float* __restrict__ x, int hl, float& r0, float& r1) {
  uint4 a[NCH], b[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) { a[i] = ldg_nc_v4(w0 + (hl + 16 * i) * 8); b[i] = ldg_nc_v4(w1 + (hl + 16 * i) * 8); }
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int i = 0; i < NCH; ++i) {
    const float4 x0 = *reinterpret_cast<const float4*>(x + (hl + 16 * i) * 8);
    const float4 x1 = *reinterpret_cast<const float4*>(x + (hl + 16 * i) * 8 + 4);
    s0 = fmaf(bf16_lo(a[i].x), x0.x, s0); s0 = fmaf(bf16_hi(a[i].x), x0.y, s0);
    s0 = fmaf(bf16_lo(a[i].y), x0.z, s0); s0 = fmaf(bf16_hi(a[i].y), x0.w, s0);
    s0 = fmaf(bf16_lo(a[i].z), x1.x, s0); s0 = fmaf(bf16_hi(a[i].z), x1.y, s0);
    s0 = fmaf(bf16_lo(a[i].w), x1.z, s0); s0 = fmaf(bf16_hi(a[i].w), x1.w, s0);
    s1 = fmaf(bf16_lo(b[i].x), x0.x, s1); s1 = fmaf(bf16_hi(b[i].x), x0.y, s1);
    s1 = fmaf(bf16_lo(b[i].y), x0.z, s1); s1 = fmaf(bf16_hi(b[i].y), x0.w, s1);
    s1 = fmaf(bf16_lo(b[i].z), x1.x, s1); s1 = fmaf(bf16_hi(b[i].z), x1.y, s1);
    s1 = fmaf(bf16_lo(b[i].w), x1.z, s1); s1 = fmaf(bf16_hi(b[i].w), x1.w, s1);
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
  r0 = s0; r1 = s1;
}

template <int NCH_J, int NCH_L, int NCH_P>
__global__ void __launch_bounds__(kDecThreads, 1)
rnnt_greedy_kernel(const DecodeDev p) {
  cg::cluster_group cluster = cg::this_cluster();
  const int CS = static_cast<int>(cluster.num_blocks());
  const int rank = static_cast<int>(cluster.block_rank());
  const int b = blockIdx.x / CS;
  const int Hj = p.Hj, Hp = p.Hp, NC = p.V + 1, blank = p.V;

  extern __shared__ __align__(16) float dsm[];
  float* s_g = dsm;                       // [Hj]      relu(enc_proj[t] + pred_proj)
  float* s_pp = s_g + Hj;                 // [Hj]      pred_proj (full, replicated in every CTA)
  float* s_x = s_pp + Hj;                 // [2][2*Hp] LSTM input (embed | h), double buffered
  float* s_c = s_x + 4 * Hp;              // [Hp/CS]   cell state of the units this CTA owns
  float* s_gate = s_c + Hp;               // [4][Hp/CS]
  float* s_uval = s_gate + 4 * Hp;        // [kDecUnits]
  int* s_uidx = reinterpret_cast<int*>(s_uval + kDecUnits);   // [kDecUnits]
  float* s_xval = reinterpret_cast<float*>(s_uidx + kDecUnits);   // [2][kMaxCluster]
  int* s_xidx = reinterpret_cast<int*>(s_xval + 2 * kMaxCluster); // [2][kMaxCluster]

  const int tid = threadIdx.x, lane = tid & 31, hl = lane & 15;
  const int unit = tid >> 4;
  const int T = p.enc_len[b];

  const int rows_j = (NC + CS - 1) / CS;
  const int j_begin = rank * rows_j, j_end = min(NC, j_begin + rows_j);
  const int us = Hp / CS;                 // LSTM units owned
  const int u_begin = rank * us;
  const int ps = Hj / CS;                 // pred_proj rows owned
  const int p_begin = rank * ps;

  for (int i = tid; i < 4 * Hp; i += kDecThreads) s_x[i] = 0.f;
  for (int i = tid; i < us; i += kDecThreads) s_c[i] = 0.f;
  cluster.sync();

  int xb = 0;                             // x buffer holding the committed h
  // One LSTM step on s_x[xb] (embed part already filled) followed by pred_proj; leaves h in s_x[xb^1].
  auto lstm_and_pred = [&]() {
    const float* x = s_x + xb * 2 * Hp;
    for (int base = 0; base < 4 * us; base += 2 * kDecUnits) {      // warp-uniform trip count, two rows per pass
      const int ra = base + unit, rb = base + kDecUnits + unit;
      const bool oka = ra < 4 * us, okb = rb < 4 * us;
      const int qa = oka ? ra : 4 * us - 1, qb = okb ? rb : 4 * us - 1;
      const int rowa = (qa / us) * Hp + u_begin + qa % us, rowb = (qb / us) * Hp + u_begin + qb % us;
      float va, vb;
      row_dot2<NCH_L>(p.w_lstm + static_cast<size_t>(rowa) * (2 * Hp), p.w_lstm + static_cast<size_t>(rowb) * (2 * Hp), x, hl, va, vb);
      if (hl == 0) {
        if (oka) s_gate[qa] = va + __ldg(p.b_lstm + rowa);
        if (okb) s_gate[qb] = vb + __ldg(p.b_lstm + rowb);
      }
    }
    __syncthreads();
    for (int u = tid; u < us; u += kDecThreads) {
      const float ig = sigmoidf_accurate(s_gate[u]), fg = sigmoidf_accurate(s_gate[us + u]);
      const float gg = tanhf(s_gate[2 * us + u]), og = sigmoidf_accurate(s_gate[3 * us + u]);
      const float c2 = fg * s_c[u] + ig * gg;
      s_c[u] = c2;
      const float h2 = og * tanhf(c2);
      for (int r = 0; r < CS; ++r) {
        float* remote = cluster.map_shared_rank(s_x, r);
        remote[(xb ^ 1) * 2 * Hp + Hp + u_begin + u] = h2;
      }
    }
    cluster.sync();
    xb ^= 1;
    const float* hvec = s_x + xb * 2 * Hp + Hp;
    for (int base = 0; base < ps; base += kDecUnits) {
      const bool ok = base + unit < ps;
      const int row = p_begin + (ok ? base + unit : ps - 1);
      const float v = row_dot<NCH_P>(p.w_pred + static_cast<size_t>(row) * Hp, hvec, hl);
      if (ok && hl == 0) {
        const float val = v + __ldg(p.b_pred + row);
        for (int rr = 0; rr < CS; ++rr) cluster.map_shared_rank(s_pp, rr)[row] = val;
      }
    }
    cluster.sync();
  };

  lstm_and_pred();                        // SOS: blank embedding == zero vector, zero state

  int t = 0, n_emit = 0, symbols = 0, par = 0;
  while (t < T) {
    const float* ep = p.enc_proj + (static_cast<size_t>(b) * p.T_max + t) * Hj;
    for (int i = tid; i < Hj; i += kDecThreads) s_g[i] = fmaxf(__ldg(ep + i) + s_pp[i], 0.f);
    __syncthreads();
    // ---- partial argmax over this CTA's vocabulary rows
    float best = -INFINITY; int best_i = 0x7fffffff;
    for (int base = j_begin; base < j_end; base += 2 * kDecUnits) {  // warp-uniform trip count, two rows per pass
      const int ra = base + unit, rb = base + kDecUnits + unit;
      const bool oka = ra < j_end, okb = rb < j_end;
      const int rowa = oka ? ra : j_end - 1, rowb = okb ? rb : j_end - 1;
      float va, vb;
      row_dot2<NCH_J>(p.w_out + static_cast<size_t>(rowa) * Hj, p.w_out + static_cast<size_t>(rowb) * Hj, s_g, hl, va, vb);
      va += __ldg(p.b_out + rowa); vb += __ldg(p.b_out + rowb);
      if (oka && va > best) { best = va; best_i = rowa; }   // rows visited in increasing order: first max wins
      if (okb && vb > best) { best = vb; best_i = rowb; }
    }
    if (hl == 0) { s_uval[unit] = best; s_uidx[unit] = best_i; }
    __syncthreads();
    if (tid < 32) {
      float v = s_uval[tid]; int ix = s_uidx[tid];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
        if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
      }
      if (tid < CS) {
        cluster.map_shared_rank(s_xval, tid)[par * kMaxCluster + rank] = v;
        cluster.map_shared_rank(s_xidx, tid)[par * kMaxCluster + rank] = ix;
      }
    }
    cluster.sync();
    float bv = s_xval[par * kMaxCluster]; int k = s_xidx[par * kMaxCluster];
    for (int r = 1; r < CS; ++r) {
      const float ov = s_xval[par * kMaxCluster + r]; const int oi = s_xidx[par * kMaxCluster + r];
      if (ov > bv || (ov == bv && oi < k)) { bv = ov; k = oi; }
    }
    par ^= 1;
    if (k == blank) { ++t; symbols = 0; continue; }
    // ---- emit
    if (rank == 0 && tid == 0 && n_emit < p.U_max) {
      p.tokens[static_cast<size_t>(b) * p.U_max + n_emit] = k;
      p.frames[static_cast<size_t>(b) * p.U_max + n_emit] = t;
    }
    ++n_emit;
    float* xe = s_x + xb * 2 * Hp;
    for (int i = tid; i < Hp; i += kDecThreads) xe[i] = __ldg(p.embed + static_cast<size_t>(k) * Hp + i);
    __syncthreads();
    lstm_and_pred();
    if (++symbols >= p.max_symbols) { ++t; symbols = 0; }
  }
  if (rank == 0 && tid == 0) p.n_tok[b] = n_emit;
  cluster.sync();                         // no CTA exits while a peer may still address its smem
}

template <int A, int B2, int C>
static cudaError_t launch_dec(const DecodeDev& p, int B, int CS, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(rnnt_greedy_kernel<A, B2, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(B * CS); cfg.blockDim = dim3(kDecThreads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, rnnt_greedy_kernel<A, B2, C>, p);
}

cudaError_t launch_rnnt_greedy(const DecodeArgs& a, int num_sms, cudaStream_t stream) {
  if (a.Hj % 128 || a.Hp % 128 || a.B <= 0) return cudaErrorInvalidValue;
  DecodeDev p{a.enc_proj, a.enc_len, static_cast<const __nv_bfloat16*>(a.w_out), a.b_out, a.embed,
              static_cast<const __nv_bfloat16*>(a.w_lstm), a.b_lstm, static_cast<const __nv_bfloat16*>(a.w_pred), a.b_pred,
              a.tokens, a.frames, a.n_tok, a.T_max, a.Hj, a.Hp, a.V, a.U_max, a.max_symbols};
  int CS = 1;
  for (int c = kMaxCluster; c >= 1; c >>= 1)
    if (a.B * c <= num_sms && a.Hp % c == 0 && a.Hj % c == 0) { CS = c; break; }
  const size_t smem = (2 * a.Hj + 4 * a.Hp + a.Hp + 4 * a.Hp + 2 * kDecUnits + 4 * kMaxCluster) * sizeof(float) + 64;
  const int nj = a.Hj / 128, nl = 2 * a.Hp / 128, np = a.Hp / 128;
  if (nj == 5 && nl == 10 && np == 5) return launch_dec<5, 10, 5>(p, a.B, CS, smem, stream);
  if (nj == 1 && nl == 2 && np == 1) return launch_dec<1, 2, 1>(p, a.B, CS, smem, stream);
  return cudaErrorInvalidValue;
}

}  // namespace rs
