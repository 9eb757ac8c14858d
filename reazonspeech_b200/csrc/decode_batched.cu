// Batched, weights-stationary RNN-T greedy decode (N8 + N9) for B >= 8 utterances.
//
// Same algorithm as decode.cu (NeMo GreedyRNNTInfer._greedy_decode, RNNTDecoder.predict,
// RNNTJoint.joint; pkg/nemo-asr/src/transcribe.py:48-53 is where the reference reaches it), but
// organised around the fact that the decode weights (joint output 3001x640, LSTM 2560x1280, joint
// pred 640x640: 11 MB in bf16) fit in the shared memory of the whole GPU: every CTA of a persistent
// cooperative grid keeps a fixed slice of each matrix in ITS shared memory for the whole decode
// (about 21 vocabulary rows, 5 LSTM units x 4 gates, 5 pred rows), and all utterances advance one
// joint evaluation per iteration.  Nothing but activations moves: per iteration a CTA reads B x 2.5 KB
// of enc_proj / pred_proj rows from L2 instead of streaming 3.8 MB of weights per utterance, so the
// cost per iteration is set by three grid barriers and is almost independent of B.
//
//   phase J  every CTA: argmax over W_out[slice] relu(enc_proj[b,t_b] + pred_proj[b]) for all b, folded into a
//            grid-wide 64-bit red.max per utterance (logit bits | complemented row index)
//   barrier  -> every CTA reads the same token k_b, updates (t_b, symbols_b), emits
//   phase L  utterances that emitted: LSTM gates of the CTA's units on (embed[k_b], h_b) -> new h slice
//   barrier
//   phase P  utterances that emitted: pred_proj rows of the CTA from the new h
//   barrier
//
// fp32 activations and accumulation, bf16 weights (exact copies of the packed checkpoint), argmax ties
// to the lower index: results are identical to the per-utterance kernel.
#include <cooperative_groups.h>

#include "common.cuh"
#include "kernels.h"

namespace rs {

constexpr int kBdThreads = 256;
constexpr int kBdWarps = kBdThreads / 32;
constexpr int kRowBlk = 8;                  // joint rows accumulated per register block
constexpr int kUpw = 4;                     // utterances per warp in phase J (W rows are read once per 4 utterances)

struct BatchedDev {
  const float* enc_proj; const int32_t* enc_len;
  const __nv_bfloat16* w_out; const float* b_out; const float* embed;
  const __nv_bfloat16* w_lstm; const float* b_lstm; const __nv_bfloat16* w_pred; const float* b_pred;
  int32_t* tokens; int32_t* frames; int32_t* n_tok;
  unsigned long long* best; // [3][B] packed (ordered logit bits << 32 | ~row): grid-wide argmax by red.max, 3-deep ring
  float* hbuf;              // [2][B][Hp]
  float* ppbuf;             // [B][Hj]
  unsigned int* counter;    // grid barrier
  long long* prof;          // [8] cycle counters of CTA 0 (phase J, barrier, reduce, phase L, barrier, phase P, barrier, iterations)
  int B, T_max, Hj, Hp, V, U_max, max_symbols;
  int rows_j, units, rows_p;   // per-CTA slice sizes
};

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int& target, unsigned int nblocks) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += nblocks;
    __threadfence();
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(counter) : "memory");
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
  }
  __syncthreads();
}

// Butterfly reduce-scatter of N (power of two, >= 32) per-lane values: on return v[0..N/32) of lane l
// hold the warp totals of the original v[(N/32)*l .. (N/32)*l + N/32).
template <int N>
__device__ __forceinline__ void warp_reduce_scatter(float (&v)[N], int lane) {
#pragma unroll
  for (int o = 16, n = N; o >= 1; o >>= 1, n >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
}

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

template <int KJ, int KP>     // floats per lane of a joint-width (Hj/32) and a pred-width (Hp/32) vector
__global__ void __launch_bounds__(kBdThreads, 1)
rnnt_greedy_batched_kernel(const BatchedDev p) {
  extern __shared__ __align__(16) uint8_t bsm[];
  const int G = gridDim.x, cta = blockIdx.x;
  const int Hj = p.Hj, Hp = p.Hp, NC = p.V + 1, blank = p.V, B = p.B;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // ---- slices owned by this CTA
  const int j0 = min(NC, cta * p.rows_j), j1 = min(NC, j0 + p.rows_j);
  const int nj = j1 - j0;
  const int u0 = min(Hp, cta * p.units), u1 = min(Hp, u0 + p.units);
  const int nu = u1 - u0;
  const int p0 = min(Hj, cta * p.rows_p), p1 = min(Hj, p0 + p.rows_p);
  const int np = p1 - p0;

  // ---- shared memory carve-up
  __nv_bfloat16* s_wout = reinterpret_cast<__nv_bfloat16*>(bsm);                    // [rows_j][Hj]
  __nv_bfloat16* s_wlstm = s_wout + static_cast<size_t>(p.rows_j) * Hj;             // [4*units][2*Hp]  (gate-major)
  __nv_bfloat16* s_wpred = s_wlstm + static_cast<size_t>(4 * p.units) * 2 * Hp;     // [rows_p][Hp]
  float* s_c = reinterpret_cast<float*>(s_wpred + static_cast<size_t>(p.rows_p) * Hp);   // [B][units]
  int* s_t = reinterpret_cast<int*>(s_c + static_cast<size_t>(B) * p.units);        // [B]
  int* s_sym = s_t + B; int* s_n = s_sym + B; int* s_par = s_n + B; int* s_tok = s_par + B;
  int* s_emit = s_tok + B;                                                            // [B] compact list
  int* s_len = s_emit + B;                                                            // [B] enc_len copy
  int* s_cnt = s_len + B;                                                            // [2]: n_emit, n_active

  for (int i = tid; i < nj * Hj / 8; i += kBdThreads)
    reinterpret_cast<uint4*>(s_wout)[i] = reinterpret_cast<const uint4*>(p.w_out + static_cast<size_t>(j0) * Hj)[i];
  for (int r = 0; r < 4 * nu; ++r) {
    const int gate = r / nu, u = r % nu;
    const uint4* src = reinterpret_cast<const uint4*>(p.w_lstm + (static_cast<size_t>(gate) * Hp + u0 + u) * 2 * Hp);
    uint4* dst = reinterpret_cast<uint4*>(s_wlstm + static_cast<size_t>(gate * p.units + u) * 2 * Hp);
    for (int i = tid; i < 2 * Hp / 8; i += kBdThreads) dst[i] = src[i];
  }
  for (int i = tid; i < np * Hp / 8; i += kBdThreads)
    reinterpret_cast<uint4*>(s_wpred)[i] = reinterpret_cast<const uint4*>(p.w_pred + static_cast<size_t>(p0) * Hp)[i];
  for (int i = tid; i < B * p.units; i += kBdThreads) s_c[i] = 0.f;
  for (int b = tid; b < B; b += kBdThreads) { s_t[b] = 0; s_sym[b] = 0; s_n[b] = 0; s_par[b] = 0; s_tok[b] = blank; s_emit[b] = b; s_len[b] = p.enc_len[b]; }
  if (tid == 0) { s_cnt[0] = B; s_cnt[1] = 0; }
  __syncthreads();

  unsigned int target = 0;
  int iter = 0;
  long long prof[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long jprof[3] = {0, 0, 0};   // inside phase J: operand loads, row loop + reduce, argmax + store
  auto tick = [&](int slot, long long& t0) { if (cta == 0 && tid == 0) { const long long t1 = clock64(); prof[slot] += t1 - t0; t0 = t1; } };
  long long tk = clock64();

  // LSTM step + pred_proj for the utterances listed in s_emit[0..n_emit): token s_tok[b], state parity s_par[b].
  auto lstm_and_pred = [&]() {
    const int n_emit = s_cnt[0];
    // ---- phase L
    if (nu > 0) {
      for (int eq = warp; eq < n_emit; eq += kBdWarps) {
        const int b = s_emit[(eq + cta) % n_emit];
        const int k = s_tok[b], par = s_par[b];
        float x[2 * KP];                          // lane slice of (embed[k] | h_b): 2*Hp/32 contiguous values
        const float* src = (lane < 16) ? p.embed + static_cast<size_t>(k) * Hp + lane * 2 * KP
                                       : p.hbuf + (static_cast<size_t>(par) * B + b) * Hp + (lane - 16) * 2 * KP;
#pragma unroll
        for (int i = 0; i < 2 * KP / 4; ++i) {
          const float4 v = ldcg4(src + 4 * i);
          x[4 * i] = v.x; x[4 * i + 1] = v.y; x[4 * i + 2] = v.z; x[4 * i + 3] = v.w;
        }
        float acc[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) acc[r] = 0.f;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
          if (r < 4 * p.units) {                  // warp-uniform
            const uint2* wr = reinterpret_cast<const uint2*>(s_wlstm + static_cast<size_t>(r) * 2 * Hp + lane * 2 * KP);
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int i = 0; i < 2 * KP / 4; ++i) {
              const uint2 w = wr[i];
              a0 = fmaf(bf16_lo(w.x), x[4 * i], a0); a1 = fmaf(bf16_hi(w.x), x[4 * i + 1], a1);
              a2 = fmaf(bf16_lo(w.y), x[4 * i + 2], a2); a3 = fmaf(bf16_hi(w.y), x[4 * i + 3], a3);
            }
            acc[r] = (a0 + a1) + (a2 + a3);
          }
        }
        warp_reduce_scatter<32>(acc, lane);       // lane l now holds the total of gate row l
        const float tot = acc[0];
        // gather the four gates of unit (lane) into lanes 0..nu-1
        const int uu = lane < nu ? lane : 0;
        const float gi = __shfl_sync(0xffffffffu, tot, 0 * p.units + uu);
        const float gf = __shfl_sync(0xffffffffu, tot, 1 * p.units + uu);
        const float gg = __shfl_sync(0xffffffffu, tot, 2 * p.units + uu);
        const float go = __shfl_sync(0xffffffffu, tot, 3 * p.units + uu);
        if (lane < nu) {
          const int unit = u0 + lane;
          const float ig = sigmoidf_accurate(gi + __ldg(p.b_lstm + unit));
          const float fg = sigmoidf_accurate(gf + __ldg(p.b_lstm + Hp + unit));
          const float cg = tanhf(gg + __ldg(p.b_lstm + 2 * Hp + unit));
          const float og = sigmoidf_accurate(go + __ldg(p.b_lstm + 3 * Hp + unit));
          const float c2 = fg * s_c[b * p.units + lane] + ig * cg;
          s_c[b * p.units + lane] = c2;
          __stcg(p.hbuf + (static_cast<size_t>(par ^ 1) * B + b) * Hp + unit, og * tanhf(c2));
        }
      }
    }
    tick(3, tk);
    grid_barrier(p.counter, target, G);
    tick(4, tk);
    // ---- phase P (state parity flips for the utterances that stepped)
    for (int e = tid; e < n_emit; e += kBdThreads) s_par[s_emit[e]] ^= 1;
    __syncthreads();
    if (np > 0) {
      for (int eq = warp; eq < n_emit; eq += kBdWarps) {
        const int b = s_emit[(eq + cta) % n_emit];
        float hv[KP];
        const float* src = p.hbuf + (static_cast<size_t>(s_par[b]) * B + b) * Hp + lane * KP;
#pragma unroll
        for (int i = 0; i < KP / 4; ++i) {
          const float4 v = ldcg4(src + 4 * i);
          hv[4 * i] = v.x; hv[4 * i + 1] = v.y; hv[4 * i + 2] = v.z; hv[4 * i + 3] = v.w;
        }
        for (int r = 0; r < np; ++r) {
          const uint2* wr = reinterpret_cast<const uint2*>(s_wpred + static_cast<size_t>(r) * Hp + lane * KP);
          float a = 0.f;
#pragma unroll
          for (int i = 0; i < KP / 4; ++i) {
            const uint2 w = wr[i];
            a = fmaf(bf16_lo(w.x), hv[4 * i], a); a = fmaf(bf16_hi(w.x), hv[4 * i + 1], a);
            a = fmaf(bf16_lo(w.y), hv[4 * i + 2], a); a = fmaf(bf16_hi(w.y), hv[4 * i + 3], a);
          }
          a = warp_sum(a);
          if (lane == 0) __stcg(p.ppbuf + static_cast<size_t>(b) * Hj + p0 + r, a + __ldg(p.b_pred + p0 + r));
        }
      }
    }
    tick(5, tk);
    grid_barrier(p.counter, target, G);
    tick(6, tk);
  };

  lstm_and_pred();                                // SOS: every utterance steps once on the blank (zero) embedding

  for (;;) {
    // ---- phase J: partial argmax over this CTA's vocabulary rows, kUpw utterances per warp
    for (int base = 0; base < B; base += kUpw * kBdWarps) {
      int bu[kUpw]; bool ok[kUpw];
      float g[kUpw][KJ];
      // every CTA needs the same B activation rows; rotating which warp takes which utterance by the CTA
      // index keeps the 148 CTAs from requesting the same L2 lines at the same instant (hot-spotting)
#pragma unroll
      for (int u = 0; u < kUpw; ++u) {
        bu[u] = base + (kUpw * warp + u + 5 * cta) % (kUpw * kBdWarps);
        ok[u] = bu[u] < B && s_t[bu[u]] < s_len[bu[u]];
      }
      bool any = false;
#pragma unroll
      for (int u = 0; u < kUpw; ++u) any |= ok[u];
      if (!any) continue;                                      // warp-uniform: nothing active in this group
#pragma unroll
      for (int u = 0; u < kUpw; ++u) {
        if (ok[u]) {                                           // warp-uniform
          const int b = bu[u], t = s_t[b];
          const float* ep = p.enc_proj + (static_cast<size_t>(b) * p.T_max + t) * Hj + lane * KJ;
          const float* pp = p.ppbuf + static_cast<size_t>(b) * Hj + lane * KJ;
#pragma unroll
          for (int i = 0; i < KJ / 4; ++i) {
            const float4 e0 = __ldg(reinterpret_cast<const float4*>(ep) + i), q0 = ldcg4(pp + 4 * i);
            g[u][4 * i] = fmaxf(e0.x + q0.x, 0.f); g[u][4 * i + 1] = fmaxf(e0.y + q0.y, 0.f);
            g[u][4 * i + 2] = fmaxf(e0.z + q0.z, 0.f); g[u][4 * i + 3] = fmaxf(e0.w + q0.w, 0.f);
          }
        } else {
#pragma unroll
          for (int i = 0; i < KJ; ++i) g[u][i] = 0.f;
        }
      }
      if (cta == 0 && tid == 0) { float sink = 0.f;
#pragma unroll
        for (int u = 0; u < kUpw; ++u) sink += g[u][0] + g[u][KJ - 1];
        if (sink == 123456.f) jprof[2] += 1;                     // forces the loads to complete before the timestamp
        const long long t1 = clock64(); jprof[0] += t1 - tk; tk = t1; }
      // after the reduce-scatter lane l holds ONE total: row (l >> 2) of the block for utterance (l & 3)
      float best = -INFINITY;
      int bi = 0x7fffffff;
      for (int rb = 0; rb < nj; rb += kRowBlk) {
        float acc[kUpw * kRowBlk];
#pragma unroll
        for (int r = 0; r < kRowBlk; ++r) {
          float a[kUpw];
#pragma unroll
          for (int u = 0; u < kUpw; ++u) a[u] = 0.f;
          if (rb + r < nj) {                                   // warp-uniform
            const uint2* wr = reinterpret_cast<const uint2*>(s_wout + static_cast<size_t>(rb + r) * Hj + lane * KJ);
#pragma unroll
            for (int i = 0; i < KJ / 4; ++i) {
              const uint2 w = wr[i];
              const float w0 = bf16_lo(w.x), w1 = bf16_hi(w.x), w2 = bf16_lo(w.y), w3 = bf16_hi(w.y);
#pragma unroll
              for (int u = 0; u < kUpw; ++u) {
                a[u] = fmaf(w0, g[u][4 * i], a[u]); a[u] = fmaf(w1, g[u][4 * i + 1], a[u]);
                a[u] = fmaf(w2, g[u][4 * i + 2], a[u]); a[u] = fmaf(w3, g[u][4 * i + 3], a[u]);
              }
            }
          }
#pragma unroll
          for (int u = 0; u < kUpw; ++u) acc[kUpw * r + u] = a[u];
        }
        warp_reduce_scatter<kUpw * kRowBlk>(acc, lane);
        const int r = rb + (lane >> 2);
        if (r < nj) {
          const int row = j0 + r;
          const float v = acc[0] + __ldg(p.b_out + row);
          if (v > best) { best = v; bi = row; }                 // blocks visited in increasing row order
        }
      }
      if (cta == 0 && tid == 0) { if (best == 123456.f) jprof[2] += 1; const long long t1 = clock64(); jprof[1] += t1 - tk; tk = t1; }
#pragma unroll
      for (int o = 16; o >= kUpw; o >>= 1) {                   // lanes with equal (lane & 3) = same utterance
        const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
#pragma unroll
      for (int u = 0; u < kUpw; ++u)
        if (lane == u && ok[u]) {
          // grid-wide argmax without a gather: order-preserving float bits in the high word, complemented row
          // index in the low word (ties -> lower row), one fire-and-forget 64-bit max per (CTA, utterance)
          unsigned int fb = __float_as_uint(best);
          fb = (fb & 0x80000000u) ? ~fb : (fb | 0x80000000u);
          const unsigned long long packed = (static_cast<unsigned long long>(fb) << 32) | (0xffffffffu - static_cast<unsigned int>(bi));
          asm volatile("red.relaxed.gpu.global.max.u64 [%0], %1;" ::"l"(p.best + static_cast<size_t>(iter % 3) * B + bu[u]), "l"(packed) : "memory");
        }
    }
    tick(0, tk);
    grid_barrier(p.counter, target, G);
    tick(1, tk);
    // ---- token per active utterance from the packed maxima; advance the (t, symbols) state (identically in every CTA)
    for (int b = tid; b < B; b += kBdThreads) {
      if (!(s_t[b] < s_len[b])) { s_tok[b] = -1; continue; }
      const unsigned long long v = __ldcg(p.best + static_cast<size_t>(iter % 3) * B + b);
      const int k = static_cast<int>(0xffffffffu - static_cast<unsigned int>(v & 0xffffffffull));
      if (k == blank) { s_t[b] += 1; s_sym[b] = 0; s_tok[b] = -1; }
      else {
        const int n = s_n[b];
        if (cta == 0 && n < p.U_max) {
          p.tokens[static_cast<size_t>(b) * p.U_max + n] = k;
          p.frames[static_cast<size_t>(b) * p.U_max + n] = s_t[b];
        }
        s_n[b] = n + 1;
        s_tok[b] = k;
        if (++s_sym[b] >= p.max_symbols) { s_t[b] += 1; s_sym[b] = 0; }
      }
    }
    // the slot that iteration iter+2 will use was last read two barriers ago: clear it now (visible through the
    // next barrier, which precedes that iteration's red.max)
    if (cta == 0) for (int b = tid; b < B; b += kBdThreads) __stcg(p.best + static_cast<size_t>((iter + 2) % 3) * B + b, 0ull);
    ++iter;
    __syncthreads();
    if (tid == 0) {                                // compact list of the utterances that emitted (ordered by b)
      int ne = 0, na = 0;
      for (int b = 0; b < B; ++b) {
        if (s_tok[b] >= 0) s_emit[ne++] = b;
        if (s_t[b] < s_len[b]) ++na;
      }
      s_cnt[0] = ne; s_cnt[1] = na;
    }
    __syncthreads();
    const int n_active = s_cnt[1];
    tick(2, tk);
    prof[7] += 1;
    if (s_cnt[0] > 0) lstm_and_pred();
    if (n_active == 0) break;
  }
  if (cta == 0) for (int b = tid; b < B; b += kBdThreads) p.n_tok[b] = s_n[b];
  if (cta == 0 && tid == 0) { for (int i = 0; i < 8; ++i) p.prof[i] = prof[i]; p.prof[8] = jprof[0]; p.prof[9] = jprof[1]; p.prof[10] = jprof[2]; }
}

size_t rnnt_batched_workspace_bytes(int B, int Hj, int Hp, int num_sms) {
  return static_cast<size_t>(2) * B * Hp * 4 + static_cast<size_t>(B) * Hj * 4 + static_cast<size_t>(3) * B * 8 + 256;
}

template <int KJ, int KP>
static cudaError_t launch_bd(BatchedDev p, int grid, size_t smem, cudaStream_t stream) {
  cudaError_t e = cudaFuncSetAttribute(rnnt_greedy_batched_kernel<KJ, KP>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rnnt_greedy_batched_kernel<KJ, KP>, kBdThreads, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {&p};
  return cudaLaunchCooperativeKernel(reinterpret_cast<void*>(rnnt_greedy_batched_kernel<KJ, KP>), dim3(grid), dim3(kBdThreads), args, smem, stream);
}

cudaError_t launch_rnnt_greedy_batched(const DecodeArgs& a, void* workspace, int num_sms, cudaStream_t stream) {
  if (a.Hj % 128 || a.Hp % 128 || a.B <= 0) return cudaErrorInvalidValue;
  const int G = num_sms;
  BatchedDev p;
  p.enc_proj = a.enc_proj; p.enc_len = a.enc_len;
  p.w_out = static_cast<const __nv_bfloat16*>(a.w_out); p.b_out = a.b_out; p.embed = a.embed;
  p.w_lstm = static_cast<const __nv_bfloat16*>(a.w_lstm); p.b_lstm = a.b_lstm;
  p.w_pred = static_cast<const __nv_bfloat16*>(a.w_pred); p.b_pred = a.b_pred;
  p.tokens = a.tokens; p.frames = a.frames; p.n_tok = a.n_tok;
  char* ws = static_cast<char*>(workspace);
  p.hbuf = reinterpret_cast<float*>(ws); ws += static_cast<size_t>(2) * a.B * a.Hp * 4;
  p.ppbuf = reinterpret_cast<float*>(ws); ws += static_cast<size_t>(a.B) * a.Hj * 4;
  p.best = reinterpret_cast<unsigned long long*>(ws); ws += static_cast<size_t>(3) * a.B * 8;
  p.counter = reinterpret_cast<unsigned int*>(ws);
  p.prof = reinterpret_cast<long long*>(ws + 64);
  p.B = a.B; p.T_max = a.T_max; p.Hj = a.Hj; p.Hp = a.Hp; p.V = a.V; p.U_max = a.U_max; p.max_symbols = a.max_symbols;
  p.rows_j = (a.V + 1 + G - 1) / G;
  p.units = (a.Hp + G - 1) / G;
  p.rows_p = (a.Hj + G - 1) / G;
  if (4 * p.units > 32 || G > 256) return cudaErrorInvalidValue;
  cudaError_t e = cudaMemsetAsync(p.hbuf, 0, static_cast<size_t>(2) * a.B * a.Hp * 4 + static_cast<size_t>(a.B) * a.Hj * 4 + static_cast<size_t>(3) * a.B * 8 + 256, stream);
  if (e != cudaSuccess) return e;
  const size_t smem = (static_cast<size_t>(p.rows_j) * a.Hj + static_cast<size_t>(4 * p.units) * 2 * a.Hp + static_cast<size_t>(p.rows_p) * a.Hp) * 2 +
                      static_cast<size_t>(a.B) * p.units * 4 + static_cast<size_t>(a.B) * 7 * 4 + 64;
  if (smem > 220 * 1024) return cudaErrorInvalidValue;
  if (a.Hj == 640 && a.Hp == 640) return launch_bd<20, 20>(p, G, smem, stream);
  if (a.Hj == 128 && a.Hp == 128) return launch_bd<4, 4>(p, G, smem, stream);
  return cudaErrorInvalidValue;
}

}  // namespace rs
