// Host side of transcribe()'s padding (pkg/nemo-asr/src/audio.py:70-83) for a whole batch: the waveforms of B utterances,
// wherever the caller's arrays live, are laid out as the rows [zeros(pad) | samples | zeros to L] of the pinned matrix the
// engine copies to the GPU.  Plain memcpy / one multiply per sample, split over a few threads by rows; no CUDA calls.
// It exists because the same loop in the host language (one numpy slice assignment per utterance) costs ~0.5 ms per 30 s
// clip under the interpreter lock -- as much as the GPU spends on the clip -- and with one engine per GPU in one process
// (nemo/asr/multi_gpu.py) that serialised staging, not the GPUs, set the throughput.
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/rs_engine.h"

namespace {

void stage_range(void* dst, int64_t L, const void* const* src, const int64_t* n, const int32_t* src_is_pcm16, int dst_is_pcm16,
                 int64_t pad, int r0, int r1) {
  for (int r = r0; r < r1; ++r) {
    const int64_t len = n[r];
    if (dst_is_pcm16) {
      int16_t* row = static_cast<int16_t*>(dst) + static_cast<int64_t>(r) * L;
      std::memset(row, 0, static_cast<size_t>(pad) * 2);
      std::memcpy(row + pad, src[r], static_cast<size_t>(len) * 2);
      std::memset(row + pad + len, 0, static_cast<size_t>(L - pad - len) * 2);
    } else {
      float* row = static_cast<float*>(dst) + static_cast<int64_t>(r) * L;
      std::memset(row, 0, static_cast<size_t>(pad) * 4);
      if (src_is_pcm16 != nullptr && src_is_pcm16[r]) {
        const int16_t* s = static_cast<const int16_t*>(src[r]);
        float* d = row + pad;
        for (int64_t i = 0; i < len; ++i) d[i] = static_cast<float>(s[i]) * (1.0f / 32768.0f);   // what a file decoder returns
      } else {
        std::memcpy(row + pad, src[r], static_cast<size_t>(len) * 4);
      }
      std::memset(row + pad + len, 0, static_cast<size_t>(L - pad - len) * 4);
    }
  }
}

}  // namespace

extern "C" int rs_stage_rows(void* dst, int64_t L, const void* const* src, const int64_t* n, const int32_t* src_is_pcm16,
                             int dst_is_pcm16, int B, int64_t pad, int threads) {
  if (dst == nullptr || src == nullptr || n == nullptr || B < 0 || pad < 0 || L < 0) return RS_ERR_INVALID_ARG;
  for (int r = 0; r < B; ++r) {
    if (n[r] < 0 || n[r] + 2 * pad > L || (n[r] > 0 && src[r] == nullptr)) return RS_ERR_INVALID_ARG;
    if (dst_is_pcm16 && (src_is_pcm16 == nullptr || !src_is_pcm16[r])) return RS_ERR_INVALID_ARG;   // int16 rows only from int16 sources
  }
  if (threads < 1) threads = 1;
  if (threads > B) threads = B;
  if (threads <= 1) {
    stage_range(dst, L, src, n, src_is_pcm16, dst_is_pcm16, pad, 0, B);
    return RS_OK;
  }
  std::vector<std::thread> pool;
  pool.reserve(threads - 1);
  for (int t = 1; t < threads; ++t)
    pool.emplace_back(stage_range, dst, L, src, n, src_is_pcm16, dst_is_pcm16, pad, static_cast<int>(static_cast<int64_t>(B) * t / threads),
                      static_cast<int>(static_cast<int64_t>(B) * (t + 1) / threads));
  stage_range(dst, L, src, n, src_is_pcm16, dst_is_pcm16, pad, 0, static_cast<int>(static_cast<int64_t>(B) / threads));
  for (auto& th : pool) th.join();
  return RS_OK;
}
