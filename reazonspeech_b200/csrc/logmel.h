// Host-side interface of the fused log-mel kernel (logmel.cu); internal to librs_engine.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rs {

constexpr int kLmTileFrames = 32;        // frames per CTA
constexpr int kLmMetaInts = 272;         // reazonspeech_b200/logmel_tables.py META_INTS
constexpr int kLmMetaStart = 16;         //   first bin of (slot, lane)
constexpr int kLmMetaOut = 16 + 128;     //   filter index of (slot, lane)

// CTAs (= partial statistic rows) per utterance of at most L_max samples: the valid frames are L_max / hop
inline int logmel_tiles(int L_max, int hop) { return (L_max / hop + kLmTileFrames - 1) / kLmTileFrames; }

struct LmTables {             // device pointers into the packed weights (logmel_tables.py)
  const float* window;        // [512]  Hann(400) centred in the FFT frame
  const float* tw_b;          // [16][16][2]  W256^(t k1) at [k1][t]
  const float* tw_x;          // [8][16][2]   W512^(t + 16 k2) at [k2][t]
  const float* mel_w;         // [n_taps][16] slot-dealt filter weights x 1/4
  const int32_t* mel_meta;    // [kLmMetaInts]
  int n_taps;
};

struct LogmelArgs {
  const void* wav;            // f32 [B, L_max], or int16 PCM [B, L_max] when wav_i16 (scaled by 2^-15 on load)
  bool wav_i16;
  const int32_t* len; int B, L_max;
  float* mel;                 // [B, L_max / hop + 1, n_mels] f32: log-mel, un-normalised unless normalise_in_place
  int32_t* mel_len;           // [B] valid frames
  float* partials;            // [B, logmel_tiles, n_mels, 2] f32 workspace
  float* stats;               // [B, n_mels, 2] f32: (mean, 1 / (std + eps)) per feature
  unsigned int* tickets;      // [B], zero before the first launch (the kernel leaves them zero)
  LmTables tb;
  int n_mels, hop, n_fft, win;
  float preemph, guard, eps;
  bool normalise_in_place;    // rs_logmel: also run the normalisation pass over `mel` (NeMo's output tensor)
};
cudaError_t launch_logmel_fused(const LogmelArgs& a, cudaStream_t stream);

}  // namespace rs
