"""``load_model`` / ``transcribe`` with the reference's signatures (pkg/nemo-asr/src/transcribe.py:9-60)
on top of the B200 engine, plus the batched ``transcribe_batch`` the reference lacks
(its evaluators leave ``_evaluate_batch`` unimplemented, pkg/evaluation/examples/rs-nemo/eval.py:31-32).

The object returned by ``load_model`` is a duck-typed stand-in for NeMo's EncDecRNNTBPEModel at
exactly the three points the reference touches it:
``model.transcribe(list_of_tensors, batch_size=..., return_hypotheses=True, verbose=...)``,
``hyp.y_sequence`` / ``hyp.timestamp`` and ``model.tokenizer.ids_to_text``; the reference's own
transcribe()/decode_hypothesis() therefore run unmodified on it."""
from __future__ import annotations

import glob
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from ...config import ModelConfig
from ...engine import Engine
from ...tokenizer import PieceTableTokenizer, SentencePieceTokenizer, synthetic_pieces
from ...weights import load_nemo_archive, random_state_dict
from .audio import SAMPLERATE, norm_audio, pad_audio
from .decode import PAD_SECONDS, build_result, decode_hypothesis
from .interface import AudioData, TranscribeConfig, TranscribeResult

HF_REPO = "reazon-research/reazonspeech-nemo-v2"
ENV_CHECKPOINT = "REAZONSPEECH_NEMO_CHECKPOINT"
ENV_SYNTHETIC = "REAZONSPEECH_B200_SYNTHETIC"


@dataclass
class Hypothesis:
    """The two NeMo Hypothesis fields the reference reads (decode.py:40,44), ALSD-shaped:
    y_sequence = [blank, tok_0, ...]; timestamp[i] = frame_i + i + 1 so that decode.py:48's
    ``step - idx - 1`` recovers the emitting encoder frame."""
    y_sequence: torch.Tensor
    timestamp: List[int]
    score: float = 0.0

    @staticmethod
    def from_greedy(tokens: Sequence[int], frames: Sequence[int], blank: int) -> "Hypothesis":
        y = torch.tensor([blank, *[int(t) for t in tokens]], dtype=torch.long)
        return Hypothesis(y, [int(f) + i + 1 for i, f in enumerate(frames)])


class HostStaging:
    """Grow-only host buffers of one in-flight batch, reused across calls and pinned when a CUDA device is there:
    allocating (cudaHostAlloc) and zero-filling a fresh 64 MB pinned batch costs more than the batch's GPU time."""

    def __init__(self, pin: bool):
        self.pin = pin
        self._wav = self._lens = self._tok = self._frm = self._ntok = None

    def _grown(self, buf, numel: int, dtype):
        if buf is None or buf.numel() < numel:
            buf = torch.empty(int(numel * 1.25) + 16, dtype=dtype)
            if self.pin:
                buf = buf.pin_memory()
        return buf

    def stage(self, waves: Sequence[np.ndarray], pad: int):
        """Rows ``[zeros(pad) | wave | zeros(pad) | zeros to L]`` (pad_audio, audio.py:70-83, written in place) -> (wav [B, L], lens [B]).

        A batch whose waveforms are ALL 16-bit PCM (int16 arrays, e.g. ``audio_from_path(..., pcm16=True)``) is staged as
        int16 and scaled by 2^-15 on the device (rs_transcribe_batch_pcm16): half the pinned memory and PCIe bytes.  A mixed
        batch is staged as float32, int16 members converted the way a file decoder would (sample / 32768)."""
        B = len(waves)
        L = max(len(w) for w in waves) + 2 * pad
        L = (L + 3) & ~3                                # rows start 8 / 16-byte aligned: the kernel's vectorised staging load
        pcm = all(w.dtype == np.int16 for w in waves)
        if pcm:
            self._wav16 = self._grown(getattr(self, "_wav16", None), B * L, torch.int16)
            wav = self._wav16[: B * L].view(B, L)
        else:
            self._wav = self._grown(self._wav, B * L, torch.float32)
            wav = self._wav[: B * L].view(B, L)
        self._lens = self._grown(self._lens, B, torch.int32)
        lens = self._lens[:B]
        # the copies run in the library (rs_stage_rows: memcpy split over a few threads, interpreter lock released): the
        # same loop as numpy slice assignments cost ~0.5 ms per 30 s clip, as much as the GPU spends on it
        import ctypes as C
        from ...engine import load_library
        srcs = [np.ascontiguousarray(w if w.dtype in (np.int16, np.float32) else w.astype(np.float32)) for w in waves]
        ptr = (C.c_void_p * B)(*[a.ctypes.data for a in srcs])
        n = (C.c_int64 * B)(*[a.shape[0] for a in srcs])
        is16 = (C.c_int32 * B)(*[int(a.dtype == np.int16) for a in srcs])
        rc = load_library().rs_stage_rows(wav.data_ptr(), L, ptr, n, is16, int(pcm), B, pad, min(4, B))
        if rc != 0:
            raise RuntimeError(f"rs_stage_rows failed ({rc})")
        lens.copy_(torch.tensor([a.shape[0] + 2 * pad for a in srcs], dtype=torch.int32))
        return wav, lens

    def outputs(self, B: int, U: int):
        self._tok = self._grown(self._tok, B * U, torch.int32)
        self._frm = self._grown(self._frm, B * U, torch.int32)
        self._ntok = self._grown(self._ntok, B, torch.int32)
        return self._tok[: B * U].view(B, U), self._frm[: B * U].view(B, U), self._ntok[:B]


class B200RnntModel:
    """Engine + tokenizer behind NeMo's model surface."""

    def __init__(self, engine: Engine, tokenizer, max_batch: int = 64, decoding: str = "greedy", beam_size: int = 4):
        self.engine = engine
        self.cfg = engine.cfg
        self.tokenizer = tokenizer
        self.max_batch = max_batch
        self.decoding, self.beam_size = decoding, beam_size       # "greedy" (north_star's parity target) or "alsd" (the checkpoint's default)
        pin = torch.cuda.is_available()
        self._staging = (HostStaging(pin), HostStaging(pin))      # double buffer: stage batch k+1 while batch k runs

    # -- token-level batched path
    def iter_token_batches(self, waveforms: Sequence[np.ndarray], pad: int = 0):
        """16 kHz mono waveforms (each gets ``pad`` zero samples on both sides) -> yields ``(indices, [(tokens, frames)])``
        batch by batch.

        Utterances are sorted by length and cut into batches of at most ``max_batch`` so padding waste stays small.
        The engine call of batch k runs on a worker thread (ctypes drops the GIL) while this thread stages batch k+1
        into the other staging set and the caller post-processes batch k-1: on a long list the host work hides behind
        the GPU.  One engine call is in flight at a time (an engine is not re-entrant)."""
        from concurrent.futures import ThreadPoolExecutor
        if len(waveforms) == 0:
            return
        order = sorted(range(len(waveforms)), key=lambda i: len(waveforms[i]))
        batches = [order[lo:lo + self.max_batch] for lo in range(0, len(order), self.max_batch)]
        eng = self.engine
        eng.ensure_workspace(len(batches[0]), (len(waveforms[order[-1]]) + 2 * pad + 3) & ~3)   # once, on this thread (stage() rounds rows up to 4 samples)

        def run(staging, idx):
            wav, lens = staging.stage([waveforms[i] for i in idx], pad)
            out = staging.outputs(len(idx), eng.u_max(wav.shape[1]))
            return wav, lens, out

        def collect(done, idx):
            tokens, frames, ntok = done
            counts = ntok.tolist()
            return idx, [(tokens[r, :n].tolist(), frames[r, :n].tolist()) for r, n in enumerate(counts[: len(idx)])]

        if len(batches) == 1:                                         # nothing to overlap: skip the thread hand-off (one-clip calls)
            wav, lens, out = run(self._staging[0], batches[0])
            yield collect(eng.transcribe_host(wav, lens, out[0].shape[1], out), batches[0])
            return
        with ThreadPoolExecutor(max_workers=1) as pool:
            in_flight = None
            for k, idx in enumerate(batches):
                wav, lens, out = run(self._staging[k & 1], idx)
                nxt = (pool.submit(eng.transcribe_host, wav, lens, out[0].shape[1], out), idx)
                if in_flight is not None:
                    yield collect(in_flight[0].result(), in_flight[1])
                in_flight = nxt
            yield collect(in_flight[0].result(), in_flight[1])

    def transcribe_tokens(self, waveforms: Sequence[np.ndarray], pad: int = 0):
        """-> [(tokens, frames)] in input order."""
        results = [None] * len(waveforms)
        for idx, items in self.iter_token_batches(waveforms, pad):
            for i, item in zip(idx, items):
                results[i] = item
        return results

    def iter_token_batches_raw(self, waves: Sequence[np.ndarray], samplerate: int, pad: int = 0):
        """Like ``iter_token_batches`` for audio that still needs ``norm_audio`` (pkg/nemo-asr/src/audio.py:54-68): waveforms
        at ``samplerate`` (any rate), mono [n] or channels-first [c, n] with the SAME channel count, float or int16 PCM.  They
        are staged as they are (pinned), copied to the GPU and resampled / down-mixed / padded there (rs_resample_mono)
        straight into the buffer the engine transcribes from: the host never touches a sample arithmetically.  (scipy's
        resample_poly, which this kernel restates, costs the host ~10 ms per 30 s 48 kHz clip -- more than the whole engine.)"""
        if len(waves) == 0:
            return
        eng = self.engine
        waves = [w if w.ndim == 2 else w[None] for w in waves]
        C = waves[0].shape[0]
        if any(w.shape[0] != C for w in waves):
            raise ValueError("iter_token_batches_raw: all waveforms of a call must have the same number of channels")
        pcm = all(w.dtype == np.int16 for w in waves)
        order = sorted(range(len(waves)), key=lambda i: waves[i].shape[1])
        for lo in range(0, len(order), self.max_batch):
            idx = order[lo:lo + self.max_batch]
            L = (max(waves[i].shape[1] for i in idx) + 3) & ~3
            raw = torch.zeros(len(idx), C, L, dtype=torch.int16 if pcm else torch.float32)
            if torch.cuda.is_available():
                raw = raw.pin_memory()
            rows = raw.numpy()
            for r, i in enumerate(idx):
                w = waves[i]
                rows[r, :, : w.shape[1]] = w if (pcm or w.dtype != np.int16) else w.astype(np.float32) * np.float32(1.0 / 32768.0)
            lens = torch.tensor([waves[i].shape[1] for i in idx], dtype=torch.int32)
            with torch.cuda.device(eng.device):
                wav, wl = eng.resample_mono(raw.to(eng.device, non_blocking=True), lens.to(eng.device, non_blocking=True), samplerate, pad)
                tokens, frames, ntok = eng.transcribe_device(wav, wl)
                tokens, frames, counts = tokens.cpu(), frames.cpu(), ntok.cpu().tolist()
            yield idx, [(tokens[r, :n].tolist(), frames[r, :n].tolist()) for r, n in enumerate(counts)]

    # -- NeMo's call shape (transcribe.py:48-53): already padded tensors
    def transcribe_alsd(self, waveforms: Sequence[np.ndarray], pad: int = 0) -> List[Hypothesis]:
        """ALSD beam search (NeMo's align_length_sync_decoding, the strategy reazonspeech-nemo-v2 ships with): hypotheses exactly
        as NeMo hands them to the reference's decode.py -- y_sequence with the leading blank, timestamp = alignment steps."""
        eng = self.engine
        out: List[Optional[Hypothesis]] = [None] * len(waveforms)
        order = sorted(range(len(waveforms)), key=lambda i: len(waveforms[i]))
        for lo in range(0, len(order), self.max_batch):
            idx = order[lo:lo + self.max_batch]
            wav, lens = self._staging[0].stage([waveforms[i] for i in idx], pad)
            with torch.cuda.device(eng.device):
                x = wav.to(eng.device, non_blocking=True)
                if x.dtype == torch.int16:
                    x = x.to(torch.float32) * (1.0 / 32768.0)
                mel, mel_len = eng.log_mel(x, lens.to(eng.device))
                enc, enc_len = eng.encode(mel, mel_len)
                y, steps, n, score = [a.cpu() for a in eng.alsd(enc, enc_len, beam=self.beam_size)]
            for r, i in enumerate(idx):
                k = int(n[r])
                out[i] = Hypothesis(y[r, : k + 1].to(torch.long), steps[r, :k].tolist(), float(score[r]))
        return out

    def transcribe(self, audio, batch_size: int = 1, return_hypotheses: bool = True, verbose: bool = True, **_):
        waves = [a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a) for a in audio]
        if self.decoding == "alsd":
            out = self.transcribe_alsd(waves)
        else:
            out = [Hypothesis.from_greedy(t, f, self.cfg.blank) for t, f in self.transcribe_tokens(waves)]
        if return_hypotheses:
            return out
        return [self.tokenizer.ids_to_text(h.y_sequence.tolist()[1:]) for h in out]


def _find_checkpoint() -> Optional[str]:
    env = os.environ.get(ENV_CHECKPOINT)
    if env:
        return env
    hub = os.path.expanduser(os.environ.get("HF_HOME", "~/.cache/huggingface"))
    hits = glob.glob(os.path.join(hub, "hub", "models--" + HF_REPO.replace("/", "--"), "snapshots", "*", "*.nemo"))
    return sorted(hits)[-1] if hits else None


def load_model(device=None, *, checkpoint: Optional[str] = None, synthetic: Optional[bool] = None,
               config: Optional[ModelConfig] = None, seed: int = 0, max_batch: int = 64, devices: Optional[Sequence] = None,
               decoding: str = "greedy", beam_size: int = 4):
    """Load the ReazonSpeech FastConformer-RNNT onto a B200.

    ``device``: None / "cuda" / "cuda:N" as in the reference (transcribe.py:9-22, eval.py:26).
    "cpu" raises: this engine has no CPU path.  ``decoding``: "greedy" (default: BASELINE.json's parity target) or "alsd", NeMo's
    align_length_sync_decoding beam search with ``beam_size`` -- the strategy the shipped checkpoint decodes with by default
    (pkg/nemo-asr/src/decode.py:29); it runs on the GPU too (csrc/decode_alsd.cu).  ``devices`` (e.g. ``range(8)`` or ``["cuda:0", "cuda:1"]``) loads one
    replica per listed GPU into THIS process and returns a model that deals every call's utterances across them
    (``multi_gpu.MultiGpuRnntModel``; same surface, results in input order).  Weights come from ``checkpoint`` (a .nemo file),
    $REAZONSPEECH_NEMO_CHECKPOINT or the local Hugging Face cache of reazonspeech-nemo-v2.
    With ``synthetic=True`` (or $REAZONSPEECH_B200_SYNTHETIC=1) seeded random weights of the same
    architecture are used instead -- the only option offline."""
    if device is None:
        device = "cuda"
    if str(device).startswith("cpu"):
        raise RuntimeError("reazonspeech_b200: device='cpu' is not supported (hand-written sm_100a kernels only)")
    if synthetic is None:
        synthetic = os.environ.get(ENV_SYNTHETIC, "") not in ("", "0")
    path = checkpoint or (None if synthetic else _find_checkpoint())
    if path is not None:
        cfg, sd, tok = load_nemo_archive(path)
        tokenizer = SentencePieceTokenizer(tok) if tok else PieceTableTokenizer(synthetic_pieces(cfg.vocab_size))
        if decoding == "greedy" and not cfg.checkpoint_decoding.startswith("greedy"):
            import warnings                 # the reference never overrides the checkpoint's strategy (transcribe.py:26-28): say what differs
            warnings.warn(f"{path}: the checkpoint is configured for '{cfg.checkpoint_decoding}' decoding; this model decodes greedily "
                          f"(transcripts can differ from the reference's default).  Pass decoding='alsd' for NeMo's ALSD beam search.", stacklevel=2)
    elif synthetic:
        cfg = config or ModelConfig()
        sd = random_state_dict(cfg, seed)
        tokenizer = PieceTableTokenizer(synthetic_pieces(cfg.vocab_size))
    else:
        raise FileNotFoundError(
            f"no .nemo checkpoint for {HF_REPO}: pass checkpoint=..., set ${ENV_CHECKPOINT}, populate the Hugging Face "
            f"cache, or request seeded synthetic weights with synthetic=True / ${ENV_SYNTHETIC}=1")
    if decoding not in ("greedy", "alsd"):
        raise ValueError(f"decoding must be 'greedy' or 'alsd', got {decoding!r}")
    if devices is None:
        return B200RnntModel(Engine(cfg, sd, str(device), alsd=decoding == "alsd"), tokenizer, max_batch=max_batch, decoding=decoding, beam_size=beam_size)
    if decoding != "greedy":
        raise ValueError("the one-process multi-GPU model decodes greedily; load one model per device for beam search")
    names = [d if isinstance(d, str) else f"cuda:{int(d)}" for d in devices]
    if len(names) == 0 or len(set(names)) != len(names):
        raise ValueError(f"devices must name distinct GPUs, got {list(devices)!r}")
    from ...engine import pack_weights
    from .multi_gpu import MultiGpuRnntModel
    packed = pack_weights(sd, cfg)                                   # repacked once, uploaded once per device
    return MultiGpuRnntModel([B200RnntModel(Engine(cfg, None, n, packed=packed), tokenizer, max_batch=max_batch) for n in names])


def _prepare(audio: AudioData) -> np.ndarray:
    wave = pad_audio(norm_audio(audio), PAD_SECONDS).waveform
    return wave if wave.dtype == np.int16 else wave.astype(np.float32, copy=False)     # int16 = PCM, scaled on the device


def transcribe(model, audio: AudioData, config: Optional[TranscribeConfig] = None) -> TranscribeResult:
    """One utterance, same contract as the reference (transcribe.py:30-60)."""
    if config is None:
        config = TranscribeConfig()
    wave = torch.from_numpy(_prepare(audio))
    hyp = model.transcribe([wave], batch_size=1, return_hypotheses=True, verbose=config.verbose)[0]
    result = decode_hypothesis(model, hyp)
    if config.raw_hypothesis:
        result.hypothesis = hyp
    return result


def transcribe_batch(model, audios: Sequence[AudioData], config: Optional[TranscribeConfig] = None) -> List[TranscribeResult]:
    """Many utterances through one or a few engine launches; results in input order.

    Same per-utterance semantics as ``transcribe`` (norm_audio, 0.5 s of silence on both sides, greedy decode,
    decode_hypothesis); the padding is written straight into the staging buffer and the post-processing of one batch
    overlaps the engine call of the next.  A model object without ``iter_token_batches`` (e.g. a real NeMo model)
    is driven through its ``transcribe`` method instead."""
    if config is None:
        config = TranscribeConfig()
    out: List[Optional[TranscribeResult]] = [None] * len(audios)

    def finish(i, hyp):
        r = decode_hypothesis(model, hyp)
        if config.raw_hypothesis:
            r.hypothesis = hyp
        out[i] = r

    if hasattr(model, "iter_token_batches") and getattr(model, "decoding", "greedy") == "greedy":
        blank = model.cfg.blank
        pad = int(PAD_SECONDS * SAMPLERATE)
        # audio that still needs norm_audio (another rate, several channels) and is uniform in both goes to the GPU as it is:
        # resampling, down-mixing and padding run there (iter_token_batches_raw); anything else is normalised on the host
        raw_ok = (hasattr(model, "iter_token_batches_raw") and len(audios) > 0 and
                  len({(a.samplerate, np.asarray(a.waveform).ndim, np.asarray(a.waveform).shape[0] if np.asarray(a.waveform).ndim == 2 else 1)
                       for a in audios}) == 1 and
                  (audios[0].samplerate != SAMPLERATE or np.asarray(audios[0].waveform).ndim == 2))
        if raw_ok:
            batches = model.iter_token_batches_raw([np.asarray(a.waveform) for a in audios], audios[0].samplerate, pad=pad)
        else:
            batches = model.iter_token_batches([np.asarray(norm_audio(a).waveform) for a in audios], pad=pad)
        for idx, items in batches:
            for i, (tokens, frames) in zip(idx, items):
                finish(i, Hypothesis.from_greedy(tokens, frames, blank))
    else:
        tensors = [torch.from_numpy(_prepare(a)) for a in audios]
        hyps = model.transcribe(tensors, batch_size=max(len(tensors), 1), return_hypotheses=True, verbose=config.verbose)
        for i, hyp in enumerate(hyps):
            finish(i, hyp)
    return out
