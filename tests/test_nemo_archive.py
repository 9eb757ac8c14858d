"""load_model's checkpoint ingestion (reference: EncDecRNNTBPEModel.from_pretrained at pkg/nemo-asr/src/transcribe.py:26-28):
a ``.nemo`` archive is a tar of model_config.yaml + model_weights.ckpt + a SentencePiece model.  The real
reazonspeech-nemo-v2 archive is not available offline, so these tests assemble archives with NeMo's layout and key names
 and check that the loader maps them one-to-one, ignores what inference does not use, and REFUSES
configurations the kernels do not implement instead of transcribing garbage."""
import copy
import io
import tarfile

import pytest
import torch
import yaml

from reazonspeech_b200.config import ModelConfig
from reazonspeech_b200.weights import load_nemo_archive, random_state_dict


def nemo_yaml(cfg: ModelConfig) -> dict:
    """A model_config.yaml as NeMo writes it for a FastConformer-Transducer-BPE model (only the keys that matter here,
    plus a few training-only ones that must be ignored)."""
    return {
        "sample_rate": cfg.sample_rate,
        "preprocessor": {"_target_": "nemo.collections.asr.modules.AudioToMelSpectrogramPreprocessor", "sample_rate": cfg.sample_rate,
                         "normalize": "per_feature", "window_size": 0.025, "window_stride": 0.01, "window": "hann",
                         "features": cfg.n_mels, "n_fft": cfg.n_fft, "log": True, "frame_splicing": 1, "dither": 1e-5,
                         "pad_to": 0, "pad_value": 0.0, "preemph": 0.97, "log_zero_guard_type": "add",
                         "log_zero_guard_value": 2.0 ** -24, "mag_power": 2.0, "lowfreq": 0, "highfreq": None},
        "spec_augment": {"freq_masks": 2, "time_masks": 10},
        "encoder": {"_target_": "nemo.collections.asr.modules.ConformerEncoder", "feat_in": cfg.n_mels, "n_layers": cfg.n_layers,
                    "d_model": cfg.d_model, "subsampling": "dw_striding", "subsampling_factor": 8,
                    "subsampling_conv_channels": cfg.sub_channels, "causal_downsampling": False, "ff_expansion_factor": 4,
                    "self_attention_model": "rel_pos_local_attn", "n_heads": cfg.n_heads,
                    "att_context_size": [cfg.att_left, cfg.att_right], "global_tokens": cfg.global_tokens,
                    "global_tokens_spacing": 1, "global_attn_separate": False, "xscaling": True, "untie_biases": True,
                    "pos_emb_max_len": 5000, "conv_kernel_size": 9, "conv_norm_type": "batch_norm", "conv_context_size": None,
                    "dropout": 0.1, "dropout_att": 0.1},
        "decoder": {"_target_": "nemo.collections.asr.modules.RNNTDecoder", "normalization_mode": None, "random_state_sampling": False,
                    "blank_as_pad": True, "vocab_size": cfg.vocab_size,
                    "prednet": {"pred_hidden": cfg.pred_hidden, "pred_rnn_layers": 1, "t_max": None, "dropout": 0.2}},
        "joint": {"_target_": "nemo.collections.asr.modules.RNNTJoint", "num_classes": cfg.vocab_size, "fuse_loss_wer": True,
                  "jointnet": {"joint_hidden": cfg.joint_hidden, "activation": "relu", "dropout": 0.2,
                               "encoder_hidden": cfg.d_model, "pred_hidden": cfg.pred_hidden}},
        "decoding": {"strategy": "greedy_batch", "greedy": {"max_symbols": 10}, "beam": {"beam_size": 2}},
        "optim": {"name": "adamw", "lr": 1e-3},
    }


def write_nemo(path, config: dict, state: dict, tokenizer: bytes = None, prefix: str = "./", mode: str = "w"):
    def add(tar, name, data: bytes):
        info = tarfile.TarInfo(prefix + name)
        info.size = len(data)
        tar.addfile(info, io.BytesIO(data))
    with tarfile.open(path, mode) as tar:
        add(tar, "model_config.yaml", yaml.safe_dump(config).encode())
        buf = io.BytesIO()
        torch.save(state, buf)
        add(tar, "model_weights.ckpt", buf.getvalue())
        if tokenizer is not None:
            add(tar, "a1b2c3_tokenizer.model", tokenizer)
            add(tar, "d4e5f6_vocab.txt", b"ignored\\n")


@pytest.fixture(scope="module")
def tiny():
    cfg = ModelConfig.tiny()
    return cfg, random_state_dict(cfg, seed=0)


def checkpoint_state(sd):
    """What a real checkpoint carries besides the tensors the engine uses: buffers and counters."""
    extra = dict(sd)
    extra["preprocessor.featurizer.window"] = torch.hann_window(400)
    extra["preprocessor.featurizer.fb"] = torch.zeros(1, 80, 257)
    extra["encoder.layers.0.conv.batch_norm.num_batches_tracked"] = torch.tensor(12345)
    extra["encoder.pos_enc.pe"] = torch.zeros(1, 33, 256)
    return {k: (v.half() if k.endswith("linear1.weight") else v) for k, v in extra.items()}       # mixed precision on disk is fine


@pytest.mark.parametrize("mode,prefix", [("w", "./"), ("w:gz", ""), ("w", "model/")])
def test_archive_round_trip(tmp_path, tiny, mode, prefix):
    cfg, sd = tiny
    path = tmp_path / "m.nemo"
    write_nemo(path, nemo_yaml(cfg), checkpoint_state(sd), tokenizer=b"spm-bytes", prefix=prefix, mode=mode)
    got_cfg, got_sd, tok = load_nemo_archive(str(path))
    assert got_cfg == cfg
    assert tok == b"spm-bytes"
    assert set(got_sd) == set(sd)                                   # buffers and counters dropped
    for k, v in sd.items():
        assert got_sd[k].dtype == torch.float32
        ref = v.half().float() if k.endswith("linear1.weight") else v
        assert torch.equal(got_sd[k], ref), k


def test_archive_without_tokenizer_and_with_defaults(tmp_path, tiny):
    cfg, sd = tiny
    y = nemo_yaml(cfg)
    for k in ("window", "normalize", "log", "mag_power", "lowfreq", "highfreq", "log_zero_guard_type", "frame_splicing"):
        del y["preprocessor"][k]                                     # NeMo's constructor defaults apply
    del y["decoding"]
    path = tmp_path / "m.nemo"
    write_nemo(path, y, sd)
    got_cfg, _, tok = load_nemo_archive(str(path))
    assert got_cfg == cfg and tok is None


def test_archive_errors(tmp_path, tiny):
    cfg, sd = tiny
    bad = dict(sd); del bad["encoder.layers.1.self_attn.pos_bias_u"]
    write_nemo(tmp_path / "a.nemo", nemo_yaml(cfg), bad)
    with pytest.raises(ValueError, match="lacks 1 tensors"):
        load_nemo_archive(str(tmp_path / "a.nemo"))
    bad = dict(sd); bad["joint.pred.weight"] = torch.zeros(3, 3)
    write_nemo(tmp_path / "b.nemo", nemo_yaml(cfg), bad)
    with pytest.raises(ValueError, match="joint.pred.weight has shape"):
        load_nemo_archive(str(tmp_path / "b.nemo"))
    with tarfile.open(tmp_path / "c.nemo", "w") as tar:
        info = tarfile.TarInfo("./readme.txt"); info.size = 2
        tar.addfile(info, io.BytesIO(b"hi"))
    with pytest.raises(ValueError, match="not a .nemo archive"):
        load_nemo_archive(str(tmp_path / "c.nemo"))


@pytest.mark.parametrize("block,key,value", [
    ("preprocessor", "normalize", "all_features"), ("preprocessor", "window", "hamming"), ("preprocessor", "log", False),
    ("preprocessor", "mag_power", 1.0), ("preprocessor", "lowfreq", 20), ("preprocessor", "highfreq", 7600),
    ("preprocessor", "log_zero_guard_type", "clamp"), ("preprocessor", "frame_splicing", 3),
    ("encoder", "self_attention_model", "rel_pos"), ("encoder", "subsampling", "striding"), ("encoder", "subsampling_factor", 4),
    ("encoder", "conv_norm_type", "layer_norm"), ("encoder", "untie_biases", False), ("encoder", "global_attn_separate", True),
    ("encoder", "global_tokens_spacing", 4), ("encoder", "conv_context_size", [8, 0]), ("encoder", "att_context_size", [-1, -1]),
    ("encoder", "feat_in", 128), ("decoder", "blank_as_pad", False), ("joint", "num_classes", 99),
])
def test_unimplemented_settings_are_refused(tiny, block, key, value):
    cfg, _ = tiny
    y = copy.deepcopy(nemo_yaml(cfg))
    y[block][key] = value
    with pytest.raises(ValueError, match="model_config.yaml"):
        ModelConfig.from_nemo_yaml(y)


def test_nested_settings_are_refused_and_variants_are_understood(tiny):
    cfg, _ = tiny
    y = copy.deepcopy(nemo_yaml(cfg)); y["decoder"]["prednet"]["pred_rnn_layers"] = 2
    with pytest.raises(ValueError, match="pred_rnn_layers"):
        ModelConfig.from_nemo_yaml(y)
    y = copy.deepcopy(nemo_yaml(cfg)); y["joint"]["jointnet"]["activation"] = "tanh"
    with pytest.raises(ValueError, match="activation"):
        ModelConfig.from_nemo_yaml(y)
    y = copy.deepcopy(nemo_yaml(cfg)); y["encoder"]["att_context_size"] = [[16, 16], [16, 4]]      # multi-lookahead list: first entry
    assert ModelConfig.from_nemo_yaml(y) == cfg
    y = copy.deepcopy(nemo_yaml(cfg)); y["preprocessor"]["highfreq"] = 8000; y["preprocessor"]["log_zero_guard_value"] = "tiny"
    got = ModelConfig.from_nemo_yaml(y)
    assert got.log_zero_guard == pytest.approx(1.1754943508222875e-38) and got.replace(log_zero_guard=cfg.log_zero_guard) == cfg
    full = ModelConfig.from_nemo_yaml(nemo_yaml(ModelConfig()))
    assert full == ModelConfig()                                    # the 619 M defaults of SURVEY.md App. A.1


def test_unsupported_arithmetic_settings_are_rejected_and_the_strategy_is_reported():
    """Settings that change the arithmetic but used to load silently: exact_pad, causal_downsampling,
    chunked attention context, encoder reduction, a joint built without dropout (its output layer is then joint_net.1);
    and the checkpoint's decoding strategy is surfaced (the shipped model asks for ALSD beam search)."""
    import copy
    cfg = ModelConfig.tiny()
    base = nemo_yaml(cfg)
    for block, key, value in (("preprocessor", "exact_pad", True), ("encoder", "causal_downsampling", True),
                              ("encoder", "att_context_style", "chunked_limited"), ("encoder", "reduction", "pooling")):
        y = copy.deepcopy(base)
        y[block][key] = value
        with pytest.raises(ValueError, match=key):
            ModelConfig.from_nemo_yaml(y)
    y = copy.deepcopy(base)
    y["joint"]["jointnet"]["dropout"] = 0.0
    with pytest.raises(ValueError, match="joint_net.1"):
        ModelConfig.from_nemo_yaml(y)
    y = copy.deepcopy(base)
    y["decoding"] = {"strategy": "beam", "beam": {"search_type": "alsd", "beam_size": 4}, "greedy": {"max_symbols": 10}}
    assert ModelConfig.from_nemo_yaml(y).checkpoint_decoding == "alsd"
    assert ModelConfig.from_nemo_yaml(base).checkpoint_decoding.startswith("greedy")
