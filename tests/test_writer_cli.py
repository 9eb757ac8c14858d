"""Writers and CLI of the drop-in package against the reference's own writer.py outputs
(tests/golden/writer_cases.json, made by tests/golden/make_writer_golden.py) and the CLI contract of
pkg/nemo-asr/src/cli.py (usage / exit status / option parsing)."""
import io
import json
import os

import pytest

from reazonspeech_b200.nemo.asr import cli, writer
from reazonspeech_b200.nemo.asr.interface import Segment

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def cases():
    with open(os.path.join(HERE, "golden", "writer_cases.json")) as f:
        return json.load(f)


class Named(io.StringIO):
    def __init__(self, name):
        super().__init__()
        self.name = name


@pytest.mark.parametrize("ext", ["vtt", "srt", "ass", "json", "tsv", "txt"])
def test_writer_output_is_byte_identical(cases, ext):
    fp = io.StringIO()
    w = writer.get_writer(fp, ext)
    assert type(w).__name__ == cases["outputs"][ext]["class"] and w.ext == (ext if ext != "txt" else "txt")
    w.write_header()
    for a, b, t in cases["segments"]:
        w.write(Segment(start_seconds=a, end_seconds=b, text=t))
    assert fp.getvalue() == cases["outputs"][ext]["text"]


def test_get_writer_dispatch_matches_reference(cases):
    for key, cls in cases["get_writer"].items():
        if key.startswith("ext="):
            got = writer.get_writer(io.StringIO(), eval(key[4:]))
        else:
            got = writer.get_writer(Named(key))
        assert type(got).__name__ == cls, key


def test_cli_without_audio_prints_usage_and_returns_1(capsys):
    assert cli.main([]) == 1
    err = capsys.readouterr().err
    assert "no audio file specified" in err and "--output" in err


def test_cli_help_goes_to_stderr(capsys):
    assert cli.main(["-h"]) is None
    cap = capsys.readouterr()
    assert "--output" in cap.err and cap.out == ""


def test_cli_rejects_unknown_option():
    import getopt
    with pytest.raises(getopt.GetoptError):
        cli.main(["--nope", "a.wav"])


def test_cli_several_inputs_give_monotonic_times(tmp_path, monkeypatch):
    """The multi-file extension shifts every file's segments by the duration of the files before it (a subtitle file
    with times restarting at 0 is invalid); a single input is written exactly as the reference writes it."""
    import numpy as np
    from scipy.io import wavfile
    import importlib
    from reazonspeech_b200.nemo.asr import cli
    T = importlib.import_module("reazonspeech_b200.nemo.asr.transcribe")     # the package re-exports a function of the same name
    from reazonspeech_b200.nemo.asr.interface import Segment, TranscribeResult
    paths = []
    for i, secs in enumerate((2.0, 3.5, 1.0)):
        p = tmp_path / f"a{i}.wav"
        wavfile.write(p, 16000, (np.zeros(int(secs * 16000)) + i).astype(np.int16))
        paths.append(str(p))
    fake = lambda n: TranscribeResult(f"t{n}", [], [Segment(0.5, 1.0, f"t{n}")])
    monkeypatch.setattr(T, "load_model", lambda *a, **k: object())
    monkeypatch.setattr(T, "transcribe", lambda model, audio, config=None: fake(0))
    monkeypatch.setattr(T, "transcribe_batch", lambda model, audios, config=None: [fake(i) for i in range(len(audios))])
    out = tmp_path / "o.tsv"
    assert cli.main(["--to=tsv", "-o", str(out), *paths]) is None
    rows = [l.split("\t") for l in out.read_text().splitlines()[1:]]
    starts = [float(r[0]) for r in rows]
    assert starts == [0.5, 2.5, 6.0]                         # 0.5 s into each file, files of 2.0 and 3.5 s before the later ones
    assert cli.main(["--to=tsv", "-o", str(out), paths[1]]) is None
    assert [float(v) for v in out.read_text().splitlines()[1].split("\t")[:2]] == [0.5, 1.0]
