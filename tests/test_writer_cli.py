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
