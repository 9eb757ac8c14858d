"""Generate tests/golden/writer_cases.json by running the REFERENCE's own writers
(/root/reference/pkg/nemo-asr/src/writer.py, imported unmodified) on seeded segments.
Build container only (needs /root/reference):  python tests/golden/make_writer_golden.py"""
import importlib.util
import io
import json
import os
import random

REF = "/root/reference/pkg/nemo-asr/src/writer.py"
HERE = os.path.dirname(os.path.abspath(__file__))


class Seg:
    def __init__(self, a, b, t):
        self.start_seconds, self.end_seconds, self.text = a, b, t


class Named(io.StringIO):
    def __init__(self, name):
        super().__init__()
        self.name = name


def main():
    spec = importlib.util.spec_from_file_location("refwriter", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    rng = random.Random(11)
    segs = [(0.0, 0.08, "あ"), (0.999, 1.0004, "こんにちは。"), (59.9996, 60.08, "元気ですか?"), (3599.5, 3600.25, "tab\there"),
            (7325.678, 7329.0049, "二時間後"), (12.345678, 12.3, "end before start")]
    t = 0.0
    for _ in range(20):
        a = t + rng.random() * 3
        b = a + rng.random() * 9
        t = b
        segs.append((round(a, rng.choice([2, 3, 6])), b, rng.choice(["今日は", "天気が、いいですね。", "hello world", "！？", ""])))
    cases = {"segments": segs, "outputs": {}, "get_writer": {}}
    for ext in ("vtt", "srt", "ass", "json", "tsv", "txt"):
        fp = io.StringIO()
        w = ref.get_writer(fp, ext)
        w.write_header()
        for s in segs:
            w.write(Seg(*s))
        cases["outputs"][ext] = {"class": type(w).__name__, "text": fp.getvalue()}
    for name in ("out.vtt", "out.srt", "sub.ass", "x.json", "x.tsv", "x.txt", "noext", ""):
        cases["get_writer"][name] = type(ref.get_writer(Named(name))).__name__
    for ext in (".vtt", "VTT", "txt", None):
        cases["get_writer"]["ext=%r" % (ext,)] = type(ref.get_writer(io.StringIO(), ext)).__name__
    with open(os.path.join(HERE, "writer_cases.json"), "w") as f:
        json.dump(cases, f, ensure_ascii=False, indent=0)
    print("wrote", len(segs), "segments x", len(cases["outputs"]), "formats")


if __name__ == "__main__":
    main()
