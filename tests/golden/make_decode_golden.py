"""Generate tests/golden/decode_cases.json by running the REFERENCE's own decode_hypothesis
(/root/reference/pkg/nemo-asr/src/decode.py, imported unmodified) on seeded hypothesis stubs.

Run in the build container only (needs /root/reference):  python tests/golden/make_decode_golden.py
The reference ships no fixtures for this function; these vectors pin our decode.py to its outputs."""
import importlib.util
import json
import os
import random
import sys
import types

REF = "/root/reference/pkg/nemo-asr/src"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def load_reference_decode():
    pkg = types.ModuleType("refnemo")
    pkg.__path__ = [REF]
    sys.modules["refnemo"] = pkg
    for name in ("interface", "decode"):
        spec = importlib.util.spec_from_file_location(f"refnemo.{name}", os.path.join(REF, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refnemo.{name}"] = mod
        spec.loader.exec_module(mod)
    return sys.modules["refnemo.decode"]


class Tok:
    def __init__(self, pieces):
        self.pieces = pieces

    def ids_to_text(self, ids):
        t = "".join(self.pieces[i] for i in ids).replace("▁", " ")
        return t[1:] if t.startswith(" ") else t


class Model:
    def __init__(self, pieces):
        self.tokenizer = Tok(pieces)


class Seq(list):
    def tolist(self):
        return list(self)


class Hyp:
    def __init__(self, y, ts):
        self.y_sequence, self.timestamp = Seq(y), ts


def main():
    ref = load_reference_decode()
    pieces = ["⁇", "▁", "。", "、", "?", "!", ",", "こ", "ん", "に", "ち", "は", "元", "気", "で", "す", "か", "▁今日", "天", "気"]
    blank = len(pieces)
    rng = random.Random(7)
    cases = []
    hand = [
        ([7, 8, 9, 10, 11, 2, 1, 12, 13, 4], [10, 12, 13, 15, 20, 21, 22, 30, 31, 40]),
        ([], []),
        ([1], [3]),
        ([2, 2, 3, 4], [0, 0, 1, 1]),
        ([7] * 25, list(range(5, 30))),
        ([7, 3, 8] * 8, [2 * i for i in range(24)]),
        ([7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 7, 8], [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 30, 31, 60, 61]),
    ]
    for toks, frames in hand:
        cases.append((toks, frames))
    for _ in range(40):
        n = rng.randint(0, 60)
        toks = [rng.choice(range(1, blank)) if rng.random() > 0.25 else rng.choice([2, 3, 4, 5, 6, 1]) for _ in range(n)]
        frames, t = [], 0
        for _ in range(n):
            t += rng.choice([0, 0, 1, 1, 2, 3, 9])
            frames.append(t)
        cases.append((toks, frames))
    out = {"pieces": pieces, "blank": blank, "cases": []}
    model = Model(pieces)
    for toks, frames in cases:
        y = [blank] + toks                                   # ALSD shape: leading blank (decode.py:38-40)
        ts = [f + i + 1 for i, f in enumerate(frames)]       # step = frame + index + 1 (decode.py:48)
        r = ref.decode_hypothesis(model, Hyp(y, ts))
        out["cases"].append({
            "y_sequence": y, "timestamp": ts, "text": r.text,
            "subwords": [[s.seconds, s.token_id, s.token] for s in r.subwords],
            "segments": [[s.start_seconds, s.end_seconds, s.text] for s in r.segments],
        })
    with open(os.path.join(HERE, "decode_cases.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=0)
    print(len(out["cases"]), "cases written")


if __name__ == "__main__":
    main()
