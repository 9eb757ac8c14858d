"""Text normalisation and character error rate, API-compatible with pkg/evaluation/src/utils.py:1-33
(``normalize``, ``calculate_cer`` -> ``{"cer", "distance", "length"}``).

The reference leans on two packages that are not dependencies here: ``editdistance`` (replaced by the
banded-free two-row Levenshtein below; same result by definition) and ``num2words(lang="ja")`` for digit
runs (replaced by ``japanese_number``, a restatement of its cardinal reading: kanji numerals with myriad
grouping, decimals read digit by digit after 点 -- marked (R): num2words cannot be imported offline, the
cases in tests/test_evaluation.py are hand-computed)."""
from __future__ import annotations

import re
from typing import TypedDict


class CERResult(TypedDict):
    cer: float
    distance: int
    length: int


# characters removed before scoring (utils.py:15) and full-width -> half-width folding (utils.py:16-18)
_PUNCTUATION = "、。「」『』，,？！!!?!?"
_FULLWIDTH = {0xFF21 + i: 0x41 + i for i in range(26)}
_FULLWIDTH.update({0xFF41 + i: 0x61 + i for i in range(26)})
_FULLWIDTH.update({0xFF10 + i: 0x30 + i for i in range(10)})
_TABLE = {**{ord(c): None for c in _PUNCTUATION}, **_FULLWIDTH}

_DIGITS = "零一二三四五六七八九"
_SMALL_UNITS = ("", "十", "百", "千")
_BIG_UNITS = ("", "万", "億", "兆", "京")


def _four_digits(n: int) -> str:
    out = []
    for power in (3, 2, 1, 0):
        d = (n // 10 ** power) % 10
        if d == 0:
            continue
        out.append(("" if d == 1 and power > 0 else _DIGITS[d]) + _SMALL_UNITS[power])
    return "".join(out)


def japanese_number(text: str) -> str:
    """'123' -> '百二十三', '10000' -> '一万', '3.14' -> '三点一四' (R)."""
    whole, _, frac = text.partition(".")
    n = int(whole) if whole else 0
    if n >= 10 ** 20:
        raise OverflowError(text)
    if n == 0:
        head = _DIGITS[0]
    else:
        parts, group = [], 0
        while n:
            n, chunk = divmod(n, 10000)
            if chunk:
                parts.append(_four_digits(chunk) + _BIG_UNITS[group])
            group += 1
        head = "".join(reversed(parts))
    if frac:
        head += "点" + "".join(_DIGITS[int(c)] for c in frac)
    return head


def normalize(s: str) -> str:
    s = s.translate(_TABLE)
    try:
        return re.sub(r"\d+\.?\d*", lambda m: japanese_number(m.group(0)), s)
    except OverflowError:
        return s


def edit_distance(a: str, b: str) -> int:
    """Levenshtein distance (unit costs), two rows."""
    if len(a) < len(b):
        a, b = b, a
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return prev[-1]


def calculate_cer(reference: str, prediction: str) -> CERResult:
    reference = normalize(reference)
    prediction = normalize(prediction)
    distance = edit_distance(reference, prediction)
    return CERResult(cer=distance / len(reference), distance=distance, length=len(reference))
