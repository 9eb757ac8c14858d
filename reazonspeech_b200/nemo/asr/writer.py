"""Subtitle / transcript writers with the reference's names, constructor and method surface
(pkg/nemo-asr/src/writer.py:4-168): ``VTTWriter, SRTWriter, ASSWriter, JSONWriter, TSVWriter,
TextWriter`` and ``get_writer(fp, ext=None)``.  Outputs are byte-identical to the reference's
(tests/golden/writer_cases.json, produced by importing the reference module unmodified).

One table drives the formats: a writer is (extension, header text, time layout, line template).
Time fields are truncated, not rounded, exactly like the reference (``int(seconds % 1 * 1000)``).

Reference quirk kept on purpose (SURVEY.md App. C style): ``get_writer`` compares the class
extension ("vtt") with what it is given, and when it derives the extension from ``fp.name`` it
keeps the leading dot (".vtt") -- so ``-o out.vtt`` without ``--to`` falls through to the plain
text writer (writer.py:158-168).  Callers that pass ``ext="vtt"`` get the VTT writer."""
from __future__ import annotations

import json
import os


def _clock(seconds: float, frac_digits: int, sep: str, pad_hours: bool) -> str:
    hours = int(seconds / 3600)
    minutes = int(seconds / 60) % 60
    secs = int(seconds % 60)
    frac = int((seconds % 1) * 10 ** frac_digits)
    head = "%02i" % hours if pad_hours else "%i" % hours
    return "%s:%02i:%02i%s%0*i" % (head, minutes, secs, sep, frac_digits, frac)


class _SegmentWriter:
    """Base: subclasses set ``ext``, ``header`` and implement ``format(segment)``."""
    ext = ""
    header = ""

    def __init__(self, fp):
        self.fp = fp

    def write_header(self):
        if self.header:
            self.fp.write(self.header)

    def format(self, segment) -> str:
        raise NotImplementedError

    def write(self, segment):
        self.fp.write(self.format(segment))


class VTTWriter(_SegmentWriter):
    """WebVTT (https://www.w3.org/TR/webvtt1/)."""
    ext = "vtt"
    header = "WEBVTT\n\n"

    @staticmethod
    def _format_time(seconds):
        return _clock(seconds, 3, ".", True)

    def format(self, segment):
        return "%s --> %s\n%s\n\n" % (self._format_time(segment.start_seconds), self._format_time(segment.end_seconds), segment.text)


class SRTWriter(_SegmentWriter):
    """SubRip: numbered cues, comma before the milliseconds."""
    ext = "srt"

    def __init__(self, fp):
        super().__init__(fp)
        self.index = 0

    @staticmethod
    def _format_time(seconds):
        return _clock(seconds, 3, ",", True)

    def format(self, segment):
        self.index += 1
        return "%i\n%s --> %s\n%s\n\n" % (self.index, self._format_time(segment.start_seconds),
                                          self._format_time(segment.end_seconds), segment.text)


class ASSWriter(_SegmentWriter):
    """Advanced SubStation Alpha (libass): centisecond times, unpadded hours."""
    ext = "ass"
    header = ("[Script Info]\nScriptType: v4.00+\nCollisions: Normal\nTimer: 100.0000\n\n"
              "[V4+ Styles]\nStyle: Default,Arial,16,&Hffffff,&Hffffff,&H0,&H0,0,0,0,0,100,100,0,0,1,1,0,2,10,10,10,0\n\n"
              "[Events]\n")

    @staticmethod
    def _format_time(seconds):
        return _clock(seconds, 2, ".", False)

    def format(self, segment):
        return "Dialogue: 0,%s,%s,Default,,0,0,0,,%s\n" % (self._format_time(segment.start_seconds),
                                                          self._format_time(segment.end_seconds), segment.text)


class JSONWriter(_SegmentWriter):
    """One JSON object per line, times rounded to milliseconds, text not ASCII-escaped."""
    ext = "json"

    def format(self, segment):
        return json.dumps({"start_seconds": round(segment.start_seconds, 3), "end_seconds": round(segment.end_seconds, 3),
                           "text": segment.text}, ensure_ascii=False) + "\n"


class TSVWriter(_SegmentWriter):
    ext = "tsv"
    header = "start_seconds\tend_seconds\ttext\n"

    def format(self, segment):
        return "%.3f\t%.3f\t%s\n" % (segment.start_seconds, segment.end_seconds, segment.text)


class TextWriter(_SegmentWriter):
    """Fallback: ``[hh:mm:ss.mmm --> hh:mm:ss.mmm] text``."""
    ext = "txt"

    @staticmethod
    def _format_time(seconds):
        return _clock(seconds, 3, ".", True)

    def format(self, segment):
        return "[%s --> %s] %s\n" % (self._format_time(segment.start_seconds), self._format_time(segment.end_seconds), segment.text)


_BY_EXT = {cls.ext: cls for cls in (VTTWriter, SRTWriter, ASSWriter, JSONWriter, TSVWriter)}


def get_writer(fp, ext=None):
    if ext is None:
        ext = os.path.splitext(getattr(fp, "name", ""))[-1]        # keeps the dot, as the reference does (see module docstring)
    return _BY_EXT.get(ext, TextWriter)(fp)
