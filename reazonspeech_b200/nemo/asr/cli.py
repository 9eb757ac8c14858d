"""Command line of the drop-in package: ``python -m reazonspeech_b200.nemo.asr.cli [options] AUDIO [AUDIO ...]``.

Same contract as the reference's ``reazonspeech-nemo-asr`` entry point (pkg/nemo-asr/src/cli.py:36-74):

  -h / --help          usage on stderr, nothing else happens
  -o FILE / --output=  where the transcript goes (default: stdout)
  --to=FMT             vtt | srt | ass | json | tsv; anything else, or no option, gives the bracketed plain-text
                       lines.  Like the reference, the format is NOT inferred from FILE's extension
                       (see writer.get_writer)
  no AUDIO argument    "no audio file specified" + usage on stderr, exit status 1
  unknown option       getopt.GetoptError propagates, as in the reference

One extension: several AUDIO arguments are transcribed as one batch on the GPU (the reference reads exactly one);
their segments are written one file after the other through the same writer, every file's times shifted by the total
duration of the files before it, i.e. the transcript of the files played back to back (subtitle formats need monotonic times).
Audio decoding: soundfile when installed, scipy for WAV, librosa for compressed containers, see audio.audio_from_path.
"""
import dataclasses
import getopt
import sys
import warnings
from dataclasses import dataclass, field
from typing import List, Optional

SHORT_OPTS = "ho:"
LONG_OPTS = ("help", "output=", "to=")


@dataclass
class Options:
    help: bool = False
    output: Optional[str] = None
    fmt: Optional[str] = None
    audio: List[str] = field(default_factory=list)


def parse(argv) -> Options:
    parsed, rest = getopt.getopt(list(argv), SHORT_OPTS, LONG_OPTS)
    opt = Options(audio=rest)
    for flag, value in parsed:
        if flag in ("-h", "--help"):
            opt.help = True
            break                                            # the reference returns at the first -h it meets
        if flag in ("-o", "--output"):
            opt.output = value
        if flag == "--to":
            opt.fmt = value
    return opt


def usage() -> None:
    print(__doc__, file=sys.stderr)


def run(opt: Options) -> None:
    from .audio import audio_from_path
    from .transcribe import load_model, transcribe, transcribe_batch
    from .writer import get_writer

    sink = sys.stdout if opt.output is None else open(opt.output, "w")
    warnings.simplefilter("ignore")
    clips = [audio_from_path(path) for path in opt.audio]
    model = load_model()
    results = transcribe_batch(model, clips) if len(clips) > 1 else [transcribe(model, clips[0])]
    with sink:
        out = get_writer(sink, opt.fmt)
        out.write_header()
        offset = 0.0
        for clip, result in zip(clips, results):
            for segment in result.segments:
                out.write(segment if offset == 0.0 else
                          dataclasses.replace(segment, start_seconds=segment.start_seconds + offset, end_seconds=segment.end_seconds + offset))
            offset += clip.seconds


def main(argv=None):
    opt = parse(sys.argv[1:] if argv is None else argv)
    if opt.help:
        usage()
        return None
    if not opt.audio:
        print("no audio file specified", file=sys.stderr)
        usage()
        return 1
    run(opt)
    return None


if __name__ == "__main__":
    sys.exit(main())
