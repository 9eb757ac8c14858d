"""USAGE

    reazonspeech-b200-nemo-asr [-h] [--to={vtt,srt,ass,json,tsv}] [-o file] audio [audio ...]

OPTIONS

    audio
        Audio file(s) to transcribe (16-bit / float WAV; anything else needs librosa,
        see audio_from_path).  Several files are transcribed as one batch.

    -h, --help
        Print this help message.

    --to={vtt,srt,ass,json,tsv}
        Output format for transcription

    -o file, --output=file
        File to write transcription

Same options and behaviour as the reference CLI (pkg/nemo-asr/src/cli.py:1-77): no audio ->
message + usage on stderr and exit status 1; -h -> usage on stderr; one output stream, header
once, then every segment.  Extension: more than one audio argument (the reference reads one).
"""
import getopt
import sys
import warnings

from .audio import audio_from_path
from .transcribe import load_model, transcribe, transcribe_batch
from .writer import get_writer


def main(argv=None):
    outpath = None
    outext = None
    opts, args = getopt.getopt(sys.argv[1:] if argv is None else list(argv), "ho:", ("help", "output=", "to="))
    for key, value in opts:
        if key in ("-h", "--help"):
            print(__doc__, file=sys.stderr)
            return
        if key in ("-o", "--output"):
            outpath = value
        elif key == "--to":
            outext = value
    if not args:
        print("no audio file specified", file=sys.stderr)
        print(__doc__, file=sys.stderr)
        return 1
    outfile = open(outpath, "w") if outpath is not None else sys.stdout
    warnings.simplefilter("ignore")
    audios = [audio_from_path(path) for path in args]
    model = load_model()
    results = [transcribe(model, audios[0])] if len(audios) == 1 else transcribe_batch(model, audios)
    with outfile:
        writer = get_writer(outfile, outext)
        writer.write_header()
        for result in results:
            for segment in result.segments:
                writer.write(segment)


if __name__ == "__main__":
    sys.exit(main())
