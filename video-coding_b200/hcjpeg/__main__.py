"""Command-line twins of the reference's two executables, running on the GPU:

    python -m hcjpeg model decode frame IN.jpg [OUT.yuv]                      jpeg/bin/model.ml:29-44
    python -m hcjpeg model decode header IN.jpg                               jpeg/bin/model.ml:18-27
    python -m hcjpeg model decode log IN.jpg                                  jpeg/bin/model.ml:46-68
    python -m hcjpeg model encode frame IN.yuv WxH OUT.jpg [-quality Q] [-chroma 420|422|444]      model.ml:84-106
    python -m hcjpeg model encode log IN.yuv WxH [-quality Q] [-chroma C] [-verbose]               model.ml:108-142
    python -m hcjpeg oyuv compare {max-difference|mean-difference|mean-square-error|psnr} {y|u|v|yuv} F1 F2 WxH [-format F]
                                                                              tools/src/ocompare.ml:83-135
    python -m hcjpeg oyuv convert IN WxH OUT [WxH] [-format F] [-out-format F]    tools/src/oconv.ml:66-133 (planar formats, one frame)

so that the reference's cram tests (jpeg/test/*.t) run as they are written with
``alias model='python -m hcjpeg model'`` and ``alias oyuv='python -m hcjpeg oyuv'`` (tests/test_cli_cram.py replays them).
Same positional arguments, flags, defaults (quality 75, 4:2:0) and stdout text; errors leave with status 1 and the
model's message.  Everything is computed on the device through the C ABI; without a GPU every command fails.
"""
import math
import sys

from . import HcjError, header_decode, sexp
from .model import Decoder, Encoder, Frame, Writer, default_context


def _size(s):  # common/src/size.ml: WIDTHxHEIGHT
    w, h = s.lower().split("x")
    return int(w), int(h)


def _flags(args, spec):
    """Split Core.Command style arguments: anonymous ones in order, `-flag value` / `-flag` anywhere."""
    anon, flags, i = [], {}, 0
    while i < len(args):
        a = args[i]
        if a in spec:
            if spec[a] is None:
                flags[a] = True
                i += 1
            else:
                flags[a] = spec[a](args[i + 1])
                i += 2
        elif a.startswith("-") and a != "-" and not a[1:2].isdigit():
            raise SystemExit("unknown flag %s" % a)
        else:
            anon.append(a)
            i += 1
    return anon, flags


def _chroma(s):
    if s not in ("420", "422", "444"):
        raise SystemExit("Invalid chroma type")
    return int(s)


def _planar_format(s):
    if s.upper() not in ("420", "422", "444"):
        raise SystemExit("Invalid YUV format" if s.upper() not in ("YUY2", "UYVY", "YVYU") else "packed formats are not offered on the device")
    return int(s)


def _read(path):
    return sys.stdin.buffer.read() if path == "-" else open(path, "rb").read()


def _write(path, data):
    if path is None or path == "-":
        sys.stdout.buffer.write(data)
        sys.stdout.buffer.flush()
    else:
        with open(path, "wb") as f:
            f.write(data)


def _ocaml_float(x):
    """sexp_of_float (Sexplib0.Sexp_conv.default_string_of_float): %.15G if it reads back to the same float, else %.17G."""
    s = "%.15G" % x
    return s if not math.isnan(x) and float(s) == x else "%.17G" % x


def model(args):
    if args[:2] == ["decode", "frame"]:
        anon, _ = _flags(args[2:], {})
        frame = Decoder.decode_a_frame(_read(anon[0]))
        _write(anon[1] if len(anon) > 1 else None, frame.tobytes())
    elif args[:2] == ["decode", "header"]:
        anon, _ = _flags(args[2:], {})
        sys.stdout.write(sexp.print_s(["header", sexp.sexp_of_header(header_decode(_read(anon[0])))]))
    elif args[:2] == ["decode", "log"]:
        anon, _ = _flags(args[2:], {})
        sys.stdout.write(sexp.decode_log(_read(anon[0])))
    elif args[:2] == ["encode", "frame"]:
        anon, fl = _flags(args[2:], {"-quality": int, "-chroma": _chroma})
        (w, h), chroma = _size(anon[1]), fl.get("-chroma", 420)
        frame = Frame.frombytes(_read(anon[0]), chroma, w, h)
        writer = Writer.create()
        {420: Encoder.encode_420, 422: Encoder.encode_422, 444: Encoder.encode_444}[chroma](frame=frame, quality=fl.get("-quality", 75), writer=writer)
        _write(anon[2], writer.get_buffer())
    elif args[:2] == ["encode", "log"]:
        anon, fl = _flags(args[2:], {"-quality": int, "-chroma": _chroma, "-verbose": None})
        (w, h), chroma = _size(anon[1]), fl.get("-chroma", 420)
        frame = Frame.frombytes(_read(anon[0]), chroma, w, h)
        sys.stdout.write(sexp.encode_log(frame.tobytes(), w, h, chroma, fl.get("-quality", 75), verbose=bool(fl.get("-verbose"))))
    else:
        raise SystemExit(__doc__)


def _planes(data, w, h, fmt):
    f = Frame.frombytes(data, fmt, w, h)
    return {"y": f.y, "u": f.u, "v": f.v}


def oyuv(args):
    if args[:1] == ["compare"] and len(args) >= 3:
        what, which = args[1], args[2]
        anon, fl = _flags(args[3:], {"-format": _planar_format})
        (w, h), fmt = _size(anon[2]), fl.get("-format", 420)
        a, b = _planes(_read(anon[0]), w, h, fmt), _planes(_read(anon[1]), w, h, fmt)
        ctx = default_context()
        for k in ("y", "u", "v") if which == "yuv" else (which,):
            pa, pb = a[k], b[k]
            sse, mx, tot = ctx.compare_planes_ex(pa.plane, pb.plane)
            n = float(pa.width) * float(pa.height)
            if what == "max-difference":
                print(mx)
            elif what == "mean-difference":
                print(_ocaml_float(tot / n))
            elif what == "mean-square-error":
                print(_ocaml_float(sse / n))
            elif what == "psnr":  # ocompare.ml:57-59: 10 log10 (255^2 / mse), a float division: inf for identical planes
                print(_ocaml_float(math.inf if sse == 0 else 10.0 * math.log10(255.0 * 255.0 / (sse / n))))
            else:
                raise SystemExit(__doc__)
    elif args[:1] == ["convert"]:
        anon, fl = _flags(args[1:], {"-format": _planar_format, "-out-format": _planar_format})
        (w, h) = _size(anon[1])
        dw, dh = _size(anon[3]) if len(anon) > 3 else (w, h)
        fmt = fl.get("-format", 420)
        out = default_context().yuv_convert(_read(anon[0]), w, h, fmt, dw, dh, fl.get("-out-format", fmt))
        _write(anon[2], out.tobytes())
    else:
        raise SystemExit(__doc__)


def main(argv):
    if len(argv) < 1 or argv[0] not in ("model", "oyuv"):
        raise SystemExit(__doc__)
    try:
        (model if argv[0] == "model" else oyuv)(argv[1:])
    except HcjError as e:
        sys.stderr.write("%s\n" % e)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
