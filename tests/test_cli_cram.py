"""The reference's cram tests (jpeg/test/*.t) replayed through `python -m hcjpeg` - the command-line twin of
jpeg/bin/model.ml and tools/bin/oyuv.ml - with the command lines and the expected stdout exactly as the .t files hold them
(tests/golden/reference_goldens.json `cram_scripts`, lifted by make_golden.py; the ffmpeg steps are left out: no ffmpeg here).
Intermediate files are checked against the oracle on the way, so a pass means: same files, same printed numbers."""
import io
import os
import shlex
import sys
from contextlib import redirect_stdout

import pytest

from conftest import GOLDEN


def run_cli(argv):
    from hcjpeg.__main__ import main

    out = io.StringIO()
    with redirect_stdout(out):
        rc = main(argv)
    return rc, out.getvalue()


def test_cli_argument_forms():
    """No device needed: argument splitting and float printing follow Core.Command / sexp_of_float."""
    from hcjpeg.__main__ import _flags, _ocaml_float, _size

    assert _size("52x44") == (52, 44)
    anon, fl = _flags(["a.yuv", "64x64", "o.jpg", "-quality", "95", "-chroma", "422"], {"-quality": int, "-chroma": int})
    assert anon == ["a.yuv", "64x64", "o.jpg"] and fl == {"-quality": 95, "-chroma": 422}
    anon, fl = _flags(["-verbose", "a", "8x8"], {"-verbose": None})
    assert anon == ["a", "8x8"] and fl == {"-verbose": True}
    assert _ocaml_float(46.76864691904693) == "46.76864691904693" and _ocaml_float(46.760132097139362) == "46.760132097139362"
    assert _ocaml_float(3.0) == "3" and _ocaml_float(0.5) == "0.5" and _ocaml_float(float("inf")) == "INF" and _ocaml_float(1e-5) == "1E-05"


def test_cli_without_a_device_fails_loudly(tmp_path, monkeypatch, capsys):
    """No CPU path behind the commands: without CUDA `model decode frame` leaves with status 1 and writes nothing; the host-only
    `model decode header` still prints the pinned text (header parsing is host code in the library, as in the reference)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import hcjpeg.model

    monkeypatch.setattr(hcjpeg.model, "_default_ctx", None)
    monkeypatch.chdir(tmp_path)
    (tmp_path / "m.jpg").write_bytes(open(os.path.join(GOLDEN, "mini.jpg"), "rb").read())
    rc, out = run_cli(["model", "decode", "frame", "m.jpg", "o.yuv"])
    assert rc == 1 and out == "" and not (tmp_path / "o.yuv").exists()
    assert capsys.readouterr().err.strip() != ""  # the library's status text
    rc, out = run_cli(["model", "decode", "header", "m.jpg"])
    assert rc == 0 and out.startswith("(header\n ((frame\n   (((length 17) (sample_precision 8) (width 64) (height 64)")
    with pytest.raises(SystemExit):
        run_cli(["model", "frobnicate"])


@pytest.mark.gpu
@pytest.mark.parametrize("script", ["model-encode-and-decode.t", "test-nonstandard-sizes.t", "mouse-decode.t"])
def test_cram_script(script, goldens, orc, tmp_path, monkeypatch):
    steps = goldens["cram_scripts"][script]["steps"]
    assert steps
    # the layout the .t files assume: cwd = jpeg/test, data in ../test_data
    (tmp_path / "test").mkdir()
    (tmp_path / "test_data").mkdir()
    for name in ("Mouse480.jpg", "mini64x64.420", "mini64x64.422", "mini64x64.444"):
        (tmp_path / "test_data" / name).write_bytes(open(os.path.join(GOLDEN, name), "rb").read())
    monkeypatch.chdir(tmp_path / "test")
    printed = 0
    for step in steps:
        argv = shlex.split(step["cmd"])
        rc, out = run_cli(argv)
        assert rc == 0, step
        assert out.splitlines() == step["out"], step
        printed += len(step["out"])
        if argv[:3] == ["model", "encode", "frame"]:  # the file is the oracle's, byte for byte
            src, (w, h), dst = argv[3], [int(v) for v in argv[4].split("x")], argv[5]
            q = int(argv[argv.index("-quality") + 1]) if "-quality" in argv else 75
            c = int(argv[argv.index("-chroma") + 1]) if "-chroma" in argv else 420
            assert open(dst, "rb").read() == orc.encode(open(src, "rb").read(), w, h, c, q), step
        if argv[:3] == ["model", "decode", "frame"]:
            assert open(argv[4], "rb").read() == orc.decode(open(argv[3], "rb").read(), restart_ext=False).yuv(), step
    if script != "mouse-decode.t":
        assert printed >= 3


@pytest.mark.gpu
def test_cli_header_log_and_metrics(goldens, orc, tmp_path, monkeypatch, capsys):
    """`model decode header` prints the reference's pinned text; `decode log` / `encode log` one record per block;
    `oyuv compare` mean-difference / mean-square-error / max-difference against numpy."""
    import numpy as np

    monkeypatch.chdir(tmp_path)
    hdr = bytes.fromhex(goldens["header_480x320_q20_420"]["hex"])
    (tmp_path / "h.jpg").write_bytes(hdr)
    rc, out = run_cli(["model", "decode", "header", "h.jpg"])
    assert rc == 0 and out == goldens["header_480x320_q20_420"]["sexp_text"]
    mini = open(os.path.join(GOLDEN, "mini64x64.420"), "rb").read()
    (tmp_path / "m.yuv").write_bytes(mini)
    assert run_cli(["model", "encode", "frame", "m.yuv", "64x64", "m.jpg"])[0] == 0
    assert open("m.jpg", "rb").read() == open(os.path.join(GOLDEN, "mini.jpg"), "rb").read()  # the reference's own file
    rc, out = run_cli(["model", "decode", "log", "m.jpg"])
    assert rc == 0 and out.startswith("(header\n ((frame") and out.count("(!block_number ") == 96
    rc, out = run_cli(["model", "encode", "log", "m.yuv", "64x64", "-verbose"])
    assert rc == 0 and out.count("((!block_number 0)") == 96 and out.count("(error") == 96
    assert run_cli(["model", "decode", "frame", "m.jpg", "d.yuv"])[0] == 0
    a = np.frombuffer(mini, np.uint8).astype(np.int64)
    b = np.frombuffer(open("d.yuv", "rb").read(), np.uint8).astype(np.int64)
    from hcjpeg.__main__ import _ocaml_float

    for what, f in (("max-difference", lambda d: str(int(np.abs(d).max()))),
                    ("mean-difference", lambda d: _ocaml_float(float(np.abs(d).sum()) / d.size)),
                    ("mean-square-error", lambda d: _ocaml_float(float((d * d).sum()) / d.size))):
        rc, out = run_cli(["oyuv", "compare", what, "yuv", "m.yuv", "d.yuv", "64x64"])
        want = [f(a[:4096] - b[:4096]), f(a[4096:5120] - b[4096:5120]), f(a[5120:] - b[5120:])]
        assert rc == 0 and out.splitlines() == want, what
    rc, out = run_cli(["oyuv", "compare", "psnr", "y", "m.yuv", "m.yuv", "64x64"])
    assert rc == 0 and out.splitlines() == ["INF"]
