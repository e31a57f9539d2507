"""`model decode log` / `model encode log` as text (hcjpeg/sexp.py; SURVEY 8(f) row 4: "sexp-compatible per-block dumps").

The layout engine is pinned by the reference's own expect tests: every multi-line `print_s` output they hold
(tests/golden/sexp_layouts.json, lifted by make_golden.py) must parse and print back to itself, and the two `Header.t`
texts (jpeg/model/test/test_encode_headers.ml, jpeg/hardcaml/test/test_codeblock_decoder.ml) must come out of the bytes
of the files.  The per-block records have no expected text in the reference: their fields are checked against the oracle
(test_block_log_tap, test_encode_block_log) and their text here is checked for shape and against the oracle's blocks."""
import json
import os

import numpy as np
import pytest

import hcjpeg
from hcjpeg import sexp

import synth
from conftest import GOLDEN


def layouts():
    with open(os.path.join(GOLDEN, "sexp_layouts.json")) as f:
        return json.load(f)["layouts"]


def test_layout_engine_reproduces_every_reference_layout():
    ls = layouts()
    assert len(ls) >= 60
    sources = {l["source"].split(":")[0] for l in ls}
    assert len(sources) >= 12  # header records, code tables, matrices, hexdumps, RTL port dumps ...
    for l in ls:
        items = sexp.of_string_many(l["text"])
        assert len(items) == 1
        assert sexp.to_string_hum(items[0]) == l["text"], l["source"]


def test_single_line_and_atoms():
    assert sexp.to_string_hum(["a", ["b", "c"], []]) == "(a (b c) ())"
    assert sexp.to_string_hum("x y") == '"x y"' and sexp.to_string_hum("") == '""'
    assert sexp.to_string_hum(["!block_number", "-3"]) == "(!block_number -3)"
    assert sexp.of_string_many('(a "b c" (d))  e') == [["a", "b c", ["d"]], "e"]
    # a list that does not fit breaks after the last element that does, one column inside its parenthesis
    row = [str(i) for i in range(40)]
    text = sexp.to_string_hum(["elements", row])
    lines = text.split("\n")
    assert all(len(line) <= 78 for line in lines) and lines[0] == "(elements" and lines[1].startswith(" (0 1 ") and lines[2].startswith("  ")
    assert sexp.of_string_many(text) == [["elements", row]]


def test_header_text_matches_the_reference(goldens, data):
    """print_s [%message (header : Decoder.Header.t)] of the 623-byte header (test_encode_headers.ml:17-95) and of
    Mouse480.jpg (test_codeblock_decoder.ml:74,123-): byte for byte."""
    g = goldens["header_480x320_q20_420"]
    h = hcjpeg.header_decode(bytes.fromhex(g["hex"]))
    assert sexp.print_s(["header", sexp.sexp_of_header(h)]) == g["sexp_text"]
    h = hcjpeg.header_decode(data("Mouse480.jpg"))
    assert sexp.print_s(["headers", sexp.sexp_of_header(h)]) == goldens["mouse480"]["sexp_text"]


def test_header_with_restart_interval(orc):
    jpg = orc.encode(synth.frame(3, 64, 48, 420), 64, 48, 420, 75, restart_interval=8)
    text = sexp.to_string_hum(sexp.sexp_of_header(hcjpeg.header_decode(jpg)))
    assert "(restart_interval (((length 4) (restart_interval 8))))" in text


def _summary_records(dec, hdr):
    """Component.Summary records from the oracle's per-block taps, in the layout of Batch.block_log."""
    rec = np.zeros(dec.nblocks, hcjpeg.BLOCK_LOG_DTYPE)
    rec["coefs"], rec["dc_pred"], rec["dequant"], rec["recon"], rec["component"] = dec.coefs, dec.dc_abs, dec.dequant, dec.recon, dec.block_comp
    return rec


def test_block_record_text(orc, data):
    """Shape of a Component.Summary / Encoder.Block record: hex digits per Util.sexp_of_block (three for coefficients, two
    for pixels, two's complement), the scan component's identifier, eight rows of eight."""
    mouse = data("Mouse480.jpg")
    dec = orc.decode(mouse, want_blocks=True)
    hdr = hcjpeg.header_decode(mouse)
    ids = [hdr.scan_components[k].selector for k in range(hdr.number_of_image_components)]
    rec = _summary_records(dec, hdr)
    for k in (0, 4, 5, dec.nblocks - 1):
        rec["idct"][k] = orc.chen_inverse(dec.dequant[k])
        s = sexp.sexp_of_component_summary(rec[k], ids)
        assert [f[0] for f in s] == ["x", "y", "dc_pred", "component.identifier", "coefs", "dequant", "idct", "recon"]
        assert s[3][1] == str(ids[int(dec.block_comp[k])])
        coefs = dict((f[0], f[1]) for f in s)["coefs"]
        assert len(coefs) == 8 and all(len(r) == 8 for r in coefs)
        assert coefs[0][0] == "%03x" % (int(dec.coefs[k][0]) & 0xFFF) and coefs[7][7] == "%03x" % (int(dec.coefs[k][63]) & 0xFFF)
        text = sexp.print_s([["!block_number", str(k)], ["component", s]])
        assert text.startswith("((!block_number %d)\n (component\n  ((x " % k)
        back = sexp.of_string_many(text)
        assert back == [[["!block_number", str(k)], ["component", s]]]
    neg = np.zeros(1, hcjpeg.BLOCK_LOG_DTYPE)
    neg["coefs"][0][1], neg["idct"][0][2] = -2, -3
    s = dict((f[0], f[1]) for f in sexp.sexp_of_component_summary(neg[0], [1]))
    assert s["coefs"][0][1] == "ffe" and s["idct"][0][2] == "fd"
    blk = np.zeros(1, hcjpeg.ENCODER_BLOCK_DTYPE)
    blk["nrle"], blk["rle_value"][0][0], blk["rle_run"][0][1] = 2, -7, 62
    plain = sexp.sexp_of_encoder_block(blk[0])
    assert [f[0] for f in plain] == ["x_pos", "y_pos", "input_pixels", "fdct", "quant", "dc_pred", "rle", "decoded"]
    assert plain[6][1] == [[["run", "0"], ["value", "-7"]], [["run", "62"], ["value", "0"]]] and plain[7][1] == []
    verbose = sexp.sexp_of_encoder_block(blk[0], verbose=True)
    assert [f[0] for f in verbose[7][1][0]] == ["dequant", "idct", "recon", "error"]


@pytest.mark.gpu
def test_decode_log_text(orc, data):
    """`model decode log Mouse480.jpg` from the device taps == the same text rendered from the oracle's blocks."""
    mouse = data("Mouse480.jpg")
    dec = orc.decode(mouse, want_blocks=True, restart_ext=False)
    hdr = hcjpeg.header_decode(mouse)
    ids = [hdr.scan_components[k].selector for k in range(hdr.number_of_image_components)]
    with hcjpeg.Context(0) as ctx:
        got = sexp.decode_log(mouse, ctx)
        with ctx.batch([mouse], hcjpeg.OUT_PLANES, 0) as b:
            b.decode()
            xy = b.block_log(0)
    rec = _summary_records(dec, hdr)
    rec["x"], rec["y"] = xy["x"], xy["y"]  # positions are checked against the planes in test_block_log_tap
    for k in range(dec.nblocks):
        rec["idct"][k] = orc.chen_inverse(dec.dequant[k])
    want = sexp.print_s(["header", sexp.sexp_of_header(hdr)]) + "".join(
        sexp.print_s([["!block_number", str(k)], ["component", sexp.sexp_of_component_summary(rec[k], ids)]]) for k in range(dec.nblocks)
    )
    assert got == want
    assert got.count("(!block_number ") == dec.nblocks == 60 * 40 * 6 // 4


@pytest.mark.gpu
def test_encode_log_text(orc, data):
    """`model encode log mini64x64.420 64x64 [-verbose]`: one record per block, quantised blocks as the oracle's."""
    f = data("mini64x64.420")
    _, quant, fdct = orc.encode(f, 64, 64, 420, 75, want_blocks=True)
    with hcjpeg.Context(0) as ctx:
        plain = sexp.encode_log(f, 64, 64, 420, 75, ctx=ctx)
        verbose = sexp.encode_log(f, 64, 64, 420, 75, verbose=True, ctx=ctx, count=3)
    recs = sexp.of_string_many(plain)
    assert len(recs) == len(quant) == 96
    for k, r in enumerate(recs):
        assert r[0] == ["!block_number", "0"]  # the command never increments it (model.ml:139-141)
        fields = dict((x[0], x[1]) for x in r[1][1])
        assert [v for row in fields["quant"] for v in row] == ["%03x" % (int(v) & 0xFFF) for v in quant[k]]
        assert [v for row in fields["fdct"] for v in row] == ["%03x" % (int(v) & 0xFFF) for v in fdct[k]]
        assert fields["decoded"] == []
    v = sexp.of_string_many(verbose)
    assert len(v) == 3 and [x[0] for x in dict((x[0], x[1]) for x in v[0][1][1])["decoded"][0]] == ["dequant", "idct", "recon", "error"]
