"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/hcjpeg.h declares, and its host half (header parse, geometry, headers) agrees with the oracle.
No compute entry point is called here (there is no GPU in this tier)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def hcj():
    import hcjpeg

    hcjpeg.build()
    return hcjpeg


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "hcjpeg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hcj_[a-z0-9_]+)\s*\(", src)))


def test_exports_every_declared_symbol(hcj):
    L = C.CDLL(hcj.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), n
    assert set(names) == set(hcj._SIGS), set(names) ^ set(hcj._SIGS)
    assert hcj.lib().hcj_version() == 100


def test_status_codes_match_oracle(hcj):
    hdr = open(os.path.join(ROOT, "include", "hcjpeg.h")).read()
    orc = open(os.path.join(ROOT, "oracle", "hcj_oracle.h")).read()
    a = dict(re.findall(r"HCJ_(ERR_[A-Z0-9_]+|OK) = (-?\d+)", hdr))
    b = dict(re.findall(r"ORC_(ERR_[A-Z0-9_]+|OK) = (-?\d+)", orc))
    assert len(b) >= 20
    for k, v in b.items():
        assert a[k] == v, k
    for code, name in hcj.STATUS.items():
        assert a[name[4:]] == str(code)
        assert hcj.lib().hcj_strerror(code)


def test_no_cuda_means_error_not_fallback(hcj):
    """Without a device the product must fail loudly (never route through a CPU path)."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    st = hcj.lib().hcj_ctx_create(0, None, C.byref(h))
    assert st <= -1000
    with pytest.raises(hcj.HcjError):
        hcj.Context(0)


def _same_header(h, o):
    assert (h.has_frame, h.width, h.height, h.sample_precision, h.number_of_components) == (
        o.frame.present, o.frame.width, o.frame.height, o.frame.sample_precision, o.frame.number_of_components)
    for i in range(h.number_of_components):
        a, b = h.components[i], o.frame.components[i]
        assert (a.identifier, a.horizontal_sampling_factor, a.vertical_sampling_factor, a.quantization_table_identifier) == (b.identifier, b.h, b.v, b.tq)
    assert h.number_of_image_components == o.scan.number_of_image_components
    for i in range(h.number_of_image_components):
        a, b = h.scan_components[i], o.scan.scan_components[i]
        assert (a.selector, a.dc_coef_selector, a.ac_coef_selector) == (b.selector, b.dc, b.ac)
    assert (h.has_restart_interval, h.restart_interval) == (o.restart_interval_present, o.restart_interval)
    assert h.n_quant_tables == o.n_quant_tables and h.n_huffman_tables == o.n_huffman_tables
    for i in range(h.n_quant_tables):
        assert list(h.quant_tables[i].elements) == list(o.quant_tables[i].elements)
        assert h.quant_tables[i].table_identifier == o.quant_tables[i].table_identifier
    for i in range(h.n_huffman_tables):
        a, b = h.huffman_tables[i], o.huffman_tables[i]
        assert (a.table_class, a.destination_identifier, a.nvalues) == (b.table_class, b.destination_identifier, b.nvalues)
        assert list(a.lengths) == list(b.lengths)
        assert list(a.values)[: a.nvalues] == list(b.values)[: b.nvalues]
    assert h.scan_byte_pos == o.scan_bit_pos // 8


@pytest.mark.parametrize("name", ["Mouse480.jpg", "mini.jpg"])
def test_header_decode_matches_oracle(hcj, orc, data, name):
    jpg = data(name)
    _same_header(hcj.header_decode(jpg), orc.header_decode(jpg))


def test_header_decode_pillow_streams(hcj, orc):
    Image = pytest.importorskip("PIL.Image")
    import io

    rng = np.random.default_rng(0)
    img = Image.fromarray(rng.integers(0, 255, (72, 100, 3), dtype=np.uint8))
    for kw in (dict(subsampling=2), dict(subsampling=0, optimize=True), dict(subsampling=1, restart_marker_blocks=4), dict(subsampling=2, quality=100)):
        buf = io.BytesIO()
        img.save(buf, "JPEG", **kw)
        jpg = buf.getvalue()
        try:
            o = orc.header_decode(jpg)
        except orc.OracleError as e:
            with pytest.raises(hcj.HcjError) as ei:
                hcj.header_decode(jpg)
            assert ei.value.status == e.status
            continue
        _same_header(hcj.header_decode(jpg), o)
        f, dec = hcj.frame_info(jpg), orc.decode(jpg)
        assert (f.mcus_wide, f.mcus_high, f.blocks_per_mcu, f.nblocks, f.chroma) == (dec.mcus_wide, dec.mcus_high, dec.blocks_per_mcu, dec.nblocks, dec.chroma)
        assert [(f.decoded_width[i], f.decoded_height[i]) for i in range(3)] == dec.decoded_size
        assert [(f.actual_width[i], f.actual_height[i]) for i in range(3)] == dec.actual_size


def test_t81_table_segments_extension(hcj, orc):
    """HCJ_FLAG_T81_TABLES (stated extension): merged DQT / DHT segments and FF fill bytes parse to the same tables
    as the one-table-per-segment file; without the flag the library fails exactly like the model's parser."""
    for chroma, ri in ((420, 0), (444, 4), (422, 0)):
        jpg = orc.encode(synth.frame(7, 72, 40, chroma), 72, 40, chroma, 60, restart_interval=ri)
        for fill in (0, 1, 3):
            merged = synth.merge_table_segments(jpg, fill)
            a, b = hcj.header_decode(merged, hcj.FLAG_T81_TABLES), orc.header_decode(merged, orc.FLAG_T81_TABLES)
            _same_header(a, b)
            ref = hcj.header_decode(jpg)
            assert a.n_quant_tables == ref.n_quant_tables == 2 and a.n_huffman_tables == ref.n_huffman_tables == 4
            for i in range(2):
                assert list(a.quant_tables[i].elements) == list(ref.quant_tables[i].elements)
            for i in range(4):
                assert list(a.huffman_tables[i].lengths) == list(ref.huffman_tables[i].lengths)
            assert orc.decode(merged, t81_tables=True).yuv() == orc.decode(jpg).yuv()
            f1, f2 = hcj.frame_info(merged, hcj.FLAG_DEFAULT | hcj.FLAG_T81_TABLES), hcj.frame_info(jpg)
            assert (f1.nblocks, f1.restart_interval, f1.yuv_bytes) == (f2.nblocks, f2.restart_interval, f2.yuv_bytes)
            # model semantics on the same bytes: whatever the oracle does, the library does
            try:
                o = orc.header_decode(merged)
            except orc.OracleError as e:
                with pytest.raises(hcj.HcjError) as ei:
                    hcj.header_decode(merged)
                assert ei.value.status == e.status
            else:
                _same_header(hcj.header_decode(merged), o)
        # the flag changes nothing for a file the model reads
        _same_header(hcj.header_decode(jpg, hcj.FLAG_T81_TABLES), orc.header_decode(jpg))


def mjpeg_stream(orc, data):
    js = [orc.encode(synth.frame(i, 64, 48, c), 64, 48, c, 75, restart_interval=ri) for i, (c, ri) in enumerate([(420, 0), (444, 2), (422, 0), (420, 1)])]
    js.append(data("Mouse480.jpg"))
    js.append(synth.merge_table_segments(js[0], fill=2))
    stream, want = b"", []
    for i, j in enumerate(js):
        stream += b"\x00\xff\x12garbage" * (i % 2)
        want.append((len(stream), len(j)))
        stream += j
    return js, stream, want


def test_mjpeg_split(hcj, orc, data):
    js, stream, want = mjpeg_stream(orc, data)
    assert hcj.mjpeg_split(stream) == want
    assert hcj.mjpeg_split(stream + b"\xff\xd8\xff\xe0\x00\x04ab") == want  # a frame cut short at the end is dropped
    assert hcj.mjpeg_split(stream[: want[-1][0] + 100]) == want[:-1]
    assert hcj.mjpeg_split(b"") == [] and hcj.mjpeg_split(b"\xff\xd8\xff\xd9") == [(0, 4)]
    n = C.c_int()
    off, ln = (C.c_size_t * 2)(), (C.c_size_t * 2)()
    assert hcj.lib().hcj_mjpeg_split(stream, len(stream), off, ln, 2, C.byref(n)) == -30 and n.value == len(want)  # BUFFER_TOO_SMALL
    assert [(off[i], ln[i]) for i in range(2)] == want[:2]


def test_size_and_magnitude_goldens(hcj, goldens):
    """jpeg/model/test/test_encode_codewords.ml through the library's own scalar functions (the ones the kernels call)."""
    L = hcj.lib()
    for r in goldens["size_ranges"]["rows"]:
        assert L.hcj_size(r["lo"]) == r["size_lo"] and L.hcj_size(r["hi"]) == r["size_hi"] and L.hcj_size(-r["hi"]) == r["size_hi"]
    for r in goldens["magnitude"]["rows"]:
        size = L.hcj_size(r["value"])
        assert size == r["size"] and L.hcj_magnitude(size, r["value"]) == r["emag"] and L.hcj_mag(size, r["emag"]) == r["dmag"]
    from hcjpeg import model

    assert model.Decoder.For_testing.mag(4, 0) == -15 and model.Decoder.For_testing.mag(4, 15) == 15


def test_quant_scale_and_code_table_goldens(hcj, goldens):
    """test_quant_tables.ml and test_tables.ml through the library's host functions (what the kernels are given)."""
    L = hcj.lib()
    for q, want in goldens["quant_scale_luma"]["by_quality"].items():
        out = np.zeros(64, np.uint16)
        assert L.hcj_quant_scale(0, int(q), out.ctypes.data) == 0 and out.tolist() == want
    t = goldens["encoder_tables"]["tables"]
    bits, length = C.c_int(), C.c_int()
    for table, name in enumerate(["dc_luma", "dc_chroma"]):
        for e in t[name]:
            assert L.hcj_encoder_code(table, 0, e["data"], C.byref(bits), C.byref(length)) == 0
            assert (length.value, bits.value) == (e["length"], e["bits"]), (name, e)
    for table, name in enumerate(["ac_luma", "ac_chroma"]):
        seen = set()
        for row in t[name]:
            for e in row:
                if e["length"] == 0:  # filler of the model's [run][size] arrays, not a code
                    continue
                assert L.hcj_encoder_code(2 + table, e["run"], e["size"], C.byref(bits), C.byref(length)) == 0
                assert (length.value, bits.value) == (e["length"], e["bits"]), (name, e)
                seen.add((e["run"], e["size"]))
        assert len(seen) == 162
        for run in range(16):  # and nothing else has a code
            for size in range(16):
                L.hcj_encoder_code(2 + table, run, size, C.byref(bits), C.byref(length))
                assert (length.value != 0) == ((run, size) in seen)


def test_header_errors_match_oracle(hcj, orc, data):
    jpg = data("mini.jpg")
    cases = [
        jpg[:100],  # truncated header: find_marker never returns in the model
        jpg[:2] + b"\xff\xc2" + jpg[4:],  # progressive SOF -> unsupported marker code
        b"",
        b"\xff\xd8\xff\xda\x00\x02",
    ]
    for c in cases:
        with pytest.raises(orc.OracleError) as eo:
            orc.decode(c)
        f = hcj.FrameInfo()
        st = hcj.lib().hcj_frame_info_get(c, len(c), C.byref(f))
        assert st == eo.value.status, (c[:8], st, eo.value.status)


def test_write_headers_golden(hcj, goldens, orc):
    assert hcj.write_headers(480, 320, 420, 20).hex() == goldens["header_480x320_q20_420"]["hex"]
    for chroma in (420, 422, 444):
        for q in (1, 50, 100):
            for ri in (0, 8):
                assert hcj.write_headers(1920, 1080, chroma, q, ri) == orc.write_headers(1920, 1080, chroma, q, ri)


def test_mirror_interface_shapes(hcj):
    """The host-side mirror exposes the reference's names (decoder.mli / encoder.mli / frame.mli)."""
    from hcjpeg import model

    for name in ("decode_a_frame", "init", "decode", "get_decoded_planes", "get_yuv_frame", "Header", "For_testing"):
        assert hasattr(model.Decoder, name)
    for name in ("encode_420", "encode_422", "encode_444", "encode_monochrome", "write_headers", "Parameters"):
        assert hasattr(model.Encoder, name)
    f = model.Frame.create(420, 64, 48)
    assert (f.u.width, f.u.height) == (32, 24) and f.width == 64 and f.height == 48
    g = model.Frame.frombytes(synth.frame(1, 64, 48, 422), 422, 64, 48)
    assert (g.u.width, g.u.height) == (32, 48)
