"""Pin the CPU oracle against every golden the reference's own tests hold for the hot path.

Each test cites the reference test it restates (paths relative to the reference root).
"""
import hashlib

import numpy as np
import pytest


def sha(b):
    return hashlib.sha256(b).hexdigest()


# ---- jpeg/model/test/test_bits.ml:24-38 ------------------------------------------------------
def test_bits_roundtrip(orc):
    import ctypes as C

    L = orc.lib()
    rng = np.random.default_rng(0)
    vals = [(int(b), int(rng.integers(0, 1 << b))) for b in rng.integers(1, 17, 10000)]
    for stuffing in (0, 1):
        w = orc.Writer()
        L.orc_writer_create(C.byref(w))
        for bits, v in vals:
            L.orc_writer_put_bits(C.byref(w), stuffing, v, bits)
        L.orc_writer_flush_with_1s(C.byref(w), stuffing)
        buf = bytes(bytearray(w.buffer[i] for i in range(w.bytes_written)))
        L.orc_writer_free(C.byref(w))
        if stuffing:
            assert b"\xff" in buf
            j = 0
            while True:  # every FF is followed by 00
                j = buf.find(b"\xff", j)
                if j < 0:
                    break
                assert buf[j + 1] == 0
                j += 2
            buf = buf.replace(b"\xff\x00", b"\xff")
        r = orc.Bits()
        L.orc_bits_create(C.byref(r), buf, len(buf))
        out = C.c_int64()
        for bits, v in vals:
            assert L.orc_bits_get(C.byref(r), bits, C.byref(out)) == 0
            assert out.value == v


def test_bits_zero_extend_and_bounds(orc):
    """bitstream_reader.ml:19-22 (reads past the end are 0) and :32 (raise iff n >= total bits)."""
    import ctypes as C

    L = orc.lib()
    r = orc.Bits()
    buf = b"\xab\xcd"
    L.orc_bits_create(C.byref(r), buf, 2)
    out = C.c_int64()
    assert L.orc_bits_show(C.byref(r), 16, C.byref(out)) == -9
    assert L.orc_bits_get(C.byref(r), 12, C.byref(out)) == 0 and out.value == 0xABC
    assert L.orc_bits_get(C.byref(r), 12, C.byref(out)) == 0 and out.value == 0xD00


# ---- jpeg/model/test/test_chen_dct.ml:47-94 --------------------------------------------------
def test_chen_example(orc, goldens):
    g = goldens["chen_example"]
    x = np.array(g["input"], np.int64)
    f = orc.chen_forward(x)
    f4 = np.where(f > 0, (f + 2) // 4, -((-(f - 2)) // 4))  # OCaml '/' truncates toward zero
    assert f4.tolist() == g["fdct"]
    assert orc.chen_inverse(f4).tolist() == g["idct"]


def test_chen_roundtrip_within_2(orc):
    rng = np.random.default_rng(1)
    for _ in range(100):
        x = rng.integers(-128, 128, 64)
        f = orc.chen_forward(x)
        f4 = np.where(f > 0, (f + 2) // 4, -((-(f - 2)) // 4))
        assert np.abs(orc.chen_inverse(f4) - x).max() <= 2


# ---- jpeg/model/test/test_encode_headers.ml:17-133 -------------------------------------------
def test_header_hexdump(orc, goldens):
    g = goldens["header_480x320_q20_420"]
    assert orc.write_headers(480, 320, 420, 20).hex() == g["hex"]


def _check_header(h, p):
    assert (h.frame.width, h.frame.height, h.frame.sample_precision, h.frame.number_of_components) == (
        p["width"],
        p["height"],
        p["precision"],
        p["ncomp"],
    )
    for i, c in enumerate(p["components"]):
        hc = h.frame.components[i]
        assert (hc.identifier, hc.h, hc.v, hc.tq) == (c["id"], c["h"], c["v"], c["tq"])
    assert h.n_quant_tables == len(p["quant_tables"])
    for i, q in enumerate(p["quant_tables"]):  # list order: newest first
        hq = h.quant_tables[i]
        assert (hq.element_precision, hq.table_identifier) == (q["precision"], q["id"])
        assert list(hq.elements) == q["elements"]
    assert h.n_huffman_tables == len(p["huffman_tables"])
    for i, t in enumerate(p["huffman_tables"]):
        ht = h.huffman_tables[i]
        assert (ht.length, ht.table_class, ht.destination_identifier) == (t["length"], t["class"], t["id"])
        assert list(ht.lengths) == t["lengths"]
        assert list(ht.values)[: ht.nvalues] == t["values"]
    for i, s in enumerate(p["scan"]):
        hs = h.scan.scan_components[i]
        assert (hs.selector, hs.dc, hs.ac) == (s["selector"], s["dc"], s["ac"])
    assert bool(h.restart_interval_present) == p["restart_interval_present"]


def test_header_reparse(orc, goldens):
    g = goldens["header_480x320_q20_420"]
    _check_header(orc.header_decode(bytes.fromhex(g["hex"])), g["parsed"])


# ---- jpeg/hardcaml/test/test_codeblock_decoder.ml:118-168 ------------------------------------
def test_mouse480_header_and_entropy(orc, goldens, data):
    g = goldens["mouse480"]
    jpg = data("Mouse480.jpg")
    h = orc.header_decode(jpg)
    _check_header(h, g["parsed"])
    ent = orc.extract_entropy_coded_bits(jpg, h.scan_bit_pos // 8)
    assert ent[:64].hex() == g["entropy_first64_hex"]
    assert len(ent) == 6281
    assert sha(ent) == "5daa43a6323e8df1e2b04baff0f22770c85e60039693dc7bbf10a59272aac56a"


# ---- jpeg/model/test/test_tables.ml:4-397 ----------------------------------------------------
def test_encoder_tables(orc, goldens):
    t = goldens["encoder_tables"]["tables"]
    for which, name in enumerate(["dc_luma", "dc_chroma"]):
        assert orc.encoder_dc_table(which) == [(e["length"], e["bits"], e["data"]) for e in t[name]]
    for which, name in enumerate(["ac_luma", "ac_chroma"]):
        want = [[(e["length"], e["bits"], e["run"], e["size"]) for e in row] for row in t[name]]
        assert orc.encoder_ac_table(which) == want


# ---- jpeg/model/test/test_quant_tables.ml ----------------------------------------------------
def test_quant_scale(orc, goldens):
    for q, want in goldens["quant_scale_luma"]["by_quality"].items():
        assert orc.quant_scale(False, int(q)).tolist() == want
    assert sorted(goldens["quant_scale_luma"]["by_quality"]) == ["1", "100", "25", "50", "75", "95"]


def test_zigzag_self_inverse(orc):
    import ctypes as C

    L = orc.lib()
    inv = list((C.c_int * 64).in_dll(L, "orc_zigzag_inverse"))
    fwd = list((C.c_int * 64).in_dll(L, "orc_zigzag_forward"))
    assert sorted(inv) == list(range(64))
    assert [fwd[inv[i]] for i in range(64)] == list(range(64))
    assert inv[:8] == [0, 1, 8, 16, 9, 2, 3, 10]


# ---- jpeg/model/test/test_encode_codewords.ml ------------------------------------------------
def test_size_and_magnitude(orc, goldens):
    L = orc.lib()
    for r in goldens["size_ranges"]["rows"]:
        assert L.orc_size(r["lo"]) == r["size_lo"] and L.orc_size(r["hi"]) == r["size_hi"]
        assert L.orc_size(-r["hi"]) == r["size_hi"]
    for r in goldens["magnitude"]["rows"]:
        size = L.orc_size(r["value"])
        assert size == r["size"]
        assert L.orc_magnitude(size, r["value"]) == r["emag"]
        assert L.orc_mag(size, r["emag"]) == r["dmag"]
    assert len(goldens["magnitude"]["rows"]) == 31


# ---- jpeg/model/test/test_rle.ml -------------------------------------------------------------
def test_rle_cases(orc, goldens):
    cases = goldens["rle_cases"]["cases"]
    assert len(cases) == 9
    for c in cases:
        q = np.zeros(64, np.int64)
        for pos, v in c["set"]:
            q[pos] = v
        got, pred = orc.rle(q)
        assert got == [tuple(x) for x in c["rle"]], c["name"]
        assert pred == q[0]


def test_rle_random_roundtrip(orc):
    """test_rle.ml:95-130: run-length decode of rle(q) reproduces q (dc relative to pred)."""
    rng = np.random.default_rng(2)
    for _ in range(2000):
        q = np.where(rng.random(64) < rng.random(), rng.integers(-50, 50, 64), 0)
        got, _ = orc.rle(q, dc_pred=3)
        out = []
        for run, v in got:
            out += [0] * run + [v]
        assert len(out) == 64
        out[0] += 3
        assert out == q.tolist()


# ---- jpeg/test_data/mini.jpg == encode_420(mini64x64.420, q75) -------------------------------
def test_mini_jpg_byte_identity(orc, data):
    assert orc.encode(data("mini64x64.420"), 64, 64, 420, 75) == data("mini.jpg")


SURVEY_HASHES = {  # SURVEY.md Appendix B.2 (probe-derived regression hashes)
    (420, 99): (3196, "de2723d7a357cb52e8994a11b088bfa9e360814b473f3ded57816658b3cd48e9"),
    (420, 95): (2122, "dde5cb3b85058f8c1b7579a62448caade5532034ff6f94fbea5a4a5c7be8f2d9"),
    (420, 75): (1255, "7e96dbcf01a8691076dfec4aa59ca18632dc48aa62f9b11bff484c585189fdbf"),
    (420, 70): (1194, "7e9363898cc65fc91b573179cc3ecdcfcdee74831c5b93f2eb1bf8cbde5c9a5e"),
    (420, 50): (1033, "9c2c941c21bc3b16dbdd1c928cbab40d0c39df0084608a00fdacabea7bd05dbf"),
    (420, 30): (924, "ec088cc041eade85e56f3227cbc9e9b4042ea2731b42470ea0658bdc1a43383d"),
    (420, 20): (855, "758a7a585db3290c3e2f572252ed9d0c7bff76a1affc47b3eb9a901148382555"),
    (420, 10): (775, "6b081d4934c3508c048e8cbd894a0f044581419d0b8ea8e4fb0019925c1ec397"),
    (422, 75): (1286, "d1f108b7a2da6069f10918578229dbf7cc68b97248b8ccfa82e6d990960bb84e"),
    (444, 75): (1376, "ebb97f71791e5708fa693d645267dd6245153eae491cf8e80bb3533cf54746f1"),
}


@pytest.mark.parametrize("chroma,q", sorted(SURVEY_HASHES))
def test_encoder_regression_hashes(orc, data, chroma, q):
    out = orc.encode(data("mini64x64.%d" % chroma), 64, 64, chroma, q)
    assert (len(out), sha(out)) == SURVEY_HASHES[(chroma, q)]


# ---- jpeg/test/model-encode-and-decode.t, test-nonstandard-sizes.t ---------------------------
def _ocaml_float(x):
    """OCaml's print of a float: %.15g if it round-trips, else %.17g."""
    s = "%.15g" % x
    return s if float(s) == x else "%.17g" % x


def test_cram_psnr(orc, goldens, data):
    for run in goldens["cram_encode_decode"]["runs"]:
        chroma = int(run["chroma"])
        src = data(run["input"])
        jpg = orc.encode(src, 64, 64, chroma, run["quality"])
        dec = orc.decode(jpg)
        assert dec.chroma == chroma
        planes = orc.split_yuv(src, 64, 64, chroma)
        got = [_ocaml_float(orc.psnr(planes[i], dec.cropped[i])) for i in range(3)]
        assert got == run["psnr"], (run, got)


def test_cram_52x44(orc, goldens, data):
    """oyuv convert 64x64 -> 52x44 (420 -> 444 -> crop -> 420), encode q95, decode, PSNR."""
    g = goldens["cram_52x44"]
    y, u, v = orc.split_yuv(data("mini64x64.420"), 64, 64, 420)
    y4, u4, v4 = orc.upsample_to_444((y, u, v), 420)
    w, h = g["size"]
    yc, uc, vc = (orc.crop_clamp(p, w, h) for p in (y4, u4, v4))
    src = yc.tobytes() + orc.subsample_hv2(uc).tobytes() + orc.subsample_hv2(vc).tobytes()
    jpg = orc.encode(src, w, h, 420, g["quality"])
    assert len(jpg) == 1923
    dec = orc.decode(jpg)
    assert dec.actual_size == [(52, 44), (26, 22), (26, 22)]
    planes = orc.split_yuv(src, w, h, 420)
    got = [_ocaml_float(orc.psnr(planes[i], dec.cropped[i])) for i in range(3)]
    assert got == g["psnr"]


# ---- jpeg/test/mouse-decode.t (ffmpeg absent; libjpeg-turbo via Pillow is the stand-in) -------
def test_mouse480_decode(orc, data):
    dec = orc.decode(data("Mouse480.jpg"), want_blocks=True)
    assert sha(dec.yuv()) == "f17981ec39aee6fb10ea5fa078b397ba4df460cab8eadbbce98a916f21b2b97d"
    assert dec.nblocks == 3600
    assert dec.coefs[0, :6].tolist() == [20, 1, -2, -1, 1, 0]
    assert int((dec.coefs[:, 1:] != 0).sum()) == 7397
    assert sha(dec.coefs.astype("<i2").tobytes()) == "a46d612bafcd7f0a94297575c8b8ca1f2fb0ac1541f5fc7a991959472db4bf9e"
    assert sha(dec.coefs_abs_dc().astype("<i2").tobytes()) == "1f750d078cb2a5a395cd24b6b34befce4d4078efbb96c1caea01fcc5ae9d1356"
    assert sha(dec.dequant.astype("<i4").tobytes()) == "9e48edcfa960ae2f54b61aa7612bbc779cd46af553720ef28472ca259ae845be"


def test_mouse480_vs_libjpeg(orc, data):
    Image = pytest.importorskip("PIL.Image")
    import io

    im = Image.open(io.BytesIO(data("Mouse480.jpg")))
    im.draft("YCbCr", im.size)
    ycc = np.asarray(im.convert("YCbCr"))
    dec = orc.decode(data("Mouse480.jpg"))
    assert orc.max_difference(dec.cropped[0], ycc[:, :, 0]) <= 1  # mouse-decode.t:10 (Y = 1)


def test_mini_jpg_decode(orc, data):
    dec = orc.decode(data("mini.jpg"))
    assert sha(dec.yuv()) == "0e85b2f317212070b18c48f0190d5ffeb80913aa9447755c9407ff4bcb6098b6"
    planes = orc.split_yuv(data("mini64x64.420"), 64, 64, 420)
    assert [orc.square_error(planes[i], dec.cropped[i]) for i in range(3)] == [32263, 5091, 4713]


# ---- tools/src/planar_444.ml:139-249 ---------------------------------------------------------
def test_planar_444_vectors(orc, goldens):
    d = goldens["planar_444"]["dumps"]
    f444, f422, back = d["444<->422"]
    planes = [np.array(f444[i * 16 : (i + 1) * 16], np.uint8).reshape(4, 4) for i in range(3)]
    sub = [planes[0]] + [orc.subsample_h2(p) for p in planes[1:]]
    assert np.concatenate([p.ravel() for p in sub]).tolist() == f422
    up = [sub[0]] + [orc.supersample_h2(p) for p in sub[1:]]
    assert np.concatenate([p.ravel() for p in up]).tolist() == back
    f444, f420, back = d["444<->420"]
    sub = [planes[0]] + [orc.subsample_hv2(p) for p in planes[1:]]
    assert np.concatenate([p.ravel() for p in sub]).tolist() == f420
    up = [sub[0]] + [orc.supersample_hv2(p) for p in sub[1:]]
    assert np.concatenate([p.ravel() for p in up]).tolist() == back


def test_oracle_against_independent_decoder(orc, data):
    """External anchor for the oracle itself: libjpeg-turbo's luma is within one LSB of the oracle's on the
    reference's own test image and on oracle-encoded frames, as the reference's mouse-decode.t reports against ffmpeg."""
    Image = pytest.importorskip("PIL.Image")
    import io

    import synth

    files = [data("Mouse480.jpg")] + [orc.encode(synth.frame(900 + i, 160, 96, c), 160, 96, c, q) for i, (c, q) in enumerate(((420, 75), (444, 95), (422, 30)))]
    for j in files:
        im = Image.open(io.BytesIO(j))
        im.draft("YCbCr", im.size)
        ref = np.asarray(im).astype(np.int32)
        y = orc.decode(j).cropped[0].astype(np.int32)
        assert np.abs(y - ref[:, :, 0]).max() <= 1


def test_rgb_formula_pinned_by_pillow_and_opencv(orc):
    """D13: YCbCr -> RGB is absent from the reference, so the stated formula (JFIF full range, libjpeg's
    16-bit constants) is pinned against two independent implementations over the WHOLE (Y, Cb, Cr) cube: Pillow's
    YCbCr -> RGB conversion and OpenCV's COLOR_YCrCb2RGB agree with it to +-1 everywhere (SURVEY 8c ii)."""
    Image = pytest.importorskip("PIL.Image")
    cv2 = pytest.importorskip("cv2")
    cb, cr = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    worst_pil = worst_cv = 0
    for y in range(256):
        Y = np.full_like(cb, y)
        got = orc.ycbcr_to_rgb24(Y, cb, cr).astype(np.int32)
        pil = np.asarray(Image.merge("YCbCr", [Image.fromarray(Y), Image.fromarray(cb), Image.fromarray(cr)]).convert("RGB"))
        ocv = cv2.cvtColor(np.dstack([Y, cr, cb]), cv2.COLOR_YCrCb2RGB)
        worst_pil = max(worst_pil, int(np.abs(got - pil.astype(np.int32)).max()))
        worst_cv = max(worst_cv, int(np.abs(got - ocv.astype(np.int32)).max()))
    assert worst_pil <= 1 and worst_cv <= 1, (worst_pil, worst_cv)
