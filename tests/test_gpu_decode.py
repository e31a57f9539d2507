"""GPU parity tests of the decode path: every call goes through the C ABI (libhcjpeg.so) and is compared
bit-for-bit with the CPU oracle on the same inputs, plus the reference's golden fixtures."""
import hashlib
import io

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu


def sha(b):
    return hashlib.sha256(bytes(b)).hexdigest()


@pytest.fixture(scope="module")
def hcj():
    import hcjpeg

    assert hcjpeg.lib() is not None
    return hcjpeg


@pytest.fixture(scope="module")
def ctx(hcj):
    c = hcj.Context(0)
    yield c
    c.close()


def oracle_rgb(orc, dec):
    y, u, v = orc.upsample_to_444(dec.cropped[:3], dec.chroma)
    h, w = y.shape

    def fit(p):  # Planar_444 leaves the odd last column / row of a fresh (zero) plane untouched
        out = np.zeros((h, w), np.uint8)
        out[: min(h, p.shape[0]), : min(w, p.shape[1])] = p[:h, :w]
        return out

    return orc.ycbcr_to_rgb24(y, fit(u), fit(v))


# ---- config 0: the reference's own fixtures -----------------------------------------------------------
def test_mouse480_golden(hcj, ctx, orc, data):
    """jpeg/test/mouse-decode.t + SURVEY B.2 hashes, through the model-shaped mirror."""
    from hcjpeg.model import Decoder

    jpg = data("Mouse480.jpg")
    frame = Decoder.decode_a_frame(jpg, ctx)
    out = io.BytesIO()
    frame.output(out)
    assert sha(out.getvalue()) == "f17981ec39aee6fb10ea5fa078b397ba4df460cab8eadbbce98a916f21b2b97d"
    assert out.getvalue() == orc.decode(jpg).yuv()
    assert frame.chroma_subsampling == 420 and (frame.width, frame.height) == (480, 320)
    ent = Decoder.For_testing.extract_entropy_coded_bits(jpg, ctx)
    assert len(ent) == 6281 and sha(ent) == "5daa43a6323e8df1e2b04baff0f22770c85e60039693dc7bbf10a59272aac56a"
    coefs = Decoder.For_testing.coefficients(jpg, ctx)
    assert sha(coefs.astype("<i2").tobytes()) == "1f750d078cb2a5a395cd24b6b34befce4d4078efbb96c1caea01fcc5ae9d1356"
    d = Decoder(jpg, ctx)
    planes = d.get_decoded_planes()
    want = orc.decode(jpg).planes
    assert [p.plane.shape for p in planes] == [w.shape for w in want]
    assert all(np.array_equal(p.plane, w) for p, w in zip(planes, want))


def test_decode_a_frame_entry(hcj, ctx, orc, data):
    jpg = data("Mouse480.jpg")
    assert bytes(ctx.decode_a_frame(jpg)) == orc.decode(jpg).yuv()
    with pytest.raises(hcj.HcjError) as e:
        ctx.decode_a_frame(jpg[:2] + b"\xff\xc2" + jpg[4:])  # progressive SOF: "unsupported marker code"
    assert e.value.status == -1


def test_mini_jpg_golden(hcj, ctx, orc, data):
    outs, st = ctx.decode_batch([data("mini.jpg")])
    assert st == [0]
    assert sha(outs[0]) == "0e85b2f317212070b18c48f0190d5ffeb80913aa9447755c9407ff4bcb6098b6"


def test_cram_psnr_flow(hcj, ctx, orc, goldens, data):
    """jpeg/test/model-encode-and-decode.t re-run as oracle-encode -> GPU-decode: exact SSE goldens."""
    sse = {(420, 95): (5605, 1404, 1166), (420, 50): (64890, 9410, 7446), (420, 30): (113632, 11096, 8747),
           (422, 75): (32263, 8169, 6414), (444, 75): (32263, 10908, 9357)}
    for run in goldens["cram_encode_decode"]["runs"]:
        chroma, q = int(run["chroma"]), run["quality"]
        src = data(run["input"])
        outs, st = ctx.decode_batch([orc.encode(src, 64, 64, chroma, q)])
        assert st == [0]
        a, b = np.frombuffer(src, np.uint8), outs[0]
        y = 64 * 64
        c = (len(src) - y) // 2
        got = tuple(int(((a[s].astype(int) - b[s].astype(int)) ** 2).sum()) for s in (slice(0, y), slice(y, y + c), slice(y + c, y + 2 * c)))
        assert got == sse[(chroma, q)]
        dev_sse, dev_max = ctx.compare_planes(a[:y], b[:y])
        assert dev_sse == got[0] and dev_max == int(np.abs(a[:y].astype(int) - b[:y].astype(int)).max())


def test_yuv_convert_on_device(hcj, ctx, orc, data):
    """`oyuv convert` (oconv.ml:111-133) on the device against the oracle's Planar_444 / Yuv.crop chain, including the
    conversion the reference's odd-size cram test does (64 x 64 4:2:0 -> 52 x 44 4:2:0 through 4:4:4)."""
    def chain(frame, w, h, chroma, dw, dh, dchroma, x, y):
        planes = orc.split_yuv(frame, w, h, chroma)
        up = orc.upsample_to_444(planes, chroma)
        full = []
        for p in up:  # a fresh (zero) 4:4:4 frame: an odd last column / row is never written
            f = np.zeros((h, w), np.uint8)
            f[: min(h, p.shape[0]), : min(w, p.shape[1])] = p[:h, :w]
            full.append(orc.crop_clamp(f, dw, dh, x, y))
        if dchroma == 420:
            full[1:] = [orc.subsample_hv2(p) for p in full[1:]]
        elif dchroma == 422:
            full[1:] = [orc.subsample_h2(p) for p in full[1:]]
        return b"".join(p.tobytes() for p in full)

    src = data("mini64x64.420")
    assert bytes(ctx.yuv_convert(src, 64, 64, 420, 52, 44, 420)) == chain(src, 64, 64, 420, 52, 44, 420, 0, 0)
    rng = np.random.default_rng(3)
    for chroma, w, h in ((420, 64, 48), (422, 61, 35), (444, 33, 20), (420, 51, 37), (422, 8, 8)):
        f = synth.frame(int(rng.integers(1000)), w, h, chroma)
        for dchroma in (420, 422, 444):
            for dw, dh, x, y in ((w, h, 0, 0), (w + 9, h + 5, -3, -2), (max(2, w - 10), max(2, h - 7), 4, 3), (17, 11, w - 5, h - 4)):
                got = ctx.yuv_convert(f, w, h, chroma, dw, dh, dchroma, x, y)
                assert bytes(got) == chain(f, w, h, chroma, dw, dh, dchroma, x, y), (chroma, w, h, dchroma, dw, dh, x, y)


def test_cram_52x44_flow(hcj, ctx, orc, goldens, data):
    """jpeg/test/test-nonstandard-sizes.t:3-15 with the encoder, the decoder and the comparison on the device: a 52 x 44
    frame (partial MCUs: zero padding on encode, crop on decode) at q95 gives the reference's 1923 bytes and PSNR strings."""
    g = goldens["cram_52x44"]
    y, u, v = orc.split_yuv(data("mini64x64.420"), 64, 64, 420)
    y4, u4, v4 = orc.upsample_to_444((y, u, v), 420)
    w, h = g["size"]
    yc, uc, vc = (orc.crop_clamp(p, w, h) for p in (y4, u4, v4))
    src = yc.tobytes() + orc.subsample_hv2(uc).tobytes() + orc.subsample_hv2(vc).tobytes()
    enc, st = ctx.encode_batch([src], w, h, 420, g["quality"])
    assert st == [0] and len(enc[0]) == 1923 and enc[0] == orc.encode(src, w, h, 420, g["quality"])
    with ctx.batch(enc, hcj.OUT_YUV) as b:
        b.decode()
        m = b.compare([src])[0]
        assert [m.samples[k] for k in range(3)] == [52 * 44, 26 * 22, 26 * 22]
        for k in range(3):
            assert abs(m.psnr(k) - float(g["psnr"][k])) < 1e-9


# ---- oracle parity on seeded synthetic images -----------------------------------------------------------
CASES = [
    # (chroma, quality, w, h, restart_interval)
    (420, 75, 256, 192, 0), (420, 75, 256, 192, 8), (422, 50, 200, 120, 0), (444, 95, 160, 96, 0),
    (420, 10, 333, 77, 0), (420, 95, 52, 44, 0), (444, 75, 17, 9, 0), (420, 75, 16, 16, 1),
    (422, 75, 130, 70, 3), (444, 100, 64, 64, 2), (420, 1, 96, 80, 0), (420, 75, 8, 8, 0),
]


def oracle_yuv444(orc, dec):
    """Planar_444.convert_from_420 / _422 of the cropped frame into a fresh (zero) 4:4:4 frame."""
    y, u, v = orc.upsample_to_444(dec.cropped[:3], dec.chroma)
    h, w = y.shape
    out = np.zeros((3, h, w), np.uint8)
    for k, p in enumerate((y, u, v)):
        out[k, : min(h, p.shape[0]), : min(w, p.shape[1])] = p[:h, :w]
    return out


@pytest.mark.parametrize("mode_name", ["yuv", "planes", "rgb", "yuv444"])
def test_batch_matches_oracle(hcj, ctx, orc, mode_name):
    mode = {"yuv": hcj.OUT_YUV, "planes": hcj.OUT_PLANES, "rgb": hcj.OUT_RGB24, "yuv444": hcj.OUT_YUV444}[mode_name]
    jpgs = [orc.encode(synth.frame(100 + i, w, h, c), w, h, c, q, restart_interval=ri) for i, (c, q, w, h, ri) in enumerate(CASES)]
    outs, st = ctx.decode_batch(jpgs, mode)
    assert st == [0] * len(jpgs)
    for jpg, out, case in zip(jpgs, outs, CASES):
        dec = orc.decode(jpg)
        if mode == hcj.OUT_YUV:
            want = dec.yuv()
        elif mode == hcj.OUT_PLANES:
            want = b"".join(p.tobytes() for p in dec.planes)
        elif mode == hcj.OUT_YUV444:
            want = oracle_yuv444(orc, dec).tobytes()
        else:
            want = oracle_rgb(orc, dec).tobytes()
        assert bytes(out) == want, case


def test_batch_compare_on_device(hcj, ctx, orc, goldens, data):
    """`oyuv compare` for a whole batch on the device (hcj_batch_compare): the reference's cram PSNR strings
    (jpeg/test/model-encode-and-decode.t) and exact integer metrics against numpy on seeded frames."""
    runs = goldens["cram_encode_decode"]["runs"]
    srcs = [data(r["input"]) for r in runs]
    jpgs = [orc.encode(s, 64, 64, int(r["chroma"]), r["quality"]) for s, r in zip(srcs, runs)]
    with ctx.batch(jpgs, hcj.OUT_YUV) as b:
        b.decode()
        for m, r in zip(b.compare(srcs), runs):
            assert m.status == 0
            for k in range(3):
                assert abs(m.psnr(k) - float(r["psnr"][k])) < 1e-9, (r, k, m.psnr(k))
    frames = [synth.frame(100 + i, w, h, c) for i, (c, q, w, h, ri) in enumerate(CASES)]
    jpgs = [orc.encode(f, w, h, c, q, restart_interval=ri) for f, (c, q, w, h, ri) in zip(frames, CASES)]
    with ctx.batch(jpgs + [b"junk"], hcj.OUT_YUV) as b:
        b.decode()
        outs, st = b.fetch()
        ms = b.compare(frames + [None])
        assert ms[-1].status != 0 and st[-1] != 0
        for f, o, m, (c, q, w, h, ri) in zip(frames, outs, ms, CASES):
            a, d = np.frombuffer(f, np.uint8).astype(np.int64), o.astype(np.int64)
            cw, ch = (w if c == 444 else w // 2), (h // 2 if c == 420 else h)
            cuts = [0, w * h, w * h + cw * ch, w * h + 2 * cw * ch]
            for k in range(3):
                dlt = np.abs(a[cuts[k]:cuts[k + 1]] - d[cuts[k]:cuts[k + 1]])
                assert m.samples[k] == cuts[k + 1] - cuts[k]
                assert (m.square_error[k], m.total_difference[k], m.max_difference[k]) == (int((dlt * dlt).sum()), int(dlt.sum()), int(dlt.max()))
    # a reference of the wrong size is reported per image
    with ctx.batch(jpgs[:2], hcj.OUT_YUV444) as b:
        b.decode()
        ms = b.compare([oracle_yuv444(orc, orc.decode(jpgs[0])).tobytes(), frames[1] + b"x"])
        assert ms[0].status == 0 and list(ms[0].square_error)[:3] == [0, 0, 0] and list(ms[0].samples)[:3] == [256 * 192] * 3
        assert ms[0].psnr(0) == float("inf")
        assert ms[1].status == -31  # HCJ_ERR_INVALID_ARG: reference of the wrong size


def test_rgb_vector_paths(hcj, ctx, orc):
    """RGB24 from sub-sampled chroma with 16-pixel-aligned rows (the register path of k_rgb): 4:2:0 and 4:2:2, even and
    odd heights (the last luma row of an odd-height 4:2:0 image has no chroma row), single and many 16-pixel groups."""
    cases = [(420, 64, 48), (420, 48, 35), (420, 16, 2), (420, 16, 3), (420, 32, 5), (422, 64, 48), (422, 32, 19), (422, 16, 3),
             (420, 320, 200), (422, 320, 203), (420, 1920, 1080)]  # (heights of 1 mod 16 are sizes the model's encoder rejects)
    jpgs = [orc.encode(synth.frame(700 + i, w, h, c), w, h, c, 75, restart_interval=(8 if w > 1000 else 0)) for i, (c, w, h) in enumerate(cases)]
    outs, st = ctx.decode_batch(jpgs, hcj.OUT_RGB24)
    assert st == [0] * len(jpgs)
    for j, o, case in zip(jpgs, outs, cases):
        assert bytes(o) == oracle_rgb(orc, orc.decode(j)).tobytes(), case
    outs, st = ctx.decode_batch(jpgs, hcj.OUT_YUV444)  # the same paths, planar 4:4:4 out
    assert st == [0] * len(jpgs)
    for j, o, case in zip(jpgs, outs, cases):
        assert bytes(o) == oracle_yuv444(orc, orc.decode(j)).tobytes(), case


def test_coefficients_and_entropy_taps(hcj, ctx, orc):
    jpgs = [orc.encode(synth.frame(200 + i, w, h, c), w, h, c, q, restart_interval=ri) for i, (c, q, w, h, ri) in enumerate(CASES[:6])]
    with ctx.batch(jpgs, hcj.OUT_PLANES) as b:
        b.decode()
        for i, jpg in enumerate(jpgs):
            dec = orc.decode(jpg, want_blocks=True)
            assert np.array_equal(b.coefficients(i), dec.coefs_abs_dc().astype(np.int16))
            assert len(b.entropy(i)) == dec.entropy_len
        assert b.kernels() >= 3


def test_block_log_tap(hcj, ctx, orc, data):
    """`model decode log` on the device: Component.Summary of every block against the oracle's per-block taps."""
    jpgs = [orc.encode(synth.frame(300 + i, w, h, c), w, h, c, q, restart_interval=ri) for i, (c, q, w, h, ri) in enumerate(CASES[:5])]
    jpgs.append(data("Mouse480.jpg"))
    with ctx.batch(jpgs, hcj.OUT_PLANES) as b:
        b.decode()
        for i, j in enumerate(jpgs):
            dec = orc.decode(j, want_blocks=True)
            log = b.block_log(i)
            assert len(log) == dec.nblocks
            assert np.array_equal(log["coefs"], dec.coefs) and np.array_equal(log["dc_pred"], dec.dc_abs)
            assert np.array_equal(log["dequant"], dec.dequant) and np.array_equal(log["recon"], dec.recon)
            assert np.array_equal(log["component"], dec.block_comp)
            for k in (0, 1, dec.nblocks // 2, dec.nblocks - 1):
                assert np.array_equal(log["idct"][k], orc.chen_inverse(dec.dequant[k]))
                c = int(dec.block_comp[k])  # the block sits where recon says in the component's padded plane
                x, y = int(log["x"][k]), int(log["y"][k])
                assert np.array_equal(dec.planes[c][y:y + 8, x:x + 8].ravel(), dec.recon[k])
            part = b.block_log(i, 3, 5)
            assert part.tobytes() == log[3:8].tobytes()


def test_restart_extension_equivalence(hcj, ctx, orc):
    """Stated extension pin: decode(DRI stream) == pure-model decode of the same blocks coded without DRI."""
    w, h = 208, 112
    yuv = synth.frame(7, w, h, 420)
    twin = orc.decode(orc.encode(yuv, w, h, 420, 75), restart_ext=False).yuv()
    for ri in (1, 5, 8, 1000):
        outs, st = ctx.decode_batch([orc.encode(yuv, w, h, 420, 75, restart_interval=ri)])
        assert st == [0] and bytes(outs[0]) == twin
    # pure model semantics (flag off) on a DRI stream: the model stops at RST0 and decodes garbage; so do we
    jpg = orc.encode(yuv, w, h, 420, 75, restart_interval=8)
    outs, st = ctx.decode_batch([jpg], flags=0)
    try:
        want = orc.decode(jpg, restart_ext=False).yuv()
        assert st == [0] and bytes(outs[0]) == want
    except orc.OracleError as e:
        assert st == [e.status]


def test_t81_table_segments_decode(hcj, ctx, orc):
    """Merged DQT / DHT segments + fill bytes (HCJ_FLAG_T81_TABLES) decode to the pixels of the plain file; the same
    bytes without the flag fail like the model."""
    plain, merged = [], []
    for i, (chroma, ri) in enumerate(((420, 0), (420, 8), (444, 0), (422, 3))):
        j = orc.encode(synth.frame(40 + i, 200, 120, chroma), 200, 120, chroma, 70, restart_interval=ri)
        plain.append(j)
        merged.append(synth.merge_table_segments(j, fill=i))
    want = [orc.decode(j).yuv() for j in plain]
    outs, st = ctx.decode_batch(merged + plain, hcj.OUT_YUV, hcj.FLAG_DEFAULT | hcj.FLAG_T81_TABLES)
    assert st == [0] * 8
    for o, w in zip(outs, want + want):
        assert bytes(o) == w
    for m in merged:
        assert bytes(orc.decode(m, t81_tables=True).yuv()) == bytes(outs[merged.index(m)])
    outs, st = ctx.decode_batch(merged + plain)
    assert st[:4] == [orc.decode_status(m) for m in merged] and st[4:] == [0] * 4
    assert all(s != 0 for s in st[:4])
    for o, w in zip(outs[4:], want):
        assert bytes(o) == w


def test_mjpeg_stream_decode(hcj, ctx, orc, data):
    """hcj_decode_stream: every frame of a Motion-JPEG stream equals the oracle's decode of that frame."""
    from test_abi import mjpeg_stream

    js, stream, want = mjpeg_stream(orc, data)
    flags = hcj.FLAG_DEFAULT | hcj.FLAG_T81_TABLES
    for mode in (hcj.OUT_YUV, hcj.OUT_RGB24):
        frames, st = ctx.decode_stream(stream, mode, flags)
        assert st == [0] * len(js)
        for j, f in zip(js, frames):
            dec = orc.decode(j, t81_tables=True)
            assert bytes(f) == (dec.yuv() if mode == hcj.OUT_YUV else oracle_rgb(orc, dec).tobytes())
    # a longer stream: 70 frames of mixed geometry, one of them corrupt
    many = [js[i % 4] for i in range(70)]
    many[33] = many[33][:-300] + b"\xff\xd9"  # scan cut short: decodes (zero-extended reader) or fails exactly like the oracle
    many[50] = many[50][:200]  # header cut short: not a frame, the splitter moves on to the next SOI
    frames, st = ctx.decode_stream(b"".join(many))
    good = [m for i, m in enumerate(many) if i != 50]
    assert len(st) == 69 and st == [orc.decode_status(m) for m in good]
    for i in (0, 1, 2, 3, 33, 34, 49, 50, 68):
        if st[i] == 0:
            assert bytes(frames[i]) == orc.decode(good[i]).yuv()


def test_pillow_streams(hcj, ctx, orc):
    """Third-party (libjpeg-turbo) streams: custom Huffman tables, restart markers."""
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(5)
    base = np.clip(rng.normal(128, 40, (120, 168, 3)), 0, 255).astype(np.uint8)
    img = Image.fromarray(base)
    jpgs = []
    for kw in (dict(subsampling=2), dict(subsampling=0, optimize=True), dict(subsampling=2, restart_marker_blocks=3),
               dict(subsampling=1, quality=95, optimize=True), dict(subsampling=2, quality=30, restart_marker_blocks=1)):
        buf = io.BytesIO()
        img.save(buf, "JPEG", **kw)
        jpgs.append(buf.getvalue())
    outs, st = ctx.decode_batch(jpgs)
    for jpg, out, s in zip(jpgs, outs, st):
        try:
            want = orc.decode(jpg).yuv()
        except orc.OracleError as e:
            assert s == e.status
            continue
        assert s == 0 and bytes(out) == want


def test_independent_decoder_cross_check(hcj, ctx, orc, data):
    """The reference's own external check (jpeg/test/mouse-decode.t:10-13: max abs difference <= 1 against ffmpeg), with
    libjpeg-turbo (Pillow) as the independent decoder: a different IDCT (jidctint), so luma may differ by one LSB, never
    more; 4:4:4 chroma likewise (subsampled chroma goes through libjpeg's own up-sampling and is not comparable)."""
    Image = pytest.importorskip("PIL.Image")
    cases = [(data("Mouse480.jpg"), 420)]
    for i, (c, q) in enumerate(((420, 75), (444, 95), (422, 30), (444, 50))):
        cases.append((orc.encode(synth.frame(900 + i, 320, 200, c), 320, 200, c, q), c))
    outs, st = ctx.decode_batch([j for j, _ in cases])
    assert st == [0] * len(cases)
    for (j, chroma), o in zip(cases, outs):
        im = Image.open(io.BytesIO(j))
        im.draft("YCbCr", im.size)
        assert im.mode == "YCbCr"
        ref = np.asarray(im).astype(np.int32)
        h, w = ref.shape[:2]
        y = o[: w * h].reshape(h, w).astype(np.int32)
        assert np.abs(y - ref[:, :, 0]).max() <= 1
        if chroma == 444:
            for k in (1, 2):
                p = o[k * w * h: (k + 1) * w * h].reshape(h, w).astype(np.int32)
                assert np.abs(p - ref[:, :, k]).max() <= 1


def test_corrupt_streams_do_not_poison_batch(hcj, ctx, orc, data):
    good = data("Mouse480.jpg")
    start = orc.header_decode(good).scan_bit_pos // 8
    rng = np.random.default_rng(11)
    batch = [good]
    for trial in range(24):
        bad = bytearray(good)
        for _ in range(3):
            v = int(rng.integers(0, 255))
            bad[int(rng.integers(start, len(bad) - 2))] = v if v != 0xFF else 0x7F
        batch.append(bytes(bad))
    batch += [good[:100], good[: len(good) - 2], b"", good[:2] + b"\xff\xc2" + good[4:], good]
    outs, st = ctx.decode_batch(batch)
    want_good = orc.decode(good).yuv()
    assert st[0] == 0 and st[-1] == 0 and bytes(outs[0]) == want_good and bytes(outs[-1]) == want_good
    statuses = set()
    for jpg, out, s in zip(batch, outs, st):
        try:
            want, ws = orc.decode(jpg).yuv(), 0
        except orc.OracleError as e:
            want, ws = None, e.status
        if s == -23 or ws == -23:
            continue
        assert s == ws, (s, ws)
        statuses.add(s)
        if ws == 0:
            assert bytes(out) == want
    assert len(statuses) >= 3


def test_idct_blocks_tap(hcj, ctx, orc):
    rng = np.random.default_rng(6)
    qt = orc.quant_scale(False, 50).astype(np.uint16)
    coefs = np.where(rng.random((5000, 64)) < 0.2, rng.integers(-300, 300, (5000, 64)), 0).astype(np.int16)
    coefs[:, 0] = rng.integers(-1024, 1024, 5000)
    coefs[-50:] = rng.integers(-32768, 32767, (50, 64))  # far above the int32 guard: 64-bit path
    got = ctx.idct_blocks(coefs, qt)
    inv = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
    for b in list(range(0, 5000, 13)) + list(range(4950, 5000)):
        d = np.zeros(64, np.int64)
        d[inv] = coefs[b].astype(np.int64) * qt
        want = np.clip(orc.chen_inverse(d), -128, 127) + 128
        assert got[b].tolist() == want.tolist(), b


def test_wide_idct_path(hcj, ctx, orc, data):
    """Quant tables patched to 255 / 16-bit tables: blocks exceed the 32-bit IDCT guard and must take the
    64-bit path (flagged by the entropy decoders) and still match the model bit for bit."""
    from test_emul_device_logic import _patch_dqt

    base = orc.encode(data("mini64x64.444"), 64, 64, 444, 100)
    big = _patch_dqt(base, 255)
    # 16-bit table: rewrite DQT segments with Pq = 1 and large entries
    b16 = bytearray()
    i = 0
    src = bytearray(base)
    while i < len(src):
        if src[i] == 0xFF and src[i + 1] == 0xDB:
            tq = src[i + 4] & 15
            b16 += bytes([0xFF, 0xDB, 0x00, 0x83, 0x10 | tq]) + b"".join(bytes([1 + (k % 3), 0x2C]) for k in range(64))
            i += 69
        else:
            b16.append(src[i])
            i += 1
    batch = [base, big, bytes(b16), data("Mouse480.jpg")]
    for restart_free in (True,):
        outs, st = ctx.decode_batch(batch)
        for jpg, out, s in zip(batch, outs, st):
            assert s == 0 and bytes(out) == orc.decode(jpg).yuv()
    yuv = synth.frame(77, 96, 64, 420)
    jr = _patch_dqt(orc.encode(yuv, 96, 64, 420, 90, restart_interval=2), 200)
    outs, st = ctx.decode_batch([jr])
    assert st == [0] and bytes(outs[0]) == orc.decode(jr).yuv()


# ---- full-size configurations ---------------------------------------------------------------------------
@pytest.mark.parametrize("ri", [0, 8])
def test_1080p_420_full_size(hcj, ctx, orc, ri):
    """BASELINE configs 2/3 at full size: a handful of distinct 1080p images against the oracle, then a
    larger batch checked through a size-independent property (identical inputs -> identical outputs,
    distinct inputs -> the oracle's checksum of checksums)."""
    w, h = 1920, 1080
    uniq = [orc.encode(synth.frame(2000 + i, w, h, 420), w, h, 420, 75, restart_interval=ri) for i in range(3)]
    want = [sha(orc.decode(j).yuv()) for j in uniq]
    batch = [uniq[i % 3] for i in range(48)]
    outs, st = ctx.decode_batch(batch)
    assert st == [0] * 48
    assert [sha(o) for o in outs] == [want[i % 3] for i in range(48)]


def test_4k_444_rgb_full_size(hcj, ctx, orc):
    """BASELINE config 4: 3840x2160 4:4:4 q95 -> RGB24."""
    w, h = 3840, 2160
    jpg = orc.encode(synth.frame(4000, w, h, 444), w, h, 444, 95)
    outs, st = ctx.decode_batch([jpg, jpg], hcj.OUT_RGB24)
    assert st == [0, 0]
    want = oracle_rgb(orc, orc.decode(jpg)).tobytes()
    assert bytes(outs[0]) == want and bytes(outs[1]) == want


# ---- the batch pipeline and the tile-parallel kernels at their seams ----------------------------------
def test_entropy_tap_many_tiles(hcj, ctx, orc):
    """A scan of hundreds of 4 KiB destuff tiles (stuffed FF 00 pairs and restart markers straddle tile and
    16-byte boundaries): the destuffed bytes and the interval count must equal the model's."""
    w, h = 1280, 720
    for ri, q in ((0, 97), (4, 97)):
        jpg = orc.encode(synth.frame(4242 + ri, w, h, 444), w, h, 444, q, restart_interval=ri)
        assert len(jpg) > 600_000
        hdr = orc.header_decode(jpg)
        with ctx.batch([jpg], hcj.OUT_PLANES) as b:
            b.decode()
            got = bytes(b.entropy(0))
            dec = orc.decode(jpg, want_blocks=True)
            assert len(got) == dec.entropy_len
            if ri == 0:  # the pure model's extract_entropy_coded_bits stops at the first marker: comparable without restarts
                assert got == bytes(orc.extract_entropy_coded_bits(jpg, hdr.scan_bit_pos // 8))
            assert np.array_equal(b.coefficients(0), dec.coefs_abs_dc().astype(np.int16))


def test_decode_batch_pinned_contiguous(hcj, ctx, orc, data):
    """hcj_decode_batch with caller buffers laid out back to back in pinned memory (the merged-copy path of the
    three-stream pipeline), a mix of restart / no-restart / corrupt images, more than one chunk."""
    import ctypes as C

    L = hcj.lib()
    w, h = 320, 176
    jpgs = []
    for i in range(40):
        jpgs.append(orc.encode(synth.frame(900 + i, w, h, 420), w, h, 420, 75, restart_interval=(8 if i % 3 else 0)))
    bad = bytearray(jpgs[7])
    bad[len(bad) // 2] ^= 0x5A
    jpgs[7] = bytes(bad)
    jpgs[21] = data("Mouse480.jpg")
    n = len(jpgs)
    sizes = [hcj.out_size(hcj.frame_info(j), hcj.OUT_YUV) for j in jpgs]
    in_bytes = sum((len(j) + 31) // 16 * 16 for j in jpgs)
    out_bytes = sum((s + 255) // 256 * 256 for s in sizes)
    pin_in, pin_out = L.hcj_host_alloc(in_bytes), L.hcj_host_alloc(out_bytes)
    assert pin_in and pin_out
    try:
        jp, lens = (C.c_void_p * n)(), (C.c_size_t * n)()
        op, caps, status = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        oi = oo = 0
        for i, j in enumerate(jpgs):
            C.memmove(pin_in + oi, j, len(j))
            jp[i], lens[i] = pin_in + oi, len(j)
            oi += (len(j) + 31) // 16 * 16
            op[i], caps[i] = pin_out + oo, (sizes[i] + 255) // 256 * 256
            oo += (sizes[i] + 255) // 256 * 256
        C.memset(pin_out, 0xA5, out_bytes)
        hcj._check(L.hcj_decode_batch(ctx._h, jp, lens, n, hcj.OUT_YUV, hcj.FLAG_DEFAULT, op, caps, status))
        for i, j in enumerate(jpgs):
            try:
                want = orc.decode(j).yuv()
            except orc.OracleError as e:
                assert status[i] == e.status, i
                continue
            assert status[i] == 0, (i, status[i])
            got = bytes(np.ctypeslib.as_array(C.cast(op[i], C.POINTER(C.c_uint8)), shape=(sizes[i],)))
            assert got == want, i
    finally:
        L.hcj_host_free(pin_in)
        L.hcj_host_free(pin_out)


def test_speculative_many_subsequences_and_rounds(hcj, ctx, orc):
    """Scans without restart markers long enough for several CTAs of subsequences per image, at a quality where
    blocks are longer than the guess window of the first pass (the fix-point rounds have work to do)."""
    cases = [(444, 98, 640, 480), (420, 12, 1024, 768), (422, 90, 800, 608)]
    jpgs = [orc.encode(synth.frame(3100 + i, w, h, c), w, h, c, q) for i, (c, q, w, h) in enumerate(cases)]
    outs, st = ctx.decode_batch(jpgs)
    assert st == [0] * len(jpgs)
    for j, o in zip(jpgs, outs):
        assert bytes(o) == orc.decode(j).yuv()


def test_long_restart_intervals(hcj, ctx, orc):
    """Restart intervals long enough for the subsequence decoder (north_star: in parallel across restart intervals and
    WITHIN an interval): one MCU row per interval (a camera's DRI), a few rows, one interval for the whole image, an
    interval count that does not divide the MCUs; next to short-interval and marker-free images in the same batch."""
    cases = [(420, 75, 640, 368, 40), (420, 75, 640, 368, 22), (444, 92, 400, 240, 50), (422, 60, 512, 200, 32), (420, 85, 800, 608, 1000),
             (420, 75, 640, 368, 920), (444, 95, 1000, 96, 125), (420, 75, 256, 192, 8), (420, 30, 333, 77, 0), (420, 50, 1920, 64, 120),
             (444, 75, 333, 177, 100)]
    jpgs = [orc.encode(synth.frame(5100 + i, w, h, c), w, h, c, q, restart_interval=ri) for i, (c, q, w, h, ri) in enumerate(cases)]
    for mode in (hcj.OUT_YUV, hcj.OUT_RGB24):
        outs, st = ctx.decode_batch(jpgs, mode)
        assert st == [0] * len(jpgs)
        for case, j, o in zip(cases, jpgs, outs):
            dec = orc.decode(j)
            assert bytes(o) == (dec.yuv() if mode == hcj.OUT_YUV else oracle_rgb(orc, dec).tobytes()), case
    with ctx.batch(jpgs, hcj.OUT_PLANES) as b:
        b.decode()
        for i, j in enumerate(jpgs):
            d = orc.decode(j, want_blocks=True)
            assert np.array_equal(b.coefficients(i), d.coefs_abs_dc().astype(np.int16)), cases[i]
    # an interval cut short / a missing marker: whatever the oracle says
    bad = []
    for j in jpgs[:6]:
        h0 = hcj.header_decode(j).scan_byte_pos
        k = j.find(b"\xff\xd1", h0)
        bad.append(j[: k - 40] + j[k:] if k > h0 + 80 else j[:-40] + b"\xff\xd9")
        bad.append(j[:k] + j[k + 2:] if k > 0 else j)
    outs, st = ctx.decode_batch(bad)
    for j, o, s_ in zip(bad, outs, st):
        assert s_ == orc.decode_status(j)
        if s_ == 0:
            assert bytes(o) == orc.decode(j).yuv()


def test_rgb_colour_vs_pillow(hcj, ctx, orc):
    """D13 on the device: the RGB24 output of 4:4:4 images against Pillow's YCbCr -> RGB applied to the device's own
    planar 4:4:4 output of the same images: +-1 (an independent colour conversion; the oracle's formula is pinned the
    same way over the whole cube in test_oracle_goldens.py), and exact against the oracle's stated formula."""
    Image = pytest.importorskip("PIL.Image")
    w, h = 200, 120
    jpgs = [orc.encode(synth.frame(700 + i, w, h, c), w, h, c, q) for i, (c, q) in enumerate(((444, 95), (444, 40), (420, 75), (422, 75)))]
    rgb, st = ctx.decode_batch(jpgs, hcj.OUT_RGB24)
    yuv, st2 = ctx.decode_batch(jpgs, hcj.OUT_YUV444)
    assert st == [0] * 4 and st2 == [0] * 4
    for r, p in zip(rgb, yuv):
        planes = [Image.fromarray(p[k * w * h:(k + 1) * w * h].reshape(h, w)) for k in range(3)]
        pil = np.asarray(Image.merge("YCbCr", planes).convert("RGB")).astype(np.int32)
        assert np.abs(r.reshape(h, w, 3).astype(np.int32) - pil).max() <= 1
        y, u, v = (p[k * w * h:(k + 1) * w * h].reshape(h, w) for k in range(3))
        assert np.array_equal(r.reshape(h, w, 3), orc.ycbcr_to_rgb24(y, u, v))


def test_single_image_truncated_at_sos(hcj, ctx, orc, data):
    """A file that ends right after its SOS header has a scan of zero bytes: no terminator, alone in a batch or not."""
    jpg = data("mini.jpg")
    cut = hcj.header_decode(jpg).scan_byte_pos
    short = jpg[:cut]
    _, st1 = ctx.decode_batch([short])
    _, st2 = ctx.decode_batch([jpg, short, jpg])
    assert st1 == [-20] and st2 == [0, -20, 0]
    assert orc.decode_status(short) == -20


def test_two_contexts_in_one_process(hcj, ctx, orc):
    """hcj_decode_batch_multi / hcj_encode_batch_multi: images sharded by index over several contexts of one process, a
    host thread per context (here on one device when only one is present: the contexts are still independent)."""
    ndev = max(1, hcj.device_count())
    ctxs = [hcj.Context(k % ndev) for k in range(max(2, ndev))]
    try:
        w, h = 160, 96
        frames = [synth.frame(800 + i, w, h, 420) for i in range(11)]
        jpgs, st = hcj.encode_batch_multi(ctxs, frames, w, h, 420, 75, 0)
        assert st == [0] * 11
        for f, j in zip(frames, jpgs):
            assert j == orc.encode(f, w, h, 420, 75)
        jr, _ = ctx.encode_batch(frames[:5], w, h, 420, 75, 4)
        batch = jpgs + jr + [b"\xff\xd8garbage"]
        outs, st = hcj.decode_batch_multi(ctxs, batch)
        assert st[:-1] == [0] * 16 and st[-1] != 0
        for j, o in zip(batch[:-1], outs):
            assert bytes(o) == orc.decode(j).yuv()
        lo_hi = [hcj.shard_range(len(batch), k, len(ctxs)) for k in range(len(ctxs))]
        assert lo_hi[0][0] == 0 and lo_hi[-1][1] == len(batch) and all(a[1] == b[0] for a, b in zip(lo_hi, lo_hi[1:]))
    finally:
        for c in ctxs:
            c.close()


def test_multi_gpu_contexts(hcj, orc):
    """The same across real devices (skipped below 2 GPUs): per-device kernel attributes and SM counts."""
    if hcj.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctxs = [hcj.Context(k) for k in range(hcj.device_count())]
    try:
        w, h = 256, 160
        frames = [synth.frame(820 + i, w, h, 420) for i in range(4 * len(ctxs))]
        for ri in (0, 8):
            jpgs, st = hcj.encode_batch_multi(ctxs, frames, w, h, 420, 75, ri)
            assert not any(st)
            outs, st = hcj.decode_batch_multi(ctxs, jpgs, hcj.OUT_RGB24)
            assert not any(st)
            for j, o in zip(jpgs, outs):
                d = orc.decode(j)
                y, u, v = orc.upsample_to_444(d.cropped, 420)
                assert np.array_equal(o.reshape(h, w, 3), orc.ycbcr_to_rgb24(y, u, v))
    finally:
        for c in ctxs:
            c.close()


def test_fused_rgb444(hcj, ctx, orc, data):
    """J4: RGB24 of 4:4:4 images is produced inside the IDCT kernel (no plane round trip through HBM).  Odd sizes (crop in
    x and y, rows that are not 8-byte aligned), several tiles per MCU row (> 42 MCUs wide), blocks flagged for the 64-bit
    IDCT (k_rgb444_fix), 16-bit quant tables (not fused), and a mixed batch with sub-sampled images (k_rgb)."""
    from test_emul_device_logic import _patch_dqt

    cases = [(444, 8, 8, 75), (444, 17, 9, 90), (444, 333, 77, 60), (444, 344, 24, 95), (444, 1000, 40, 100), (420, 64, 48, 75),
             (444, 3840, 16, 95), (422, 80, 40, 75), (444, 61, 61, 30)]
    jpgs = [orc.encode(synth.frame(1300 + i, w, h, c), w, h, c, q) for i, (c, w, h, q) in enumerate(cases)]
    base = orc.encode(data("mini64x64.444"), 64, 64, 444, 100)
    jpgs += [_patch_dqt(base, 255), _patch_dqt(orc.encode(synth.frame(5, 96, 40, 444), 96, 40, 444, 100, restart_interval=3), 200)]
    b16 = bytearray()
    i, src = 0, bytearray(base)
    while i < len(src):
        if src[i] == 0xFF and src[i + 1] == 0xDB:
            b16 += bytes([0xFF, 0xDB, 0x00, 0x83, 0x10 | (src[i + 4] & 15)]) + b"".join(bytes([1 + (k % 3), 0x2C]) for k in range(64))
            i += 69
        else:
            b16.append(src[i])
            i += 1
    jpgs.append(bytes(b16))
    want = [oracle_rgb(orc, orc.decode(j)).tobytes() for j in jpgs]
    outs, st = ctx.decode_batch(jpgs, hcj.OUT_RGB24)
    assert st == [0] * len(jpgs)
    for k, (o, w_) in enumerate(zip(outs, want)):
        assert bytes(o) == w_, k
    # the resident three-step form takes the same kernels
    with ctx.batch(jpgs, hcj.OUT_RGB24) as b:
        b.decode()
        outs, st = b.fetch()
        assert st == [0] * len(jpgs) and [bytes(o) for o in outs] == want


def test_fused_rgb_subsampled(hcj, ctx, orc, data):
    """J4 for 4:2:0 / 4:2:2 (HCJ_FUSED_SUB=1): RGB24 of even-sized images comes out of the IDCT kernel (Planar_444
    up-sampling and the colour conversion inside the tile; the units that need another tile's chroma go through
    k_rgb_deferred), and the default two-kernel form with two pixel rows per thread.  Widths
    that are not a multiple of 16 (crop inside the last unit), several tiles per MCU row (> 21 / 32 MCUs wide), one
    MCU row, tiny images, blocks flagged for the 64-bit IDCT, and odd sizes (not fused) in the same batch."""
    from test_emul_device_logic import _patch_dqt

    cases = [(420, 256, 192, 75), (420, 1000, 40, 90), (420, 1004, 38, 60), (420, 52, 44, 95), (420, 16, 16, 75), (420, 2, 2, 75),
             (420, 34, 18, 50), (422, 200, 120, 75), (422, 1030, 24, 85), (422, 18, 9, 75), (420, 1920, 64, 75), (420, 640, 368, 30),
             (420, 333, 77, 75), (422, 131, 70, 75), (444, 64, 48, 75), (420, 700, 34, 100), (420, 354, 64, 75)]
    jpgs = [orc.encode(synth.frame(1700 + i, w, h, c), w, h, c, q, restart_interval=(8 if i % 2 else 0)) for i, (c, w, h, q) in enumerate(cases)]
    jpgs += [_patch_dqt(orc.encode(synth.frame(9, 96, 64, 420), 96, 64, 420, 100, restart_interval=2), 200),
             _patch_dqt(orc.encode(synth.frame(10, 704, 32, 422), 704, 32, 422, 100), 255)]
    want = [oracle_rgb(orc, orc.decode(j)).tobytes() for j in jpgs]
    os_env = __import__("os").environ
    os_env["HCJ_FUSED_SUB"] = "1"  # the fused form is opt-in for sub-sampled images (slower than the two kernels: hcj_api.cu)
    try:
        outs, st = ctx.decode_batch(jpgs, hcj.OUT_RGB24)
    finally:
        del os_env["HCJ_FUSED_SUB"]
    assert st == [0] * len(jpgs)
    for k, (o, w_) in enumerate(zip(outs, want)):
        assert bytes(o) == w_, (k, (cases + ["wide420", "wide422"])[k])
    # the default: k_idct_persistent to planes, then k_rgb_sub_pairs (even sizes) / k_rgb (odd sizes)
    outs, st = ctx.decode_batch(jpgs, hcj.OUT_RGB24)
    assert st == [0] * len(jpgs) and [bytes(o) for o in outs] == want


def test_destuff_three_kernel_form(hcj, ctx, orc, data):
    """K1 has two forms: the chained scan with decoupled look-back (one kernel, the default) and count / scan / write
    (HCJ_DESTUFF_3PASS=1).  Same results: files of many tiles, restart markers, stuffed bytes at tile boundaries,
    terminators in the middle of a scan, truncated files."""
    jpgs = [orc.encode(synth.frame(7300 + i, w, h, c), w, h, c, q, restart_interval=ri)
            for i, (c, w, h, q, ri) in enumerate([(444, 640, 480, 98, 0), (420, 800, 608, 75, 8), (422, 512, 200, 100, 1), (420, 64, 48, 50, 0)])]
    jpgs += [jpgs[0][: len(jpgs[0]) // 2] + b"\xff\xd9" + jpgs[0][len(jpgs[0]) // 2:], jpgs[1][:-300], data("Mouse480.jpg")]
    os_env = __import__("os").environ
    res = []
    for three in (False, True):
        if three:
            os_env["HCJ_DESTUFF_3PASS"] = "1"
        try:
            with ctx.batch(jpgs, hcj.OUT_YUV) as b:
                b.decode()
                outs, st = b.fetch()
                res.append((st, [None if o is None else bytes(o) for o in outs], [b.entropy(i) for i in range(len(jpgs)) if st[i] == 0]))
        finally:
            os_env.pop("HCJ_DESTUFF_3PASS", None)
    assert res[0] == res[1]
    for j, o, s_ in zip(jpgs, res[0][1], res[0][0]):
        assert s_ == orc.decode_status(j)
        if s_ == 0:
            assert o == orc.decode(j).yuv()
