"""GPU parity tests of the encoder mirror: byte-identical files versus the oracle / the reference's
mini.jpg, through the C ABI."""
import hashlib

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


def test_mini_jpg_byte_identity(hcj, ctx, data):
    """jpeg/test_data/mini.jpg == Encoder.encode_420 mini64x64.420 q75, via the model-shaped mirror."""
    from hcjpeg.model import Encoder, Frame, Writer

    frame = Frame.frombytes(data("mini64x64.420"), 420, 64, 64)
    writer = Writer.create()
    Encoder.encode_420(frame=frame, quality=75, writer=writer, ctx=ctx)
    assert writer.get_buffer() == data("mini.jpg")


def test_survey_hashes(hcj, ctx, data):
    from test_oracle_goldens import SURVEY_HASHES

    for (chroma, q), (n, h) in sorted(SURVEY_HASHES.items()):
        outs, st = ctx.encode_batch([data("mini64x64.%d" % chroma)], 64, 64, chroma, q)
        assert st == [0] and (len(outs[0]), sha(outs[0])) == (n, h), (chroma, q)


CASES = [
    # (chroma, quality, w, h, restart_interval)
    (420, 75, 256, 192, 0), (420, 75, 256, 192, 8), (422, 50, 200, 120, 0), (444, 95, 160, 96, 0),
    (420, 10, 336, 80, 0), (420, 95, 52, 44, 0), (444, 75, 17, 9, 0), (420, 75, 16, 16, 1),
    (422, 75, 130, 70, 3), (444, 100, 64, 64, 2), (420, 1, 96, 80, 0), (420, 100, 64, 48, 0),
]


@pytest.mark.parametrize("case", CASES)
def test_encode_matches_oracle(hcj, ctx, orc, case):
    chroma, q, w, h, ri = case
    frames = [synth.frame(300 + i, w, h, chroma) for i in range(3)]
    outs, st = ctx.encode_batch(frames, w, h, chroma, q, ri)
    assert st == [0, 0, 0]
    for f, o in zip(frames, outs):
        assert o == orc.encode(f, w, h, chroma, q, restart_interval=ri)


def test_quantized_tap(hcj, ctx, orc):
    w, h = 128, 80
    f = synth.frame(5, w, h, 420)
    _, quant, _ = orc.encode(f, w, h, 420, 60, want_blocks=True)
    got = ctx.encode_quantized(f, w, h, 420, 60)
    assert np.array_equal(got, quant.astype(np.int16))


def test_extreme_content(hcj, ctx, orc):
    """Saturated / checkerboard content: long codes, many stuffed FF bytes, ZRL runs."""
    w, h = 64, 64
    rng = np.random.default_rng(9)
    frames = [
        bytes(rng.choice([0, 255], w * h * 3).astype(np.uint8)),
        bytes(np.full(w * h * 3, 255, np.uint8)),
        bytes(np.zeros(w * h * 3, np.uint8)),
        bytes((np.indices((h * 3, w)).sum(0) % 2 * 255).astype(np.uint8)),
    ]
    for q in (100, 75, 5):
        outs, st = ctx.encode_batch(frames, w, h, 444, q)
        assert st == [0] * 4
        for f, o in zip(frames, outs):
            assert o == orc.encode(f, w, h, 444, q), q


def test_dense_frame_exceeds_default_capacity(hcj, ctx, orc):
    """Saturated noise, 4:4:4, quality 100 is more than 3 bytes per pixel (found by tools/fuzz_soak.sh, seed 30): the C ABI
    answers HCJ_ERR_BUFFER_TOO_SMALL with the length the frame needs, the front-end's default capacity retries at that size."""
    w, h = 245, 137
    rng = np.random.default_rng(30)
    noise = bytes(rng.choice([0, 255], w * h * 3).astype(np.uint8))
    flat = bytes(w * h * 3)
    want = [orc.encode(f, w, h, 444, 100) for f in (noise, flat)]
    assert len(want[0]) > w * h * 3 + (1 << 16)
    assert len(want[0]) <= hcj.lib().hcj_encode_bound(w, h, 444)
    outs, st = ctx.encode_batch([noise, flat], w, h, 444, 100)
    assert st == [0, 0] and outs == want
    outs, st = ctx.encode_batch([noise, flat], w, h, 444, 100, capacity=len(want[0]) - 1)
    assert st == [-30, 0]  # HCJ_ERR_BUFFER_TOO_SMALL
    assert outs[0] is None and outs[1] == want[1]
    outs, st = ctx.encode_batch([noise], w, h, 444, 100, capacity=len(want[0]))
    assert st == [0] and outs == want[:1]


def test_plane_bounds_status(hcj, ctx, orc):
    """W = 17 at 4:2:0: per-component rounding disagrees and the model raises."""
    f = synth.frame(1, 17, 16, 420)
    with pytest.raises(orc.OracleError) as e:
        orc.encode(f, 17, 16, 420, 75)
    with pytest.raises(hcj.HcjError) as g:
        ctx.encode_batch([f], 17, 16, 420, 75)
    assert g.value.status == e.value.status == -10


@pytest.mark.parametrize("ri", [0, 8])
def test_1080p_encode_full_size(hcj, ctx, orc, ri):
    """BASELINE config 5 at full size + encode -> decode round trip through both device paths."""
    w, h = 1920, 1080
    frames = [synth.frame(5000 + i, w, h, 420) for i in range(2)]
    want = [orc.encode(f, w, h, 420, 75, restart_interval=ri) for f in frames]
    outs, st = ctx.encode_batch([frames[i % 2] for i in range(8)], w, h, 420, 75, ri)
    assert st == [0] * 8
    assert [sha(o) for o in outs] == [sha(want[i % 2]) for i in range(8)]
    dec, st = ctx.decode_batch(outs[:2])
    assert st == [0, 0]
    for j, d in zip(want, dec):
        assert bytes(d) == orc.decode(j).yuv()


def test_encode_device_time_and_long_single_segment(hcj, ctx, orc):
    """A frame whose scan is one long segment (no restart markers): the byte stuffing is split into 1 KiB units;
    the result must still be the model's bytes, and the device time of the call is reported."""
    import ctypes as C

    w, h = 1024, 768
    frame = synth.frame(77, w, h, 444)
    outs, st = ctx.encode_batch([frame], w, h, 444, 96)
    assert st == [0] and outs[0] == orc.encode(frame, w, h, 444, 96)
    ms = C.c_float()
    hcj._check(hcj.lib().hcj_encode_last_device_ms(ctx._h, C.byref(ms)))
    assert 0.0 < ms.value < 1000.0


def test_encode_monochrome(hcj, ctx, orc):
    """Encoder.encode_monochrome (encoder.ml:543-552): one component, luma tables only, and the model's linear
    Plane.blit into the padded plane (widths that are not a multiple of 8 shear the image) - byte-identical files;
    the one-component file decodes through get_decoded_planes like the oracle's."""
    rng = np.random.default_rng(11)
    for w, h, q, ri in ((64, 48, 75, 0), (61, 35, 50, 0), (8, 8, 90, 2), (200, 120, 95, 0), (131, 77, 20, 5), (1, 1, 75, 0)):
        frames = [synth.plane(rng, w, h).tobytes() for _ in range(3)]
        outs, st = ctx.encode_batch(frames, w, h, 400, q, ri)
        assert st == [0, 0, 0], (w, h, q, ri)
        for f, o in zip(frames, outs):
            assert o == orc.encode(f, w, h, 400, q, restart_interval=ri), (w, h, q, ri)
        planes, st = ctx.decode_batch(outs, hcj.OUT_PLANES)
        assert st == [0, 0, 0]
        for o, pl in zip(outs, planes):
            assert bytes(pl) == b"".join(p.tobytes() for p in orc.decode(o).planes)
        _, st = ctx.decode_batch(outs[:1], hcj.OUT_YUV)  # Decoder.get_yuv_frame needs three components (decoder.ml:415-420)
        assert st == [-12]


def test_chen_example_on_device(hcj, ctx, goldens):
    """jpeg/model/test/test_chen_dct.ml:47-94 through the kernels: the example block as the luma block of an 8 x 8 4:4:4
    frame at quality 100 (every quant entry 1, so Block.quant is the test's (f +/- 2) / 4) gives the pinned FDCT values,
    and hcj_idct_blocks of those the pinned IDCT values (after the decoder's clip and level shift)."""
    g = goldens["chen_example"]
    x = np.array(g["input"], np.int64)
    assert x.min() >= -128 and x.max() <= 127
    y = (x + 128).astype(np.uint8).tobytes()
    quant = ctx.encode_quantized(y + bytes(128), 8, 8, 444, 100)
    zz = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                   35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
    nat = np.zeros(64, np.int64)
    nat[zz] = quant[0]
    assert nat.tolist() == g["fdct"]
    recon = ctx.idct_blocks(quant[:1], np.ones(64, np.uint16))
    assert recon[0].tolist() == (np.clip(np.array(g["idct"]), -128, 127) + 128).tolist()


@pytest.mark.parametrize("case", [(420, 75, 80, 48, 0), (422, 40, 72, 40, 0), (444, 95, 33, 17, 0), (420, 75, 64, 64, 3), (400, 60, 43, 29, 0)])
def test_encode_block_log(hcj, ctx, orc, case):
    """`model encode log -verbose`: Encoder.Block.t of every block (x_pos, y_pos, input_pixels, fdct, quant, dc_pred, rle,
    decoded) against the oracle's per-block taps and the oracle's rle / inverse transform."""
    chroma, q, w, h, ri = case
    f = synth.frame(40, w, h, 444)[: w * h] if chroma == 400 else synth.frame(40, w, h, chroma)
    _, quant, fdct = orc.encode(f, w, h, chroma, q, restart_interval=ri, want_blocks=True)
    log = ctx.encode_block_log(f, w, h, chroma, q, ri)
    assert len(log) == len(quant)
    assert np.array_equal(log["quant"], quant.astype(np.int16))
    assert np.array_equal(log["fdct"], fdct)
    info = hcj.frame_info(orc.encode(f, w, h, chroma, q, restart_interval=ri))
    bpm, ncomp = info.blocks_per_mcu, info.ncomp
    order = [(c, bx, by) for c in range(ncomp) for by in range(info.vs[c]) for bx in range(info.hs[c])]
    qts = [orc.quant_scale(False, q)] + [orc.quant_scale(True, q)] * 2
    preds = [0] * ncomp
    planes = orc.split_yuv(f, w, h, chroma)
    for i, b in enumerate(log):
        mcu, k = divmod(i, bpm)
        c, bx, by = order[k]
        my, mx = divmod(mcu, info.mcus_wide)
        assert (b["component"], b["x_pos"], b["y_pos"]) == (c, (mx * info.hs[c] + bx) * 8, (my * info.vs[c] + by) * 8)
        if ri and k == 0 and mcu % ri == 0:
            preds = [0] * ncomp
        pairs, preds[c] = orc.rle(quant[i], preds[c])
        n = len(pairs)
        assert b["nrle"] == n and list(zip(b["rle_run"][:n].tolist(), b["rle_value"][:n].tolist())) == pairs, i
        assert b["dc_pred"] == preds[c] == int(quant[i][0])
        # input pixels: the zero-padded plane (linear blit for monochrome)
        src = np.asarray(planes[c])
        ph, pw = info.decoded_height[c], info.decoded_width[c]
        padded = np.zeros((ph, pw), np.uint8)
        if chroma == 400:
            padded.reshape(-1)[: src.size] = src.reshape(-1)
        else:
            padded[: src.shape[0], : src.shape[1]] = src
        blk = padded[b["y_pos"]: b["y_pos"] + 8, b["x_pos"]: b["x_pos"] + 8].reshape(64)
        assert np.array_equal(b["input_pixels"], blk)
        # Block.decoded
        zz = np.array(ZIGZAG_INVERSE)
        deq = np.zeros(64, np.int64)
        deq[zz] = quant[i].astype(np.int64) * np.asarray(qts[c], np.int64)
        assert np.array_equal(b["dequant"], deq)
        idct = orc.chen_inverse(deq)
        assert np.array_equal(b["idct"], idct)
        rec = np.clip(np.asarray(idct) + 128, 0, 255)
        assert np.array_equal(b["recon"], rec)
        assert np.array_equal(b["error"], np.abs(rec - blk.astype(np.int64)))


ZIGZAG_INVERSE = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                  35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47,
                  55, 62, 63]


@pytest.mark.parametrize("chunk", [1, 2, 3])
def test_encode_chunk_pipeline(hcj, ctx, orc, chunk):
    """hcj_encode_batch moves its frames through the device in chunks (upload, kernels and download of neighbouring
    chunks overlap; source frames and finished files are double-buffered): more chunks than buffers, a last chunk that
    is short, and byte buffers that have to grow for a later chunk (flat frames first, noise at quality 100 last)."""
    import os

    w, h, chroma = 176, 96, 420
    rng = np.random.default_rng(5)
    n = len(synth.frame(0, w, h, chroma))
    frames = [bytes([90 + i]) * n for i in range(3)] + [synth.frame(40 + i, w, h, chroma) for i in range(3)] + [rng.integers(0, 256, n, dtype=np.uint8).tobytes()]
    old = os.environ.get("HCJ_ENC_CHUNK")
    os.environ["HCJ_ENC_CHUNK"] = str(chunk)
    try:
        for q, ri in ((100, 0), (60, 5)):
            outs, st = ctx.encode_batch(frames, w, h, chroma, q, ri)
            assert st == [0] * len(frames)
            for f, o in zip(frames, outs):
                assert o == orc.encode(f, w, h, chroma, q, restart_interval=ri)
    finally:
        if old is None:
            del os.environ["HCJ_ENC_CHUNK"]
        else:
            os.environ["HCJ_ENC_CHUNK"] = old
