"""Randomised GPU parity sweeps: seeded geometry / content / corruption, every result bit-exact against the oracle
(decoded frames, statuses, coefficient taps, encoded files).  The cases are drawn from one PCG64 stream per test, so a
failure reproduces from the printed case.  HCJ_FUZZ_SEED (default 0) shifts every stream: `tools/fuzz_soak.sh` walks
it for a soak run on a GPU box."""
import os

import numpy as np
import pytest

import synth

pytestmark = pytest.mark.gpu

SEED = int(os.environ.get("HCJ_FUZZ_SEED", "0"))
# HCJ_FUZZ_MAX_W / _MAX_H / _COUNT enlarge the decode sweep (many sub-sequences and tiles per scan) for a soak run
MAX_W = int(os.environ.get("HCJ_FUZZ_MAX_W", "320"))
MAX_H = int(os.environ.get("HCJ_FUZZ_MAX_H", "200"))
COUNT = int(os.environ.get("HCJ_FUZZ_COUNT", "120"))


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


@pytest.fixture(params=["routing_default", "all_intervals_speculative"])
def routing(request):
    """Which entropy kernel decodes images with restart intervals: by default short intervals go one per thread (K2)
    and long ones through the subsequence decoder (K3, every interval a unit); HCJ_LONG_RI_BLOCKS=1 sends every
    interval, however short (degenerate ones of <= 16 bits included), through K3."""
    old = os.environ.get("HCJ_LONG_RI_BLOCKS")
    if request.param == "all_intervals_speculative":
        os.environ["HCJ_LONG_RI_BLOCKS"] = "1"
    yield request.param
    if old is None:
        os.environ.pop("HCJ_LONG_RI_BLOCKS", None)
    else:
        os.environ["HCJ_LONG_RI_BLOCKS"] = old


def random_frame(rng, w, h, chroma, kind):
    cw = w if chroma == 444 else w // 2
    ch = h // 2 if chroma == 420 else h
    n = w * h + 2 * cw * ch
    if kind == 0:  # the bench's sinusoids + noise
        return synth.frame(int(rng.integers(1 << 30)), w, h, chroma)
    if kind == 1:  # white noise: dense blocks, long codes, ZRL runs at low quality
        return rng.integers(0, 256, n, dtype=np.uint8).tobytes()
    if kind == 2:  # flat: one symbol per block
        return bytes([int(rng.integers(256))]) * n
    # extremes in 8 x 8 checkerboards: the largest coefficients the FDCT produces
    yy, xx = np.mgrid[0:h, 0:w]
    y = (((xx // int(rng.integers(1, 9))) + (yy // int(rng.integers(1, 9)))) % 2 * 255).astype(np.uint8)
    u = rng.integers(0, 2, (ch, cw), dtype=np.uint8) * 255
    v = rng.integers(0, 2, (ch, cw), dtype=np.uint8) * 255
    return y.tobytes() + u.tobytes() + v.tobytes()


def random_cases(rng, count, max_w=320, max_h=200):
    cases = []
    for _ in range(count):
        chroma = int(rng.choice([420, 422, 444]))
        w, h = int(rng.integers(2, max_w)), int(rng.integers(2, max_h))
        q = int(rng.choice([1, 5, 25, 50, 75, 90, 100, int(rng.integers(1, 101))]))
        ri = int(rng.choice([0, 0, 1, 2, 3, 8, int(rng.integers(1, 40))]))
        cases.append((chroma, w, h, q, ri, int(rng.integers(4))))
    return cases


def encode_cases(orc, rng, cases):
    """Oracle-encoded files of the cases.  Sizes the model's encoder cannot take (its padded chroma plane is narrower
    than the MCU grid when the luma size is 1 mod 16: Plane bounds, status -10) are bumped by one sample."""
    out_cases, jpgs = [], []
    for c, w, h, q, ri, kind in cases:
        while True:
            try:
                jpgs.append(orc.encode(random_frame(rng, w, h, c, kind), w, h, c, q, restart_interval=ri))
                break
            except orc.OracleError as e:
                assert e.status == -10
                w, h = w + (w % 16 == 1), h + (h % 16 == 1)
        out_cases.append((c, w, h, q, ri, kind))
    return out_cases, jpgs


def test_fuzz_decode_all_modes(hcj, ctx, orc, routing):
    rng = np.random.default_rng(20261018 + SEED)
    cases, jpgs = encode_cases(orc, rng, random_cases(rng, COUNT, MAX_W, MAX_H))
    # (a restart interval of <= 16 bits makes the model's `show` bound observable: status -9, also a defined result)
    want_st = [orc.decode_status(j) for j in jpgs]
    decs = [orc.decode(j) if s == 0 else None for j, s in zip(jpgs, want_st)]
    assert sum(s == 0 for s in want_st) > COUNT * 5 // 6
    from test_gpu_decode import oracle_rgb, oracle_yuv444

    want = {
        hcj.OUT_YUV: lambda d: d.yuv(),
        hcj.OUT_PLANES: lambda d: b"".join(p.tobytes() for p in d.planes),
        hcj.OUT_RGB24: lambda d: oracle_rgb(orc, d).tobytes(),
        hcj.OUT_YUV444: lambda d: oracle_yuv444(orc, d).tobytes(),
    }
    for mode, f in want.items():
        outs, st = ctx.decode_batch(jpgs, mode)
        assert st == want_st
        for case, o, d in zip(cases, outs, decs):
            if d is not None:
                assert bytes(o) == f(d), (mode, case)


def test_fuzz_coefficients(hcj, ctx, orc, routing):
    rng = np.random.default_rng(7 + SEED)
    cases, jpgs = encode_cases(orc, rng, random_cases(rng, 40, 200, 120))
    with ctx.batch(jpgs, hcj.OUT_PLANES) as b:
        b.decode()
        for i, (case, j) in enumerate(zip(cases, jpgs)):
            if orc.decode_status(j) != 0:
                continue
            d = orc.decode(j, want_blocks=True)
            assert np.array_equal(b.coefficients(i), d.coefs_abs_dc().astype(np.int16)), case


def test_fuzz_corrupt_streams(hcj, ctx, orc, routing):
    """Byte flips, insertions and cuts in the entropy-coded segment: the model mostly decodes garbage without raising;
    whatever the oracle does (frame or status), the library does, image by image."""
    rng = np.random.default_rng(99 + SEED)
    cases, good = encode_cases(orc, rng, random_cases(rng, 60, 160, 120))
    bad = []
    for j in good:
        j = bytearray(j)
        hdr = hcj.header_decode(bytes(j)).scan_byte_pos
        op = int(rng.integers(4))
        if op == 0 and len(j) - hdr > 4:  # flip bytes
            for _ in range(int(rng.integers(1, 6))):
                j[int(rng.integers(hdr, len(j) - 2))] = int(rng.integers(256))
        elif op == 1 and len(j) - hdr > 8:  # cut the scan short, keep EOI
            j = j[: int(rng.integers(hdr + 1, len(j) - 2))] + b"\xff\xd9"
        elif op == 2:  # a stray marker in the scan
            p = int(rng.integers(hdr, len(j) - 2))
            j = j[:p] + bytes([0xFF, int(rng.choice([0xD0, 0xD3, 0xD9, 0xC4, 0x01]))]) + j[p:]
        else:  # no terminator at all
            j = j[:-2]
        bad.append(bytes(j))
    for flags in (hcj.FLAG_DEFAULT, 0):
        outs, st = ctx.decode_batch(bad, hcj.OUT_YUV, flags)
        for case, j, o, s in zip(cases, bad, outs, st):
            want = orc.decode_status(j, restart_ext=bool(flags & hcj.FLAG_RESTART_EXT))
            assert s == want, (case, s, want)
            if s == 0:
                assert bytes(o) == orc.decode(j, restart_ext=bool(flags & hcj.FLAG_RESTART_EXT)).yuv(), case


def test_fuzz_encode(hcj, ctx, orc):
    rng = np.random.default_rng(5 + SEED)
    for _ in range(40):
        chroma = int(rng.choice([420, 422, 444]))
        w, h = int(rng.integers(2, 260)), int(rng.integers(2, 180))
        q = int(rng.choice([1, 10, 50, 75, 95, 100]))
        ri = int(rng.choice([0, 0, 1, 5, 8]))
        frames = [random_frame(rng, w, h, chroma, int(rng.integers(4))) for _ in range(3)]
        try:
            outs, st = ctx.encode_batch(frames, w, h, chroma, q, ri)
        except hcj.HcjError as e:  # a geometry the model's encoder rejects fails the call (all frames share it)
            for f in frames:
                with pytest.raises(orc.OracleError) as eo:
                    orc.encode(f, w, h, chroma, q, restart_interval=ri)
                assert eo.value.status == e.status, (chroma, w, h, q, ri)
            continue
        for f, o, s in zip(frames, outs, st):
            try:
                want = orc.encode(f, w, h, chroma, q, restart_interval=ri)
            except orc.OracleError as e:  # sizes the model's encoder rejects (Plane bounds): same status
                assert s == e.status, (chroma, w, h, q, ri, s, e.status)
                continue
            assert s == 0 and o == want, (chroma, w, h, q, ri)
