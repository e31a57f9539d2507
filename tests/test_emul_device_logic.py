"""CPU checks of the kernels' per-thread logic (video-coding_b200/csrc/hcj_device.cuh compiled with g++,
driven by tests/emul/emul.cpp) against the oracle.  The real parity tests are tests/test_gpu_*.py; these
exist so that logic errors are found without spending GPU time."""
import numpy as np
import pytest

import emul
import synth


def _scan_start(orc, jpg):
    return orc.header_decode(jpg).scan_bit_pos // 8


# ---- IDCT: 32-bit path + L1 guard vs the model's 63-bit arithmetic ---------------------------------
def test_idct_random_blocks(orc):
    rng = np.random.default_rng(0)
    qt = orc.quant_scale(False, 75).astype(np.uint16)
    coefs = np.zeros((4000, 64), np.int16)
    for b in range(4000):
        nz = rng.integers(1, 64)
        idx = rng.choice(64, nz, replace=False)
        coefs[b, idx] = rng.integers(-60, 60, nz)
        coefs[b, 0] = rng.integers(-1024, 1024)
    got = emul.reconstruct_blocks(coefs, qt)
    inv = [0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
           35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63]
    for b in range(0, 4000, 7):
        d = np.zeros(64, np.int64)
        d[inv] = coefs[b].astype(np.int64) * qt
        want = np.clip(orc.chen_inverse(d), -128, 127) + 128
        assert got[b].tolist() == want.tolist()


def test_idct_guard_is_sufficient(orc):
    """Adversarial blocks with sum|dequant| just under the guard: the 32-bit path must equal int64."""
    limit = emul.lib().emu_idct_l1_limit()
    rng = np.random.default_rng(1)
    n = 3000
    d = np.zeros((n, 64), np.int64)
    for b in range(n):
        k = int(rng.choice([1, 2, 3, 8, 16, 64]))
        idx = rng.choice(64, k, replace=False)
        w = rng.random(k)
        mag = np.floor(w / w.sum() * (limit - 1)).astype(np.int64)
        d[b, idx] = mag * rng.choice([-1, 1], k)
    # worst cases by construction: all mass on one odd-frequency coefficient / one row / one column
    for j, pos in enumerate([1, 8, 9, 7, 56, 63, 3, 24, 5, 40]):
        for sgn in (-1, 1):
            blk = np.zeros(64, np.int64)
            blk[pos] = sgn * (limit - 1)
            d = np.vstack([d, blk[None]])
    got = emul.idct32_unguarded(d.astype(np.int32))
    for b in range(d.shape[0]):
        assert got[b].tolist() == orc.chen_inverse(d[b]).tolist(), b


def test_idct_wide_path(orc):
    """Blocks above the guard (and 16-bit quant tables) take the int64 path and still match."""
    rng = np.random.default_rng(2)
    qt = np.full(64, 255, np.uint16)
    coefs = rng.integers(-2000, 2000, (200, 64)).astype(np.int16)
    got = emul.reconstruct_blocks(coefs, qt)
    qt16 = rng.integers(256, 65535, 64).astype(np.uint16)
    got16 = emul.reconstruct_blocks(coefs, qt16, force_wide=True)
    inv = np.array([0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])
    for b in range(200):
        for q, g in ((qt, got), (qt16, got16)):
            d = np.zeros(64, np.int64)
            d[inv] = coefs[b].astype(np.int64) * q.astype(np.int64)
            want = np.clip(orc.chen_inverse(d), -128, 127) + 128
            assert g[b].tolist() == want.tolist()


# ---- entropy decode ----------------------------------------------------------------------------------
def _oracle_coefs(orc, jpg, restart_ext=True):
    dec = orc.decode(jpg, restart_ext=restart_ext, want_blocks=True)
    return dec, dec.coefs_abs_dc().astype(np.int16)


@pytest.mark.parametrize("name", ["Mouse480.jpg", "mini.jpg"])
def test_sequential_decode_fixtures(orc, data, name):
    jpg = data(name)
    dec, want = _oracle_coefs(orc, jpg)
    st, got = emul.decode_segments(jpg, dec.nblocks, _scan_start(orc, jpg), restart=False)
    assert st == 0
    assert np.array_equal(got, want)


@pytest.mark.parametrize("T,S", [(64, 1024), (7, 256), (512, 512), (3, 64)])
@pytest.mark.parametrize("name", ["Mouse480.jpg", "mini.jpg"])
def test_speculative_decode_fixtures(orc, data, name, T, S):
    jpg = data(name)
    dec, want = _oracle_coefs(orc, jpg)
    st, got, rounds = emul.decode_speculative(jpg, dec.nblocks, _scan_start(orc, jpg), T=T, S=S)
    assert st == 0
    assert np.array_equal(got, want)
    assert rounds >= 1


@pytest.mark.parametrize("chroma,q,w,h", [(420, 75, 256, 192), (444, 95, 160, 96), (422, 30, 200, 120), (420, 10, 333, 77)])
def test_speculative_decode_synthetic(orc, chroma, q, w, h):
    yuv = synth.frame(1000 + q, w, h, chroma)
    jpg = orc.encode(yuv, w, h, chroma, q)
    dec, want = _oracle_coefs(orc, jpg)
    for T, S in ((128, 1024), (16, 128)):
        st, got, rounds = emul.decode_speculative(jpg, dec.nblocks, _scan_start(orc, jpg), T=T, S=S)
        assert st == 0
        assert np.array_equal(got, want), (T, S, rounds)


@pytest.mark.parametrize("ri", [1, 3, 8])
def test_restart_interval_decode(orc, ri):
    w, h = 208, 112
    yuv = synth.frame(7, w, h, 420)
    jpg = orc.encode(yuv, w, h, 420, 75, restart_interval=ri)
    dec, want = _oracle_coefs(orc, jpg)
    st, got = emul.decode_segments(jpg, dec.nblocks, _scan_start(orc, jpg), restart=True)
    assert st == 0
    assert np.array_equal(got, want)
    # stated-extension pin: same pixels as the restart-free twin decoded with pure model semantics
    twin = orc.decode(orc.encode(yuv, w, h, 420, 75), restart_ext=False)
    assert dec.yuv() == twin.yuv()


@pytest.mark.parametrize("ri,S", [(1, 128), (3, 128), (13, 256), (40, 512), (200, 1024)])
def test_restart_intervals_as_speculative_units(orc, ri, S):
    """Long restart intervals go through the subsequence decoder, every interval a unit with its own block range and
    DC predictors (north_star: in parallel across restart intervals AND within an interval)."""
    w, h = 208, 112
    for chroma, q in ((420, 75), (444, 92)):
        yuv = synth.frame(11 + ri, w, h, chroma)
        jpg = orc.encode(yuv, w, h, chroma, q, restart_interval=ri)
        dec, want = _oracle_coefs(orc, jpg)
        st, got, _ = emul.decode_units(jpg, dec.nblocks, _scan_start(orc, jpg), T=32, S=S)
        assert st == 0
        assert np.array_equal(got, want), (chroma, ri)


def test_truncated_stream_zero_extension(orc, data):
    """A scan cut short still decodes (the reader zero-extends, bitstream_reader.ml:19-22); same garbage."""
    jpg = data("mini.jpg")
    cut = jpg[: len(jpg) - 300] + b"\xff\xd9"
    try:
        dec, want = _oracle_coefs(orc, cut)
    except orc.OracleError as e:
        st, _, _ = emul.decode_speculative(cut, 96, _scan_start(orc, cut))
        assert st == e.status
        return
    st, got, _ = emul.decode_speculative(cut, dec.nblocks, _scan_start(orc, cut), T=32, S=256)
    assert st == 0 and np.array_equal(got, want)
    st, got = emul.decode_segments(cut, dec.nblocks, _scan_start(orc, cut), restart=False)
    assert st == 0 and np.array_equal(got, want)


def test_corrupt_short_scans_speculative(orc, data):
    """Damaged scans of small images, where what is left of the data ends long before the blocks do: the blocks that
    begin beyond the end of the data (the model's reader delivers zeros there) belong to the thread that gets there,
    whichever subsequence it started in.  `corrupt_flat_48x98.jpg` is the case the GPU fuzz sweep found;
    `corrupt_undefined_dc_at_mcu_start.jpg` the soak's: an undefined DC code exactly where a subsequence's first MCU
    begins (the thread that starts there must raise it, so the start is where the DC symbol is looked for)."""
    cases = [data("corrupt_flat_48x98.jpg"), data("corrupt_undefined_dc_at_mcu_start.jpg")]
    rng = np.random.default_rng(17)
    for i in range(30):
        chroma = int(rng.choice([420, 422, 444]))
        w, h = int(rng.integers(8, 120)), int(rng.integers(8, 100))
        kind = int(rng.integers(3))
        n = len(synth.frame(0, w, h, chroma))
        yuv = synth.frame(i, w, h, chroma) if kind == 0 else bytes([int(rng.integers(256))]) * n if kind == 1 else rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        try:
            j = bytearray(orc.encode(yuv, w, h, chroma, int(rng.choice([20, 75, 95])), restart_interval=int(rng.choice([0, 0, 5, 40]))))
        except orc.OracleError:
            continue
        start = _scan_start(orc, bytes(j))
        op = int(rng.integers(3))
        if op == 0 and len(j) - start > 6:
            for _ in range(int(rng.integers(1, 5))):
                j[int(rng.integers(start, len(j) - 2))] = int(rng.integers(0xFF))
        elif op == 1 and len(j) - start > 8:
            j = j[: int(rng.integers(start + 1, len(j) - 2))] + b"\xff\xd9"
        else:
            p = int(rng.integers(start, len(j) - 2))
            j = j[:p] + j[p + int(rng.integers(1, 6)):]
        cases.append(bytes(j))
    checked = 0
    for k, j in enumerate(cases):
        start = _scan_start(orc, j)
        try:
            dec, want = _oracle_coefs(orc, j)
            ost, nblocks = 0, dec.nblocks
        except orc.OracleError as e:
            ost, nblocks = e.status, 4096
        if ost not in (0, -2, -3, -4):
            continue
        for T, S in ((16, 128), (64, 1024)):
            st, got, _ = emul.decode_units(j, nblocks, start, T=T, S=S)
            assert st == ost, (k, T, S, st, ost)
            if ost == 0:
                assert np.array_equal(got, want), (k, T, S)
                checked += 1
    assert checked >= 10


def test_corrupt_stream_status(orc, data):
    """Flip bytes in the scan: wherever the model raises, the device logic reports the same status."""
    jpg = bytearray(data("Mouse480.jpg"))
    start = _scan_start(orc, bytes(jpg))
    rng = np.random.default_rng(3)
    seen = set()
    for trial in range(40):
        bad = bytearray(jpg)
        for _ in range(3):
            pos = int(rng.integers(start, len(bad) - 2))
            v = int(rng.integers(0, 255))
            bad[pos] = v if v != 0xFF else 0x7F
        bad = bytes(bad)
        try:
            dec, want = _oracle_coefs(orc, bad)
            ost = 0
        except orc.OracleError as e:
            ost = e.status
        st, got = emul.decode_segments(bad, 3600, start, restart=False)
        st2, got2, _ = emul.decode_speculative(bad, 3600, start, T=64, S=512)
        if ost == -23 or st == -23:
            continue  # DC beyond int16: documented domain limit of the coefficient store
        assert st == ost and st2 == ost, (trial, ost, st, st2)
        seen.add(ost)
        if ost == 0:
            assert np.array_equal(got, want) and np.array_equal(got2, want)
    assert len(seen) >= 2  # both clean decodes and raised statuses were exercised


# ---- encoder ------------------------------------------------------------------------------------------
def test_quantize_reciprocal_exhaustive():
    assert emul.lib().emu_quantize_check(40000) == 0


def test_fdct_quant_blocks(orc):
    rng = np.random.default_rng(4)
    for q in (1, 10, 50, 75, 95, 100):
        for chroma in (False, True):
            qt = orc.quant_scale(chroma, q).astype(np.uint16)
            fwd = [0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31, 40, 44, 53,
                   10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63]
            for _ in range(50):
                pix = rng.integers(0, 256, 64).astype(np.uint8) if rng.random() < 0.5 else rng.choice([0, 255], 64).astype(np.uint8)
                got, fd = emul.fdct_quant(pix, qt)
                f = orc.chen_forward(pix.astype(np.int64) - 128)
                assert fd.tolist() == f.tolist()
                want = np.zeros(64, np.int64)
                for i in range(64):
                    qq = int(qt[fwd[i]])
                    fi = int(f[i])
                    want[fwd[i]] = -((-(fi - 2 * qq)) // (4 * qq)) if fi < 0 else (fi + 2 * qq) // (4 * qq)
                assert got.tolist() == want.tolist()


@pytest.mark.parametrize("chroma,ri", [(420, 0), (422, 0), (444, 0), (420, 2), (444, 5)])
def test_entropy_encode_matches_oracle(orc, data, chroma, ri):
    src = data("mini64x64.%d" % chroma)
    jpg, quant, _ = orc.encode(src, 64, 64, chroma, 75, restart_interval=ri, want_blocks=True)
    hdr = orc.write_headers(64, 64, chroma, 75, ri)
    body = emul.entropy_encode(quant.astype(np.int16), 64, 64, chroma, ri)
    assert hdr + body + b"\xff\xd9" == jpg


def _patch_dqt(jpg, value):
    """Overwrite every 8-bit quant table entry: blows the dequantised magnitudes past the int32 guard."""
    b = bytearray(jpg)
    i = 0
    while True:
        i = b.find(b"\xff\xdb", i)
        if i < 0:
            break
        b[i + 5 : i + 5 + 64] = bytes([value]) * 64
        i += 69
    return bytes(b)


def test_wide_block_flags(orc, data):
    """Entropy decoders flag blocks whose sum(|dequantised coef|) may reach the IDCT guard; an unflagged
    block is certainly below it (so k_idct may skip its own check)."""
    limit = emul.lib().emu_idct_l1_limit()
    for jpg in (data("Mouse480.jpg"), _patch_dqt(orc.encode(data("mini64x64.444"), 64, 64, 444, 100), 255)):
        dec = orc.decode(jpg, want_blocks=True)
        l1 = np.abs(dec.dequant.astype(np.int64)).sum(1)
        start = _scan_start(orc, jpg)
        st, got = emul.decode_segments(jpg, dec.nblocks, start, restart=False)
        flags_seq = emul.decode_segments.wide
        st2, got2, _ = emul.decode_speculative(jpg, dec.nblocks, start, T=16, S=1024)
        flags_spec = emul.decode_speculative.wide
        assert st == 0 and st2 == 0
        for flags in (flags_seq, flags_spec):
            bit = np.array([(flags[b >> 5] >> (b & 31)) & 1 for b in range(dec.nblocks)])
            assert not np.any((l1 >= limit) & (bit == 0))
        if l1.max() < limit // 4:
            assert flags_seq.sum() == 0 and flags_spec.sum() == 0
        else:
            assert flags_seq.sum() > 0
        # and the reconstruction is exact either way
        qts = [np.array(list(orc.header_decode(jpg).quant_tables[i].elements), np.int64) for i in range(2)]
        assert np.array_equal(got, dec.coefs_abs_dc().astype(np.int16))
