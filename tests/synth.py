"""Synthetic inputs shared by tests and bench.py (SURVEY.md 8d recipe): planar YCbCr frames
p(x, y) = clip(a + sum_k A_k sin(fx_k x + fy_k y + phi_k) + N(0, 5^2)), PCG64 seeded per image."""
import numpy as np


def plane(rng, w, h):
    a = rng.uniform(60, 190)
    x = np.arange(w, dtype=np.float64)
    y = np.arange(h, dtype=np.float64)
    p = np.full((h, w), a, np.float32)
    for _ in range(6):
        amp = rng.uniform(5, 40)
        f = rng.choice([0.002, 0.01, 0.05]) * 2 * np.pi
        fx, fy = rng.uniform(-1, 1, 2) * f
        phi = rng.uniform(0, 2 * np.pi)
        # sin(fx x + fy y + phi) by the angle-sum identity: two outer products instead of a 2-D sin
        sx, cx = np.sin(fx * x + phi), np.cos(fx * x + phi)
        sy, cy = np.sin(fy * y), np.cos(fy * y)
        p += (amp * (np.outer(cy, sx) + np.outer(sy, cx))).astype(np.float32)
    p += rng.standard_normal((h, w), dtype=np.float32) * 5
    return np.clip(np.rint(p), 0, 255).astype(np.uint8)


def frame(seed, w, h, chroma=420):
    """Raw planar frame bytes as Frame.input reads them (frame.ml:72-76)."""
    rng = np.random.default_rng(seed)
    cw = w if chroma == 444 else w // 2
    ch = h // 2 if chroma == 420 else h
    return plane(rng, w, h).tobytes() + plane(rng, cw, ch).tobytes() + plane(rng, cw, ch).tobytes()


def jpeg_segments(jpg):
    """[(marker code, payload without the length field)] of the header up to and including SOS, then the rest."""
    segs, p = [], 2
    assert jpg[:2] == b"\xff\xd8"
    while True:
        assert jpg[p] == 0xFF, p
        code = jpg[p + 1]
        n = int.from_bytes(jpg[p + 2:p + 4], "big")
        segs.append((code, jpg[p + 4:p + 2 + n]))
        p += 2 + n
        if code == 0xDA:
            return segs, jpg[p:]


def merge_table_segments(jpg, fill=2):
    """The same image with all DQT tables in one segment, all DHT tables in one segment (what ffmpeg and many
    cameras write) and `fill` 0xFF fill bytes in front of every marker: T.81-legal, unreadable by the model's parser."""
    segs, rest = jpeg_segments(jpg)
    dqt = b"".join(p for c, p in segs if c == 0xDB)
    dht = b"".join(p for c, p in segs if c == 0xC4)
    out, done = [b"\xff\xd8"], set()
    for c, p in segs:
        if c in (0xDB, 0xC4):
            if c in done:
                continue
            done.add(c)
            p = dqt if c == 0xDB else dht
        out.append(b"\xff" * fill + bytes([0xFF, c]) + (len(p) + 2).to_bytes(2, "big") + p)
    return b"".join(out) + rest
