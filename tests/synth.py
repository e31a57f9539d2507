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
