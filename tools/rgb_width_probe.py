"""Stage times of RGB24 decode for image widths that are / are not multiples of 16 (the colour kernel's register paths
serve both; only the short last group of a row runs the edge variant).  Run on a GPU box: python tools/rgb_width_probe.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "video-coding_b200")]
import hcjpeg  # noqa: E402
import synth  # noqa: E402

ctx = hcjpeg.Context(0)
for w, h, c in ((1360, 768, 420), (1366, 768, 420), (1365, 767, 420), (1366, 768, 422), (1366, 768, 444), (1360, 768, 444)):
    frames = [synth.frame(i, w, h, c) for i in range(8)]
    jpgs, st = ctx.encode_batch(frames, w, h, c, 75, 8)
    assert st == [0] * 8
    with ctx.batch([jpgs[i % 8] for i in range(512)], hcjpeg.OUT_RGB24) as b:
        for _ in range(3):
            b.decode()
        t = b.decode_stages()
        print("%dx%d %d: rgb %.3f ms, idct %.3f ms per 512 frames (%.1f MP)" % (w, h, c, t["rgb"], t["idct"], 512 * w * h / 1e6))
