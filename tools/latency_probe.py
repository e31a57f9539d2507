"""Single-image latency of the public calls (decode_a_frame to YUV / RGB24, encode_batch of one frame) for the three
bench geometries, pageable host buffers, wall clock around the Python call.  Run on a GPU box: python tools/latency_probe.py"""
import sys, time; sys.path[:0]=['/root/repo','/root/repo/tests','/root/repo/video-coding_b200']
import numpy as np, synth, hcjpeg
ctx = hcjpeg.Context(0)
for name,(w,h,c,q,ri) in {'1080p_ri8':(1920,1080,420,75,8),'1080p_nori':(1920,1080,420,75,0),'4k444':(3840,2160,444,95,0)}.items():
    f = synth.frame(1,w,h,c)
    j,_ = ctx.encode_batch([f],w,h,c,q,ri); j=j[0]
    for mode,mn in ((hcjpeg.OUT_YUV,'yuv'),(hcjpeg.OUT_RGB24,'rgb')):
        for _ in range(3): ctx.decode_a_frame(j, mode)
        t=time.perf_counter(); n=20
        for _ in range(n): ctx.decode_a_frame(j, mode)
        dt=(time.perf_counter()-t)/n*1e3
        print(name, mn, len(j), 'bytes', round(dt,3), 'ms per frame (python call incl. numpy alloc)')
    t=time.perf_counter()
    for _ in range(10): ctx.encode_batch([f],w,h,c,q,ri)
    print(name,'encode', round((time.perf_counter()-t)/10*1e3,3),'ms')
