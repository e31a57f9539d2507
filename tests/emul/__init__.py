"""TEST INFRASTRUCTURE: CPU emulation of the kernels' per-thread logic (see emul.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "video-coding_b200", "csrc")
_OUT = os.path.join(_ROOT, "tests", "_build", "libemul.so")
_lib = None


def build():
    srcs = [os.path.join(_HERE, "emul.cpp")] + [os.path.join(_CSRC, f) for f in ("hcj_device.cuh", "hcj_host.cpp", "hcj_host.h", "hcj_common.h")]
    if os.path.exists(_OUT) and all(os.path.getmtime(_OUT) >= os.path.getmtime(s) for s in srcs):
        return _OUT
    os.makedirs(os.path.dirname(_OUT), exist_ok=True)
    subprocess.check_call(
        ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-I", _CSRC, "-x", "c++",
         os.path.join(_HERE, "emul.cpp"), os.path.join(_CSRC, "hcj_host.cpp"), "-o", _OUT]
    )
    return _OUT


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.emu_reconstruct_blocks.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_void_p]
        L.emu_idct32_unguarded.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.emu_decode_segments.argtypes = [C.c_char_p, C.c_int64, C.c_uint, C.c_char_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.emu_decode_speculative.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.POINTER(C.c_int), C.c_void_p]
        L.emu_decode_units.argtypes = [C.c_char_p, C.c_int64, C.c_uint, C.c_char_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p,
                                       C.POINTER(C.c_int), C.c_void_p]
        L.emu_fdct_quant.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.emu_quantize_check.argtypes = [C.c_int]
        L.emu_quantize_check.restype = C.c_int64
        L.emu_entropy_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64]
        L.emu_entropy_encode.restype = C.c_int64
        _lib = L
    return _lib


def reconstruct_blocks(coefs, qt, force_wide=False):
    coefs = np.ascontiguousarray(coefs, np.int16).reshape(-1, 64)
    qt = np.ascontiguousarray(qt, np.uint16)
    out = np.zeros((coefs.shape[0], 64), np.uint8)
    lib().emu_reconstruct_blocks(coefs.ctypes.data, coefs.shape[0], qt.ctypes.data, int(force_wide), out.ctypes.data)
    return out


def idct32_unguarded(dequant):
    d = np.ascontiguousarray(dequant, np.int32).reshape(-1, 64)
    out = np.zeros_like(d)
    lib().emu_idct32_unguarded(d.ctypes.data, d.shape[0], out.ctypes.data)
    return out


def split_entropy(jpeg, scan_start, restart):
    """Reference state machine of decoder.ml:261-281 (+ RSTn splitting): destuffed bytes and interval offsets."""
    out, segs, prev, i = bytearray(), [0], 0, scan_start
    while i < len(jpeg):
        c = jpeg[i]
        if prev == 0xFF:
            if c == 0:
                out.append(0xFF)
                prev = 0
            elif restart and 0xD0 <= c <= 0xD7:
                segs.append(len(out))
                prev = 0
            else:
                break
        elif c == 0xFF:
            prev = c
        else:
            out.append(c)
            prev = c
        i += 1
    segs.append(len(out))
    return bytes(out), segs


def decode_segments(jpeg, nblocks, scan_start, restart, flags=1):
    ent, segs = split_entropy(jpeg, scan_start, restart)
    so = np.array(segs, np.uint32)
    coefs = np.full((nblocks, 64), 0x5A5A, np.int16)  # garbage: the decoders clear the blocks themselves
    wide = np.zeros(nblocks // 32 + 2, np.uint32)
    st = lib().emu_decode_segments(jpeg, len(jpeg), flags, ent, so.ctypes.data, len(segs) - 1, coefs.ctypes.data, wide.ctypes.data)
    decode_segments.wide = wide
    return st, coefs


def decode_speculative(jpeg, nblocks, scan_start, T=64, S=1024):
    ent, _ = split_entropy(jpeg, scan_start, False)
    coefs = np.full((nblocks, 64), 0x5A5A, np.int16)  # garbage: the decoders clear the blocks themselves
    rounds = C.c_int()
    wide = np.zeros(nblocks // 32 + 2, np.uint32)
    st = lib().emu_decode_speculative(jpeg, len(jpeg), ent, len(ent), T, S, coefs.ctypes.data, C.byref(rounds), wide.ctypes.data)
    decode_speculative.wide = wide
    return st, coefs, rounds.value


def decode_units(jpeg, nblocks, scan_start, T=64, S=1024, flags=1):
    """Restart intervals decoded as units of the speculative decoder (long intervals, k_spec_*)."""
    ent, segs = split_entropy(jpeg, scan_start, b"\xff\xdd" in jpeg[:scan_start])  # RSTn only splits files that carry a DRI segment
    so = np.array(segs, np.uint32)
    coefs = np.full((nblocks, 64), 0x5A5A, np.int16)
    rounds = C.c_int()
    wide = np.zeros(nblocks // 32 + 2, np.uint32)
    st = lib().emu_decode_units(jpeg, len(jpeg), flags, ent, so.ctypes.data, len(segs) - 1, T, S, coefs.ctypes.data, C.byref(rounds), wide.ctypes.data)
    decode_units.wide = wide
    return st, coefs, rounds.value


def fdct_quant(pix, qt):
    pix = np.ascontiguousarray(pix, np.uint8).reshape(64)
    qt = np.ascontiguousarray(qt, np.uint16)
    out = np.zeros(64, np.int16)
    fd = np.zeros(64, np.int32)
    lib().emu_fdct_quant(pix.ctypes.data, qt.ctypes.data, out.ctypes.data, fd.ctypes.data)
    return out, fd


def entropy_encode(quant, width, height, chroma, restart_interval=0):
    q = np.ascontiguousarray(quant, np.int16)
    out = np.zeros(q.size * 4 + 1024, np.uint8)
    n = lib().emu_entropy_encode(q.ctypes.data, width, height, chroma, restart_interval, out.ctypes.data, out.size)
    assert n >= 0, n
    return out[:n].tobytes()
