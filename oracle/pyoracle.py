"""ctypes binding of the CPU oracle (oracle/hcj_oracle.c).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / ``--impl reference`` legs, never from the product
package.  The oracle restates hardcamls/video-coding's OCaml JPEG model on the
CPU (see hcj_oracle.h for the reference file:line of every function).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")

FLAG_RESTART_EXT = 1
FLAG_T81_TABLES = 2


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc only)."""
    src = [os.path.join(_HERE, f) for f in ("hcj_oracle.c", "hcj_oracle.h", "Makefile")]
    if (
        not force
        and os.path.exists(_LIB)
        and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in src)
    ):
        return _LIB
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB


class Decoded(C.Structure):
    _fields_ = [
        ("status", C.c_int),
        ("ncomp", C.c_int),
        ("width", C.c_int),
        ("height", C.c_int),
        ("mcus_wide", C.c_int),
        ("mcus_high", C.c_int),
        ("blocks_per_mcu", C.c_int),
        ("nblocks", C.c_int64),
        ("hs", C.c_int * 4),
        ("vs", C.c_int * 4),
        ("decoded_width", C.c_int * 4),
        ("decoded_height", C.c_int * 4),
        ("actual_width", C.c_int * 4),
        ("actual_height", C.c_int * 4),
        ("plane", C.POINTER(C.c_uint8) * 4),
        ("cropped", C.POINTER(C.c_uint8) * 4),
        ("coefs", C.POINTER(C.c_int32)),
        ("dc_abs", C.POINTER(C.c_int32)),
        ("dequant", C.POINTER(C.c_int32)),
        ("recon", C.POINTER(C.c_uint8)),
        ("block_comp", C.POINTER(C.c_int8)),
        ("entropy_len", C.c_int64),
        ("yuv_status", C.c_int),
        ("chroma", C.c_int),
    ]


class Encoded(C.Structure):
    _fields_ = [
        ("status", C.c_int),
        ("bytes", C.POINTER(C.c_uint8)),
        ("len", C.c_int64),
        ("nblocks", C.c_int64),
        ("quant", C.POINTER(C.c_int32)),
        ("fdct", C.POINTER(C.c_int32)),
    ]


class Code(C.Structure):
    _fields_ = [("length", C.c_int), ("bits", C.c_int), ("data", C.c_int)]


class Component(C.Structure):
    _fields_ = [("identifier", C.c_int), ("h", C.c_int), ("v", C.c_int), ("tq", C.c_int)]


class Sof(C.Structure):
    _fields_ = [
        ("present", C.c_int),
        ("length", C.c_int),
        ("sample_precision", C.c_int),
        ("width", C.c_int),
        ("height", C.c_int),
        ("number_of_components", C.c_int),
        ("components", Component * 255),
    ]


class ScanComponent(C.Structure):
    _fields_ = [("selector", C.c_int), ("dc", C.c_int), ("ac", C.c_int)]


class Sos(C.Structure):
    _fields_ = [
        ("present", C.c_int),
        ("length", C.c_int),
        ("number_of_image_components", C.c_int),
        ("scan_components", ScanComponent * 255),
        ("ss", C.c_int),
        ("se", C.c_int),
        ("ah", C.c_int),
        ("al", C.c_int),
    ]


class Dqt(C.Structure):
    _fields_ = [
        ("length", C.c_int),
        ("element_precision", C.c_int),
        ("table_identifier", C.c_int),
        ("elements", C.c_int64 * 64),
    ]


class Dht(C.Structure):
    _fields_ = [
        ("length", C.c_int),
        ("table_class", C.c_int),
        ("destination_identifier", C.c_int),
        ("lengths", C.c_int * 16),
        ("nvalues", C.c_int),
        ("values", C.c_int * (16 * 255)),
    ]


class Header(C.Structure):
    _fields_ = [
        ("frame", Sof),
        ("scan", Sos),
        ("restart_interval_present", C.c_int),
        ("restart_interval_length", C.c_int),
        ("restart_interval", C.c_int),
        ("n_quant_tables", C.c_int),
        ("quant_tables", Dqt * 64),
        ("n_huffman_tables", C.c_int),
        ("huffman_tables", Dht * 64),
        ("scan_bit_pos", C.c_int64),
    ]


class Bits(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("len", C.c_int64), ("length_in_bits", C.c_int64), ("bit_pos", C.c_int64)]


class Writer(C.Structure):
    _fields_ = [
        ("word_buffer", C.c_uint64),
        ("word_bits", C.c_int),
        ("buffer", C.POINTER(C.c_uint8)),
        ("bytes_written", C.c_int64),
        ("capacity", C.c_int64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.orc_decode.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.c_int, C.POINTER(Decoded)]
        L.orc_decoded_free.argtypes = [C.POINTER(Decoded)]
        L.orc_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Encoded)]
        L.orc_encoded_free.argtypes = [C.POINTER(Encoded)]
        L.orc_header_decode.argtypes = [C.c_char_p, C.c_int64, C.POINTER(Header)]
        L.orc_header_decode_ex.argtypes = [C.c_char_p, C.c_int64, C.c_int, C.POINTER(Header)]
        L.orc_extract_entropy_coded_bits.argtypes = [C.c_char_p, C.c_int64, C.c_int64, C.c_void_p, C.POINTER(C.c_int64)]
        L.orc_write_headers.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_chen_inverse_8x8.argtypes = [C.c_void_p]
        L.orc_chen_forward_8x8.argtypes = [C.c_void_p]
        L.orc_quant_scale.argtypes = [C.c_int, C.c_int, C.c_void_p]
        L.orc_size.argtypes = [C.c_int64]
        L.orc_magnitude.argtypes = [C.c_int, C.c_int64]
        L.orc_magnitude.restype = C.c_int64
        L.orc_mag.argtypes = [C.c_int, C.c_int64]
        L.orc_mag.restype = C.c_int64
        L.orc_rle.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_void_p]
        L.orc_encoder_dc_table.argtypes = [C.c_int, C.POINTER(Code), C.POINTER(C.c_int)]
        L.orc_encoder_ac_table.argtypes = [C.c_int, C.POINTER(Code), C.POINTER(C.c_int)]
        for f in ("orc_supersample_h2", "orc_supersample_hv2", "orc_subsample_h2", "orc_subsample_hv2"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_crop_clamp.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_square_error.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_square_error.restype = C.c_int64
        L.orc_max_difference.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_max_difference.restype = C.c_int64
        L.orc_psnr.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.orc_psnr.restype = C.c_double
        L.orc_ycbcr_to_rgb24.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_time_decode.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.c_int, C.c_int, C.c_int]
        L.orc_time_decode.restype = C.c_double
        L.orc_time_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        L.orc_time_encode.restype = C.c_double
        L.orc_bits_create.argtypes = [C.POINTER(Bits), C.c_char_p, C.c_int64]
        L.orc_bits_get.argtypes = [C.POINTER(Bits), C.c_int, C.POINTER(C.c_int64)]
        L.orc_bits_show.argtypes = [C.POINTER(Bits), C.c_int, C.POINTER(C.c_int64)]
        L.orc_writer_create.argtypes = [C.POINTER(Writer)]
        L.orc_writer_free.argtypes = [C.POINTER(Writer)]
        L.orc_writer_put_bits.argtypes = [C.POINTER(Writer), C.c_int, C.c_int64, C.c_int]
        L.orc_writer_flush_with_1s.argtypes = [C.POINTER(Writer), C.c_int]
        _lib = L
    return _lib


class OracleError(Exception):
    def __init__(self, status):
        super().__init__("oracle status %d" % status)
        self.status = status


def _np(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype).reshape(shape).copy()


class DecodeResult:
    """What Decoder.decode_a_frame / get_decoded_planes / Sequenced.decode expose."""

    def __init__(self, d, want_blocks):
        self.ncomp = d.ncomp
        self.width, self.height = d.width, d.height
        self.mcus_wide, self.mcus_high, self.blocks_per_mcu = d.mcus_wide, d.mcus_high, d.blocks_per_mcu
        self.nblocks = d.nblocks
        self.hs = list(d.hs)[: d.ncomp]
        self.vs = list(d.vs)[: d.ncomp]
        self.decoded_size = [(d.decoded_width[i], d.decoded_height[i]) for i in range(d.ncomp)]
        self.actual_size = [(d.actual_width[i], d.actual_height[i]) for i in range(d.ncomp)]
        self.planes = [_np(d.plane[i], (d.decoded_height[i], d.decoded_width[i]), np.uint8) for i in range(d.ncomp)]
        self.cropped = [_np(d.cropped[i], (d.actual_height[i], d.actual_width[i]), np.uint8) for i in range(d.ncomp)]
        self.yuv_status = d.yuv_status
        self.chroma = d.chroma
        self.entropy_len = d.entropy_len
        if want_blocks:
            nb = d.nblocks
            self.coefs = _np(d.coefs, (nb, 64), np.int32)
            self.dc_abs = _np(d.dc_abs, (nb,), np.int32)
            self.dequant = _np(d.dequant, (nb, 64), np.int32)
            self.recon = _np(d.recon, (nb, 64), np.uint8)
            self.block_comp = _np(d.block_comp, (nb,), np.int8)

    def yuv(self):
        """Frame.output order: Y then U then V, cropped (frame.ml:66-70)."""
        if self.yuv_status != 0:
            raise OracleError(self.yuv_status)
        return b"".join(p.tobytes() for p in self.cropped[:3])

    def coefs_abs_dc(self):
        """int32 zig-zag blocks with coefs[:,0] replaced by the resolved (absolute) DC."""
        c = self.coefs.copy()
        c[:, 0] = self.dc_abs
        return c


def decode(jpeg, restart_ext=True, want_blocks=False, t81_tables=False):
    d = Decoded()
    flags = (FLAG_RESTART_EXT if restart_ext else 0) | (FLAG_T81_TABLES if t81_tables else 0)
    st = lib().orc_decode(jpeg, len(jpeg), flags, int(want_blocks), C.byref(d))
    if st != 0:
        raise OracleError(st)
    try:
        return DecodeResult(d, want_blocks)
    finally:
        lib().orc_decoded_free(C.byref(d))


def decode_status(jpeg, restart_ext=True, t81_tables=False):
    d = Decoded()
    flags = (FLAG_RESTART_EXT if restart_ext else 0) | (FLAG_T81_TABLES if t81_tables else 0)
    st = lib().orc_decode(jpeg, len(jpeg), flags, 0, C.byref(d))
    if st == 0:
        lib().orc_decoded_free(C.byref(d))
    return st


def split_yuv(yuv, width, height, chroma):
    cw = width if chroma == 444 else width // 2
    ch = height // 2 if chroma == 420 else height
    a = np.frombuffer(yuv, np.uint8)
    if chroma == 400:
        return a[: width * height].reshape(height, width), None, None
    y = a[: width * height].reshape(height, width)
    u = a[width * height : width * height + cw * ch].reshape(ch, cw)
    v = a[width * height + cw * ch : width * height + 2 * cw * ch].reshape(ch, cw)
    return y, u, v


def encode(yuv, width, height, chroma=420, quality=75, restart_interval=0, want_blocks=False):
    """Encoder.encode_420/422/444/monochrome over a raw planar frame (bytes)."""
    y, u, v = split_yuv(yuv, width, height, chroma)
    y = np.ascontiguousarray(y)
    u = np.ascontiguousarray(u) if u is not None else None
    v = np.ascontiguousarray(v) if v is not None else None
    e = Encoded()
    st = lib().orc_encode(
        y.ctypes.data,
        u.ctypes.data if u is not None else None,
        v.ctypes.data if v is not None else None,
        width,
        height,
        chroma,
        quality,
        restart_interval,
        int(want_blocks),
        C.byref(e),
    )
    if st != 0:
        raise OracleError(st)
    try:
        out = bytes(_np(e.bytes, (e.len,), np.uint8))
        if want_blocks:
            return out, _np(e.quant, (e.nblocks, 64), np.int32), _np(e.fdct, (e.nblocks, 64), np.int32)
        return out
    finally:
        lib().orc_encoded_free(C.byref(e))


def header_decode(jpeg, flags=0):
    h = Header()
    st = lib().orc_header_decode_ex(jpeg, len(jpeg), flags, C.byref(h))
    if st != 0:
        raise OracleError(st)
    return h


def extract_entropy_coded_bits(jpeg, start_byte):
    out = np.zeros(len(jpeg) + 16, np.uint8)
    n = C.c_int64()
    st = lib().orc_extract_entropy_coded_bits(jpeg, len(jpeg), start_byte, out.ctypes.data, C.byref(n))
    if st != 0:
        raise OracleError(st)
    return out[: n.value].tobytes()


def write_headers(width, height, chroma, quality, restart_interval=0):
    out = np.zeros(4096, np.uint8)
    n = C.c_int64()
    st = lib().orc_write_headers(width, height, chroma, quality, restart_interval, out.ctypes.data, out.size, C.byref(n))
    if st != 0:
        raise OracleError(st)
    return out[: n.value].tobytes()


def chen_inverse(block):
    b = np.ascontiguousarray(block, np.int64).reshape(64).copy()
    lib().orc_chen_inverse_8x8(b.ctypes.data)
    return b


def chen_forward(block):
    b = np.ascontiguousarray(block, np.int64).reshape(64).copy()
    lib().orc_chen_forward_8x8(b.ctypes.data)
    return b


def quant_scale(chroma, quality):
    out = np.zeros(64, np.int64)
    lib().orc_quant_scale(int(chroma), quality, out.ctypes.data)
    return out


def rle(quant, dc_pred=0):
    q = np.ascontiguousarray(quant, np.int64)
    runs = np.zeros(65, np.int64)
    vals = np.zeros(65, np.int64)
    p = C.c_int64(dc_pred)
    n = lib().orc_rle(q.ctypes.data, C.byref(p), runs.ctypes.data, vals.ctypes.data)
    return [(int(runs[i]), int(vals[i])) for i in range(n)], p.value


def encoder_dc_table(which):
    out = (Code * 256)()
    n = C.c_int()
    lib().orc_encoder_dc_table(which, out, C.byref(n))
    return [(out[i].length, out[i].bits, out[i].data) for i in range(n.value)]


def encoder_ac_table(which):
    out = (Code * 256)()
    rl = (C.c_int * 16)()
    st = lib().orc_encoder_ac_table(which, out, rl)
    assert st == 0
    return [
        [(out[r * 16 + k].length, out[r * 16 + k].bits, out[r * 16 + k].data >> 4, out[r * 16 + k].data & 15) for k in range(rl[r])]
        for r in range(16)
        if rl[r]
    ]


def _plane_op(name, src, scale_w, scale_h):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros((int(h * scale_h), int(w * scale_w)), np.uint8)
    getattr(lib(), name)(src.ctypes.data, w, h, dst.ctypes.data)
    return dst


def supersample_h2(src):
    return _plane_op("orc_supersample_h2", src, 2, 1)


def supersample_hv2(src):
    return _plane_op("orc_supersample_hv2", src, 2, 2)


def subsample_h2(src):
    return _plane_op("orc_subsample_h2", src, 0.5, 1)


def subsample_hv2(src):
    return _plane_op("orc_subsample_hv2", src, 0.5, 0.5)


def crop_clamp(src, dw, dh, x_pos=0, y_pos=0):
    src = np.ascontiguousarray(src, np.uint8)
    h, w = src.shape
    dst = np.zeros((dh, dw), np.uint8)
    lib().orc_crop_clamp(src.ctypes.data, w, h, x_pos, y_pos, dst.ctypes.data, dw, dh)
    return dst


def square_error(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    assert a.shape == b.shape
    return lib().orc_square_error(a.ctypes.data, b.ctypes.data, a.size)


def max_difference(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    assert a.shape == b.shape
    return lib().orc_max_difference(a.ctypes.data, b.ctypes.data, a.size)


def psnr(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    assert a.shape == b.shape
    return lib().orc_psnr(a.ctypes.data, b.ctypes.data, a.shape[1], a.shape[0])


def ycbcr_to_rgb24(y, cb, cr):
    y = np.ascontiguousarray(y, np.uint8)
    cb = np.ascontiguousarray(cb, np.uint8)
    cr = np.ascontiguousarray(cr, np.uint8)
    out = np.zeros(y.shape + (3,), np.uint8)
    lib().orc_ycbcr_to_rgb24(y.ctypes.data, cb.ctypes.data, cr.ctypes.data, y.size, out.ctypes.data)
    return out


def upsample_to_444(planes, chroma):
    """Planar_444.convert_from_420 / convert_from_422 (tools/src/planar_444.ml:52-61,122-131)."""
    y, u, v = planes
    if chroma == 420:
        return y, supersample_hv2(u), supersample_hv2(v)
    if chroma == 422:
        return y, supersample_h2(u), supersample_h2(v)
    return y, u, v


def time_decode(jpegs, restart_ext=True, reps=1):
    arr = (C.c_char_p * len(jpegs))(*jpegs)
    lens = (C.c_int64 * len(jpegs))(*[len(j) for j in jpegs])
    return lib().orc_time_decode(arr, lens, len(jpegs), FLAG_RESTART_EXT if restart_ext else 0, reps)


def time_encode(yuv, width, height, chroma=420, quality=75, restart_interval=0, reps=1):
    a = np.frombuffer(yuv, np.uint8)
    return lib().orc_time_encode(a.ctypes.data, width, height, chroma, quality, restart_interval, reps)
