"""Drop-in mirror of the reference's software-model interface, running on the GPU.

Same names and argument meaning as ``Hardcaml_jpeg_model.Decoder`` / ``.Encoder`` and
``Hardcaml_video_common.Plane`` / ``.Frame`` (jpeg/model/src/decoder.mli, encoder.mli,
common/src/plane.mli, frame.mli) so that parity tests read like the reference's own tests:

    frame = Decoder.decode_a_frame(bits)          # jpeg/bin/model.ml:37
    frame.output(out_channel)                     # jpeg/bin/model.ml:43
    Encoder.encode_420(frame=frame, quality=75, writer=writer); writer.get_buffer()

Where the model raises, these raise ``HcjError`` with the matching status (hcjpeg.STATUS).
"""
import numpy as np

from . import (
    FLAG_DEFAULT,
    OUT_PLANES,
    OUT_YUV,
    Context,
    HcjError,
    frame_info,
    header_decode,
    write_headers,
)

_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


class Plane:
    """common/src/plane.mli: width x height bytes, zero initialised."""

    def __init__(self, width, height, data=None):
        self.width, self.height = width, height
        self.plane = np.zeros((height, width), np.uint8) if data is None else np.asarray(data, np.uint8).reshape(height, width)

    @staticmethod
    def create(width, height):
        return Plane(width, height)

    def output(self, out_channel):  # plane.ml:63-69
        out_channel.write(self.plane.tobytes())

    def input(self, in_channel):  # plane.ml:73-82
        data = in_channel.read(self.width * self.height)
        if len(data) != self.width * self.height:
            raise EOFError("End_of_image")
        self.plane = np.frombuffer(data, np.uint8).reshape(self.height, self.width).copy()


class Frame:
    """common/src/frame.mli."""

    def __init__(self, y, u, v, chroma_subsampling):
        self.y, self.u, self.v, self.chroma_subsampling = y, u, v, chroma_subsampling

    @staticmethod
    def create(chroma_subsampling, width, height):  # frame.ml:32-40
        cw = width if chroma_subsampling == 444 else width // 2
        ch = height // 2 if chroma_subsampling == 420 else height
        return Frame(Plane(width, height), Plane(cw, ch), Plane(cw, ch), chroma_subsampling)

    @property
    def width(self):
        return self.y.width

    @property
    def height(self):
        return self.y.height

    def output(self, out_channel):  # frame.ml:66-70
        for p in (self.y, self.u, self.v):
            p.output(out_channel)

    def input(self, in_channel):  # frame.ml:72-76
        for p in (self.y, self.u, self.v):
            p.input(in_channel)

    def tobytes(self):
        return b"".join(p.plane.tobytes() for p in (self.y, self.u, self.v))

    @staticmethod
    def frombytes(data, chroma_subsampling, width, height):
        import io

        f = Frame.create(chroma_subsampling, width, height)
        f.input(io.BytesIO(data))
        return f


class Writer:
    """common/src/bitstream_writer.mli as far as the encoder's callers use it (get_buffer)."""

    def __init__(self):
        self._buf = bytearray()

    @staticmethod
    def create():
        return Writer()

    def get_buffer(self):
        return bytes(self._buf)

    def bytes_written(self):
        return len(self._buf)


class Decoder:
    """jpeg/model/src/decoder.mli."""

    class Header:
        decode = staticmethod(header_decode)  # decoder.ml:37-70

    def __init__(self, bits, ctx=None, flags=FLAG_DEFAULT):
        """Decoder.init (decoder.ml:304-345): ``bits`` is the whole file."""
        self.bits = bits
        self.ctx = ctx or default_context()
        self.flags = flags
        self.info = frame_info(bits)
        self._planes = None

    @staticmethod
    def init(header, bits, ctx=None):
        return Decoder(bits, ctx)

    def decode(self):  # decoder.ml:397
        outs, st = self.ctx.decode_batch([self.bits], OUT_PLANES, self.flags, raise_on_error=True)
        self._planes = outs[0]

    def get_decoded_planes(self):  # decoder.ml:399-401
        if self._planes is None:
            self.decode()
        f, out, off = self.info, [], 0
        for i in range(f.ncomp):
            w, h = f.decoded_width[i], f.decoded_height[i]
            out.append(Plane(w, h, self._planes[off : off + w * h]))
            off += w * h
        return out

    def get_yuv_frame(self):  # decoder.ml:403-420
        f = self.info
        if f.ncomp < 3:
            raise HcjError(-12, "get_yuv_frame")
        if f.chroma == 0:
            raise HcjError(-11, "Frame.of_planes")
        planes = self.get_decoded_planes()
        crop = [Plane(f.actual_width[i], f.actual_height[i], planes[i].plane[: f.actual_height[i], : f.actual_width[i]]) for i in range(3)]
        return Frame(crop[0], crop[1], crop[2], f.chroma)

    @staticmethod
    def decode_a_frame(bits, ctx=None, flags=FLAG_DEFAULT):
        """decoder.ml:422-427: header, init, decode, cropped frame — in one device pass."""
        ctx = ctx or default_context()
        outs, st = ctx.decode_batch([bits], OUT_YUV, flags, raise_on_error=True)
        f = frame_info(bits)
        y = f.actual_width[0] * f.actual_height[0]
        c = f.actual_width[1] * f.actual_height[1]
        o = outs[0]
        return Frame(
            Plane(f.actual_width[0], f.actual_height[0], o[:y]),
            Plane(f.actual_width[1], f.actual_height[1], o[y : y + c]),
            Plane(f.actual_width[2], f.actual_height[2], o[y + c : y + 2 * c]),
            f.chroma,
        )

    class For_testing:
        @staticmethod
        def mag(cat, code):  # decoder.mli:64-65 (mag' of decoder.ml:73-79)
            from . import lib

            return lib().hcj_mag(cat, code)

        @staticmethod
        def extract_entropy_coded_bits(bits, ctx=None):  # decoder.ml:261-281
            ctx = ctx or default_context()
            with ctx.batch([bits], OUT_PLANES, 0) as b:
                b.decode()
                return b.entropy(0)

        @staticmethod
        def coefficients(bits, ctx=None, flags=FLAG_DEFAULT):
            """Component.coefs of Sequenced.decode with the DC resolved (decoder.ml:167-204,433-435)."""
            ctx = ctx or default_context()
            with ctx.batch([bits], OUT_PLANES, flags) as b:
                b.decode()
                return b.coefficients(0)


class Encoder:
    """jpeg/model/src/encoder.mli."""

    @staticmethod
    def write_headers(params, writer):  # encoder.ml:371-418
        writer._buf += write_headers(params["width"], params["height"], params["chroma"], params["quality"])

    class Parameters:
        @staticmethod
        def c420(width, height, quality):
            return dict(width=width, height=height, quality=quality, chroma=420)

        @staticmethod
        def c422(width, height, quality):
            return dict(width=width, height=height, quality=quality, chroma=422)

        @staticmethod
        def c444(width, height, quality):
            return dict(width=width, height=height, quality=quality, chroma=444)

    @staticmethod
    def _encode(frame, quality, writer, chroma, ctx=None, restart_interval=0):
        ctx = ctx or default_context()
        outs, st = ctx.encode_batch([np.frombuffer(frame.tobytes(), np.uint8)], frame.width, frame.height, chroma, quality, restart_interval)
        if st[0] != 0:
            raise HcjError(st[0], "Encoder.encode_%d" % chroma)
        writer._buf += outs[0]

    @staticmethod
    def encode_420(frame, quality, writer, ctx=None):  # encoder.ml:522-527
        Encoder._encode(frame, quality, writer, 420, ctx)

    @staticmethod
    def encode_422(frame, quality, writer, ctx=None):  # encoder.ml:529-534
        Encoder._encode(frame, quality, writer, 422, ctx)

    @staticmethod
    def encode_444(frame, quality, writer, ctx=None):  # encoder.ml:536-541
        Encoder._encode(frame, quality, writer, 444, ctx)

    @staticmethod
    def encode_monochrome(frame, quality, writer, ctx=None):  # encoder.ml:543-552; ``frame`` is a Plane
        ctx = ctx or default_context()
        outs, st = ctx.encode_batch([np.ascontiguousarray(frame.plane, np.uint8).ravel()], frame.width, frame.height, 400, quality, 0)
        if st[0] != 0:
            raise HcjError(st[0], "Encoder.encode_monochrome")
        writer._buf += outs[0]
