"""The text of `model decode log` / `model encode log` (jpeg/bin/model.ml:46-68,108-142) from the device taps.

The reference prints its block-level logs with ``print_s [%message ...]``: s-expressions laid out by
``Sexp.to_string_hum`` (sexplib0 ``pp_hum_indent``: every list in a ``Format`` box of indent 1, a break hint of one
blank between its elements, margin 78).  RTL bring-up users diff these logs, so the layout is part of the interface.
This module restates

* the ``Format`` pretty-printing engine as far as that printer uses it (token queue, size resolution through the
  scan stack, the line-break rule of ``pp_open_box`` boxes) -- :func:`to_string_hum`;
* ``sexp_of`` of the records the two commands print: ``Decoder.Header.t`` (decoder.ml:6-13 over markers.ml),
  ``Decoder.Component.Summary.t`` (decoder.ml:189-203), ``Encoder.Block.t`` (encoder.ml:17-66), the hex blocks of
  ``Util.sexp_of_block`` (util.ml:3-28).

Pinned by the reference's own expect tests (tests/test_sexp_log.py; fixtures under tests/golden/ lifted by
make_golden.py): the ``(header ...)`` text of jpeg/model/test/test_encode_headers.ml and the ``(headers ...)`` text of
jpeg/hardcaml/test/test_codeblock_decoder.ml are reproduced from the bytes of the files, and every multi-line
``print_s`` output any expect test of the reference holds (68 of them) parses and prints back to itself.  The reference
holds no expected text for the per-block records; they go through the same printer.
"""
import collections

_INFINITY = 1000000010


class _Format:
    """OCaml's Format engine (stdlib format.ml) for text, ``pp_open_box``, ``pp_print_space`` and ``pp_close_box``."""

    def __init__(self, margin=78):
        self.margin = margin
        self.max_indent = margin - 10
        self.out = []
        self.is_new_line = True
        self._rinit()

    def _rinit(self):
        self.queue = collections.deque()
        self.left_total = self.right_total = 1
        self.scan_stack = [(-1, [-1, ("text", ""), 0])]
        self.format_stack = []
        self.current_indent = 0
        self.curr_depth = 0
        self.space_left = self.margin
        self._open_box_gen(0, "hov")  # pp_open_sys_box

    # ---- output side -------------------------------------------------------------------------
    def _break_new_line(self, offset, width):
        self.out.append("\n")
        self.is_new_line = True
        indent = self.margin - width + offset
        self.current_indent = min(self.max_indent, indent)
        self.space_left = self.margin - self.current_indent
        self.out.append(" " * self.current_indent)

    def _break_same_line(self, width):
        self.space_left -= width
        self.out.append(" " * width)

    def _format_token(self, size, token):
        kind = token[0]
        if kind == "text":
            self.space_left -= size
            self.out.append(token[1])
            self.is_new_line = False
        elif kind == "begin":
            _, off, ty = token
            if self.margin - self.space_left > self.max_indent and self.format_stack:  # pp_force_break_line
                bty, bwidth = self.format_stack[-1]
                if bwidth > self.space_left and bty != "fits":
                    self._break_new_line(0, bwidth)
            width = self.space_left - off
            self.format_stack.append((ty if size > self.space_left else "fits", width))
        elif kind == "end":
            if self.format_stack:
                self.format_stack.pop()
        else:  # break
            _, fits_width, off = token
            if not self.format_stack:
                return
            ty, width = self.format_stack[-1]
            if ty == "hov":
                if size > self.space_left:
                    self._break_new_line(off, width)
                else:
                    self._break_same_line(fits_width)
            elif ty == "box":
                if self.is_new_line:
                    self._break_same_line(fits_width)
                elif size > self.space_left:
                    self._break_new_line(off, width)
                elif self.current_indent > self.margin - width + off:
                    self._break_new_line(off, width)
                else:
                    self._break_same_line(fits_width)
            else:  # fits
                self._break_same_line(fits_width)

    def _advance_left(self):
        while self.queue:
            elem = self.queue[0]
            size = elem[0]
            if size < 0 and self.right_total - self.left_total < self.space_left:
                return
            self.queue.popleft()
            self._format_token(size if size >= 0 else _INFINITY, elem[1])
            self.left_total += elem[2]

    # ---- scanning side -----------------------------------------------------------------------
    def _enqueue(self, elem):
        self.right_total += elem[2]
        self.queue.append(elem)

    def _set_size(self, is_break):
        if not self.scan_stack:
            return
        left_total, elem = self.scan_stack[-1]
        if left_total < self.left_total:
            self.scan_stack = [(-1, [-1, ("text", ""), 0])]
            return
        kind = elem[1][0]
        if (kind == "break" and is_break) or (kind == "begin" and not is_break):
            elem[0] += self.right_total
            self.scan_stack.pop()

    def _scan_push(self, set_break_size, elem):
        self._enqueue(elem)
        if set_break_size:
            self._set_size(True)
        self.scan_stack.append((self.right_total, elem))

    def _open_box_gen(self, indent, ty):
        self.curr_depth += 1
        self._scan_push(False, [-self.right_total, ("begin", indent, ty), 0])

    def open_box(self, indent):
        self._open_box_gen(indent, "box")

    def close_box(self):
        if self.curr_depth > 1:
            self._enqueue([0, ("end",), 0])
            self._set_size(True)
            self._set_size(False)
            self.curr_depth -= 1

    def print_string(self, s):
        self._enqueue([len(s), ("text", s), len(s)])
        self._advance_left()

    def print_space(self):
        self._scan_push(True, [-self.right_total, ("break", 1, 0), 1])

    def flush(self):
        while self.curr_depth > 1:
            self.close_box()
        self.right_total = _INFINITY
        self._advance_left()
        text = "".join(self.out)
        self.out = []
        self._rinit()
        return text


def _must_escape(s):
    """sexplib0 ``must_escape``: atoms that print quoted."""
    if s == "":
        return True
    for i, c in enumerate(s):
        if c in '"();\\' or ord(c) <= 32 or ord(c) >= 127:
            return True
        if c == "|" and i > 0 and s[i - 1] == "#":
            return True
        if c == "#" and i > 0 and s[i - 1] == "|":
            return True
    return False


def _atom(s):
    if not _must_escape(s):
        return s
    esc = {'"': '\\"', "\\": "\\\\", "\n": "\\n", "\t": "\\t", "\r": "\\r", "\b": "\\b"}
    return '"' + "".join(esc.get(c, c if 32 <= ord(c) < 127 else "\\%03d" % ord(c)) for c in s) + '"'


def to_string_hum(sexp, indent=1, margin=78):
    """``Sexp.to_string_hum``: ``sexp`` is a str (atom) or a list / tuple of s-expressions."""
    f = _Format(margin)

    def pp(x):
        if isinstance(x, str):
            f.print_string(_atom(x))
        elif not x:
            f.print_string("()")
        else:
            f.open_box(indent)
            f.print_string("(")
            pp(x[0])
            for y in x[1:]:
                f.print_space()
                pp(y)
            f.print_string(")")
            f.close_box()

    pp(sexp)
    return f.flush()


def of_string_many(text):
    """``Sexp.of_string_many``: the s-expressions of ``text`` (atoms as str, lists as list); quoted atoms are unescaped."""
    n = len(text)

    def skip(p):
        while p < n and text[p] in " \n\t\r":
            p += 1
        return p

    def parse(p):
        p = skip(p)
        if p >= n:
            raise ValueError("unexpected end of input")
        c = text[p]
        if c == "(":
            p, items = p + 1, []
            while True:
                p = skip(p)
                if p >= n:
                    raise ValueError("unterminated list")
                if text[p] == ")":
                    return items, p + 1
                item, p = parse(p)
                items.append(item)
        if c == ")":
            raise ValueError("unbalanced ')' at %d" % p)
        if c == '"':
            q, out = p + 1, []
            while text[q] != '"':
                if text[q] != "\\":
                    out.append(text[q])
                    q += 1
                    continue
                e = text[q + 1]
                if e.isdigit():
                    out.append(chr(int(text[q + 1 : q + 4])))
                    q += 4
                elif e == "\n":  # line continuation: the newline and the blanks behind it are skipped
                    q += 2
                    while text[q] in " \t":
                        q += 1
                else:
                    out.append({"n": "\n", "t": "\t", "r": "\r", "b": "\b"}.get(e, e))
                    q += 2
            return "".join(out), q + 1
        q = p
        while q < n and text[q] not in ' \n\t\r()"':
            q += 1
        return text[p:q], q

    out, pos = [], skip(0)
    while pos < n:
        item, pos = parse(pos)
        out.append(item)
        pos = skip(pos)
    return out


def print_s(sexp):
    """What ``print_s`` writes: the human layout and a newline."""
    return to_string_hum(sexp) + "\n"


# ---- sexp_of the records --------------------------------------------------------------------------
def _ints(xs):
    return [str(int(x)) for x in xs]


def _record(obj, names):
    return [[n, str(int(getattr(obj, n)))] for n in names]


def sexp_of_header(h):
    """``Decoder.Header.sexp_of_t`` of an :class:`hcjpeg.Header` (field order of decoder.ml:6-13; the table lists are in the
    model's list order, i.e. the reverse of the file order, as hcj_header_decode reports them)."""
    frame = []
    if h.has_frame:
        comps = [
            _record(h.components[i], ("identifier", "horizontal_sampling_factor", "vertical_sampling_factor", "quantization_table_identifier"))
            for i in range(h.number_of_components)
        ]
        frame = [
            [["length", str(h.sof_length)]]
            + _record(h, ("sample_precision", "width", "height", "number_of_components"))
            + [["components", comps]]
        ]
    qts = []
    for i in range(h.n_quant_tables):
        q = h.quant_tables[i]
        qts.append(_record(q, ("length", "element_precision", "table_identifier")) + [["elements", _ints(q.elements)]])
    hts = []
    for i in range(h.n_huffman_tables):
        t = h.huffman_tables[i]
        hts.append(
            _record(t, ("length", "table_class", "destination_identifier"))
            + [["lengths", _ints(t.lengths)], ["values", _ints(t.values[: t.nvalues])]]
        )
    dri = [[["length", str(h.dri_length)], ["restart_interval", str(h.restart_interval)]]] if h.has_restart_interval else []
    scan = []
    if h.has_scan:
        sc = [_record(h.scan_components[i], ("selector", "dc_coef_selector", "ac_coef_selector")) for i in range(h.number_of_image_components)]
        scan = [
            [["length", str(h.sos_length)], ["number_of_image_components", str(h.number_of_image_components)], ["scan_components", sc]]
            + _record(h, ("start_of_predictor_selection", "end_of_predictor_selection", "successive_approximation_bit_high",
                          "successive_approximation_bit_low"))
        ]
    return [["frame", frame], ["quant_tables", qts], ["huffman_tables", hts], ["restart_interval", dri], ["scan", scan]]


def _hex_block(values, digits):
    """``Util.sexp_of_block`` (util.ml:3-20): 8 rows of 8, the low ``digits`` hex digits of every value."""
    mask = (1 << (4 * digits)) - 1
    v = [int(x) & mask for x in values]
    return [["%0*x" % (digits, v[8 * y + x]) for x in range(8)] for y in range(8)]


def sexp_of_component_summary(rec, identifiers):
    """``Decoder.Component.Summary.sexp_of_t`` (decoder.ml:189-203) of one record of ``Batch.block_log``.  ``coefs`` as the
    model holds them: slot 0 is the DC differential.  ``identifiers``: the component identifier of every scan component
    (``Header.scan_components[k].selector``; the record carries the scan component's index)."""
    return [
        ["x", str(int(rec["x"]))],
        ["y", str(int(rec["y"]))],
        ["dc_pred", str(int(rec["dc_pred"]))],
        ["component.identifier", str(int(identifiers[int(rec["component"])]))],
        ["coefs", _hex_block(rec["coefs"], 3)],
        ["dequant", _hex_block(rec["dequant"], 3)],
        ["idct", _hex_block(rec["idct"], 2)],
        ["recon", _hex_block(rec["recon"], 2)],
    ]


def sexp_of_encoder_block(rec, verbose=False):
    """``Encoder.Block.sexp_of_t`` (encoder.ml:17-66) of one record of ``Context.encode_block_log``; ``verbose`` = the
    encoder was created with ``~compute_reconstruction_error:true`` (`-verbose`)."""
    n = int(rec["nrle"])
    rle = [[["run", str(int(rec["rle_run"][i]))], ["value", str(int(rec["rle_value"][i]))]] for i in range(n)]
    decoded = []
    if verbose:
        decoded = [[
            ["dequant", _hex_block(rec["dequant"], 3)],
            ["idct", _hex_block(rec["idct"], 3)],
            ["recon", _hex_block(rec["recon"], 2)],
            ["error", _hex_block(rec["error"], 2)],
        ]]
    return [
        ["x_pos", str(int(rec["x_pos"]))],
        ["y_pos", str(int(rec["y_pos"]))],
        ["input_pixels", _hex_block(rec["input_pixels"], 2)],
        ["fdct", _hex_block(rec["fdct"], 3)],
        ["quant", _hex_block(rec["quant"], 3)],
        ["dc_pred", str(int(rec["dc_pred"]))],
        ["rle", rle],
        ["decoded", decoded],
    ]


# ---- the two commands -----------------------------------------------------------------------------
def decode_log(bits, ctx=None, first=0, count=None):
    """stdout of ``model decode log INPUT-BITS`` (jpeg/bin/model.ml:46-68): the header, then one record per block in
    ``Sequenced.decode`` order.  Pure model semantics (restart markers are not interpreted, like the model)."""
    from . import OUT_PLANES, header_decode
    from .model import default_context

    ctx = ctx or default_context()
    header = header_decode(bits)
    ids = [header.scan_components[k].selector for k in range(header.number_of_image_components)]
    out = [print_s(["header", sexp_of_header(header)])] if first == 0 else []
    with ctx.batch([bits], OUT_PLANES, 0) as b:
        b.decode()
        for k, rec in enumerate(b.block_log(0, first, count)):
            out.append(print_s([["!block_number", str(first + k)], ["component", sexp_of_component_summary(rec, ids)]]))
    return "".join(out)


def encode_log(frame, width, height, chroma=420, quality=75, verbose=False, ctx=None, first=0, count=None):
    """stdout of ``model encode log INPUT-FRAME WxH [-quality Q] [-chroma C] [-verbose]`` (jpeg/bin/model.ml:108-142).
    The command never increments its block counter (model.ml:139-141), so every record says 0."""
    from .model import default_context

    ctx = ctx or default_context()
    recs = ctx.encode_block_log(frame, width, height, chroma, quality, 0, first, count)
    return "".join(print_s([["!block_number", "0"], ["block", sexp_of_encoder_block(r, verbose)]]) for r in recs)
