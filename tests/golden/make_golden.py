#!/usr/bin/env python3
"""Regenerate tests/golden/ from the read-only reference checkout.

Run in the build container only (``/root/reference`` does not exist on the GPU
box).  It copies the reference's *test data* (not sources) and lifts the
expected outputs ("goldens") out of the reference's own expect / cram tests
into ``reference_goldens.json``, recording the file:line each one came from so
that tests can cite it.  Nothing here is executed at test time.

    python tests/golden/make_golden.py [/root/reference]
"""
import hashlib
import json
import os
import re
import shutil
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def read(rel):
    with open(os.path.join(REF, rel)) as f:
        return f.read()


def line_of(text, needle):
    return text[: text.index(needle)].count("\n") + 1


def ints(s):
    return [int(x) for x in re.findall(r"-?\d+", s)]


def hexdump_bytes(block):
    """Bytes of a Core ``String.Hexdump`` rendering (offset, 16 hex bytes, ascii)."""
    out = bytearray()
    for m in re.finditer(r'"([0-9a-f]{8})  ((?:[0-9a-f]{2} ?| )+?)\s*\|', block):
        out += bytes(int(b, 16) for b in m.group(2).split())
    return bytes(out)


def expect_blocks(text):
    return re.findall(r"\[%expect\s*\{\|(.*?)\|\}\]", text, re.S)


def dedent_sexp(block, start):
    """The text `print_s` wrote: from the line that begins with `start`, without the indentation ppx_expect adds."""
    i = block.index(start)
    pad = i - (block.rfind("\n", 0, i) + 1)
    lines = block[i - pad :].rstrip().split("\n")
    assert all(l[:pad].strip() == "" for l in lines)
    return "\n".join(l[pad:] for l in lines) + "\n"


g = {"_about": "goldens lifted from hardcamls/video-coding's own tests by make_golden.py"}

# ---- data files (jpeg/test_data) ---------------------------------------------------------
files = {}
for name in ["Mouse480.jpg", "mini.jpg", "mini64x64.420", "mini64x64.422", "mini64x64.444"]:
    src = os.path.join(REF, "jpeg/test_data", name)
    shutil.copyfile(src, os.path.join(OUT, name))
    os.chmod(os.path.join(OUT, name), 0o644)
    with open(src, "rb") as f:
        data = f.read()
    files[name] = {"bytes": len(data), "sha256": hashlib.sha256(data).hexdigest()}
g["test_data"] = {"source": "jpeg/test_data/", "files": files}

# ---- Chen DCT example (test_chen_dct.ml) ---------------------------------------------------
t = read("jpeg/model/test/test_chen_dct.ml")
blk = expect_blocks(t)[0]
parts = re.split(r"\b(input|fdct|idct)\b", blk)
chen = {"source": "jpeg/model/test/test_chen_dct.ml:%d" % line_of(t, '"example transform"')}
for i in range(1, len(parts), 2):
    chen[parts[i]] = ints(parts[i + 1])
    assert len(chen[parts[i]]) == 64
g["chen_example"] = chen

# ---- header hexdump + parse (test_encode_headers.ml) ---------------------------------------
t = read("jpeg/model/test/test_encode_headers.ml")
blk = expect_blocks(t)[0]
hdr = hexdump_bytes(blk)
assert len(hdr) == 623, len(hdr)
sexp = blk[blk.index("(header") :]


def sexp_tables(sexp):
    qts = []
    for m in re.finditer(
        r"\(element_precision (\d+)\) \(table_identifier (\d+)\)\s*\(elements\s*\(([^)]*)\)", sexp
    ):
        qts.append({"precision": int(m.group(1)), "id": int(m.group(2)), "elements": ints(m.group(3))})
    hts = []
    for m in re.finditer(
        r"\(length (\d+)\) \(table_class (\d+)\) \(destination_identifier (\d+)\)\s*"
        r"\(lengths \(([^)]*)\)\)\s*\(values\s*\(([^)]*)\)",
        sexp,
    ):
        hts.append(
            {
                "length": int(m.group(1)),
                "class": int(m.group(2)),
                "id": int(m.group(3)),
                "lengths": ints(m.group(4)),
                "values": ints(m.group(5)),
            }
        )
    fr = re.search(r"\(sample_precision (\d+)\) \(width (\d+)\) \(height (\d+)\)\s*\(number_of_components (\d+)\)", sexp)
    comps = [
        {"id": int(a), "h": int(b), "v": int(c), "tq": int(d)}
        for a, b, c, d in re.findall(
            r"\(identifier (\d+)\) \(horizontal_sampling_factor (\d+)\)\s*"
            r"\(vertical_sampling_factor (\d+)\) \(quantization_table_identifier (\d+)\)",
            sexp,
        )
    ]
    scan = [
        {"selector": int(a), "dc": int(b), "ac": int(c)}
        for a, b, c in re.findall(
            r"\(selector (\d+)\) \(dc_coef_selector (\d+)\) \(ac_coef_selector (\d+)\)", sexp
        )
    ]
    return {
        "precision": int(fr.group(1)),
        "width": int(fr.group(2)),
        "height": int(fr.group(3)),
        "ncomp": int(fr.group(4)),
        "components": comps,
        "quant_tables": qts,  # in Header.t list order (= reverse file order)
        "huffman_tables": hts,
        "scan": scan,
        "restart_interval_present": "(restart_interval ())" not in sexp,
    }


g["header_480x320_q20_420"] = {
    "source": "jpeg/model/test/test_encode_headers.ml:%d" % line_of(t, '"example header"'),
    "hex": hdr.hex(),
    "parsed": sexp_tables(sexp),
    # the text itself: print_s [%message (header : Decoder.Header.t)] as Sexp.to_string_hum lays it out
    "sexp_text": dedent_sexp(blk, "(header"),
}

# ---- Mouse480 header + first 64 destuffed entropy bytes (hardcaml/test/test_codeblock_decoder.ml)
t = read("jpeg/hardcaml/test/test_codeblock_decoder.ml")
i0 = t.index('("String.subo entropy_bits ~len:64"')
i1 = t.index("┌Signals", i0)
blk = t[i0:i1]
first64 = hexdump_bytes(blk)
assert len(first64) == 64
g["mouse480"] = {
    "source": "jpeg/hardcaml/test/test_codeblock_decoder.ml:%d" % line_of(t, '("String.subo entropy_bits ~len:64"'),
    "entropy_first64_hex": first64.hex(),
    "parsed": sexp_tables(blk[blk.index("(headers") :]),
    "sexp_text": dedent_sexp(blk, "(headers"),  # print_s [%message (headers : Model.Header.t)]
}

# ---- encoder code tables (test_tables.ml) --------------------------------------------------
t = read("jpeg/model/test/test_tables.ml")
tabs = {}
for blk in expect_blocks(t):
    name = re.search(r"Tables\.Default\.(\w+)", blk).group(1)
    if name.startswith("dc"):
        tabs[name] = [
            {"length": int(a), "bits": int(b), "data": int(c)}
            for a, b, c in re.findall(r"\(length (\d+)\) \(bits (\d+)\) \(data (\d+)\)", blk)
        ]
    else:
        # array (by run) of arrays (by size): every row starts with its size-0 entry (a real
        # code for run 0 / 15, the zero-length dummy otherwise), so split rows there.
        rows = []
        for a, b, c, d in re.findall(
            r"\(length (\d+)\) \(bits (\d+)\)\s*\(data \(\(run (\d+)\) \(size (\d+)\)\)\)", blk
        ):
            e = {"length": int(a), "bits": int(b), "run": int(c), "size": int(d)}
            if e["size"] == 0:
                rows.append([])
            rows[-1].append(e)
        tabs[name] = rows
g["encoder_tables"] = {"source": "jpeg/model/test/test_tables.ml:4-397", "tables": tabs}

# ---- quant table scaling (test_quant_tables.ml) --------------------------------------------
t = read("jpeg/model/test/test_quant_tables.ml")
q = {}
for blk in expect_blocks(t):
    m = re.search(r'"Quant\.scale Quant\.luma (\d+)"\s*\(([^)]*)\)', blk)
    if m:
        q[m.group(1)] = ints(m.group(2))
        assert len(q[m.group(1)]) == 64
g["quant_scale_luma"] = {"source": "jpeg/model/test/test_quant_tables.ml:4-64", "by_quality": q}

# ---- size / magnitude (test_encode_codewords.ml) -------------------------------------------
t = read("jpeg/model/test/test_encode_codewords.ml")
blks = expect_blocks(t)
g["size_ranges"] = {
    "source": "jpeg/model/test/test_encode_codewords.ml:10-32",
    "rows": [
        {"i": int(a), "lo": int(b), "hi": int(c), "size_lo": int(d), "size_hi": int(e)}
        for a, b, c, d, e in re.findall(
            r"\(i (\d+)\) \(lo (\d+)\) \(hi (\d+)\) \(size_lo (\d+)\) \(size_hi (\d+)\)", blks[0]
        )
    ],
}
g["magnitude"] = {
    "source": "jpeg/model/test/test_encode_codewords.ml:34-77",
    "rows": [
        {"value": int(a), "size": int(b), "emag": int(c), "dmag": int(d)}
        for a, b, c, d in re.findall(
            r"\(value (-?\d+)\) \(size (\d+)\) \(emag (\d+)\) \(dmag (-?\d+)\)", blks[1]
        )
    ],
}

# ---- RLE hand cases (test_rle.ml) ----------------------------------------------------------
t = read("jpeg/model/test/test_rle.ml")
cases = []
for m in re.finditer(r'let%expect_test "([^"]+)" =(.*?)\[%expect\s*\{\|(.*?)\|\}\]', t, re.S):
    name, body, exp = m.groups()
    if "block.rle" not in exp:
        continue
    sets = [(int(a), int(b)) for a, b in re.findall(r"block\.quant\.\((\d+)\) <- (\d+)", body)]
    rle = [(int(a), int(b)) for a, b in re.findall(r"\(run (\d+)\) \(value (-?\d+)\)", exp)]
    cases.append({"name": name, "set": sets, "rle": rle})
g["rle_cases"] = {"source": "jpeg/model/test/test_rle.ml:4-93", "cases": cases}

# ---- cram tests: PSNR goldens --------------------------------------------------------------
t = read("jpeg/test/model-encode-and-decode.t")
runs = []
for m in re.finditer(
    r"model encode frame \S*/(mini64x64\.\d+) 64x64 model\.jpg -quality (\d+)(?: -chroma (\d+))?.*?"
    r"compare max-difference[^\n]*\n((?:\s+\d+\n){3}).*?compare psnr[^\n]*\n((?:\s+[\d.]+\n){3})",
    t,
    re.S,
):
    runs.append(
        {
            "input": m.group(1),
            "quality": int(m.group(2)),
            "chroma": m.group(3) or "420",
            "maxdiff_vs_ffmpeg": ints(m.group(4)),
            "psnr": [s for s in m.group(5).split()],
        }
    )
assert len(runs) == 5, runs
g["cram_encode_decode"] = {"source": "jpeg/test/model-encode-and-decode.t:7-72", "runs": runs}
t = read("jpeg/test/test-nonstandard-sizes.t")
m = re.search(r"compare psnr[^\n]*\n((?:\s+[\d.]+\n?){3})", t)
g["cram_52x44"] = {
    "source": "jpeg/test/test-nonstandard-sizes.t:3-15",
    "quality": 95,
    "size": [52, 44],
    "psnr": m.group(1).split(),
}
t = read("jpeg/test/mouse-decode.t")
g["cram_mouse"] = {"source": "jpeg/test/mouse-decode.t:8-13", "maxdiff_vs_ffmpeg": [1, 0, 0]}

# ---- up/down sampling 4x4 vectors (tools/src/planar_444.ml) --------------------------------
t = read("tools/src/planar_444.ml")
ups = {}
for m in re.finditer(r'let%expect_test "([^"]+)" =(.*?)\n;;', t, re.S):
    name, body = m.groups()
    ups[name] = [ints(b) for b in expect_blocks(body)]
g["planar_444"] = {"source": "tools/src/planar_444.ml:139-249", "dumps": ups}

# ---- the cram tests as they are written: the `model` / `oyuv` command lines and the output they must print ----
# (ffmpeg is not in this image: its commands and the comparisons against its output are left out and counted)
cram = {}
for name in ["model-encode-and-decode.t", "test-nonstandard-sizes.t", "mouse-decode.t"]:
    t = read("jpeg/test/" + name)
    steps, skipped = [], 0
    for line in t.split("\n"):
        if line.startswith("  $ "):
            cmd = line[4:].strip()
            keep = cmd.split()[0] in ("model", "oyuv") and "out_ffmpeg" not in cmd
            skipped += 0 if keep or cmd.startswith(".") else 1
            steps.append({"cmd": cmd, "out": []}) if keep else steps.append(None)
        elif line.startswith("  ") and steps and steps[-1] is not None:
            steps[-1]["out"].append(line[2:])
    cram[name] = {"source": "jpeg/test/" + name, "steps": [x for x in steps if x], "skipped_ffmpeg_steps": skipped}
g["cram_scripts"] = cram

# ---- every multi-line s-expression any expect test of the reference holds, as laid out by Sexp.to_string_hum ------
# (pins the layout engine of hcjpeg/sexp.py: parse the text, print it again, compare)
def top_level_sexps(block):
    """(start, end) of the balanced top-level lists of an expect block; None if the block is not made of s-expressions."""
    spans, depth, start, i, n = [], 0, None, 0, len(block)
    while i < n:
        c = block[i]
        if c == '"':
            i += 1
            while i < n and block[i] != '"':
                i += 2 if block[i] == "\\" else 1
        elif c == "(":
            if depth == 0:
                start = i
            depth += 1
        elif c == ")":
            depth -= 1
            if depth < 0:
                return None
            if depth == 0:
                spans.append((start, i + 1))
        i += 1
    return spans if depth == 0 else None


layouts = []
for root, _, names in sorted(os.walk(REF)):
    for name in sorted(names):
        if not name.endswith(".ml"):
            continue
        rel = os.path.relpath(os.path.join(root, name), REF)
        t = read(rel)
        for m in re.finditer(r"\[%expect\s*\{\|(.*?)\|\}\]", t, re.S):
            blk = m.group(1)
            spans = top_level_sexps(blk)
            for a, b in spans or []:
                src = blk[a:b]
                if "\n" not in src:
                    continue
                col = a - (blk.rfind("\n", 0, a) + 1)
                lines = src.split("\n")
                if any(l[:col].strip() for l in lines[1:]):
                    continue  # not the output of one print_s
                layouts.append({"source": "%s:%d" % (rel, line_of(t, src)), "text": "\n".join([lines[0]] + [l[col:] for l in lines[1:]])})
with open(os.path.join(OUT, "sexp_layouts.json"), "w") as f:
    json.dump({"_about": "multi-line print_s outputs of the reference's expect tests (make_golden.py)", "layouts": layouts}, f, indent=1)
print("wrote", f.name, len(layouts), "layouts")

with open(os.path.join(OUT, "reference_goldens.json"), "w") as f:
    json.dump(g, f, indent=1, sort_keys=True)
print("wrote", os.path.join(OUT, "reference_goldens.json"))
