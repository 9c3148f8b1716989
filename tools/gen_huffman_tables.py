#!/usr/bin/env python3
"""Generates swift-mp3_b200/csrc/iso_huffman.inc and tests/iso_huffman.json: the ISO 11172-3 Table B.7 Huffman code tables
1-3, 5-13, 15, 16 (16-23) and 24 (24-31) plus the two count1 (quadruple) tables A / B, for the opt-in ISO mode.

Source of the numbers: the reference holds literals only for tables 1-3, 5-10, 13 and 15 (dead code except 15,
Sources/SwiftMP3/MP3Encoder.swift:2288-2506), and its table 10 has two typos — entries (5,4) and (5,5) read (11, 53) / (11, 52)
where ISO has (11, 21) / (11, 20); as written the table is not a prefix code.  The complete set is therefore read from the
decoder tables of the libavcodec that ships in this image's opencv wheel (FFmpeg's mpegaudiodec_common.c stores, per table,
the code lengths and the (x << 4 | y) symbols in code order; codes follow canonically), and is accepted only if
  * every table is a complete prefix code (Kraft sum exactly 1),
  * tables 1-3, 5-9 and 15 equal the reference's literals entry for entry, table 10 in all but the two typos, table 13 in all
    256 lengths (32 of the reference's dead table-13 code words deviate from ISO) — checked when /root/reference is present.
The generated files are data, committed; this script documents where they came from."""
import glob
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIZES = [4, 9, 9, 16, 16, 36, 36, 36, 64, 64, 64, 256, 256, 256, 256]
IDS = [1, 2, 3, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15, 16, 24]
LINBITS = {16: 1, 17: 2, 18: 3, 19: 4, 20: 6, 21: 8, 22: 10, 23: 13, 24: 4, 25: 5, 26: 6, 27: 7, 28: 8, 29: 9, 30: 11, 31: 13}


def from_libavcodec():
    import cv2  # noqa: F401  (locates the wheel)
    so = glob.glob(os.path.join(os.path.dirname(os.path.dirname(cv2.__file__)), "opencv_python_headless.libs", "libavcodec*"))[0]
    b = open(so, "rb").read()
    i = b.find(bytes([3, 3, 2, 1, 6, 6, 5, 5, 5, 3, 3, 3, 1, 6, 6, 5, 5, 5, 3, 2, 2, 2]))          # lengths of tables 1, 2, 3
    j = b.find(bytes([0x11, 0x01, 0x10, 0x00]), i - 4096)                                          # symbols of table 1
    assert i > 0 and 0 < i - j < 4096, "Huffman tables not found in " + so
    qb = b.find(bytes([1, 4, 4, 5, 4, 6, 5, 6, 4, 5, 5, 6, 5, 6, 6, 6]))
    qc = b.find(bytes([1, 5, 4, 5, 6, 5, 4, 4, 7, 3, 6, 0, 7, 2, 3, 1]))
    assert qb > 0 and qc > 0
    tabs, o = {}, 0
    for s, t in zip(SIZES, IDS):
        dim = int(round(s ** 0.5))
        ln = [[0] * dim for _ in range(dim)]; cd = [[0] * dim for _ in range(dim)]
        code = 0
        for l, sy in zip(b[i + o:i + o + s], b[j + o:j + o + s]):
            ln[sy >> 4][sy & 15] = l; cd[sy >> 4][sy & 15] = code >> (32 - l); code += 1 << (32 - l)
        assert code == 1 << 32, "table %d is not a complete prefix code" % t
        tabs[t] = {"dim": dim, "len": ln, "code": cd}
        o += s
    quad = {"len": [list(b[qb:qb + 16]), list(b[qb + 16:qb + 32])], "code": [list(b[qc:qc + 16]), list(b[qc + 16:qc + 32])]}
    for k in range(2):
        assert abs(sum(2.0 ** -l for l in quad["len"][k]) - 1.0) < 1e-12
    return tabs, quad, os.path.basename(so)


def check_against_reference(tabs):
    path = "/root/reference/Sources/SwiftMP3/MP3Encoder.swift"
    if not os.path.exists(path):
        return "reference not present: not cross-checked in this run"
    src = open(path).read()
    notes = []
    for m in re.finditer(r"static let table(\d+) = HuffmanTable\(\s*maxValue: (\d+),\s*table: \[(.*?)\n    \]\s*\)", src, re.S):
        n = int(m.group(1))
        rows = [[(int(a), int(c)) for a, c in re.findall(r"\((\d+),\s*(\d+)\)", line.split("//")[0])] for line in m.group(3).split("\n")]
        rows = [r for r in rows if r]
        bad = [(x, y) for x, r in enumerate(rows) for y, (l, c) in enumerate(r) if (tabs[n]["len"][x][y], tabs[n]["code"][x][y]) != (l, c)]
        assert bad == ([(5, 4), (5, 5)] if n == 10 else []), "table %d differs from the reference at %r" % (n, bad)
        notes.append("table %d: %d entries equal" % (n, sum(len(r) for r in rows) - len(bad)) + (" (+ the 2 reference typos)" if bad else ""))
    for n, fn in ((13, "buildTable13"), (15, "buildTable15")):
        k = src.find("static func " + fn)
        if k < 0:
            continue
        arrs = re.findall(r"let (lengths|codes): \[Int\] = \[(.*?)\]", src[k:k + 6000], re.S)
        got = {}
        for name, body in arrs:
            got.setdefault(name, [int(v) for v in body.replace("\n", " ").split(",") if v.strip()])
        if len(got.get("lengths", [])) == 256:
            assert got["lengths"] == [v for r in tabs[n]["len"] for v in r], "table %d lengths" % n
            bad = sum(a != c for a, c in zip(got["codes"], [v for r in tabs[n]["code"] for v in r]))
            # table 15 is the reference's live table and must agree completely; its table 13 is dead code whose code words (not
            # lengths) deviate from ISO in 32 places — FFmpeg's, which decode every MP3 in the wild, are taken
            assert bad == 0 or n == 13, "table %d codes differ in %d places" % (n, bad)
            notes.append("table %d: 256 lengths equal, %d code words equal" % (n, 256 - bad))
    return "; ".join(notes)


def main():
    tabs, quad, so = from_libavcodec()
    note = check_against_reference(tabs)
    print(note)
    base, off = {}, 0
    flat_len, flat_code = [], []
    for t in IDS:
        base[t] = off
        for x in range(tabs[t]["dim"]):
            for y in range(tabs[t]["dim"]):
                flat_len.append(tabs[t]["len"][x][y]); flat_code.append(tabs[t]["code"][x][y])
        off += tabs[t]["dim"] ** 2
    tb, td, tl = [0] * 32, [0] * 32, [0] * 32
    for t in range(32):
        src_t = t if t in tabs else 16 if 16 <= t <= 23 else 24 if t >= 24 else None
        if src_t is None:
            continue
        tb[t], td[t], tl[t] = base[src_t], tabs[src_t]["dim"], LINBITS.get(t, 0)
    with open(os.path.join(ROOT, "swift-mp3_b200", "csrc", "iso_huffman.inc"), "w") as f:
        f.write("// GENERATED by tools/gen_huffman_tables.py — ISO 11172-3 Table B.7 (Huffman tables 1-3, 5-13, 15, 16, 24; count1 tables A, B).\n")
        f.write("// Data, not logic.  Read from %s, accepted after: %s.\n" % (so, note))
        f.write("constexpr int kHuffEntries = %d;\n" % off)
        f.write("// code | length << 24, tables concatenated; entry of (x, y) in table t: kHuffBase[t] + x * kHuffDim[t] + y (x, y clamped to 15)\n")
        f.write("__device__ const uint32_t kHuffPacked[kHuffEntries] = {\n")
        for k in range(0, off, 8):
            f.write("  " + ", ".join("0x%08xu" % (flat_code[i] | flat_len[i] << 24) for i in range(k, min(k + 8, off))) + ",\n")
        f.write("};\n")
        f.write("__device__ const uint8_t kHuffLenFlat[kHuffEntries] = {\n")
        for k in range(0, off, 32):
            f.write("  " + ", ".join("%d" % flat_len[i] for i in range(k, min(k + 32, off))) + ",\n")
        f.write("};\n")
        f.write("// per table_select value 0...31 (4 and 14 do not exist; 0 codes nothing): first entry, row length, linbits\n")
        f.write("__constant__ uint16_t kHuffBaseC[32] = {%s};\n" % ", ".join(map(str, tb)))
        f.write("__constant__ uint8_t kHuffDimC[32] = {%s};\n" % ", ".join(map(str, td)))
        f.write("__constant__ uint8_t kHuffLinbitsC[32] = {%s};\n" % ", ".join(map(str, tl)))
        f.write("// linbits of tables 16...23 / 24...31, four bits each (entry i = (packed >> 4 i) & 15)\n")
        f.write("constexpr uint32_t kLinbits16Packed = 0x%08xu, kLinbits24Packed = 0x%08xu;\n" %
                (sum(LINBITS[16 + i] << (4 * i) for i in range(8)), sum(LINBITS[24 + i] << (4 * i) for i in range(8))))
        f.write("// count1 tables A (count1table_select 0) and B (1), index v << 3 | w << 2 | x << 1 | y; table B is the 4-bit code 15 - index\n")
        f.write("constexpr unsigned long long kQuadLenAPacked = 0x%016xull;   // four bits per length\n" % sum(quad["len"][0][i] << (4 * i) for i in range(16)))
        f.write("constexpr unsigned long long kQuadCodeAPacked = 0x%016xull;  // four bits per code word\n" % sum(quad["code"][0][i] << (4 * i) for i in range(16)))
        assert quad["len"][1] == [4] * 16 and quad["code"][1] == [15 - i for i in range(16)]
    json.dump({"tables": {str(t): tabs[t] for t in IDS}, "linbits": {str(k): v for k, v in LINBITS.items()}, "quad": quad,
               "source": so, "checked": note}, open(os.path.join(ROOT, "tests", "iso_huffman.json"), "w"))
    print("wrote iso_huffman.inc (%d entries) and tests/iso_huffman.json" % off)


if __name__ == "__main__":
    sys.exit(main())
