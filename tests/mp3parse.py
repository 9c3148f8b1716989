"""Test-only parser for the streams this encoder family produces (reference semantics): walks the
frames, reads header + side info, rebuilds the main-data FIFO the way the reference filled the slots (SRC:2110-2121),
and Huffman-decodes every granule-channel with table 15 back to ix[576].  Pure Python / numpy: small inputs only."""
import numpy as np

BITRATES = [0, 32, 40, 48, 56, 64, 80, 96, 112, 128, 160, 192, 224, 256, 320, 0]
SRATES = [44100, 48000, 32000, 0]


class Bits:
    def __init__(self, data, bitpos=0):
        self.d, self.p = data, bitpos

    def get(self, n):
        v = 0
        for _ in range(n):
            v = (v << 1) | ((self.d[self.p >> 3] >> (7 - (self.p & 7))) & 1)
            self.p += 1
        return v


def parse_frames(stream):
    """-> list of dicts (header fields, side info per granule-channel, slot bytes)."""
    frames, pos = [], 0
    while pos + 4 <= len(stream):
        h = int.from_bytes(stream[pos:pos + 4], "big")
        assert h >> 21 == 0x7FF, "lost sync at byte %d" % pos
        prot, bri, sri, pad = (h >> 16) & 1, (h >> 12) & 15, (h >> 10) & 3, (h >> 9) & 1
        mode, mode_ext, copyright, original = (h >> 6) & 3, (h >> 4) & 3, (h >> 3) & 1, (h >> 2) & 1
        ch = 1 if mode == 3 else 2
        size = 144 * BITRATES[bri] * 1000 // SRATES[sri] + pad
        crc = 0 if prot else 2
        side_n = 17 if ch == 1 else 32
        b = Bits(stream, (pos + 4 + crc) * 8)
        mdb = b.get(9); b.get(5 if ch == 1 else 3)
        scfsi = [b.get(4) for _ in range(ch)]
        gcs = []
        for gr in range(2):
            for c in range(ch):
                g = dict(part23=b.get(12), big_values=b.get(9), global_gain=b.get(8), scalefac_compress=b.get(4), ws=b.get(1))
                if g["ws"]:
                    g.update(block_type=b.get(2), mixed=b.get(1), table_select=[b.get(5), b.get(5)], sbg=[b.get(3), b.get(3), b.get(3)])
                else:
                    g.update(block_type=0, mixed=0, table_select=[b.get(5), b.get(5), b.get(5)], region0=b.get(4), region1=b.get(3))
                g.update(preflag=b.get(1), scalefac_scale=b.get(1), count1table=b.get(1))
                gcs.append(g)
        hdr = 4 + crc + side_n
        frames.append(dict(pos=pos, size=size, bitrate_index=bri, sr_index=sri, padding=pad, mode=mode, mode_ext=mode_ext,
                           copyright=copyright, original=original, protection=prot, ch=ch, mdb=mdb, scfsi=scfsi, gc=gcs,
                           slot=bytes(stream[pos + hdr:pos + size]), header=bytes(stream[pos:pos + hdr])))
        pos += size
    assert pos == len(stream), "trailing bytes"
    return frames


def main_data_per_frame(frames):
    """Undo the slot filling: frame n's Huffman bytes were appended to the FIFO before slot n-1 was filled."""
    huff = [(sum(g["part23"] for g in f["gc"]) + 7) // 8 for f in frames]
    fifo = bytearray(); out = []; consumed_slots = 0
    # replay: when frame n is encoded its bytes join the FIFO, then slot n-1 takes min(slot, len) bytes (zero padded)
    # Inverse: walk the slots in order and hand their non-pad bytes back to the frames in order.
    pending = bytearray()          # FIFO content reconstructed from the slots
    need = list(huff)
    data = [bytearray() for _ in frames]
    fi = 0                         # frame currently being refilled
    avail = 0                      # bytes present in the FIFO at the encoder when the slot was filled
    appended = 0
    for n, f in enumerate(frames):
        # slot n is filled after frame n+1 was appended (or at flush): FIFO then holds appended(n+1) - taken bytes
        appended = sum(huff[:min(n + 2, len(frames))])
        taken = sum(len(d) for d in data) + len(pending)
        take = min(len(f["slot"]), appended - taken)
        pending += f["slot"][:take]
        while fi < len(frames) and len(pending) >= need[fi] - len(data[fi]):
            k = need[fi] - len(data[fi]); data[fi] += pending[:k]; del pending[:k]; fi += 1
    return [bytes(d) for d in data], huff


def decode_table15(data, bitpos, big_values, len15, code15, lut=None):
    """-> (ix[576], next bit position).  len15 / code15: 256-entry tables."""
    if lut is None:
        lut = {(int(len15[i]), int(code15[i])): i for i in range(256)}
    b = Bits(data, bitpos); ix = np.zeros(576, np.int32)
    for p in range(big_values):
        code, n = 0, 0
        while True:
            code = (code << 1) | b.get(1); n += 1
            if (n, code) in lut:
                break
            assert n < 14, "bad Huffman code"
        v = lut[(n, code)]; x, y = v >> 4, v & 15
        if x and b.get(1): x = -x
        if y and b.get(1): y = -y
        ix[2 * p], ix[2 * p + 1] = x, y
    return ix, b.p


def decode_stream(stream, len15, code15):
    """-> (frames, ix array [n_gc][576]) decoded from the bytes alone."""
    frames = parse_frames(stream)
    data, _ = main_data_per_frame(frames)
    lut = {(int(len15[i]), int(code15[i])): i for i in range(256)}
    out = []
    for f, d in zip(frames, data):
        pos = 0
        for g in f["gc"]:
            ix, end = decode_table15(d, pos, g["big_values"], len15, code15, lut)
            assert end - pos == g["part23"], "part2_3_length mismatch"
            pos = end; out.append(ix)
    return frames, np.array(out)
