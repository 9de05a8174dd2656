#!/usr/bin/env python
"""Walk a <prefix>.1.cfr file by the reference's grammar (FMIndex::Save, FMIndex.hpp:571-586 and the
Save() of every member) and print every scalar field -- the `_space` bookkeeping included -- plus a
CRC of every array.  Used to compare an index written by this repo's builder with one written by the
reference's centrifuger-build field by field (tests/test_builder.py), and to debug a mismatch.

    python tools/cfr_dump.py PREFIX [PREFIX2]     # one dump, or a field-by-field diff of two
"""
import struct
import sys
import zlib


class Cur:
    def __init__(self, buf):
        self.b, self.p = buf, 0

    def u64(self):
        v = struct.unpack_from("<Q", self.b, self.p)[0]
        self.p += 8
        return v

    def i32(self):
        v = struct.unpack_from("<i", self.b, self.p)[0]
        self.p += 4
        return v

    def u8(self):
        v = self.b[self.p]
        self.p += 1
        return v

    def arr(self, nbytes):
        v = self.b[self.p:self.p + nbytes]
        assert len(v) == nbytes, "truncated file"
        self.p += nbytes
        return v


def crc(x):
    return "crc32:%08x/%d" % (zlib.crc32(x) & 0xffffffff, len(x))


def alphabet(c, out, name):
    out.append((name + ".space", c.u64()))
    out.append((name + ".method", c.i32()))
    n = c.u64()
    out.append((name + ".n", n))
    if n:
        out.append((name + ".list", bytes(c.arr(n)).decode()))
        out.append((name + ".code", crc(c.arr(256 * 4))))
        out.append((name + ".codeLen", crc(c.arr(256 * 2))))
    return n


def bitvector(c, out, name, keep=None):
    out.append((name + ".space", c.u64()))
    n = c.u64()
    out.append((name + ".n", n))
    for f in ("rb", "sb", "selectSpeed", "selectTypeSupport"):
        out.append((name + "." + f, c.i32()))
    if n > 0:
        words = (n + 63) // 64
        B = c.arr(words * 8)
        out.append((name + ".B", crc(B)))
        out.append((name + ".rank.space", c.u64()))
        wc = c.u64()
        out.append((name + ".rank.wordCnt", wc))
        R = c.arr(((wc + 7) // 8) * 2 * 8)
        out.append((name + ".rank.R", crc(R)))
        out.append((name + ".select.space", c.u64()))
        out.append((name + ".select.n", c.u64()))
        out.append((name + ".select.speed", c.i32()))
        if keep is not None:
            keep[name] = (bytes(B), bytes(R))


def wavelet(c, out, name, keep=None):
    out.append((name + ".space", c.u64()))
    out.append((name + ".n", c.u64()))
    an = alphabet(c, out, name + ".alphabet")
    nodes = c.i32()
    out.append((name + ".tNodeCnt", nodes))
    out.append((name + ".selectSpeed", c.i32()))
    if an == 0:
        return
    for i in range(nodes):
        nm = "%s.node%d" % (name, i)
        out.append((nm + ".prefix", c.u64()))
        out.append((nm + ".prefixLen", c.i32()))
        out.append((nm + ".children", (c.i32(), c.i32())))
        bitvector(c, out, nm + ".v", keep)


def dump(prefix, keep=None):
    buf = memoryview(open(prefix + ".1.cfr", "rb").read())
    c, out = Cur(buf), []
    out.append(("n", c.u64()))
    out.append(("alphabetBits", c.u64()))
    out.append(("firstISA", c.u64()))
    out.append(("lastChr", chr(c.u8())))
    out.append(("bwt.space", c.u64()))
    out.append(("bwt.n", c.u64()))
    alphabet(c, out, "bwt.alphabet")
    out.append(("bwt.b", c.u64()))
    out.append(("bwt.blockCnt", c.u64()))
    bitvector(c, out, "bwt.useRunBlock", keep)
    wavelet(c, out, "bwt.waveletSeq", keep)
    wavelet(c, out, "bwt.runBlockSeq", keep)
    alphabet(c, out, "alphabets")
    alphabet(c, out, "plainAlphabetCoder")
    out.append(("C", tuple(c.u64() for _ in range(5))))
    out.append(("aux.n", c.u64()))
    out.append(("aux.sampleStrategy", c.i32()))
    out.append(("aux.sampleRate", c.i32()))
    out.append(("aux.sampleSize", c.u64()))
    pw = c.u64()
    out.append(("aux.precomputeWidth", pw))
    ps = c.u64()
    out.append(("aux.precomputeSize", ps))
    out.append(("aux.adjustedSA0", c.u64()))
    out.append(("aux.sampledSA.size", c.u64()))
    l = c.i32()
    out.append(("aux.sampledSA.l", l))
    sn = c.u64()
    out.append(("aux.sampledSA.n", sn))
    W = c.arr(((sn * l + 63) // 64) * 8)
    out.append(("aux.sampledSA.W", crc(W)))
    P = c.arr(ps * 16)
    out.append(("aux.precomputedRange", crc(P)))
    if keep is not None:
        keep["sampledSA"] = bytes(W)
        keep["precomputedRange"] = bytes(P)
    ml = c.u64()
    out.append(("aux.maxLcp", ml))
    assert ml == 0, "semiLcp arrays are not handled here"
    sc = c.u64()
    out.append(("aux.selectedSA.count", sc))
    out.append(("aux.selectedSAFilterSampleRate", c.i32()))
    sel = c.arr(sc * 16)
    out.append(("aux.selectedSA", crc(sel)))
    if keep is not None:
        keep["selectedSA"] = bytes(sel)
    if c.p < len(buf):
        he = c.u8()
        out.append(("aux.hasEndMarker", he))
    out.append(("file.size", len(buf)))
    out.append(("file.cursor", c.p))
    return out


def main():
    a = dump(sys.argv[1])
    if len(sys.argv) == 2:
        for k, v in a:
            print("%-44s %s" % (k, v))
        return 0
    b = dump(sys.argv[2])
    bad = 0
    for (ka, va), (kb, vb) in zip(a, b):
        flag = "" if (ka, va) == (kb, vb) else "   <-- DIFF"
        bad += bool(flag)
        print("%-44s %-34s %s%s" % (ka, va, vb if (ka, va) != (kb, vb) else "", flag))
    print("%d differing fields" % bad)
    return 1 if bad or len(a) != len(b) else 0


if __name__ == "__main__":
    sys.exit(main())
