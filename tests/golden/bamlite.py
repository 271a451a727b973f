"""Tiny pure-Python BGZF/BAM reader used only by the fixture tooling in tests/golden/.

Not part of the product path.  Decodes alignment records (pos, cigar, seq, qual, MD tag)
so that `make_fixtures.py` can rebuild the reference bases the hg38 FASTA would have
supplied (the reference's tests download hg38 at test time).
"""
import struct
import zlib

_SEQ = "=ACMGRSVTWYHKDBN"
_CIG = "MIDNSHP=X"


def bgzf_decompress(path):
    data = open(path, "rb").read()
    out = []
    o = 0
    while o < len(data):
        assert data[o:o + 4] == b"\x1f\x8b\x08\x04", "not BGZF"
        xlen = struct.unpack_from("<H", data, o + 10)[0]
        extra = data[o + 12:o + 12 + xlen]
        bsize = None
        e = 0
        while e + 4 <= len(extra):
            slen = struct.unpack_from("<H", extra, e + 2)[0]
            if extra[e:e + 2] == b"BC":
                bsize = struct.unpack_from("<H", extra, e + 4)[0]
            e += 4 + slen
        cstart = o + 12 + xlen
        cend = o + bsize + 1 - 8
        out.append(zlib.decompress(data[cstart:cend], -15))
        o += bsize + 1
    return b"".join(out)


class Rec:
    __slots__ = ("tid", "pos", "mapq", "flag", "qname", "cigar", "seq", "qual", "md", "l_seq")

    def end_pos(self):
        return self.pos + sum(l for op, l in self.cigar if op in "MDN=X")


def read_bam(path):
    raw = bgzf_decompress(path)
    assert raw[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<i", raw, 4)[0]
    o = 8 + l_text
    n_ref = struct.unpack_from("<i", raw, o)[0]
    o += 4
    refs = []
    for _ in range(n_ref):
        l_name = struct.unpack_from("<i", raw, o)[0]
        name = raw[o + 4:o + 4 + l_name - 1].decode()
        l_ref = struct.unpack_from("<i", raw, o + 4 + l_name)[0]
        refs.append((name, l_ref))
        o += 8 + l_name
    recs = []
    while o < len(raw):
        bs = struct.unpack_from("<i", raw, o)[0]
        p = o + 4
        tid, pos, l_rn, mapq, _bin, n_cig, flag, l_seq, _ntid, _npos, _tlen = struct.unpack_from("<iiBBHHHiiii", raw, p)
        q = p + 32
        r = Rec()
        r.tid, r.pos, r.mapq, r.flag, r.l_seq = tid, pos, mapq, flag, l_seq
        r.qname = raw[q:q + l_rn - 1].decode()
        q += l_rn
        r.cigar = []
        for i in range(n_cig):
            c = struct.unpack_from("<I", raw, q + 4 * i)[0]
            r.cigar.append((_CIG[c & 15], c >> 4))
        q += 4 * n_cig
        sb = raw[q:q + (l_seq + 1) // 2]
        r.seq = "".join(_SEQ[(sb[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
        q += (l_seq + 1) // 2
        r.qual = raw[q:q + l_seq]
        q += l_seq
        r.md = None
        end = p + bs
        while q < end:
            tag = raw[q:q + 2]
            ty = chr(raw[q + 2])
            q += 3
            if ty in "AcC":
                q += 1
            elif ty in "sS":
                q += 2
            elif ty in "iIf":
                q += 4
            elif ty in "ZH":
                z = raw.index(b"\0", q)
                if tag == b"MD":
                    r.md = raw[q:z].decode()
                q = z + 1
            elif ty == "B":
                sub = chr(raw[q])
                cnt = struct.unpack_from("<i", raw, q + 1)[0]
                q += 5 + cnt * {"c": 1, "C": 1, "s": 2, "S": 2, "i": 4, "I": 4, "f": 4}[sub]
            else:
                raise ValueError("bad aux type " + ty)
        recs.append(r)
        o = end
    return refs, recs


def ref_from_md(r):
    """Yield (ref_pos, base) for every reference position the alignment covers (M/=/X/D)."""
    if r.md is None or r.flag & 4:
        return
    # expand MD into per-reference-position tokens: None = match, 'X' = that ref base
    md = []
    i = 0
    s = r.md
    while i < len(s):
        if s[i].isdigit():
            j = i
            while j < len(s) and s[j].isdigit():
                j += 1
            md.extend([None] * int(s[i:j]))
            i = j
        elif s[i] == "^":
            j = i + 1
            while j < len(s) and s[j].isalpha():
                j += 1
            md.extend(("D", c) for c in s[i + 1:j])
            i = j
        else:
            md.append(s[i])
            i += 1
    m = 0
    rp = r.pos
    qp = 0
    for op, l in r.cigar:
        if op in "M=X":
            for _ in range(l):
                t = md[m]
                m += 1
                yield rp, (r.seq[qp] if t is None else t)
                rp += 1
                qp += 1
        elif op in "IS":
            qp += l
        elif op == "D":
            for _ in range(l):
                t = md[m]
                m += 1
                assert isinstance(t, tuple), (r.qname, r.md, r.cigar)
                yield rp, t[1]
                rp += 1
        elif op == "N":
            rp += l
