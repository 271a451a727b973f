"""Minimal VCF -> BCF2.2 encoder for the tests (test infrastructure only): turns a text VCF fixture into the binary
form `bcf::Reader::from_path` also accepts, so the BCF decoder of csrc/io/hts_io.hpp can be checked against the same
golden outputs. Written from the VCF specification (section 6, "BCF specification"), independently of the decoder:
sample columns are dropped, every INFO field is encoded according to its header Type."""
import struct
import zlib

INT_MISSING = {1: -128, 2: -32768, 3: -2147483648}


def _bgzf(data):
    out = []
    for o in range(0, len(data), 0xFF00):
        chunk = data[o:o + 0xFF00]
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        comp = co.compress(chunk) + co.flush()
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                   struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    out.append(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
    return b"".join(out)


def _desc(n, t):
    if n < 15:
        return bytes([(n << 4) | t])
    return bytes([0xF0 | t]) + _typed_ints([n])


def _int_type(vals):
    lo = min([v for v in vals if v is not None] or [0])
    hi = max([v for v in vals if v is not None] or [0])
    if lo >= -120 and hi <= 127:
        return 1
    if lo >= -32760 and hi <= 32767:
        return 2
    return 3


def _typed_ints(vals):
    t = _int_type(vals)
    fmt = {1: "<b", 2: "<h", 3: "<i"}[t]
    return _desc(len(vals), t) + b"".join(struct.pack(fmt, INT_MISSING[t] if v is None else v) for v in vals)


def _typed_floats(vals):
    body = b""
    for v in vals:
        body += struct.pack("<I", 0x7F800001) if v is None else struct.pack("<f", v)
    return _desc(len(vals), 5) + body


def _typed_str(s):
    b = s.encode()
    return _desc(len(b), 7) + b


def vcf_to_bcf(vcf_path, bcf_path):
    header, records = [], []
    for line in open(vcf_path):
        line = line.rstrip("\n")
        if line.startswith("##"):
            if not line.startswith("##FORMAT"):
                header.append(line)
        elif line.startswith("#"):
            header.append("\t".join(line.split("\t")[:8]))
        elif line:
            records.append(line.split("\t")[:8])
    contigs, dictionary, info_type = [], ["PASS"], {}
    for h in header:
        if h.startswith("##contig=<"):
            contigs.append(h.split("ID=")[1].split(",")[0].split(">")[0])
        elif h.startswith("##INFO=<") or h.startswith("##FILTER=<"):
            ident = h.split("ID=")[1].split(",")[0].split(">")[0]
            if ident not in dictionary:
                dictionary.append(ident)
            if h.startswith("##INFO"):
                info_type[ident] = h.split("Type=")[1].split(",")[0].split(">")[0]
    for rec in records:  # contigs used but not declared: declare them (bcftools does the same on conversion)
        if rec[0] not in contigs:
            contigs.append(rec[0])
            header.insert(-1, "##contig=<ID=%s>" % rec[0])
    text = ("\n".join(header) + "\n").encode() + b"\0"
    out = bytearray(b"BCF\x02\x02" + struct.pack("<I", len(text)) + text)
    for chrom, pos, ident, ref, alt, qual, flt, info in records:
        alleles = [ref] + ([] if alt == "." else alt.split(","))
        items = [] if info in (".", "") else info.split(";")
        shared = bytearray()
        shared += struct.pack("<iii", contigs.index(chrom), int(pos) - 1, len(ref))
        shared += struct.pack("<I", 0x7F800001) if qual == "." else struct.pack("<f", float(qual))
        shared += struct.pack("<I", (len(alleles) << 16) | len(items))
        shared += struct.pack("<I", 0)  # n_fmt << 24 | n_sample
        shared += _typed_str("" if ident == "." else ident) if ident != "." else _desc(0, 7)
        for a in alleles:
            shared += _typed_str(a)
        shared += _desc(0, 0) if flt in (".", "") else _typed_ints([dictionary.index(f) for f in flt.split(";")])
        for it in items:
            key, _, val = it.partition("=")
            shared += _typed_ints([dictionary.index(key)])
            typ = info_type.get(key, "String")
            if typ == "Flag" or val == "":
                shared += _desc(0, 0)
            elif typ == "Integer":
                shared += _typed_ints([None if v == "." else int(v) for v in val.split(",")])
            elif typ == "Float":
                shared += _typed_floats([None if v == "." else float(v) for v in val.split(",")])
            else:
                shared += _typed_str(val)
        out += struct.pack("<II", len(shared), 0) + shared
    with open(bcf_path, "wb") as f:
        f.write(_bgzf(bytes(out)))
