"""The oracle shares the product's I/O and formatting headers (csrc/io/hts_io.hpp, fmt_util.hpp), so CUDA-vs-oracle
agreement cannot vouch for them. These tests pin them to independent implementations instead:
  * BAM decoding (BGZF inflate, record fields, 4-bit bases, qualities, CIGAR, end_pos) record by record against the
    pure-Python reader tests/golden/bamlite.py, on every fixture BAM, sequential and threaded loader;
  * CigarStringView::read_pos as restated for the host (hts_io.hpp) and for the kernels (phase_core.h) against a
    straightforward Python walk of the CIGAR;
  * the `freq` text (ryu shortest round-trip, src/common.rs via csv/serde) against digits from Python's repr();
  * record ids against hashlib.sha1 over Rust's `{:?}` rendering of the byte vector (src/microphasing.rs:667-675);
  * TSV field quoting against Python's csv module."""
import csv
import hashlib
import io
import os
import random
import struct
import subprocess
import sys

import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, GOLDEN)
import bamlite  # noqa: E402


@pytest.fixture(scope="session")
def io_dump(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("io") / "io_dump")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wno-missing-field-initializers", "-o", exe, os.path.join(ROOT, "tests", "units", "io_dump.cpp"),
                    "-lz", "-lpthread"], check=True)
    return exe


def py_read_pos(cigar, start, ref_pos):
    """rust-htslib CigarStringView::read_pos(ref_pos, false, false): (1, qpos) Some, (0, 0) None, (-1, 0) Err."""
    j = None
    n = len(cigar)
    for i, (op, _) in enumerate(cigar):
        if op in "MX=IS":
            j = i
            break
        if op in "DN":
            return (-1, 0)
        if op == "H" and 0 < i < n - 1:
            return (-1, 0)
        if op in "PH" and i == n - 1:
            return (0, 0)
    if j is None:
        return (0, 0)
    rpos, qpos = start, 0
    while rpos <= ref_pos and j < n:
        op, l = cigar[j]
        if op in "MX=":
            if rpos + l > ref_pos:
                return (1, qpos + ref_pos - rpos)
            rpos += l
            qpos += l
        elif op in "SI":
            qpos += l
        elif op in "DN":
            rpos += l
        elif op == "P":
            pass
        elif op == "H":
            return (-1, 0) if j < n - 1 else (0, 0)
        j += 1
    return (0, 0)


BAMS = sorted(os.path.join(GOLDEN, d, "reads.bam") for d in os.listdir(GOLDEN) if os.path.exists(os.path.join(GOLDEN, d, "reads.bam")))


@pytest.mark.parametrize("threads", [1, 3])
@pytest.mark.parametrize("bam", BAMS, ids=[os.path.basename(os.path.dirname(b)) for b in BAMS])
def test_bam_decode_matches_independent_python_reader(io_dump, bam, threads):
    out = subprocess.run([io_dump, "bam", bam, str(threads)], capture_output=True, text=True, check=True).stdout.splitlines()
    refs, recs = bamlite.read_bam(bam)
    assert len(out) == len(recs)
    n_probe = 0
    for line, r in zip(out, recs):
        f = line.split(" ")
        cig = "".join("%d%s" % (l, op) for op, l in r.cigar) or "*"
        assert [int(f[0]), int(f[1]), int(f[2]), int(f[3]), int(f[4]), int(f[5])] == [r.tid, r.pos, r.end_pos(), r.mapq, r.flag, r.l_seq]
        assert f[6] == r.qname and f[7] == cig and f[8] == r.seq and f[9] == r.qual.hex()
        for probe, got in zip([r.pos, r.pos + 7, r.pos + 50, r.end_pos() - 1, r.end_pos()], f[10:15]):
            a, qa, b, qb = (int(x) for x in got.split(":"))
            want = py_read_pos(r.cigar, r.pos, probe)
            assert (a, qa) == want, (cig, probe)
            # the kernels' statement folds Err into None (the caller treats both as "no support") and bounds q by l_seq
            # (an empty CIGAR means "one M over the whole read" to the kernels - the packer's short form; unmapped records never get there)
            if r.cigar:
                assert (b == 1) == (want[0] == 1) and (b != 1 or qb == want[1]), (cig, probe)
            n_probe += 1
    assert n_probe == 5 * len(recs)


def ryu_text(v):
    """Rust's `ryu` pretty printing of an f64 from Python's shortest round-trip digits."""
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "inf" if v > 0 else "-inf"
    sign = "-" if struct.pack(">d", v)[0] & 0x80 else ""
    v = abs(v)
    if v == 0:
        return sign + "0.0"
    mant, _, exp = ("%r" % v).partition("e")
    if exp:
        ip, _, fp = mant.partition(".")
        digits = (ip + fp).lstrip("0") or "0"
        e10 = int(exp) - len(fp)
        digits_stripped = digits.rstrip("0")
        e10 += len(digits) - len(digits_stripped)
        digits = digits_stripped
    else:
        ip, _, fp = mant.partition(".")
        fp = fp.rstrip("0")
        raw = ip + fp
        digits = raw.lstrip("0")
        e10 = -len(fp)
        stripped = digits.rstrip("0")
        e10 += len(digits) - len(stripped)
        digits = stripped
    k = len(digits)
    kk = k + e10
    if 0 <= e10 and kk <= 16:
        return sign + digits + "0" * e10 + ".0"
    if 0 < kk <= 16:
        return sign + digits[:kk] + "." + digits[kk:]
    if -5 < kk <= 0:
        return sign + "0." + "0" * (-kk) + digits
    if k == 1:
        return sign + digits + "e" + str(kk - 1)
    return sign + digits[0] + "." + digits[1:] + "e" + str(kk - 1)


def test_float_record_id_and_csv_formatting_match_independent_implementations(io_dump):
    rng = random.Random(12345)
    floats = [1.0, 0.5, 0.15, 0.5833333333333334, 1e-7, 1e-5, 9.999e-6, 1e16, 9999999999999998.0, 1e15, 123456.789, 0.1 + 0.2, 2.0 / 3.0, float("nan"),
              0.0, 5e-324, 1.7976931348623157e308, 1e21, 1e22, 0.001, 0.0001, 0.00001234]
    floats += [a / b for a in range(1, 40) for b in range(a, 60, 7)]          # count / depth ratios as the path produces them
    floats += [(a / b) * (c / 17.0) for a in range(1, 9) for b in range(9, 14) for c in range(1, 5)]
    floats += [rng.random() * 10 ** rng.randint(-12, 18) for _ in range(400)]
    ids = []
    for _ in range(60):
        n = rng.choice([0, 1, 26, 27, 28, 33, 64])
        seq = bytes(rng.choice(b"ACGTacgtN") for _ in range(n))
        ids.append(("ENST%011d" % rng.randrange(10 ** 9), rng.randrange(0, 3 * 10 ** 8), rng.choice("FR"), seq))
    fields = [b"", b"plain", b"with\ttab", b'with"quote', b"line\nbreak", b"cr\rhere", b"p.Gly12Asp|p.Ala4Val", b'"', b"a,b"]
    fields += [bytes(rng.choice(b'ab\t"\n\r,| ') for _ in range(rng.randint(1, 12))) for _ in range(60)]
    lines = ["f %016x" % struct.unpack(">Q", struct.pack(">d", v))[0] for v in floats]
    lines += ["id %s %d %s %s" % (tx, off, st, seq.hex() or "00"[:0] or "-") for tx, off, st, seq in ids if seq] 
    lines += ["csv %s" % (f.hex() or "-") for f in fields]
    out = subprocess.run([io_dump, "fmt"], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.splitlines()
    want = [ryu_text(v) for v in floats]
    for tx, off, st, seq in ids:
        if not seq:
            continue
        msg = "[" + ", ".join(str(b) for b in seq) + "]" + tx + str(off)
        want.append(hashlib.sha1(msg.encode()).hexdigest()[:15] + st)
    for f in fields:
        buf = io.StringIO()
        csv.writer(buf, delimiter="\t", lineterminator="\n", quoting=csv.QUOTE_MINIMAL).writerow([f.decode("latin-1"), "x"])
        text = buf.getvalue()
        want.append(text[:-3].encode("latin-1").hex())  # drop "\tx\n"
    assert len(out) == len(want)
    for i, (g, w) in enumerate(zip(out, want)):
        assert g == w, (i, lines[i], g, w)


def test_threaded_bam_loader_matches_sequential_loader_on_a_multi_batch_file(tmp_path):
    """The file drivers load the BAM with the threaded loader: records point into the inflated BGZF batches it keeps, a pool
    parses the headers while the next batch is framed, a record that straddles two batches gets a buffer of its own. The
    fixture BAMs fit one batch, so a 30 MB synthetic file (two batches of 256 blocks, straddlers between them) is loaded both
    ways and compared record by record (fields, bases, qualities, CIGAR)."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    d = str(tmp_path / "sample")
    os.makedirs(d)
    subprocess.run([os.path.join(ROOT, "oracle", "_build", "mph_synth_files"), d, "77", "160", "100", "1", "1", "0.1", "0.1"], check=True,
                   stdout=subprocess.DEVNULL)
    assert os.path.getsize(os.path.join(d, "reads.bam")) > 20 << 20
    exe = str(tmp_path / "loader_cmp")
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", exe, os.path.join(ROOT, "tests", "units", "loader_cmp.cpp"), "-lz", "-lpthread"], check=True)
    r = subprocess.run([exe, os.path.join(d, "reads.bam")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ", 0 differences" in r.stdout
    # every guessed segment boundary discarded: the checker frames all segments again from the real boundaries
    r = subprocess.run([exe, os.path.join(d, "reads.bam")], capture_output=True, text=True, env=dict(os.environ, MPH_IO_BAD_GUESS="1", MPH_IO_TRACE="1"))
    assert r.returncode == 0, r.stdout + r.stderr
    assert ", 0 differences" in r.stdout and "framed again" in r.stderr and ", 0 framed again" not in r.stderr


def test_device_deflate_decoder_matches_zlib(tmp_path):
    """core/inflate_core.h is what one GPU thread runs per BGZF block (kernels/inflate_kernels.cu). Here it runs on the CPU
    against zlib: every block of the fixture BAMs, zlib-written streams with stored / fixed / dynamic blocks at several levels
    and strategies, and corrupted streams (an error code, never a write outside the output slice)."""
    exe = str(tmp_path / "inflate_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-o", exe, os.path.join(ROOT, "tests", "units", "inflate_check.cpp"), "-lz"], check=True)
    bams = sorted(os.path.join(GOLDEN, d, "reads.bam") for d in os.listdir(GOLDEN) if os.path.exists(os.path.join(GOLDEN, d, "reads.bam")))
    assert bams
    r = subprocess.run([exe] + bams, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("inflate ok")
