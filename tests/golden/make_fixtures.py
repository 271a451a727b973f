#!/usr/bin/env python3
"""Regenerate tests/golden/ from the reference's own test resources.

Run in the build container only (needs /root/reference, which does not exist on the GPU
box).  For every live integration test of reference tests/lib.rs it

  * copies the inputs (BAM / VCF / GTF) and the checked-in expected outputs, and
  * writes `ref_patch.tsv`: the reference-genome bases the hg38 chromosome FASTA would
    have supplied.  The reference's tests download hg38 at test time (tests/lib.rs:79-104)
    and there is no network here, so the bases are rebuilt from the BAM's CIGAR+MD tags
    (zero conflicts on every fixture), soft-mask case is taken from the expected outputs,
    and — for `normal` mode, where every window is printed — uncovered CDS positions are
    filled from the expected FASTA, with the SHA-1 record ids (which hash sequence,
    transcript and offset) as the cross-check that the fill is right.

`tests/conftest.py::materialize_reference` turns ref_patch.tsv + the reference's own .fai
geometry into a sparse chrN.fa at the original coordinates, so the CLI under test is called
exactly like the reference's tests call it.
"""
import itertools
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import bamlite  # noqa: E402

RES = "/root/reference/tests/resources"

# case -> (subcommand, chrom, bam, vcf, gtf, {expected name: reference file})
CASES = {
    "forward_somatic": ("somatic", "chr14", "test_forward/forward_test.bam", "test_forward/forward_test.vcf",
                        "test_forward/forward_test.gtf",
                        {"out.fa": "test_forward/expected_output/forward_test.fa",
                         "out.tsv": "test_forward/expected_output/forward_test.tsv",
                         "out.normal.fa": "test_forward/expected_output/forward_test.normal.fa"}),
    "empty": ("somatic", "chr14", "test_forward/forward_test.bam", "test_empty/empty_test.vcf",
              "test_forward/forward_test.gtf",
              {"out.fa": "test_empty/expected_output/empty_test.fa",
               "out.tsv": "test_empty/expected_output/empty_test.tsv",
               "out.normal.fa": "test_empty/expected_output/empty_test.normal.fa"}),
    "reverse_somatic": ("somatic", "chr1", "test_reverse/reverse_test.bam", "test_reverse/reverse_test.vcf",
                        "test_reverse/reverse_test.gtf",
                        {"out.fa": "test_reverse/expected_output/reverse_test.fa",
                         "out.tsv": "test_reverse/expected_output/reverse_test.tsv",
                         "out.normal.fa": "test_reverse/expected_output/reverse_test.normal.fa"}),
    "splice_forward_somatic": ("somatic", "chr7", "splice_forward_test/INSIG1.test.bam",
                               "splice_forward_test/INSIG1.test.vcf", "splice_forward_test/INSIG1.test.gtf",
                               {"out.fa": "splice_forward_test/expected_output/splice_forward_test.fa",
                                "out.tsv": "splice_forward_test/expected_output/splice_forward_test.tsv",
                                "out.normal.fa": "splice_forward_test/expected_output/splice_forward_test.normal.fa"}),
    "splice_reverse_somatic": ("somatic", "chr6", "splice_reverse_test/MMS22L.test.bam",
                               "splice_reverse_test/MMS22L.test.vcf", "splice_reverse_test/MMS22L.test.gtf",
                               {"out.fa": "splice_reverse_test/expected_output/splice_reverse_test.fa",
                                "out.tsv": "splice_reverse_test/expected_output/splice_reverse_test.tsv",
                                "out.normal.fa": "splice_reverse_test/expected_output/splice_reverse_test.normal.fa"}),
    "forward_normal": ("normal", "chr14", "test_forward/forward_test.bam", "test_forward/forward_test.germline.vcf",
                       "test_forward/forward_test.gtf",
                       {"out.fa": "test_forward/expected_output/forward_test.germline.fa"}),
    "splice_forward_normal": ("normal", "chr7", "splice_forward_test/INSIG1.test.bam",
                              "splice_forward_test/INSIG1.test.germline.vcf", "splice_forward_test/INSIG1.test.gtf",
                              {"out.fa": "splice_forward_test/expected_output/splice_forward_test.germline.fa"}),
}
# inputs that share a chromosome share one reference patch (like the reference's chrN.fa)
UNSORTED = ("test_unsorted_gtf/chr14.unsorted.BDKRB2_DHRS2.gtf", "test_unsorted_gtf/chr14.sorted.DHRS2_BDKRB2.gtf",
            "test_unsorted_gtf/empty.vcf", "test_unsorted_gtf/forward_test.bam")


def read_vcf(path):
    out = []
    for line in open(path):
        if line.startswith("#"):
            continue
        t = line.rstrip("\n").split("\t")
        out.append((int(t[1]) - 1, t[3], t[4].split(",")))
    return out


def gene_spans(gtf):
    spans = []
    for line in open(gtf):
        t = line.split("\t")
        if len(t) >= 9 and t[2] == "gene":
            spans.append((int(t[3]) - 1, int(t[4])))
    return spans


def build_ref(bam_path):
    _, recs = bamlite.read_bam(bam_path)
    ref = {}
    for r in recs:
        for p, b in bamlite.ref_from_md(r):
            if p in ref:
                assert ref[p] == b, ("MD conflict", p)
            ref[p] = b
    return ref


def case_evidence_from_tsv(tsv_path, ref, variants, lower, upper):
    """Collect soft-mask (case) evidence for reference positions from the expected somatic TSV."""
    if not os.path.exists(tsv_path) or os.path.getsize(tsv_path) == 0:
        return
    rows = [l.rstrip("\n").split("\t") for l in open(tsv_path)]
    hdr = rows[0]
    io, im, inn = hdr.index("offset"), hdr.index("mutant_sequence"), hdr.index("normal_sequence")
    varpos = {p for p, _, _ in variants}
    dels = [(p, len(r) - 1) for p, r, alts in variants if len(r) > 1]
    for row in rows[1:]:
        ws = int(row[io]) - 1
        for seq in (row[inn], row[im]):
            if not seq:
                continue
            cands = [d for d in dels if ws <= d[0] < ws + 40]
            done = False
            for k in range(len(cands) + 1):
                for applied in itertools.combinations(cands, k):
                    skip = {}
                    for p, l in applied:
                        skip[p] = l
                    pos = []
                    i = ws
                    for _ in seq:
                        pos.append(i)
                        i += 1 + skip.get(i, 0)
                    ok = all(p in varpos or p not in ref or ref[p].upper() == c.upper() for p, c in zip(pos, seq))
                    if ok:
                        for p, c in zip(pos, seq):
                            if p in varpos:
                                continue
                            (lower if c.islower() else upper).add(p)
                            ref.setdefault(p, c.upper())
                        done = True
                        break
                if done:
                    break


def apply_case(ref, lower, upper, variants):
    out = dict(ref)
    if not lower:
        return out
    # soft-masked runs are contiguous: a position with no direct evidence inherits the case of
    # the nearest position that has evidence (only inside / next to an evidenced lowercase run)
    ev = sorted([(p, True) for p in lower] + [(p, False) for p in upper])
    import bisect
    keys = [p for p, _ in ev]
    lo, hi = min(lower) - 64, max(lower) + 64
    for p in list(out):
        if p < lo or p > hi:
            continue
        if p in lower:
            out[p] = out[p].lower()
        elif p in upper:
            continue
        else:
            k = bisect.bisect_left(keys, p)
            best = None
            for kk in (k - 1, k):
                if 0 <= kk < len(keys):
                    d = abs(keys[kk] - p)
                    if best is None or d < best[0]:
                        best = (d, ev[kk][1])
            if best and best[1] and best[0] <= 40:
                out[p] = out[p].lower()
    return out


def write_patch(path, chrom, ref):
    ps = sorted(ref)
    with open(path, "w") as f:
        f.write("#chrom\tstart0\tbases\n")
        i = 0
        while i < len(ps):
            j = i
            while j + 1 < len(ps) and ps[j + 1] == ps[j] + 1:
                j += 1
            f.write("%s\t%d\t%s\n" % (chrom, ps[i], "".join(ref[p] for p in ps[i:j + 1])))
            i = j + 1


def fill_from_normal_fasta(case_dir, chrom, ref, oracle):
    """normal mode prints every window: run the oracle on the N-padded reference to learn each
    record's offset, then copy bases for still-unknown positions out of the reference's expected
    FASTA (record k of ours <-> record k of theirs).  Verified afterwards through the SHA-1 ids."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from conftest import materialize_reference  # noqa: E402
    import tempfile
    exp = [l.rstrip("\n") for l in open(os.path.join(case_dir, "expected", "out.fa"))]
    exp_recs = list(zip(exp[0::2], exp[1::2]))
    for _ in range(3):
        write_patch(os.path.join(case_dir, "ref_patch.tsv"), chrom, ref)
        with tempfile.TemporaryDirectory() as td:
            fa = materialize_reference(case_dir, td)
            tsv = os.path.join(td, "o.tsv")
            res = subprocess.run([oracle, "normal", os.path.join(case_dir, "reads.bam"), "-r", fa, "-b",
                                  os.path.join(case_dir, "variants.vcf"), "-t", tsv],
                                 stdin=open(os.path.join(case_dir, "annotation.gtf")), capture_output=True)
            if res.returncode != 0:
                print("  normal fill skipped: oracle failed:", res.stderr.decode()[-300:].strip())
                return False
            rows = [l.rstrip("\n").split("\t") for l in open(tsv)]
        hdr, rows = rows[0], rows[1:]
        got = res.stdout.decode().split("\n")
        got_recs = list(zip(got[0::2], got[1::2]))
        if len(got_recs) != len(exp_recs):
            print("  normal fill: record count differs: ours %d theirs %d" % (len(got_recs), len(exp_recs)))
            return False
        changed = 0
        for (gid, gseq), (eid, eseq) in zip(got_recs, exp_recs):
            if gid == eid:
                continue
            if len(gseq) != len(eseq):
                continue
        # offsets: normal-mode TSV offset column is 0-based window start (normal_microphasing.rs:572)
        io = hdr.index("offset")
        for row, (eid, eseq), (gid, gseq) in zip(rows, exp_recs, got_recs):
            if gid == eid or len(gseq) != len(eseq):
                continue
            ws = int(row[io])
            for k, (gc, ec) in enumerate(zip(gseq, eseq)):
                if gc == "N" and ec != "N":
                    ref[ws + k] = ec
                    changed += 1
        if changed == 0:
            break
    return True


def main():
    oracle = os.path.join(HERE, "..", "..", "oracle", "_build", "mph_oracle")
    refs = {}
    for case, (sub, chrom, bam, vcf, gtf, expected) in CASES.items():
        d = os.path.join(HERE, case)
        os.makedirs(os.path.join(d, "expected"), exist_ok=True)
        shutil.copyfile(os.path.join(RES, bam), os.path.join(d, "reads.bam"))
        shutil.copyfile(os.path.join(RES, vcf), os.path.join(d, "variants.vcf"))
        shutil.copyfile(os.path.join(RES, gtf), os.path.join(d, "annotation.gtf"))
        shutil.copyfile(os.path.join(RES, chrom + ".fa.fai"), os.path.join(d, "ref.fa.fai"))
        for name, src in expected.items():
            shutil.copyfile(os.path.join(RES, src), os.path.join(d, "expected", name))
        with open(os.path.join(d, "case.txt"), "w") as f:
            f.write("subcommand\t%s\nchrom\t%s\nsource\t%s\n" % (sub, chrom, os.path.dirname(bam)))
        key = (chrom, bam)
        if key not in refs:
            ref = build_ref(os.path.join(RES, bam))
            refs[key] = ref
        ref = dict(refs[key])
        variants = read_vcf(os.path.join(RES, vcf))
        lower, upper = set(), set()
        if sub == "somatic":
            case_evidence_from_tsv(os.path.join(d, "expected", "out.tsv"), ref, variants, lower, upper)
        ref = apply_case(ref, lower, upper, variants)
        if sub == "normal" and os.path.exists(oracle):
            fill_from_normal_fasta(d, chrom, ref, oracle)
        write_patch(os.path.join(d, "ref_patch.tsv"), chrom, ref)
        print(case, "positions", len(ref), "lower", len(lower))
    # unsorted-GTF exit-status test
    d = os.path.join(HERE, "unsorted_gtf")
    os.makedirs(d, exist_ok=True)
    shutil.copyfile(os.path.join(RES, UNSORTED[0]), os.path.join(d, "unsorted.gtf"))
    shutil.copyfile(os.path.join(RES, UNSORTED[1]), os.path.join(d, "sorted.gtf"))
    shutil.copyfile(os.path.join(RES, UNSORTED[2]), os.path.join(d, "variants.vcf"))
    shutil.copyfile(os.path.join(RES, UNSORTED[3]), os.path.join(d, "reads.bam"))
    shutil.copyfile(os.path.join(RES, "chr14.fa.fai"), os.path.join(d, "ref.fa.fai"))
    write_patch(os.path.join(d, "ref_patch.tsv"), "chr14", build_ref(os.path.join(RES, UNSORTED[3])))
    with open(os.path.join(d, "case.txt"), "w") as f:
        f.write("subcommand\tsomatic\nchrom\tchr14\nsource\ttest_unsorted_gtf\n")
    # filter / build_reference fixtures are self-contained
    for name in ("test_filter", "test_filter_long", "test_filter_fs", "test_build"):
        dst = os.path.join(HERE, name)
        if os.path.exists(dst):
            shutil.rmtree(dst)
        shutil.copytree(os.path.join(RES, name), dst)


if __name__ == "__main__":
    main()
