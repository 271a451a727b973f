"""microphaser_b200 — B200-native (sm_100a) implementation of microphaser's per-window phasing path.

Thin ctypes binding of the C ABI in include/microphaser_gpu.h. The library is built in-tree by
`python -m microphaser_b200.build`; importing the package never falls back to a CPU path: if the
CUDA library is missing `load()` raises, and every phase call fails without a CUDA device.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libmicrophaser_gpu.so")
CLI_PATH = os.path.join(_HERE, "_lib", "microphaser")

MPH_OK, MPH_ERR_CUDA, MPH_ERR_INPUT, MPH_ERR_PANIC, MPH_ERR_UNSUPPORTED, MPH_ERR_INTERNAL = 0, -1, -2, -3, -4, -5


class MphError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("microphaser_gpu error %d: %s" % (code, msg))
        self.code = code


class Timing(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("h2d_ms", "k1_ms", "k2_ms", "k3_ms", "k4_ms", "d2h_ms", "residue_ms", "total_ms")] + \
               [(n, C.c_uint64) for n in ("h2d_bytes", "d2h_bytes", "windows", "read_windows", "windows_enumerated", "n_interesting", "n_records")] + \
               [("kernel_launches", C.c_uint32), ("n_replay_units", C.c_uint32), ("replay_ms", C.c_double), ("k5_ms", C.c_double), ("pack_ms", C.c_double), ("ingest_ms", C.c_double), ("write_ms", C.c_double), ("kernels_ms", C.c_double)]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class SynthParams(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("n_transcripts", C.c_uint32), ("exons_per_transcript", C.c_uint32), ("exon_len_min", C.c_uint32),
                ("exon_len_max", C.c_uint32), ("read_len", C.c_uint32), ("coverage", C.c_double), ("germline_per_kb", C.c_double),
                ("somatic_per_kb", C.c_double), ("lowq_frac", C.c_double), ("indel_read_frac", C.c_double),
                ("ins_var_frac", C.c_double), ("del_var_frac", C.c_double)]


class Record(C.Structure):
    _fields_ = [("id", C.c_char_p), ("transcript", C.c_char_p), ("gene_id", C.c_char_p), ("gene_name", C.c_char_p), ("chrom", C.c_char_p),
                ("offset", C.c_uint64), ("frame", C.c_uint64), ("freq", C.c_double), ("depth", C.c_uint32), ("nvar", C.c_uint32),
                ("nsomatic", C.c_uint32), ("nvariant_sites", C.c_uint32), ("nsomvariant_sites", C.c_uint32), ("reverse", C.c_int),
                ("variant_sites", C.c_char_p), ("somatic_positions", C.c_char_p), ("somatic_aa_change", C.c_char_p),
                ("germline_positions", C.c_char_p), ("germline_aa_change", C.c_char_p), ("normal_sequence", C.c_char_p),
                ("mutant_sequence", C.c_char_p), ("fasta_mutant", C.c_char_p), ("fasta_normal", C.c_char_p)]


class BatchView(C.Structure):
    _fields_ = [("window_len", C.c_uint32)] + [(n, C.c_uint64) for n in ("n_reads", "n_vars", "n_segments", "n_chunks", "n_windows", "n_transcripts", "n_genes")] + \
               [(n, C.c_void_p) for n in ("read_start", "read_end", "read_flags")] + [("n_variant_reads", C.c_uint64)] + \
               [(n, C.c_void_p) for n in ("vr_read", "vr_vlo", "vr_seq_off", "vr_cig_off", "vr_lseq", "vr_ncig", "vr_nv")] + \
               [("bases", C.c_void_p), ("bases_bytes", C.c_uint64), ("cigars", C.c_void_p), ("n_cigar_ops", C.c_uint64), ("vars", C.c_void_p),
                ("segments", C.c_void_p), ("chunks", C.c_void_p), ("ref", C.c_void_p), ("ref_bytes", C.c_uint64), ("h2d_bytes", C.c_uint64)]


# every symbol include/microphaser_gpu.h declares
EXPORTS = ["mph_ctx_create", "mph_ctx_destroy", "mph_last_error", "mph_packer_create", "mph_packer_destroy", "mph_packer_add_gene",
           "mph_packer_finish", "mph_batch_destroy", "mph_batch_get_view", "mph_phase_batch", "mph_batch_upload", "mph_phase_resident",
           "mph_phase_collect", "mph_ctx_timing", "mph_result_destroy", "mph_result_count", "mph_result_get", "mph_result_write",
           "mph_run_somatic", "mph_run_normal", "mph_run_somatic_multi", "mph_run_normal_multi", "mph_translate", "mph_set_load", "mph_set_probe", "mph_run_filter",
           "mph_run_build_reference", "mph_synth_batch", "mph_synth_write_files"]

_lib = None


def load():
    """dlopen the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libmicrophaser_gpu.so is missing: run `python -m microphaser_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    P = C.c_void_p
    lib.mph_ctx_create.argtypes = [C.c_int, C.POINTER(P)]
    lib.mph_ctx_destroy.argtypes = [P]
    lib.mph_ctx_destroy.restype = None
    lib.mph_last_error.argtypes = [P]
    lib.mph_last_error.restype = C.c_char_p
    lib.mph_batch_destroy.argtypes = [P]
    lib.mph_batch_destroy.restype = None
    lib.mph_batch_get_view.argtypes = [P, C.POINTER(BatchView)]
    lib.mph_phase_batch.argtypes = [P, P, C.POINTER(P)]
    lib.mph_batch_upload.argtypes = [P, P]
    lib.mph_phase_resident.argtypes = [P]
    lib.mph_phase_collect.argtypes = [P, C.POINTER(P)]
    lib.mph_ctx_timing.argtypes = [P, C.POINTER(Timing)]
    lib.mph_result_destroy.argtypes = [P]
    lib.mph_result_destroy.restype = None
    lib.mph_result_count.argtypes = [P]
    lib.mph_result_count.restype = C.c_uint64
    lib.mph_result_get.argtypes = [P, C.c_uint64, C.POINTER(Record)]
    lib.mph_result_write.argtypes = [P, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    lib.mph_run_somatic.argtypes = [P] + [C.c_char_p] * 7 + [C.c_uint32, C.c_int]
    lib.mph_run_normal.argtypes = [P] + [C.c_char_p] * 6 + [C.c_uint32, C.c_int]
    lib.mph_run_somatic_multi.argtypes = [C.POINTER(P), C.c_int] + [C.c_char_p] * 7 + [C.c_uint32, C.c_int]
    lib.mph_run_normal_multi.argtypes = [C.POINTER(P), C.c_int] + [C.c_char_p] * 6 + [C.c_uint32, C.c_int]
    lib.mph_translate.argtypes = [P, P, P, P, C.c_uint64, P, P, P]
    lib.mph_set_load.argtypes = [P, P, C.c_uint32, C.c_uint64]
    lib.mph_set_probe.argtypes = [P, P, C.c_uint32, C.c_uint64, P]
    lib.mph_run_filter.argtypes = [P] + [C.c_char_p] * 7 + [C.c_uint32]
    lib.mph_run_build_reference.argtypes = [P] + [C.c_char_p] * 3 + [C.c_uint32]
    lib.mph_synth_batch.argtypes = [C.POINTER(SynthParams), C.c_uint32, C.c_int, C.POINTER(P)]
    lib.mph_synth_write_files.argtypes = [C.POINTER(SynthParams), C.c_uint32, C.c_char_p]
    _lib = lib
    return lib


def _check(rc, ctx=None):
    if rc != MPH_OK:
        raise MphError(rc, (load().mph_last_error(ctx) or b"").decode(errors="replace"))


def synth_write_files(outdir, n_transcripts=450, coverage=30.0, read_len=150, exons=8, exon_len=(90, 250), germline_per_kb=1.0,
                      somatic_per_kb=1.0, lowq_frac=0.02, indel_read_frac=0.03, seed=0x4D500002, window_len=27, ins_var_frac=0.0, del_var_frac=0.0):
    """Write the workload of Batch.synthetic(...) with the same arguments as FASTA / GTF / VCF / BAM files."""
    os.makedirs(outdir, exist_ok=True)
    sp = SynthParams(seed, n_transcripts, exons, exon_len[0], exon_len[1], read_len, coverage, germline_per_kb, somatic_per_kb,
                     lowq_frac, indel_read_frac, ins_var_frac, del_var_frac)
    _check(load().mph_synth_write_files(C.byref(sp), window_len, outdir.encode()))


class Context:
    """One CUDA device (mph_ctx)."""

    def __init__(self, device=0):
        self.lib = load()
        self.h = C.c_void_p()
        _check(self.lib.mph_ctx_create(device, C.byref(self.h)))

    def close(self):
        if self.h:
            self.lib.mph_ctx_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_somatic(self, bam, ref, variants, gtf, fasta_out, tsv, normal_out, window_len=27, warn_only=False):
        """`microphaser somatic` on files (reference src/microphasing.rs:1943 `phase`)."""
        enc = [s.encode() for s in (bam, ref, variants, gtf, fasta_out, tsv, normal_out)]
        _check(self.lib.mph_run_somatic(self.h, *enc, window_len, int(warn_only)), self.h)

    def run_normal(self, bam, ref, variants, gtf, fasta_out, tsv, window_len=27, warn_only=False):
        """`microphaser normal` on files (reference src/normal_microphasing.rs:1281 `phase`)."""
        enc = [s.encode() for s in (bam, ref, variants, gtf, fasta_out, tsv)]
        _check(self.lib.mph_run_normal(self.h, *enc, window_len, int(warn_only)), self.h)

    # ---- secondary path (reference src/peptides.rs)
    def translate(self, seqs, frames):
        """to_protein for a list of byte strings; returns (peptides, bad flags)."""
        import numpy as np
        n = len(seqs)
        off = np.zeros(n + 1, dtype=np.uint64)
        aoff = np.zeros(n + 1, dtype=np.uint64)
        for i, s in enumerate(seqs):
            off[i + 1] = off[i] + len(s)
            aoff[i + 1] = aoff[i] + (len(s) // 3 if len(s) >= 2 else 0)
        flat = np.frombuffer(b"".join(seqs) + b"\0", dtype=np.uint8).copy()
        fr = np.asarray(frames, dtype=np.int8)
        aa = np.zeros(int(aoff[n]) + 1, dtype=np.uint8)
        bad = np.zeros(max(n, 1), dtype=np.uint8)
        _check(self.lib.mph_translate(self.h, flat.ctypes.data, off.ctypes.data, fr.ctypes.data, n, aa.ctypes.data, aoff.ctypes.data, bad.ctypes.data), self.h)
        raw = aa.tobytes()
        return [raw[int(aoff[i]):int(aoff[i + 1])] for i in range(n)], bad[:n].tolist()

    def set_load(self, peptides, k):
        """peptides: numpy uint8 array of shape (n, k) or bytes of length n*k."""
        import numpy as np
        arr = np.ascontiguousarray(np.frombuffer(peptides, dtype=np.uint8) if isinstance(peptides, (bytes, bytearray)) else peptides, dtype=np.uint8)
        n = arr.size // k
        _check(self.lib.mph_set_load(self.h, arr.ctypes.data, k, n), self.h)

    def set_probe(self, queries, k):
        import numpy as np
        arr = np.ascontiguousarray(np.frombuffer(queries, dtype=np.uint8) if isinstance(queries, (bytes, bytearray)) else queries, dtype=np.uint8)
        n = arr.size // k
        hit = np.zeros(max(n, 1), dtype=np.uint8)
        _check(self.lib.mph_set_probe(self.h, arr.ctypes.data, k, n, hit.ctypes.data), self.h)
        return hit[:n]

    def run_filter(self, reference_bin, tsv_in, fasta_out, normal_out, tsv_out, removed_tsv, removed_fasta, peptide_length=9):
        enc = [s.encode() for s in (reference_bin, tsv_in, fasta_out, normal_out, tsv_out, removed_tsv, removed_fasta)]
        _check(self.lib.mph_run_filter(self.h, *enc, peptide_length), self.h)

    def run_build_reference(self, reference_fasta, binary_out, fasta_out, peptide_length=9):
        enc = [s.encode() for s in (reference_fasta, binary_out, fasta_out)]
        _check(self.lib.mph_run_build_reference(self.h, *enc, peptide_length), self.h)

    def phase_batch(self, batch):
        res = C.c_void_p()
        _check(self.lib.mph_phase_batch(self.h, batch.h, C.byref(res)), self.h)
        return Result(res)

    def upload(self, batch):
        _check(self.lib.mph_batch_upload(self.h, batch.h), self.h)

    def phase_resident(self):
        _check(self.lib.mph_phase_resident(self.h), self.h)

    def collect(self):
        res = C.c_void_p()
        _check(self.lib.mph_phase_collect(self.h, C.byref(res)), self.h)
        return Result(res)

    def timing(self):
        t = Timing()
        _check(self.lib.mph_ctx_timing(self.h, C.byref(t)))
        return t.as_dict()


def run_somatic_multi(contexts, bam, ref, variants, gtf, fasta_out, tsv, normal_out, window_len=27, warn_only=False):
    """`microphaser somatic` sharded by gene range over several devices (one Context each); ordered concatenation, no collective."""
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    enc = [s.encode() for s in (bam, ref, variants, gtf, fasta_out, tsv, normal_out)]
    _check(load().mph_run_somatic_multi(arr, len(contexts), *enc, window_len, int(warn_only)), contexts[0].h)


def run_normal_multi(contexts, bam, ref, variants, gtf, fasta_out, tsv, window_len=27, warn_only=False):
    """`microphaser normal` sharded by gene range over several devices (one Context each)."""
    arr = (C.c_void_p * len(contexts))(*[c.h for c in contexts])
    enc = [s.encode() for s in (bam, ref, variants, gtf, fasta_out, tsv)]
    _check(load().mph_run_normal_multi(arr, len(contexts), *enc, window_len, int(warn_only)), contexts[0].h)


class Batch:
    def __init__(self, handle):
        self.h = handle

    @classmethod
    def synthetic(cls, n_transcripts=450, coverage=30.0, read_len=150, exons=8, exon_len=(90, 250), germline_per_kb=1.0,
                  somatic_per_kb=1.0, lowq_frac=0.02, indel_read_frac=0.03, seed=0x4D500002, window_len=27, pin=True, ins_var_frac=0.0,
                  del_var_frac=0.0, mode="somatic"):
        sp = SynthParams(seed, n_transcripts, exons, exon_len[0], exon_len[1], read_len, coverage, germline_per_kb, somatic_per_kb,
                         lowq_frac, indel_read_frac, ins_var_frac, del_var_frac)
        h = C.c_void_p()
        old = os.environ.get("MPH_SYNTH_MODE")
        if mode == "normal":
            os.environ["MPH_SYNTH_MODE"] = "1"
        try:
            _check(load().mph_synth_batch(C.byref(sp), window_len, int(pin), C.byref(h)))
        finally:
            if mode == "normal":
                if old is None:
                    del os.environ["MPH_SYNTH_MODE"]
                else:
                    os.environ["MPH_SYNTH_MODE"] = old
        return cls(h)

    def view(self):
        v = BatchView()
        _check(load().mph_batch_get_view(self.h, C.byref(v)))
        return v

    def close(self):
        if self.h:
            load().mph_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Result:
    def __init__(self, handle):
        self.h = handle

    def __len__(self):
        return int(load().mph_result_count(self.h))

    def record(self, i):
        r = Record()
        _check(load().mph_result_get(self.h, i, C.byref(r)))
        return {n: (getattr(r, n).decode() if isinstance(getattr(r, n), bytes) else getattr(r, n)) for n, _ in Record._fields_}

    def write(self, fasta_path, tsv_path, normal_path):
        fds = [os.open(p, os.O_WRONLY | os.O_CREAT | os.O_TRUNC, 0o644) for p in (fasta_path, tsv_path, normal_path)]
        try:
            hw = C.c_int(0)
            _check(load().mph_result_write(self.h, fds[0], fds[1], fds[2], C.byref(hw)))
        finally:
            for fd in fds:
                os.close(fd)

    def close(self):
        if self.h:
            load().mph_result_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
