"""N > 1 path on CPU (gloo, world_size 2): the path shards by gene range with no data-path collective,
so what has to hold is (1) shards are independent and their ordered concatenation equals the
single-process output, (2) bench.py's rank aggregation = max of times, sum of work."""
import os
import subprocess
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, build_emu

sys.path.insert(0, ROOT)


def _split_gtf(gtf_path, n_shards):
    """Contiguous gene ranges, whole genes only (the product balances by read count; any contiguous cut is valid)."""
    genes, cur = [], []
    for line in open(gtf_path):
        t = line.split("\t")
        if len(t) >= 3 and t[2] == "gene" and cur:
            genes.append(cur)
            cur = []
        cur.append(line)
    if cur:
        genes.append(cur)
    per = (len(genes) + n_shards - 1) // n_shards
    return ["".join("".join(g) for g in genes[i * per:(i + 1) * per]) for i in range(n_shards)]


def _worker(rank, world, port, indir, outdir, emu):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    # (2) aggregation: time = max over ranks, work = sum over ranks
    got = bench.aggregate_over_ranks(dist, "cpu", 1.0 + rank, 10.0 - rank, 100 * (rank + 1), 1000 * (rank + 1))
    assert got == (float(world), 10.0, 100.0 * world * (world + 1) / 2, 1000.0 * world * (world + 1) / 2), got
    assert bench.shard_seed(rank) != bench.shard_seed((rank + 1) % world)
    # (1) every rank phases its own gene range
    shard = _split_gtf(os.path.join(indir, "annotation.gtf"), world)[rank]
    d = os.path.join(outdir, "r%d" % rank)
    os.makedirs(d)
    with open(os.path.join(d, "out.fa"), "wb") as fo:
        r = subprocess.run([emu, "somatic", os.path.join(indir, "reads.bam"), "-r", os.path.join(indir, "ref.fa"), "-b",
                            os.path.join(indir, "variants.vcf"), "-t", os.path.join(d, "out.tsv"), "-n", os.path.join(d, "out.normal.fa")],
                           input=shard.encode(), stdout=fo, stderr=subprocess.PIPE)
    assert r.returncode == 0, r.stderr.decode()
    parts = [None] * world
    dist.all_gather_object(parts, {n: open(os.path.join(d, n), "rb").read() for n in ("out.fa", "out.tsv", "out.normal.fa")})
    if rank == 0:
        for name in ("out.fa", "out.normal.fa"):
            open(os.path.join(outdir, name), "wb").write(b"".join(p[name] for p in parts))
        # the TSV header goes out with the first row only
        tsv, seen = b"", False
        for p in parts:
            body = p["out.tsv"]
            if body and seen:
                body = body.split(b"\n", 1)[1]
            seen = seen or bool(body)
            tsv += body
        open(os.path.join(outdir, "out.tsv"), "wb").write(tsv)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    from microphaser_b200 import synth
    emu = build_emu()
    indir = str(tmp_path / "in")
    synth.generate(indir, synth.Params(seed=77, n_genes=6, coverage=25.0, indel_frac=0.1))
    single = tmp_path / "single"
    single.mkdir()
    with open(os.path.join(indir, "annotation.gtf")) as gin, open(single / "out.fa", "wb") as fo:
        r = subprocess.run([emu, "somatic", os.path.join(indir, "reads.bam"), "-r", os.path.join(indir, "ref.fa"), "-b",
                            os.path.join(indir, "variants.vcf"), "-t", str(single / "out.tsv"), "-n", str(single / "out.normal.fa")],
                           stdin=gin, stdout=fo, stderr=subprocess.PIPE)
    if r.returncode == 3:
        pytest.skip("input needs the serial replay path")
    assert r.returncode == 0, r.stderr.decode()
    outdir = tmp_path / "multi"
    outdir.mkdir()
    mp.spawn(_worker, args=(2, 29533, indir, str(outdir), emu), nprocs=2, join=True)
    for name in ("out.fa", "out.tsv", "out.normal.fa"):
        assert open(outdir / name, "rb").read() == open(single / name, "rb").read(), name
    assert os.path.getsize(single / "out.tsv") > 0
