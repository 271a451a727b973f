"""Full-size checks of the CUDA path through size-independent properties (the oracle would need
minutes to hours at these sizes): determinism, prefix consistency (records of the first N genes do
not depend on what follows — shards are independent), closed-form window counts."""
import pytest

pytestmark = pytest.mark.gpu


def records(res, limit=None):
    n = len(res) if limit is None else min(limit, len(res))
    return [tuple(sorted(res.record(i).items())) for i in range(n)]


def test_repeatable_and_prefix_consistent(product):
    import microphaser_b200 as m
    ctx = m.Context(0)
    full = m.Batch.synthetic(n_transcripts=450, coverage=30.0, seed=0x4D500002)  # BASELINE.json config 2 shape
    half = m.Batch.synthetic(n_transcripts=225, coverage=30.0, seed=0x4D500002)
    r1 = ctx.phase_batch(full)
    t1 = ctx.timing()
    r2 = ctx.phase_batch(full)
    t2 = ctx.timing()
    rh = ctx.phase_batch(half)
    th = ctx.timing()
    assert len(r1) == len(r2) and len(r1) > 100
    a, b, h = records(r1), records(r2), records(rh)
    assert a == b, "two runs over the same batch differ"
    assert a[:len(h)] == h, "records of the first 225 genes depend on the genes that follow"
    assert t1["windows"] == t2["windows"] and t1["read_windows"] == t2["read_windows"]
    assert 0 < th["windows"] < t1["windows"]
    # every main-ORF window up to and including the stop codon: (cds_nt - 27)/3 + 1 in-frame windows per exon
    v = full.view()
    assert t1["windows"] <= v.n_windows and t1["windows"] > 0.9 * v.n_windows
    ctx.close()


def test_whole_exome_shape_resident_run(product):
    """A shard of BASELINE.json config 3 (100x, 150 bp) — resident kernels are repeatable and agree with the one-shot call."""
    import microphaser_b200 as m
    ctx = m.Context(0)
    batch = m.Batch.synthetic(n_transcripts=2000, coverage=100.0, seed=0x4D500003)
    one = ctx.phase_batch(batch)
    t_one = ctx.timing()
    ctx.upload(batch)
    ctx.phase_resident()
    ctx.phase_resident()
    two = ctx.collect()
    t_two = ctx.timing()
    assert len(one) == len(two) > 1000
    assert records(one, 500) == records(two, 500)
    assert t_one["windows"] == t_two["windows"] and t_one["read_windows"] == t_two["read_windows"]
    assert t_one["read_windows"] > 50 * t_one["windows"]  # ~100x * (150-27)/150 reads per window
    ctx.close()


@pytest.mark.parametrize("mode", ["somatic", "normal"])
def test_pipelined_stages_match_single_stage(product, mode, monkeypatch):
    """mph_phase_batch cuts a batch at gene boundaries into stages (copy / kernels / residue overlap); the records must not
    depend on where the cuts fall. MPH_STAGES forces the stage count (large batches pick it from the read count)."""
    import microphaser_b200 as m
    if mode == "normal":
        monkeypatch.setenv("MPH_SYNTH_MODE", "1")
    ctx = m.Context(0)
    batch = m.Batch.synthetic(n_transcripts=600 if mode == "normal" else 2000, coverage=60.0, seed=0x4D500004)
    monkeypatch.setenv("MPH_STAGES", "1")
    one = ctx.phase_batch(batch)
    t_one = ctx.timing()
    monkeypatch.setenv("MPH_STAGES", "7")
    many = ctx.phase_batch(batch)
    t_many = ctx.timing()
    assert len(one) == len(many) > 1000
    assert records(one) == records(many)
    assert t_one["windows"] == t_many["windows"] and t_one["read_windows"] == t_many["read_windows"]
    # every stage runs the chain (the single-stage call may have run it twice: the first call on a context grows the arenas)
    assert t_many["kernel_launches"] >= 3 * t_one["kernel_launches"]
    # pageable host buffers take the other copy schedule (stage s + 1 is queued after the kernels of stage s)
    pageable = m.Batch.synthetic(n_transcripts=600 if mode == "normal" else 2000, coverage=60.0, seed=0x4D500004, pin=False)
    again = ctx.phase_batch(pageable)
    assert records(again) == records(one)
    if mode == "somatic":
        assert t_one["n_replay_units"] > 0, "the synthetic workload is expected to contain irregular transcripts"
    ctx.close()
