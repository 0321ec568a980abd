"""Planner of per-chromosome shards from .bai statistics (breakdancer_b200/shard.py)."""
import numpy as np

from breakdancer_b200 import api, shard


def test_plan_keeps_contigs_whose_index_has_no_record_counts(monkeypatch):
    """An index without samtools' pseudo-bin gives no record counts; a small contig whose chunks lie inside one BGZF member spans
    a single byte. It must still be assigned to a rank (weight >= 1), not dropped like a sequence without chunks."""
    rec = np.array([-1, -1, -1, -1], np.int64)
    byt = np.array([6_400_000, 1, 0, 640], np.int64)       # big contig, tiny contig inside one member, no chunks at all, small contig
    monkeypatch.setattr(api, "bai_reference_stats", lambda path: (rec, byt))
    weights, bins = shard.plan_from_index(["a.bam"], 2)
    assert weights.tolist() == [100000, 1, 0, 10]
    assigned = sorted(t for b in bins for t in b)
    assert assigned == [0, 1, 3]                           # only the sequence without any chunk is left out


def test_plan_prefers_record_counts_where_the_index_has_them(monkeypatch):
    rec = np.array([500, -1, 0], np.int64)
    byt = np.array([64, 6400, 64], np.int64)
    monkeypatch.setattr(api, "bai_reference_stats", lambda path: (rec, byt))
    weights, bins = shard.plan_from_index(["a.bam", "b.bam"], 3)
    assert weights.tolist() == [1000, 200, 0]
    assert sorted(t for b in bins for t in b) == [0, 1]
