"""bench.py's workload plans (BASELINE.json configs[1..4]): every read of the job belongs to exactly one rank's batches, in
order, whatever the world size; the reference arm takes the native arm's first batch."""
import importlib.util
import os

import pytest

from conftest import ROOT

spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
bench = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bench)


@pytest.mark.parametrize("name", sorted(bench.WORKLOADS))
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_plans_partition_the_reads(name, world, monkeypatch):
    for k in ("MLG_BENCH_G", "MLG_BENCH_READS", "MLG_BENCH_TOTAL_READS", "MLG_BENCH_BATCH_READS"):
        monkeypatch.delenv(k, raising=False)
    nxt, total_seen = 0, None
    for rank in range(world):
        G, paired, batches, total, desc = bench.workload_plan(name, world, rank)
        total_seen = total if total_seen is None else total_seen
        assert total == total_seen and G == bench.WORKLOADS[name]["G"]
        for r0, n in batches:
            assert r0 == nxt and 0 < n <= bench.BATCH_READS
            nxt += n
        assert "configs[" in desc
    assert nxt == total_seen
    w = bench.WORKLOADS[name]
    assert total_seen == (w["total"] or w["per_gpu"] * world)


def test_env_overrides(monkeypatch):
    monkeypatch.setenv("MLG_BENCH_TOTAL_READS", "1e6")
    monkeypatch.setenv("MLG_BENCH_BATCH_READS", "3e5")
    monkeypatch.setenv("MLG_BENCH_G", "1234")
    G, paired, batches, total, desc = bench.workload_plan("stream", 3, 1)
    assert G == 1234 and total == 1_000_000
    assert batches[0][0] == 333_334 and sum(n for _, n in batches) == 333_333 and max(n for _, n in batches) == 300_000
