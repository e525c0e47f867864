"""CPU: host-side logic — batch marshalling, synthetic generator determinism, byte model, tag names, sharding."""
import numpy as np
import kiwi_b200
from kiwi_b200 import bytemodel, shard
from kiwi_b200.synth import synth_batch, u16len, SEED


def test_encode_batch_offsets_and_surrogates():
    texts = ["", "가", "a😀b", "\ud83d", "안녕 하세요"]
    blob, off = kiwi_b200.encode_batch(texts)
    assert off.dtype == np.uint32 and off.tolist() == [0, 0, 1, 5, 6, 12]
    assert blob.dtype == np.dtype("<u2") and len(blob) == 12
    assert blob[2] == 0xD83D and blob[3] == 0xDE00 and blob[5] == 0xD83D
    b2, o2 = kiwi_b200.encode_batch([])
    assert len(b2) == 0 and o2.tolist() == [0]


def test_synth_is_deterministic_and_web_shaped():
    a = synth_batch(512, SEED); b = synth_batch(512, SEED); c = synth_batch(512, SEED + 1)
    assert a == b and a != c
    lens = [u16len(s) for s in a]
    assert 50 < sum(lens) / len(lens) < 100 and max(lens) <= 184


def test_tag_names_follow_reference_table():
    assert kiwi_b200.tag_to_string(1) == "NNG" and kiwi_b200.tag_to_string(53) == "Z_CODA" and kiwi_b200.tag_to_string(0x84) == "VV-I"
    assert kiwi_b200.tag_to_string(60) == "P"


def test_bytemodel_is_linear_in_counters():
    c = dict(sentences=2, rawUnits=10, normUnits=14, trieVisits=5, trieProbes=7, trieHits=3, candForms=4, nodesBuilt=6, nodesFinal=5,
             candEntries=8, candEvals=9, lmSteps=10, lmHops=20, lmProbes=30, pairs=40, pathsWritten=11, pathsKept=6, tokens=4)
    c2 = {k: 2 * v for k, v in c.items()}
    assert bytemodel.total_bytes(c2) == 2 * bytemodel.total_bytes(c)
    assert bytemodel.viterbi_bytes(c) == 32 * 5 + 4 * 8 + 32 * 9 + 20 * 24 + 4 * 30 + 40 * 11 + 40 * 40 + 15 * 4


def test_round_robin_shards_restore_order():
    n, w = 37, 4
    parts = [shard.shard_indices(n, r, w) for r in range(w)]
    assert sorted(i for p in parts for i in p) == list(range(n))
    assert all(p == list(range(r, n, w)) for r, p in enumerate(parts))
    results = [[("res", i) for i in p] for p in parts]
    assert shard.merge_round_robin(results, n) == [("res", i) for i in range(n)]


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference (the reference's own CPU path, oracle/_ref/ref_bench) must print ONE JSON line with the
    contract's keys for both model types; runs here on a tiny sample."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.exists(os.path.join(root, "oracle", "_ref", "ref_bench")):
        import pytest
        pytest.skip("oracle/_ref/ref_bench missing: run __graft_entry__.build() where /root/reference exists")
    for model in ("knlm", "cong"):
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--model", model, "--steps", "1", "--warmup", "1",
                              "--cpu-sample", "128"], capture_output=True, text=True, timeout=600).stdout
        lines = [l for l in out.splitlines() if l.startswith("{")]
        assert len(lines) == 1
        d = json.loads(lines[0])
        assert d["impl"] == "reference" and d["unit"] == "sentences/s" and d["higher_is_better"] is True and d["value"] > 0
        assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
        assert d["e2e"] == {"value": d["value"], "unit": "sentences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        assert "workload" in d["config"] and model in d["config"]["workload"].lower()
