"""GPU (-m gpu): the two analysis options added at the end of round 2 - AnalyzeOption::openEnding and AnalyzeOption::blocklist - through the
C ABI against vectors of the unmodified reference (tests/golden/open_*, block_*).  Both are bit-exact in the host simulation of the kernel
sources (tests/test_hostsim_pipeline.py); this file holds their first hardware runs, which is why it sorts behind the other GPU tests."""
import os
import numpy as np
import pytest
import kiwi_b200

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_open_ending_matches_reference_golden(kiwi):
    """AnalyzeOption::openEnding through the C ABI (kiwi_analyze_option_t::open_ending, capi.h:662-670): tokens and float scores of
    inputs_written against the unmodified reference's open-ending vectors (tests/golden/open_inputs_written)."""
    from tests.goldenio import read_golden, read_inputs
    texts = read_inputs("inputs_written"); gold = read_golden("open_inputs_written")
    res = kiwi.analyze_batch(texts, kiwi_b200.default_option(open_ending=True))
    plain = kiwi.analyze_batch(texts)
    differs = 0
    for i, (t, g) in enumerate(zip(texts, gold)):
        got = res.sentence(i)
        assert [(int(k["morph_id"]), int(k["tag"]), int(k["position"]), int(k["length"])) for k in got] == [x[:4] for x in g["tokens"]], (i, t)
        assert np.float32(res.scores[i]) == np.float32(g["score"]), (i, t, float(res.scores[i]), g["score"])
        differs += int(res.scores[i] != plain.scores[i])
    assert differs > len(texts) // 2      # (the end-of-sentence step is really gone)


def test_blocklist_matches_reference_golden(kiwi):
    """kiwi_new_morphset / kiwi_morphset_add / kiwi_analyze_option_t::blocklist (capi.h:660, 1243-1263): the (form, tag) list of
    tests/golden/MANIFEST.json resolves to the reference's morpheme count, and inputs_written analysed with that blocklist equals the
    unmodified reference's vectors (tests/golden/block_inputs_written); a later call without the blocklist is unaffected."""
    import json
    from tests.goldenio import read_golden, read_inputs
    man = json.load(open(os.path.join(ROOT, "tests", "golden", "MANIFEST.json")))["blocklist"]
    ms = kiwi_b200.MorphSet(kiwi)
    added = sum(ms.add(*item.rsplit("/", 1)) for item in man["spec"].split(";"))
    assert added == len(man["morpheme_ids"])
    texts = read_inputs("inputs_written"); gold = read_golden("block_inputs_written")
    before = kiwi.analyze_batch(texts)
    res = kiwi.analyze_batch(texts, kiwi_b200.default_option(blocklist=ms))
    for i, (t, g) in enumerate(zip(texts, gold)):
        got = res.sentence(i)
        assert [(int(k["morph_id"]), int(k["tag"]), int(k["position"]), int(k["length"])) for k in got] == [x[:4] for x in g["tokens"]], (i, t)
        assert np.float32(res.scores[i]) == np.float32(g["score"]), (i, t, float(res.scores[i]), g["score"])
    after = kiwi.analyze_batch(texts)
    assert after.tokens.tobytes() == before.tokens.tobytes() and (after.scores == before.scores).all()
    assert int((res.scores != before.scores).sum()) > len(texts) // 2
    ms.close()
