"""GPU (-m gpu): typo-tolerant analysis (BASELINE.json config 4: basicTypoSet prepared inverse, typoThreshold 2.5,
typoCostWeight 6) through the C ABI — kiwi_typo_get_default + kiwi_typo_prepare + kiwi_analyze_option_t::typo_transformer —
against the golden vectors of the UNMODIFIED reference (tests/golden/typo6_*, made by make_golden.py with KB_TYPO=basic).
The device typo graph / walk (lattice.cu genTypoGraph, searchTypo) is also checked without a GPU by the host simulation
of the kernel source (tests/test_hostsim_lattice.py); this file is the proof for the real 32-lane build and for the
Viterbi / emit kernels on lattices that carry typo costs.
NOTE (round 1): written after the round's GPU budget was spent — first hardware run is the driver's round-end run."""
import os
import numpy as np
import pytest
import kiwi_b200
from tests.goldenio import read_golden, read_inputs
from tests.orc import TYPO_IMAGES

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _tok4(arr):
    return [(int(k["morph_id"]), int(k["tag"]), int(k["position"]), int(k["length"])) for k in arr]


def _close(a, b):
    return abs(a - b) <= RTOL * max(1.0, abs(b))


@pytest.fixture(scope="module")
def typo(kiwi):
    if not os.path.exists(TYPO_IMAGES["basic"]):
        pytest.skip("typo image missing: run __graft_entry__.build() where /root/reference exists")
    t = kiwi_b200.PreparedTypo(path=TYPO_IMAGES["basic"])             # kiwi_b200_typo_load
    yield t
    t.close()


@pytest.mark.parametrize("name", ["inputs_web", "inputs_written", "inputs_ref_tests", "inputs_dialect_typos"])
def test_typo_lattice_matches_reference_golden(kiwi, typo, name):
    opt = kiwi_b200.default_option(typo=typo, typo_threshold=2.5)
    texts = read_inputs(name); gold = read_golden("typo6_" + name)
    for t, g in list(zip(texts, gold))[::2]:
        assert kiwi.debug_lattice(t, opt, max_rows=1 << 17).tolist() == g["lattice"], t


@pytest.mark.parametrize("name", ["inputs_web", "inputs_written", "inputs_ref_tests", "inputs_dialect_typos"])
def test_typo_tokens_and_scores_match_reference_golden(kiwi, typo, name):
    opt = kiwi_b200.default_option(typo=typo, typo_threshold=2.5)
    texts = read_inputs(name); gold = read_golden("typo6_" + name); plain = read_golden(name)
    res = kiwi.analyze_batch(texts, opt)
    exact = corrected = 0
    for i, (t, g, p) in enumerate(zip(texts, gold, plain)):
        got = res.sentence(i)
        assert _tok4(got) == [x[:4] for x in g["tokens"]], (i, t)
        assert _close(float(res.scores[i]), g["score"]), (i, t, float(res.scores[i]), g["score"])
        for k, x in zip(got, g["tokens"]):
            assert _close(float(k["score"]), x[4]), (i, t)
        exact += int(np.float32(res.scores[i]) == np.float32(g["score"]))
        # TokenInfo::typoCost: node cost (flags bits 1-3, units of 0.5) / tokens of the node (bits 4-7, minus 1)
        costs = [np.float32(((int(k["flags"]) >> 1) & 7) * 0.5) / np.float32((int(k["flags"]) >> 4) + 1) if (int(k["flags"]) >> 1) & 7 else np.float32(0) for k in got]
        assert costs == [np.float32(x) for x in g["typo_costs"]], (i, t)
        corrected += [x[:4] for x in g["tokens"]] != [x[:4] for x in p["tokens"]]
    print("%s: %d/%d sentence scores bit-exact, %d sentences corrected by the typo lattice" % (name, exact, len(texts), corrected))
    assert exact >= 0.99 * len(texts)
    # the option is per call: the same handle without it gives the plain analysis again
    res0 = kiwi.analyze_batch(texts[:32])
    for i, p in enumerate(plain[:32]):
        assert _tok4(res0.sentence(i)) == [x[:4] for x in p["tokens"]]


def test_typo_threshold_zero_equals_plain_tokens(kiwi, typo):
    """with typoThreshold 0 every replacement node is over budget: the tokens are those of the plain lattice"""
    texts = read_inputs("inputs_web")
    res_t = kiwi.analyze_batch(texts, kiwi_b200.default_option(typo=typo, typo_threshold=0.0))
    res_p = kiwi.analyze_batch(texts)
    for i in range(len(texts)):
        assert _tok4(res_t.sentence(i)) == _tok4(res_p.sentence(i)), texts[i]


def test_typo_prepare_default_set(kiwi, typo):
    """kiwi_typo_prepare(kiwi_typo_get_default(KIWI_TYPO_BASIC_TYPO_SET)) finds typo_basic.img next to the model image opened
    by kiwi_init and gives the same analysis as the explicitly loaded image"""
    t2 = kiwi_b200.PreparedTypo(default_set=kiwi_b200.TYPO_BASIC)
    try:
        texts = read_inputs("inputs_dialect_typos")[:64]
        a = kiwi.analyze_batch(texts, kiwi_b200.default_option(typo=typo))
        b = kiwi.analyze_batch(texts, kiwi_b200.default_option(typo=t2))
        assert a.tokens.tobytes() == b.tokens.tobytes() and (a.scores == b.scores).all()
    finally:
        t2.close()


def test_kiwi_res_typo_cost_through_the_reference_calls(kiwi, typo):
    """kiwi_analyze_w with kiwi_analyze_option_t::typo_transformer, then kiwi_res_typo_cost (capi.h:927) per token: the
    reference's TokenInfo::typoCost (node cost / tokens of the node, PathEvaluator.hpp:1066-1074) from the golden dump"""
    import ctypes as C
    lib = kiwi._lib
    lib.kiwi_res_typo_cost.restype = C.c_float
    lib.kiwi_res_typo_cost.argtypes = [C.c_void_p, C.c_int, C.c_int]
    opt = kiwi_b200.default_option(typo=typo, typo_threshold=2.5)
    texts = read_inputs("inputs_dialect_typos"); gold = read_golden("typo6_inputs_dialect_typos")
    nonzero = 0
    for i in (11, 131, 162, 201, 438):
        u = np.frombuffer((texts[i] + "\0").encode("utf-16-le"), dtype="<u2").copy()
        res = lib.kiwi_analyze_w(kiwi._h, u.ctypes.data, 1, opt, None)
        assert res, lib.kiwi_error()
        try:
            n = lib.kiwi_res_word_num(res, 0)
            assert n == len(gold[i]["tokens"])
            costs = [np.float32(lib.kiwi_res_typo_cost(res, 0, k)) for k in range(n)]
            assert costs == [np.float32(x) for x in gold[i]["typo_costs"]], (i, texts[i])
            nonzero += sum(1 for c in costs if c != 0)
        finally:
            lib.kiwi_res_close(res)
    assert nonzero >= 5
