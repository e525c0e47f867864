"""GPU (-m gpu): the CoNg path (BASELINE.json config 3: quantized CoNg model, int8 scorer on tensor-core tiles), called
through the C ABI, against the golden vectors of the unmodified reference (ModelType::cong, tests/golden/cong_*) and the
oracle restatement.  Stage level first (int8 dot products, float epilogues, context-trie steps, the mma tile), then
whole analyses.  Integer / index fields bit-exact; scores within 1e-4 relative (and >= 99 % bit-exact)."""
import numpy as np
import pytest
from tests.goldenio import read_golden, read_inputs, read_cong_qgemm

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _tok4(arr):
    return [(int(k["morph_id"]), int(k["tag"]), int(k["position"]), int(k["length"])) for k in arr]


def _close(a, b):
    return abs(a - b) <= RTOL * max(1.0, abs(b))


def test_cong_scorer_pieces_match_reference_vectors(kiwi_cong, oracle_cong):
    """dp4a dot + the three epilogues + context-trie transitions against the reference's own outputs (cong_probe dump)."""
    assert kiwi_cong.model_type() == 4
    ctx = []; wid = []; node = []; want_ll = []; want_next = []
    for rec in read_cong_qgemm():
        if rec[0] == "P":
            _, nd, cx, w, ll, nd2, cx2 = rec
            ctx.append(cx); wid.append(w); node.append(nd); want_ll.append(np.float32(ll)); want_next.append((nd2, cx2))
    out = kiwi_cong.debug_cong(ctx, wid, node)
    assert [np.float32(x) for x in out["eps"][:, 0]] == want_ll                       # scalar epilogue, bit exact
    assert list(zip(out["node"].tolist(), out["ctx"].tolist())) == want_next          # context trie
    for i in range(0, len(ctx), 7):
        acc, eps = oracle_cong.cong_pair(ctx[i], wid[i])
        assert int(out["dot"][i]) == acc
        assert [np.float32(x) for x in out["eps"][i]] == eps
    # gather-GEMM known answers: every shape of the dump through the dp4a path + the shape's epilogue
    for rec in read_cong_qgemm():
        if rec[0] != "Q":
            continue
        _, m, n, a, b, c = rec
        ep = oracle_cong.cong_epilogue(m, n)
        cc = [a[i] for i in range(m) for j in range(n)]; ww = [b[j] for i in range(m) for j in range(n)]
        o = kiwi_cong.debug_cong(cc, ww, [0] * len(cc))
        assert [np.float32(x) for x in o["eps"][:, ep]] == [np.float32(x) for x in c], (m, n)


def test_cong_tensor_core_tile_is_exact(kiwi_cong, oracle_cong):
    """mma.sync m16n8k32 (u8 x s8 -> s32) tile over 64 contexts x 32 outputs == exact integer dot products."""
    rs = np.random.RandomState(11)
    for n in (64, 40, 17, 5):
        ctx = rs.randint(0, 8192, size=n); wid = rs.randint(0, 60000, size=n)
        o = kiwi_cong.debug_cong(ctx, wid, np.zeros(n, np.int32))
        nu, nw = min(n, 64), min(n, 32)
        want = np.array([[oracle_cong.cong_pair(int(ctx[r]), int(wid[c]))[0] for c in range(nw)] for r in range(nu)], np.int32)
        assert (o["tile"] == want).all(), n
        assert (o["dot"] == np.array([oracle_cong.cong_pair(int(ctx[i]), int(wid[i]))[0] for i in range(n)], np.int32)).all()


@pytest.mark.parametrize("name", ["inputs_ref_tests", "inputs_web", "inputs_written", "inputs_dialect_typos"])
def test_cong_tokens_and_scores_match_reference_golden(kiwi_cong, name):
    texts = read_inputs(name); gold = read_golden("cong_" + name)
    res = kiwi_cong.analyze_batch(texts)
    exact = 0
    for i, (t, g) in enumerate(zip(texts, gold)):
        got = res.sentence(i)
        assert _tok4(got) == [x[:4] for x in g["tokens"]], (i, t)
        assert _close(float(res.scores[i]), g["score"]), (i, t, float(res.scores[i]), g["score"])
        for k, x in zip(got, g["tokens"]):
            assert _close(float(k["score"]), x[4]), (i, t)
        exact += int(np.float32(res.scores[i]) == np.float32(g["score"]))
    print("cong %s: %d/%d sentence scores bit-exact" % (name, exact, len(texts)))
    assert exact >= 0.99 * len(texts)


def test_cong_batch_8192_properties_and_sampled_oracle(kiwi_cong, oracle_cong):
    from kiwi_b200.synth import synth_batch, u16len
    texts = synth_batch(8192)
    r1 = kiwi_cong.analyze_batch(texts)
    r2 = kiwi_cong.analyze_batch(texts)
    assert (r1.token_offsets == r2.token_offsets).all() and r1.tokens.tobytes() == r2.tokens.tobytes() and (r1.scores == r2.scores).all()
    perm = np.random.RandomState(7).permutation(len(texts))[:512]
    r3 = kiwi_cong.analyze_batch([texts[i] for i in perm])
    for k, i in enumerate(perm):
        assert r3.sentence(k).tobytes() == r1.sentence(int(i)).tobytes() and r3.scores[k] == r1.scores[i]
    for i, t in enumerate(texts):
        s = r1.sentence(i)
        assert len(s) > 0
        pos = s["position"].astype(np.int64); end = pos + s["length"]
        assert (np.diff(pos) >= 0).all()
        assert int(end.max()) == u16len(t.rstrip(" "))
    for i in range(0, len(texts), 32):
        otoks, oscore = oracle_cong.analyze(texts[i])
        assert _tok4(r1.sentence(i)) == [x[:4] for x in otoks], texts[i]
        assert _close(float(r1.scores[i]), oscore)


def test_knlm_and_cong_handles_coexist(kiwi, kiwi_cong, oracle, oracle_cong):
    """Two resident models in one process: the constant-memory model view follows the handle that launches."""
    t = "키위는 형태소 분석기입니다. 두 모델을 번갈아 씁니다."
    for _ in range(2):
        a = kiwi.analyze_batch([t]); b = kiwi_cong.analyze_batch([t])
        assert _tok4(a.sentence(0)) == [x[:4] for x in oracle.analyze(t)[0]]
        assert _tok4(b.sentence(0)) == [x[:4] for x in oracle_cong.analyze(t)[0]]
