"""CPU: host simulation of the WHOLE device pipeline with 32-lane warps.  tests/hostsim/libpipeline_sim32.so is lattice.cu +
viterbi.cu (Knlm and CoNg builds) + emit.cu + model.cu — the device SOURCES — compiled as C++ (tests/hostsim/shim32/cuda_runtime.h): each
kernel is run by 32 OS threads, one per lane, under a cooperative scheduler that switches lanes at every warp collective
(ballot / shuffle / match_any / reduce / __syncwarp), so lane arithmetic, ballots, shared memory and lane-0 stores behave as
in a warp.  Tokens and scores are compared with the golden vectors of the UNMODIFIED reference, plain and with the typo
lattice (BASELINE config 4) — the only check of the typo path's Viterbi / emit side that runs without a GPU.

What it is not: a proof for the hardware build (timing, real memory-model behaviour, multi-warp blocks — the lockstep barrier is
compiled out —, tensor-core PTX of the CoNg build).  It is slow (one condition-variable hand-over per lane and collective),
hence a SAMPLE of short sentences.
Because a lane runs alone until its next collective, the simulator is harsher than a converged warp on shared state that is
read by all lanes and rewritten later without a barrier in between.  It found two such write-after-read hazards (harmless
while the warp stays converged, races by the CUDA memory model), both fixed with one __syncwarp():
  * lattice.cu appendNewNode: lane 0 changed endPosMap / out[] while slower lanes were still in the scans that decided the append;
  * viterbi.cu fixupGroup: a lane that left the per-candidate loop early returned to the caller, which re-initialises
    sm->cand / sm->candNew for the next group, while other lanes were still reading them (inputs_dialect_typos[98], [469])."""
import ctypes as C, os
import numpy as np
import pytest
from tests.goldenio import read_golden, read_inputs
from tests.orc import IMAGE, TYPO_IMAGES, CONG_IMAGE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "hostsim", "libpipeline_sim32.so")
MATCH_ALL_WITH_NORMALIZING = (1 << 0) | (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 5) | (1 << 23) | (1 << 16)
# (input file, line): short sentences; the typo list holds sentences whose analysis the typo lattice really changes
PLAIN = [("inputs_ref_tests", 5), ("inputs_ref_tests", 162), ("inputs_dialect_typos", 573), ("inputs_written", 0),
         ("inputs_dialect_typos", 98), ("inputs_dialect_typos", 469)]      # the last two: medium-mode containers (> 128 incoming paths)
TYPO = [("inputs_dialect_typos", 201), ("inputs_dialect_typos", 438), ("inputs_dialect_typos", 573), ("inputs_dialect_typos", 75),
        ("inputs_ref_tests", 259), ("inputs_ref_tests", 60), ("inputs_ref_tests", 361), ("inputs_ref_tests", 367), ("inputs_ref_tests", 5),
        ("inputs_ref_tests", 162), ("inputs_dialect_typos", 131), ("inputs_dialect_typos", 162)]      # 131 / 162: typo cost split over 4 / 2 tokens


CONG = [("inputs_written", 0), ("inputs_dialect_typos", 98), ("inputs_dialect_typos", 201), ("inputs_ref_tests", 5)]


def _open(image):
    if not os.path.exists(LIB) or not os.path.exists(image):
        pytest.skip("tests/hostsim/libpipeline_sim32.so or the model image missing: run __graft_entry__.build()")
    lib = C.CDLL(LIB)
    lib.hs32_open.restype = C.c_void_p; lib.hs32_open.argtypes = [C.c_char_p]
    lib.hs32_close.argtypes = [C.c_void_p]
    lib.hs32_set_typo.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
    lib.hs32_analyze.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32] + [C.c_void_p] * 5 + [C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p]
    h = lib.hs32_open(os.fsencode(image))
    assert h
    cap = 4096
    morph = np.zeros(cap, np.uint32); tag = np.zeros(cap, np.uint8); pos = np.zeros(cap, np.uint32); ln = np.zeros(cap, np.uint16); sc = np.zeros(cap, np.float32); fl = np.zeros(cap, np.uint8)

    def analyze(text):
        u = np.ascontiguousarray(np.frombuffer(text.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
        s = C.c_float(0); nn = C.c_int(0)
        n = lib.hs32_analyze(h, u.ctypes.data, len(u), MATCH_ALL_WITH_NORMALIZING, morph.ctypes.data, tag.ctypes.data, pos.ctypes.data, ln.ctypes.data,
                             sc.ctypes.data, cap, C.byref(s), C.byref(nn), fl.ctypes.data)
        assert n >= 0, (n, text)
        # TokenInfo::typoCost from the device row's flags: node cost (bits 1-3, units of 0.5) / tokens of the node (bits 4-7, minus 1)
        analyze.typo_costs = [float(np.float32(((int(fl[k]) >> 1) & 7) * 0.5) / np.float32((int(fl[k]) >> 4) + 1)) if (int(fl[k]) >> 1) & 7 else 0.0 for k in range(n)]
        return [(int(morph[k]), int(tag[k]), int(pos[k]), int(ln[k]), np.float32(sc[k])) for k in range(n)], np.float32(s.value)

    def set_typo(path):
        assert lib.hs32_set_typo(h, os.fsencode(path) if path else None, 2.5) == 0

    return analyze, set_typo, lambda: lib.hs32_close(h)


@pytest.fixture(scope="module")
def sim():
    analyze, set_typo, close = _open(IMAGE)
    yield analyze, set_typo
    close()


def _check(analyze, cases, prefix):
    for name, i in cases:
        t = read_inputs(name)[i]; g = read_golden(prefix + name)[i]
        toks, score = analyze(t)
        assert [x[:4] for x in toks] == [x[:4] for x in g["tokens"]], (name, i, t)
        assert [x[4] for x in toks] == [np.float32(x[4]) for x in g["tokens"]], (name, i, t)      # bit exact
        assert score == np.float32(g["score"]), (name, i, t)
        if "typo_costs" in g:
            assert [np.float32(x) for x in analyze.typo_costs] == [np.float32(x) for x in g["typo_costs"]], (name, i, t)


def test_simulated_pipeline_matches_reference(sim):
    analyze, set_typo = sim
    set_typo(None)
    _check(analyze, PLAIN, "")


def test_simulated_pipeline_matches_reference_with_typo_lattice(sim):
    if not os.path.exists(TYPO_IMAGES["basic"]):
        pytest.skip("typo image missing")
    analyze, set_typo = sim
    set_typo(TYPO_IMAGES["basic"])
    try:
        _check(analyze, TYPO, "typo6_")
        # every sentence of the list is one the typo lattice corrects or re-segments
        changed = sum([x[:4] for x in read_golden("typo6_" + n)[i]["tokens"]] != [x[:4] for x in read_golden(n)[i]["tokens"]] for n, i in TYPO)
        assert changed >= len(TYPO) - 2
    finally:
        set_typo(None)


def test_simulated_cong_pipeline_matches_reference():
    """the CoNg build of viterbi.cu (viterbi_cong_kernel): context trie, dp4a rows and the tensor-core tiles, whose
    mma.sync.m16n8k32 PTX is replaced by its fragment semantics in C++ (shim32 mma_m16n8k32_u8s8) — bit-exact scores against
    the reference's CoNg vectors"""
    analyze, set_typo, close = _open(CONG_IMAGE)
    try:
        set_typo(None)
        _check(analyze, CONG, "cong_")
    finally:
        close()


@pytest.mark.parametrize("mode", ["plain", "typo", "cong"])
def test_hostsim_full_golden_sweep(mode):
    """Round 2: with fibers instead of OS threads the simulator is fast enough for EVERY golden sentence (1273 per mode):
    tokens identical and scores bit-exact against the unmodified reference's vectors, for the Knlm, typo-lattice and CoNg builds."""
    import subprocess, sys
    if not os.path.exists(LIB) or not os.path.exists(IMAGE):
        pytest.skip("tests/hostsim/libpipeline_sim32.so or the model image missing: run __graft_entry__.build()")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "hostsim_sweep.py"), mode], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "1273 sentences, 0 mismatches" in r.stdout, r.stdout


def test_hostsim_sbg_golden_subset():
    """The SkipBigram build of viterbi.cu (viterbi_sbg_kernel: Knlm + 8-token history, sbg_math.h, the item-by-item path container
    exactInsertRound) through the simulator on every twelfth sentence of inputs_written / inputs_web (the shorter ones): tokens identical and scores bit-exact
    against the unmodified reference's ModelType::sbg vectors.  (The whole files - 33 + 158 sentences, 0 mismatches - take ten minutes:
    `python scripts/hostsim_sweep.py sbg --files inputs_written,inputs_web`.)"""
    import subprocess, sys
    from tests.orc import SBG_IMAGE
    if not os.path.exists(LIB) or not os.path.exists(SBG_IMAGE):
        pytest.skip("tests/hostsim/libpipeline_sim32.so or the SkipBigram model image missing: run __graft_entry__.build()")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "hostsim_sweep.py"), "sbg", "--files", "inputs_written,inputs_web", "--stride", "12", "--maxlen", "70"], capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0 and "0 mismatches" in r.stdout and "sbg: 0 sentences" not in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_knlm_kernel_follows_the_reference_container_behaviour():
    """once a bucket of a path container holds 64 states the reference's insertOptimized stops finding its
    first 64 entries and appends states again.  On the bench sentences where that changes the per-node path counts, the simulated Knlm
    kernel (parallel insert + the item-by-item redo of such groups) must give the counts of the oracle's default (reference-behaviour)
    mode, not those of its ORC_BUCKET_SEARCH_ALL mode (scripts/container_check.py)."""
    import subprocess, sys
    if not os.path.exists(LIB) or not os.path.exists(IMAGE):
        pytest.skip("tests/hostsim/libpipeline_sim32.so or the model image missing: run __graft_entry__.build()")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "container_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "6/6 sentences follow the reference's behaviour" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_forced_group_redo_keeps_every_golden_sentence():
    """tests/hostsim/libpipeline_sim32_redo.so = the Knlm kernel built with KB_EXACT_FROM=4: every group whose containers hold 4 states is
    evaluated a second time item by item (redoGroupExact / exactInsertRound - below 64 states per bucket that must reproduce the parallel
    insert exactly).  All 1 273 golden sentences: tokens identical, scores bit-exact."""
    import subprocess, sys
    redo = os.path.join(ROOT, "tests", "hostsim", "libpipeline_sim32_redo.so")
    if not os.path.exists(redo) or not os.path.exists(IMAGE):
        pytest.skip("tests/hostsim/libpipeline_sim32_redo.so or the model image missing: run __graft_entry__.build()")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "hostsim_sweep.py"), "plain"], capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, HS32_LIB="libpipeline_sim32_redo.so"))
    assert r.returncode == 0 and "1273 sentences, 0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_hostsim_open_ending_matches_reference():
    """AnalyzeOption::openEnding through the simulated kernels (bit 31 of the option word): all 33 + 158 sentences of inputs_written /
    inputs_web against the unmodified reference's open-ending vectors, tokens identical and scores bit-exact."""
    import subprocess, sys
    if not os.path.exists(LIB) or not os.path.exists(IMAGE):
        pytest.skip("tests/hostsim/libpipeline_sim32.so or the model image missing: run __graft_entry__.build()")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "hostsim_sweep.py"), "open", "--files", "inputs_written,inputs_web"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "191 sentences, 0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_hostsim_blocklist_matches_reference():
    """AnalyzeOption::blocklist through the simulated kernels: the candidate table patched by Model::blockedCands (what the engine uploads for
    such a call; the kernels themselves are unchanged) - all 191 sentences of inputs_written / inputs_web against the reference's vectors."""
    import subprocess, sys
    if not os.path.exists(LIB) or not os.path.exists(IMAGE):
        pytest.skip("tests/hostsim/libpipeline_sim32.so or the model image missing: run __graft_entry__.build()")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "hostsim_sweep.py"), "block", "--files", "inputs_written,inputs_web"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "191 sentences, 0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_hostsim_fuzz_strings_match_oracle():
    """1 500 random strings that look nothing like the bench text (random syllables, jamo, emoji and surrogate pairs, URLs, odd whitespace,
    repeated characters, empty strings: scripts/hostsim_sweep.py _fuzz_texts) through the simulated Knlm kernels against the oracle - which
    itself equals the unmodified reference on 1 065 of these strings (checked when the generator was written)."""
    import subprocess, sys
    if not os.path.exists(LIB) or not os.path.exists(IMAGE):
        pytest.skip("tests/hostsim/libpipeline_sim32.so or the model image missing: run __graft_entry__.build()")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "hostsim_sweep.py"), "plain", "--files", "fuzz:1500"], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "1500 sentences, 0 mismatches" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]
