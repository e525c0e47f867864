"""CPU: the product's host-side result assembly (kiwi_b200/csrc/assemble.h: paired brackets / bullets, sentence, line and
sub-sentence numbers, sentence-relative word index) run on the reference's own golden token lists must reproduce the
reference's TokenInfo fields (fillPairedTokenInfo + fillSentLineInfo, src/Kiwi.cpp:98-143, 322-415).  No GPU needed:
the inputs are the golden tokens, which the CUDA path reproduces bit-exactly (tests/test_gpu_parity.py)."""
import ctypes as C, os
import numpy as np
import pytest
from tests.goldenio import read_golden, read_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "hostsim", "libassemble_check.so")
SPACES = set(" \f\n\r\t\v\xa0              ⠀　")   # include/kiwi/Utils.h:294-325


def word_positions(units):      # getWordPositions, src/Kiwi.cpp:464-485
    out = []; position = 0; cont = False
    for u in units:
        out.append(position)
        if chr(u) in SPACES:
            if not cont: position += 1
            cont = True
        else:
            cont = False
    out.append(position)
    return out


@pytest.mark.parametrize("name", ["inputs_web", "inputs_written", "inputs_dialect_typos", "inputs_ref_tests", "cong_inputs_web"])
def test_assembly_matches_reference_token_info(name):
    if not os.path.exists(LIB):
        pytest.skip("tests/hostsim/libassemble_check.so missing: run __graft_entry__.build()")
    lib = C.CDLL(LIB)
    lib.kb_asm_run.argtypes = [C.c_int] + [C.c_void_p] * 8 + [C.c_uint32, C.c_void_p]
    texts = read_inputs(name.replace("cong_", "")); gold = read_golden(name)
    checked = 0
    for t, g in zip(texts, gold):
        toks = g["tokens"]; forms = g.get("forms")
        if not toks or forms is None:
            continue
        units = np.frombuffer(t.encode("utf-16-le", "surrogatepass"), dtype="<u2")
        if any(0xD800 <= int(u) < 0xE000 for u in units):
            continue      # the dump replaces unpaired surrogates, forms would not round-trip
        wp = word_positions(units)
        n = len(toks)
        pos = np.array([x[2] for x in toks], np.uint32); ln = np.array([x[3] for x in toks], np.uint32); tag = np.array([x[1] for x in toks], np.uint8)
        fu = [np.frombuffer(f[1].encode("utf-16-le"), dtype="<u2") for f in forms]
        off = np.zeros(n + 1, np.uint32); off[1:] = np.cumsum([len(x) for x in fu])
        blob = np.ascontiguousarray(np.concatenate(fu + [np.zeros(1, "<u2")]))
        is_yo = np.array([1 if f[1] == "요" else 0 for f in forms], np.uint8)
        wpin = np.array([wp[p] for p in pos], np.uint32)
        out = np.zeros((n, 5), np.int32)
        units = np.ascontiguousarray(units)
        lib.kb_asm_run(n, pos.ctypes.data, ln.ctypes.data, tag.ctypes.data, blob.ctypes.data, off.ctypes.data, is_yo.ctypes.data, wpin.ctypes.data,
                       units.ctypes.data if len(units) else None, len(units), out.ctypes.data)
        assert [tuple(int(v) for v in r) for r in out] == [f[0] for f in forms], (t, out.tolist(), [f[0] for f in forms])
        checked += 1
    assert checked > 0.9 * len(gold)


@pytest.mark.parametrize("name", ["inputs_web", "inputs_written", "inputs_dialect_typos"])
def test_own_substring_forms(name):
    """Tokens that keep their own substring (symbols, foreign words, numbers, unknown nouns) take it from the NORMALISED text;
    a token may start at the coda of a raw character ("몈ㅋㅋㅋ" -> "며" + "ㅋㅋㅋㅋ").  The host reconstruction
    (assemble.h ownSubstringForm) must reproduce TokenInfo::str of the reference for every such token."""
    if not os.path.exists(LIB):
        pytest.skip("tests/hostsim/libassemble_check.so missing: run __graft_entry__.build()")
    lib = C.CDLL(LIB)
    lib.kb_own_form.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_int]
    OWN_TAGS = set(range(21, 39)) | {0}      # sf..w_emoji and unknown: never dictionary forms in these inputs
    texts = read_inputs(name); gold = read_golden(name)
    checked = split = 0
    buf = np.zeros(4096, np.uint16)
    for t, g in zip(texts, gold):
        toks = g["tokens"]; forms = g.get("forms")
        units = np.ascontiguousarray(np.frombuffer(t.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
        if not toks or forms is None or any(0xD800 <= int(u) < 0xE000 for u in units):
            continue
        for i, (tok, f) in enumerate(zip(toks, forms)):
            if (tok[1] & 0x7F) not in OWN_TAGS:
                continue
            begins = i > 0 and toks[i - 1][2] + toks[i - 1][3] > tok[2]
            ends = i + 1 < len(toks) and toks[i + 1][2] < tok[2] + tok[3]
            n = lib.kb_own_form(units.ctypes.data, len(units), 1, tok[2], tok[3], int(begins), int(ends), buf.ctypes.data, len(buf))
            got = bytes(buf[:n].astype("<u2").tobytes()).decode("utf-16-le")
            assert got == f[1], (t, tok, got, f[1])
            checked += 1; split += int(begins or ends)
    assert checked > 100
    print("%s: %d own-substring tokens, %d of them split a raw character" % (name, checked, split))
