"""CPU: host simulation of the lattice kernel.  tests/hostsim/liblattice_sim.so is kiwi_b200/csrc/lattice.cu + model.cu — the
device SOURCES, not a restatement — compiled as plain C++ with a one-lane "warp" (tests/hostsim/shim/cuda_runtime.h).  It checks
the kernel's logic (normalisation, pattern spans, trie walk, typo graph, search states, connectivity pass) against the
lattices of the UNMODIFIED reference without a GPU; the parallel execution itself (32 lanes, __syncwarp ordering) is only
covered by the -m gpu parity tests.  Test infrastructure: nothing here is linked into libkiwi_b200.so."""
import ctypes as C, os
import numpy as np
import pytest
from tests.goldenio import read_golden, read_inputs
from tests.orc import IMAGE, TYPO_IMAGES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tests", "hostsim", "liblattice_sim.so")
MATCH_ALL_WITH_NORMALIZING = (1 << 0) | (1 << 1) | (1 << 2) | (1 << 3) | (1 << 4) | (1 << 5) | (1 << 23) | (1 << 16)
INPUTS = ["inputs_ref_tests", "inputs_web", "inputs_written", "inputs_dialect_typos"]


@pytest.fixture(scope="module")
def sim():
    if not os.path.exists(LIB) or not os.path.exists(IMAGE):
        pytest.skip("tests/hostsim/liblattice_sim.so or the model image missing: run __graft_entry__.build()")
    lib = C.CDLL(LIB)
    lib.hs_open.restype = C.c_void_p; lib.hs_open.argtypes = [C.c_char_p]
    lib.hs_close.argtypes = [C.c_void_p]
    lib.hs_lattice.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_int]
    lib.hs_set_typo.argtypes = [C.c_void_p, C.c_char_p, C.c_float, C.c_uint32, C.c_uint32]
    h = lib.hs_open(os.fsencode(IMAGE))
    assert h
    rows = np.zeros((1 << 17, 9), np.int32)

    def lattice(text):
        u = np.ascontiguousarray(np.frombuffer(text.encode("utf-16-le", "surrogatepass"), dtype="<u2"))
        n = lib.hs_lattice(h, u.ctypes.data, len(u), MATCH_ALL_WITH_NORMALIZING, rows.ctypes.data, len(rows))
        return n, (rows[:n].tolist() if n >= 0 else None)

    def set_typo(path, threshold=2.5, graph_per_unit=6, states_per_unit=6):
        assert lib.hs_set_typo(h, os.fsencode(path) if path else None, threshold, graph_per_unit, states_per_unit) == 0

    yield lattice, set_typo
    lib.hs_close(h)


@pytest.mark.parametrize("name", INPUTS)
def test_kernel_source_reproduces_reference_lattices(sim, name):
    lattice, set_typo = sim
    set_typo(None)
    for t, g in zip(read_inputs(name), read_golden(name)):
        n, rows = lattice(t)
        assert n >= 0 and rows == g["lattice"], t


@pytest.mark.parametrize("name", INPUTS)
def test_kernel_source_reproduces_reference_typo_lattices(sim, name):
    """BASELINE config 4: genTypoGraph + searchTypo of lattice.cu with the engine's default scratch capacities (6 graph nodes /
    6 search states per normalised-unit slot) against the reference's lattices under basicTypoSet, typoThreshold 2.5."""
    if not os.path.exists(TYPO_IMAGES["basic"]):
        pytest.skip("typo image missing")
    lattice, set_typo = sim
    set_typo(TYPO_IMAGES["basic"])
    try:
        for t, g in zip(read_inputs(name), read_golden("typo6_" + name)):
            n, rows = lattice(t)
            assert n >= 0 and rows == g["lattice"], t
    finally:
        set_typo(None)


def test_typo_scratch_overflow_is_reported(sim):
    """a scratch arena that is too small must end in ST_TYPO_OVERFLOW (7), which sends the sentence to the retry arena"""
    if not os.path.exists(TYPO_IMAGES["basic"]):
        pytest.skip("typo image missing")
    lattice, set_typo = sim
    set_typo(TYPO_IMAGES["basic"], 2.5, 1, 1)
    try:
        n, _ = lattice("가게를 운영하는 사장님께서 대신 구매 해달라고 하셔서 구매하게되었는데 가게에 달아놓으니까 이쁘네요")
        assert n == -7
    finally:
        set_typo(None)
