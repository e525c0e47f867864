"""CPU: the C-ABI library loads, exports every symbol that include/kiwi_b200.h declares, and refuses to work
without a GPU (no CPU fallback).  No compute calls here."""
import ctypes, os, re
import pytest
import kiwi_b200

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "kiwi_b200.h"), encoding="utf-8").read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(kiwi_[a-z0-9_]+)\s*\(", src)
    return sorted(set(n for n in names if not n.endswith("_t")))


def test_library_exports_every_declared_symbol():
    lib = kiwi_b200.load_library()
    names = declared_functions()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "libkiwi_b200.so does not export " + n


def test_version_and_error_convention():
    lib = kiwi_b200.load_library()
    assert lib.kiwi_version().startswith(b"0.23.1")
    lib.kiwi_clear_error()
    assert lib.kiwi_error() is None
    h = lib.kiwi_init(b"/nonexistent/model/path", 0, 0, 0)
    assert not h
    assert b"cannot open model image" in lib.kiwi_error()
    lib.kiwi_clear_error()
    assert lib.kiwi_close(None) == -2          # KIWIERR_INVALID_HANDLE


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from tests.orc import IMAGE
    if not os.path.exists(IMAGE):
        pytest.skip("model image missing")
    with pytest.raises(kiwi_b200.KiwiError, match="no CPU fallback"):
        kiwi_b200.Kiwi(IMAGE)


def test_image_header_layout_is_stable():
    # the flatten tool (oracle side) and the loader (product side) share include/kiwi_b200_image.h only
    src = open(os.path.join(ROOT, "include", "kiwi_b200_image.h")).read()
    assert "KB2_IMAGE_VERSION 6u" in src
    from tests.orc import IMAGE
    if os.path.exists(IMAGE):
        import struct
        with open(IMAGE, "rb") as f:
            magic, version = struct.unpack("<QI", f.read(12))
        assert magic == 0x31474D4932424B and version == 6


def test_global_config_by_value_plumbing():
    """kiwi_config_t crosses the ABI by value in both directions (capi.h:607-615); a NULL handle yields the zeroed struct."""
    lib = kiwi_b200.load_library()
    c = lib.kiwi_get_global_config(None)
    assert c.cut_off_threshold == 0 and c.max_unk_form_size == 0 and ctypes.sizeof(kiwi_b200.Config) == 52
    lib.kiwi_set_global_config(None, c)      # no-op on a NULL handle


def test_typo_handle_conventions_without_gpu():
    """kiwi_typo_get_default returns shared handles that must not be closed (capi.h:499,568); a flat typo image is uploaded to
    the device on preparation, so without a GPU preparing fails loudly; a missing file is an error, not a fallback."""
    lib = kiwi_b200.load_library()
    h = lib.kiwi_typo_get_default(kiwi_b200.TYPO_BASIC)
    assert h and h == lib.kiwi_typo_get_basic()
    assert not lib.kiwi_typo_get_default(99)
    assert lib.kiwi_typo_close(h) == -1
    lib.kiwi_clear_error()
    assert not lib.kiwi_b200_typo_load(b"/nonexistent/typo.img")
    assert b"cannot open typo image" in lib.kiwi_error()
    import torch
    from tests.orc import TYPO_IMAGES
    if not torch.cuda.is_available() and os.path.exists(TYPO_IMAGES["basic"]):
        with pytest.raises(kiwi_b200.KiwiError, match="CUDA error"):
            kiwi_b200.PreparedTypo(path=TYPO_IMAGES["basic"])
