"""CPU: the C-ABI library loads and exports every symbol include/bzhalo2.h declares (no compute calls)."""
import ctypes, os
import pytest
import battlezips_halo2_b200 as bz


def test_library_built_and_exports_every_declared_symbol():
    path = bz.lib_path()
    assert os.path.exists(path), "libbzhalo2.so missing: run __graft_entry__.build()"
    lib = ctypes.CDLL(path)
    assert len(bz.EXPORTS) >= 20
    missing = [s for s in bz.EXPORTS if not hasattr(lib, s)]
    assert not missing, f"declared in include/bzhalo2.h but not exported: {missing}"


def test_version_string():
    lib = bz.load_library()
    assert b"sm_100a" in lib.bz_version()


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product refuses to construct a context instead of computing on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(bz.BzError):
        bz.Context(0)


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the product package may reference it."""
    pkg = os.path.dirname(bz.lib_path().rsplit("/lib/", 1)[0] + "/x")
    for dirpath, _, files in os.walk(pkg):
        if "/build" in dirpath or "/lib" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cc", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
