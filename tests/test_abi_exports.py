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


def test_rust_shim_binds_every_header_symbol():
    """ffi/src/sys.rs is generated from include/bzhalo2.h (scripts/gen_rust_bindings.py): regenerating it changes nothing, every
    declared symbol has its `pub fn` with the header's arity, and lib.rs / the halo2_proofs patch only name symbols that exist."""
    import re, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    assert subprocess.run([sys.executable, os.path.join(root, "scripts", "gen_rust_bindings.py"), "--check"]).returncode == 0, \
        "ffi/src/sys.rs is stale: run python scripts/gen_rust_bindings.py"
    sys_rs = open(os.path.join(root, "battlezips-halo2_b200", "ffi", "src", "sys.rs")).read()
    declared = set(re.findall(r"pub fn (bz_[a-z0-9_]+)\(", sys_rs))
    assert declared == set(bz.EXPORTS)
    header = re.sub(r"/\*.*?\*/", "", open(bz.binding.header_path()).read(), flags=re.S)
    for name, args in re.findall(r"\b(bz_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;", header, flags=re.S):
        arity = 0 if args.strip() == "void" else args.count(",") + 1
        rust_args = re.search(r"pub fn %s\((.*?)\)( ->|;)" % name, sys_rs).group(1)
        assert (rust_args.count(":") if rust_args else 0) == arity, name
    for f in ("src/lib.rs", "halo2_proofs-0.2.0-bz.patch"):
        used = set(re.findall(r"\b(bz_[a-z0-9_]+)\(", open(os.path.join(root, "battlezips-halo2_b200", "ffi", f)).read()))
        used -= {"bz_token"}
        assert used <= declared, (f, used - declared)
