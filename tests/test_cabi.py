"""The C-ABI library loads without a GPU and exports every symbol include/ebm_b200.h declares."""

import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ebm_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ebm_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def lib():
    from torchebm_b200 import _lib, build

    build.build()
    return _lib.load()


def test_header_symbols_are_exported_and_bound(lib):
    from torchebm_b200 import _lib

    declared = _declared_symbols()
    assert len(declared) >= 15
    assert sorted(_lib.PROTOTYPES) == declared, "binding table and header disagree"
    for name in declared:
        assert getattr(lib, name) is not None


def test_abi_version_and_argument_errors_need_no_gpu(lib):
    from torchebm_b200 import _lib

    assert lib.ebm_abi_version() == _lib.EBM_ABI_VERSION
    # null descriptor -> EBM_ERR_INVALID before anything touches the device
    rc = lib.ebm_energy_f32(None, None, 1, None, None)
    assert rc == _lib.ERR_INVALID
    assert b"descriptor" in lib.ebm_last_error()
    d = _lib.EbmEnergyDesc()
    d.kind, d.dim = _lib.ENERGY_DOUBLE_WELL, 4
    rc = lib.ebm_langevin_burst_f32(ctypes.byref(d), None, None, 8, 3, None, None, 1, None, 1, 0, 0, None, None, 1, None)
    assert rc == _lib.ERR_INVALID
    with pytest.raises(ValueError):
        _lib.check(rc, "ebm_langevin_burst_f32")
    # ABI v9: a three-hidden-layer MLP descriptor must carry W3 / b3 in buf[7..8]; the struct layout the header documents
    assert ctypes.sizeof(_lib.EbmEnergyDesc) == 10 * 4 + 4 * 4 + 10 * 8
    m = _lib.EbmEnergyDesc()
    m.kind, m.dim, m.hidden1, m.hidden2, m.hidden3 = _lib.ENERGY_MLP, 8, 8, 8, 8
    for i in range(6):
        m.buf[i] = 64   # (never dereferenced: validation fails first)
    rc = lib.ebm_energy_f32(ctypes.byref(m), ctypes.c_void_p(64), 1, ctypes.c_void_p(64), None)
    assert rc == _lib.ERR_INVALID and b"W3" in lib.ebm_last_error()


def test_missing_library_raises_instead_of_falling_back(monkeypatch, tmp_path):
    from torchebm_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_lib.EbmLibraryError, match="no CPU fallback"):
        _lib.load()


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "torchebm_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_descriptor_layout_in_python_equals_the_header_compiled_as_c(tmp_path):
    """include/ebm_b200.h is plain C (no CUDA or torch types): compile a probe with gcc and compare sizeof / offsetof of
    EbmEnergyDesc and the ABI version with the ctypes mirror every Python call marshals through."""
    import shutil
    import subprocess

    from torchebm_b200 import _lib

    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    fields = [f[0] for f in _lib.EbmEnergyDesc._fields_]
    src = tmp_path / "probe.c"
    src.write_text(
        "#include <stddef.h>\n#include <stdio.h>\n#include \"ebm_b200.h\"\nint main(void) {\n"
        "  printf(\"sizeof %zu\\n\", sizeof(EbmEnergyDesc));\n  printf(\"abi %d\\n\", EBM_ABI_VERSION);\n"
        + "".join(f"  printf(\"{f} %zu\\n\", offsetof(EbmEnergyDesc, {f}));\n" for f in fields)
        + "  return 0;\n}\n")
    exe = tmp_path / "probe"
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = dict(line.split() for line in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["sizeof"]) == ctypes.sizeof(_lib.EbmEnergyDesc)
    assert int(out["abi"]) == _lib.EBM_ABI_VERSION
    for f in fields:
        assert int(out[f]) == getattr(_lib.EbmEnergyDesc, f).offset, f
