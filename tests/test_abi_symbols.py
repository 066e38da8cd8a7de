"""The C-ABI library loads without a GPU and exports every symbol include/nls_b200.h declares."""

import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "nls_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nls_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from neo_ls_svm_b200 import _lib

    lib = _lib.load()
    names = _declared()
    assert len(names) >= 20
    for name in names:
        fn = getattr(lib, name)  # raises AttributeError if the symbol is missing
        assert fn.argtypes is not None, f"{name} has no ctypes signature in _lib.py"
    assert lib.nls_version() >= 100


def test_no_gpu_means_loud_failure_not_fallback():
    """Without a CUDA device context creation fails with a message; nothing silently runs on the CPU."""
    import torch

    from neo_ls_svm_b200 import _lib

    if torch.cuda.is_available():
        return
    lib = _lib.load()
    handle = ctypes.c_void_p()
    status = lib.nls_ctx_create(0, None, ctypes.byref(handle))
    assert status != 0 and lib.nls_last_error()
    try:
        _lib.Context(0)
    except _lib.NlsError as exc:
        assert "no CPU fallback" in str(exc)
    else:
        raise AssertionError("Context() must raise without a GPU")


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "neo_ls_svm_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


def test_integration_stub_signatures_match_the_binding():
    """The ctypes stub shown in INTEGRATION.md has the argument count of the real binding for every entry point."""
    integ = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    binding = open(os.path.join(ROOT, "neo_ls_svm_b200", "_lib.py")).read()
    found = re.findall(r"lib\.(nls_\w+)\.argtypes = \[([^\]]*)\]", integ)
    assert len(found) >= 8
    for name, args in found:
        m = re.search(r'"%s": \(\[([^\]]*)\]' % name, binding)
        assert m, f"{name} is not bound in _lib.py"
        count = lambda text: len([a for a in text.split(",") if a.strip()])  # noqa: E731
        assert count(args) == count(m.group(1)), name
