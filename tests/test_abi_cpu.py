"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/wae.h declares,
validates arguments with the reference's error text, and refuses to run without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(pkg):
    import __graft_entry__ as ge
    ge.build()
    header = open(os.path.join(ROOT, "include", "wae.h")).read()
    declared = set(re.findall(r"WAE_API\s+[\w\s\*]+?\b(wae_\w+)\s*\(", header))
    assert len(declared) >= 40
    lib = ctypes.CDLL(pkg.LIB_PATH)
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert missing == []
    assert set(pkg._binding.WAE_SYMBOLS) <= declared


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.WaeError) as e:
        pkg.Engine(0)
    assert e.value.status == 7  # WAE_NO_DEVICE
    assert "no CPU fallback" in str(e.value)


def test_product_never_references_the_oracle():
    # the oracle is test infrastructure: nothing under the package may import, link or name it
    pkg_dir = os.path.join(ROOT, "web-audio-api-rs_b200")
    for dp, _, files in os.walk(pkg_dir):
        if os.path.basename(dp) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                text = open(os.path.join(dp, f), errors="ignore").read()
                assert "liboracle" not in text and "oracle/" not in text and "wao_core" not in text, f
