"""The C-ABI library loads, exports every symbol include/metalign_b200.h declares, and fails loudly
(no CPU fallback) when there is no CUDA device.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

from conftest import HAS_GPU, ROOT


def _declared():
    with open(os.path.join(ROOT, "include", "metalign_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mlg_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from metalign_b200 import _lib
    L = C.CDLL(_lib.build())
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), "missing export: " + n
    assert sorted(_lib.SIGNATURES) == names, "python signatures and header disagree"
    assert _lib.lib().mlg_version() >= 100


def test_stats_struct_matches_header():
    from metalign_b200._lib import Stats
    assert C.sizeof(Stats) == 8 * 7 + 4 * 2 + 8 * 2 + 8 * 2 + 4 * 2 + 8 + 4 * 2


@pytest.mark.skipif(HAS_GPU, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    from metalign_b200.api import Context, MlgError
    with pytest.raises(MlgError) as ei:
        Context(0)
    assert "no CPU fallback" in str(ei.value) or "CUDA" in str(ei.value)


def test_synth_libraries_build():
    import synth
    synth.cpu_lib()
    assert os.path.exists(os.path.join(ROOT, "synth", "synth_cuda.cu"))
