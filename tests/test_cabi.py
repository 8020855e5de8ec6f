"""CPU: the C-ABI boundary. The shared library must load and export every function include/echo_b200.h declares;
no compute call is made here (there is no GPU), except to check that the product refuses to run without one."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "echo_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    names = re.findall(r"^\s*(?:int|const char\s*\*|void)\s+(echo_[a-z0-9_]+)\s*\(", src, flags=re.M)
    assert len(names) >= 20
    return sorted(set(names))


def test_library_exports_every_declared_symbol():
    from echo_tts_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libecho_b200.so not built: run python __graft_entry__.py"
    lib = C.CDLL(str(_lib.LIB_PATH))
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, f"declared in include/echo_b200.h but not exported: {missing}"
    _lib.load(strict=True)


def test_ctypes_structs_match_header_sizes():
    """The ctypes mirrors must have the C layout: spot-check sizes that would shift with a missing field."""
    from echo_tts_b200 import _lib
    assert C.sizeof(_lib.DitConfig) == 18 * 4
    assert C.sizeof(_lib.SamplerArgs) % 8 == 0 and C.sizeof(_lib.SamplerArgs) >= 17 * 4
    assert C.sizeof(_lib.AttnDesc) > 4 * C.sizeof(_lib.AttnSegment)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """Without a CUDA device the product path fails loudly (no oracle / CPU fallback behind the API)."""
    from echo_tts_b200 import _lib
    from echo_tts_b200.config import DitConfig
    from echo_tts_b200.model import B200EchoDiT
    with pytest.raises(_lib.EchoError):
        B200EchoDiT(DitConfig.tiny(), "cuda:0")
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.echo_create(C.byref(h), 0)
    assert rc != 0 and b"no CPU fallback" in lib.echo_last_error()


def test_product_does_not_import_the_oracle():
    """oracle/ is test infrastructure: nothing under echo_tts_b200/ may import or reference it."""
    pkg = os.path.join(ROOT, "echo_tts_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "echo_oracle" not in text and "host_oracle" not in text, f
