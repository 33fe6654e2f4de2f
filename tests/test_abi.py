"""CPU: the C-ABI library loads and exports every symbol include/*.h declares; without a GPU it fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

from kolibrie_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for hdr in ("kolibrie_b200.h", "cudajoin.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(kb_[a-z0-9_]+|perform_hash_join_cuda)\s*\(", text))
    names -= {"kb_filter_opcode"}
    return names


def test_library_exports_every_declared_symbol():
    L = C.CDLL(capi.LIB_PATH)
    decl = declared_symbols()
    assert len(decl) >= 39
    for name in sorted(decl):
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert set(capi.EXPORTED_SYMBOLS) == decl


def test_legacy_alias_library_exports_the_reference_symbol():
    """Kolibrie links `cudajoin` (kolibrie/build.rs:75-79) and binds perform_hash_join_cuda (cuda_join.rs:14-26)."""
    L = C.CDLL(capi.LEGACY_LIB_PATH)
    assert hasattr(L, "perform_hash_join_cuda")


def test_version_and_loud_failure_without_gpu():
    L = capi.lib()
    assert b"sm_100a" in L.kb_version()
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.KolibrieError) as e:
        capi.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_struct_layouts_match_header():
    assert C.sizeof(capi.KbTerm) == 8 and C.sizeof(capi.KbPattern) == 24 and C.sizeof(capi.KbFilterOp) == 24
    assert C.sizeof(capi.KbAgg) == 8 and C.sizeof(capi.KbRuleFilter) == 24 and C.sizeof(capi.KbRule) == 48
    assert C.sizeof(capi.KbFixpointStats) == 8 + 8 + 8 + 64 * 8 + 8
    assert capi.lib().kb_shard_of(12345, 8) < 8


def test_every_entry_point_is_documented_for_the_integrator():
    """INTEGRATION.md is the binding guide: every function the header declares must at least be named there"""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "kolibrie_b200.h")).read()
    doc = open(os.path.join(root, "INTEGRATION.md")).read()
    fns = re.findall(r"KB_API\s+[\w\s\*]+?\b(kb_\w+)\s*\(", header)
    assert len(fns) >= 45
    names = set(re.findall(r"kb_\w+", doc))
    # `kb_groups_info/keys/values/counts/free` style lists name a family by its prefix
    for fam in re.findall(r"(kb_\w+?_)(\w+(?:/\w+)+)", doc):
        for tail in fam[1].split("/"):
            names.add(fam[0] + tail)
    missing = [f for f in fns if f not in names]
    assert not missing, missing
