"""CPU: the C-ABI library is built, loads, and exports every symbol include/distdiff_sm100.h declares."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "distdiff_sm100.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import __graft_entry__ as entry
    entry.build()
    from distdiff_b200 import _lib
    handle = _lib.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(handle, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(names)
    assert handle.dd_abi_version() == 1
    assert handle.dd_proto_workspace_bytes(2048, 100, 3) > 0
    assert handle.dd_agglo_workspace_bytes(30, 100) >= 100 * 30 * 30 * 8
    # peer arena: five exchanged buffers + the exchange kernel's inbox (one fp64 row + count per sender and owned row)
    hdr = handle.dd_peer_header_bytes()
    assert hdr >= 1024 + 4 * 65536 and hdr % 256 == 0                      # flag rows / status / stamps + one inbox flag per row
    for R, D, P in ((300, 2048, 8), (1000, 2048, 8), (301, 512, 3), (7, 64, 1)):
        rows_max = -(-R // P)
        need = handle.dd_peer_arena_bytes(R, D, P)
        assert need >= R * (D * 12 + 20) + P * rows_max * (D * 8 + 8)
        assert need <= R * (D * 12 + 20) + P * rows_max * (D * 8 + 8) + 16 * 256
    assert handle.dd_peer_arena_bytes(0, 2048, 8) == 0
    assert handle.dd_energy_workspace_bytes(65536, 100) >= 2 * 65536 * 4 + 102 * 8


def test_no_cpu_fallback():
    import torch
    from distdiff_b200 import ops
    from distdiff_b200._lib import DistDiffError
    x = torch.randn(1, 4, 8, 8)
    with pytest.raises(DistDiffError):
        ops.cfg_ddim_step(torch.randn(2, 4, 8, 8), x, 7.5, 0.5, 0.6)
    with pytest.raises(DistDiffError):
        ops.add_noise(x, x, 0.5)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "distdiff_b200")
    for base, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(base, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
