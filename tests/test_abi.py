"""The C-ABI library loads and exports every symbol include/evoworld_b200.h declares (no compute)."""
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def declared_symbols():
    text = (ROOT / "include" / "evoworld_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(evw_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    from evoworld_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(built_lib, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in evoworld_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(names)


def test_version_and_sizes(built_lib):
    assert built_lib.evw_abi_version() >= 1
    assert built_lib.evw_splat_workspace(4, 512) == 4 * 6 * 512 * 512 * 8
    assert built_lib.evw_splat_workspace_flags(4, 512, 2) == 2 * 4 * 6 * 512 * 512 * 8
    assert built_lib.evw_splat_workspace_flags(4, 512, 1) == 4 * 6 * 512 * 512 * 8
    assert built_lib.evw_conf_select_workspace(5_076_400) > 0
    assert built_lib.evw_last_error() is not None


def test_library_is_sm100a_only():
    import subprocess

    from evoworld_b200 import _lib
    from evoworld_b200.build import build_cuda

    build_cuda()
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
