"""CPU tier: the C-ABI library exists, loads, and exports exactly what include/hagrid_b200.h
declares; the ctypes mirror covers every declared entry point. No compute call is made."""
import re
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = ROOT / "include" / "hagrid_b200.h"


def declared_symbols():
    text = HEADER.read_text()
    return sorted(set(re.findall(r"HGB_API[^;]*?\b(hgb_\w+)\s*\(", text)))


def test_header_declares_the_seven_reference_entry_points():
    names = declared_symbols()
    for fn in ("hgb_build_grid", "hgb_merge_grid", "hgb_flatten_grid", "hgb_expand_grid", "hgb_compress_grid",
               "hgb_setup_traversal", "hgb_traverse_grid"):
        assert fn in names


def test_library_builds_and_exports_every_declared_symbol():
    from hagrid_b200.build import build_library
    lib = build_library()
    out = subprocess.run(["nm", "-D", "--defined-only", str(lib)], check=True, capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (hgb_\w+)", out))
    assert set(declared_symbols()) == exported


def test_ctypes_mirror_matches_header():
    from hagrid_b200 import api
    assert sorted(api.EXPORTED_SYMBOLS) == declared_symbols()
    lib = api.Library()          # loading must work without a GPU
    assert lib.impl == "hagrid_b200"
    assert lib.device_count() >= 0


def test_no_cpu_fallback_without_a_device():
    """On a box without a GPU the product must fail loudly, never fall back to the oracle."""
    from hagrid_b200 import HagridError, Library, Scene, scenes
    lib = Library()
    if lib.device_count() > 0:
        pytest.skip("GPU present")
    with pytest.raises(HagridError):
        Scene(scenes.cornell32(), lib=lib)


def test_product_does_not_import_the_oracle():
    for path in list((ROOT / "hagrid_b200").rglob("*.py")) + list((ROOT / "hagrid_b200" / "csrc").glob("*")):
        if path.is_file() and path.suffix in (".py", ".cu", ".cuh", ".cpp", ".h"):
            text = path.read_text()
            assert "oracle" not in text.replace("oracle/build_ref.sh", "").replace("oracle/_ref", "") or path.name == "c_api.cpp", path


def test_struct_layouts():
    from hagrid_b200 import api
    assert api.TRI_DTYPE.itemsize == 48 and api.RAY_DTYPE.itemsize == 32 and api.HIT_DTYPE.itemsize == 16
    assert api.CELL_DTYPE.itemsize == 32 and api.SMALL_CELL_DTYPE.itemsize == 16
    assert api.CELL_DTYPE.fields["begin"][1] == 12 and api.CELL_DTYPE.fields["max"][1] == 16
    assert api.SMALL_CELL_DTYPE.fields["max"][1] == 6 and api.SMALL_CELL_DTYPE.fields["begin"][1] == 12


def test_missing_library_is_a_loud_error(tmp_path):
    from hagrid_b200 import HagridError, Library
    with pytest.raises(HagridError, match="no CPU fallback"):
        Library(tmp_path / "libhagrid_b200.so")


def test_only_tests_bench_and_smoke_use_the_oracle():
    """oracle/ is test infrastructure: nothing outside tests/, bench.py and __graft_entry__.py imports it."""
    for path in list((ROOT / "hagrid_b200").rglob("*.py")) + list((ROOT / "tools").glob("*.py")):
        text = path.read_text()
        assert "from oracle" not in text and "import oracle" not in text, path
