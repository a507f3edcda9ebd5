"""CPU-side checks of the boundary: the library loads and exports every symbol include/*.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "riptable_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rtb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    import riptable_b200.build as b

    b.build()
    from riptable_b200 import _lib

    names = _declared()
    assert len(names) >= 20
    assert _lib.MISSING == []
    for n in names:
        assert hasattr(_lib.lib, n), f"{n} declared in include/riptable_b200.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert _lib.lib.rtb_version() >= 100


def test_out_dtype_table_matches_contract():
    # SURVEY 8a': sum ints -> int64 (uint64 stays), floats keep; mean/var/std -> f64; min/max keep dtype
    from riptable_b200 import _lib
    from riptable_b200.enums import RTB_BOOL, RTB_F32, RTB_F64, RTB_I8, RTB_I64, RTB_U64, RTB_BYTES

    f = _lib.lib.rtb_reduce_out_dtype
    assert f(0, RTB_I8) == RTB_I64 and f(0, RTB_BOOL) == RTB_I64 and f(0, RTB_U64) == RTB_U64
    assert f(0, RTB_F32) == RTB_F32 and f(50, RTB_F64) == RTB_F64
    assert f(1, RTB_I8) == RTB_F64 and f(55, RTB_F32) == RTB_F64
    assert f(2, RTB_I8) == RTB_I8 and f(53, RTB_F32) == RTB_F32
    assert f(0, RTB_BYTES) < 0 and f(100, RTB_I8) < 0


def test_every_exported_entry_point_is_declared():
    """The reverse direction: nothing is exported under the rtb_ prefix that include/riptable_b200.h does not declare
    (one debugging aid excepted), so the header is the whole boundary."""
    import subprocess
    import riptable_b200.build as b

    b.build()
    out = subprocess.run(["nm", "-D", "--defined-only", b.LIB], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if line.split() and line.split()[-1].startswith("rtb_")}
    undeclared = exported - set(_declared()) - {"rtb_debug_sort_codes"}
    assert not undeclared, sorted(undeclared)


def test_cumop_workspace_query_follows_the_form_a_call_takes():
    """Host arithmetic only (no device): GB_CUMSUM / GB_CUM[NAN]MIN / MAX without a reset filter, from 4 Mi rows and up to
    16 Ki bins, take the sort-free blocked form, whose workspace is the chunk matrix; everything else the packed layout,
    which grows with the row count.  rtb_cumsum_workspace_bytes covers either."""
    from riptable_b200 import _lib

    q = _lib.lib.rtb_cumop_workspace_bytes
    n = 64 << 20
    packed = q(302, n, 1000, 0)  # cumprod: always the packed layout
    assert packed > 20 * n
    for func in (300, 306, 307, 308, 309):
        blocked = q(func, n, 1000, 0)
        assert blocked < packed // 8, func
        assert q(func, n, 1000, 1) == packed          # a reset filter
        assert q(func, n, 1_000_000, 0) == packed     # too many bins
        assert q(func, 1 << 20, 1000, 0) == q(302, 1 << 20, 1000, 0)  # too few rows
    assert q(300, n, 16_383, 0) < packed and q(300, n, 16_384, 0) == packed
    assert _lib.lib.rtb_cumsum_workspace_bytes(n, 1000) >= packed
