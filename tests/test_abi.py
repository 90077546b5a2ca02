"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/algodsp_cuda.h declares, its pure helpers follow the reference's rules, and it fails
loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import pytest

from algo_dsp_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "algodsp_cuda.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"ADSP_API\s+[\w\s\*]+?\b(adsp_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (adsp_\w+)", out))
    assert set(names) <= exported
    # ctypes prototypes cover the whole header too
    assert set(names) == set(L.PROTOTYPES)


def test_header_cites_reference_for_every_entry_group():
    src = open(HEADER).read()
    for ref in ("conv.go:76", "conv.go:194", "overlap_save.go:53", "overlap_add.go:44", "correlate.go:16", "correlate.go:200",
                "partitioned.go:212"):
        assert ref in src


def test_pure_helpers_follow_reference_rules():
    lib = L.load()
    for n, want in [(1, 1), (2, 2), (3, 4), (5, 8), (7, 8), (8, 8), (9, 16), (100, 128), (0, 1), (-3, 1)]:
        assert lib.adsp_next_pow2(n) == want          # conv.go:250-261, conv_test.go:343-362
    assert lib.adsp_is_pow2(64) == 1 and lib.adsp_is_pow2(100) == 0 and lib.adsp_is_pow2(0) == 0
    f, s = C.c_int64(), C.c_int64()
    assert lib.adsp_ols_sizes(96000, 0, C.byref(f), C.byref(s)) == L.OK and (f.value, s.value) == (262144, 166145)
    assert lib.adsp_ols_sizes(3, 100, C.byref(f), C.byref(s)) == L.ERR_INVALID_BLOCK_SIZE   # conv_test.go:675-682
    assert "power of 2" in L.last_error()
    assert lib.adsp_ols_sizes(0, 0, C.byref(f), C.byref(s)) == L.ERR_EMPTY_KERNEL
    assert lib.adsp_ols_sizes(300, 256, C.byref(f), C.byref(s)) == L.OK and f.value == 1024   # silently raised
    b = C.c_int64()
    assert lib.adsp_ola_sizes(64, 256, C.byref(b), C.byref(f)) == L.OK and (b.value, f.value) == (256, 512)  # example_test.go:79-81
    assert lib.adsp_ola_sizes(96000, 0, C.byref(b), C.byref(f)) == L.OK and (b.value, f.value) == (131072, 262144)
    st, ln = C.c_int64(), C.c_int64()
    for mode, want in [(0, (0, 7)), (1, (1, 5)), (2, (2, 3))]:     # trimToMode conv.go:229-247 on lenA=5, lenB=3
        lib.adsp_trim_mode(5, 3, mode, C.byref(st), C.byref(ln))
        assert (st.value, ln.value) == want
    lib.adsp_trim_mode(3, 5, 2, C.byref(st), C.byref(ln))
    assert (st.value, ln.value) == (2, 3)
    for lag in range(-9, 10):                                       # conv_test.go:373-385
        assert lib.adsp_lag_from_index(lib.adsp_index_from_lag(lag, 10), 10) == lag
    assert lib.adsp_status_string(L.ERR_EMPTY_INPUT) == b"conv: empty input"


def test_no_cpu_fallback_without_gpu():
    lib = L.load()
    if lib.adsp_device_count() > 0:
        pytest.skip("a CUDA device is present")
    h = C.c_void_p()
    assert lib.adsp_ctx_create(0, C.byref(h)) == L.ERR_CUDA
    assert "no CPU fallback" in L.last_error()
    from algo_dsp_b200 import conv
    with pytest.raises(conv.ConvError) as ei:
        conv.Convolve([1.0, 2.0], [1.0])
    assert conv.errors_is(ei.value, conv.ErrCUDA)


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "algo_dsp_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f == "siggen.py", f"{f} mentions the oracle"
    # development helpers outside tests/ must not use the checker either (those that do live in tests/tools/)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tools")):
        for f in files:
            if f.endswith((".py", ".sh", ".cu")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "from oracle" not in txt and "import oracle" not in txt, f"tools/{f} imports the oracle"


def test_partitioned_delay_line_layout_invariants():
    """Host arithmetic of the streaming engine's stage layout (csrc/fdl.cu fdl_layout), no GPU: partitions cover the IR,
    are contiguous, and every stage can wait for whole blocks and still meet the latency (size <= offset + latency + 1)."""
    import ctypes as C
    from algo_dsp_b200 import _lib as L
    lib = L.load()
    for K in (1, 7, 128, 129, 5000, 96000, 288000, 1 << 20):
        for mn in (3, 5, 7, 11, 12, 13):
            for mx in (mn, mn + 2, 13, 20):
                if mx < mn:
                    continue
                ps, cnt, off = (C.c_int * 32)(), (C.c_int * 32)(), (C.c_int64 * 32)()
                k = lib.adsp_partitioned_plan_layout(K, mn, mx, ps, cnt, off, 32)
                assert 1 <= k <= 32
                lat, pos = 1 << mn, 0
                for i in range(k):
                    assert off[i] == pos and ps[i] <= off[i] + lat + 1 and ps[i] <= max(8, min(2048, 1 << mx))
                    assert ps[i] & (ps[i] - 1) == 0 and cnt[i] >= 1
                    pos += ps[i] * cnt[i]
                assert pos >= K and pos - K < ps[k - 1]
    assert lib.adsp_partitioned_plan_layout(100, 2, 5, None, None, None, 0) == 0      # minBlockOrder < 3: no delay-line engine
