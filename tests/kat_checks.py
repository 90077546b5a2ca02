"""Known-answer checks of tests/golden/reference_kats.json against an implementation adapter
(tests/adapters.py).  Each check mirrors one test of the reference (file:line in the JSON)."""
import json
import os

import numpy as np
import pytest

import kat_inputs as KI
from adapters import ImplError

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "reference_kats.json")))
MODE = {"full": 0, "same": 1, "valid": 2}


def _err(fn, want):
    with pytest.raises(ImplError) as ei:
        fn()
    assert ei.value.sentinel == want


def check_direct(A):
    for k in KATS["direct"]:
        got = A.direct(k["a"], k["b"])
        assert len(got) == len(k["want"])
        assert np.max(np.abs(got - np.array(k["want"], float))) <= k["tol"], k["src"]
    e = KATS["direct_example"]
    got = A.direct(e["a"], e["b"])
    assert len(got) == e["want_len"]
    assert [f"{v:.2f}" for v in got[:3]] == [f"{v:.2f}" for v in e["want_head"]]
    c = KATS["direct_circular"]
    assert np.max(np.abs(A.direct_circular(c["a"], c["b"]) - np.array(c["want"], float))) <= c["tol"]


def check_helpers(A):
    for n, want in KATS["next_power_of_2"]["cases"]:
        assert A.next_power_of_2(n) == want
    lr = KATS["lag_roundtrip"]
    for lag in lr["lags"]:
        assert A.lag_from_index(A.index_from_lag(lag, lr["len_b"]), lr["len_b"]) == lag
    fe = KATS["find_peak_empty"]
    assert A.find_peak([]) == (fe["index"], fe["value"])


def check_lengths_and_modes(A):
    L = KATS["lengths"]
    sig = np.sin(2 * np.pi * np.arange(L["signal_len"]) / 50)
    assert len(A.convolve(sig, L["short_kernel"])) == L["short_len"]
    assert len(A.convolve(sig, np.exp(-np.arange(L["long_kernel_len"]) / 20))) == L["long_len"]
    m = KATS["modes"]
    for fn in (A.convolve_mode, A.correlate_mode):
        for name in ("full", "same", "valid"):
            assert len(fn(m["a"], m["b"], MODE[name])) == m[name]
    oa = KATS["overlap_add_example"]
    kern = np.exp(-np.arange(oa["kernel_len"]) / 10)
    kern[0] = 1
    assert A.ola_sizes(oa["kernel_len"], oa["block_size_arg"]) == (oa["block_size"], oa["fft_size"])
    sig = np.sin(2 * np.pi * np.arange(oa["signal_len"]) / 20)
    for _ in range(3):
        assert len(A.overlap_add(kern, oa["block_size_arg"], sig)) == oa["result_len"]


def check_cross_impl(A):
    for k in KATS["cross_impl"]:
        x = KI.gen(k["gen"])
        h = np.array(k["kernel"], float) if "kernel" in k else KI.gen(k["kernel_gen"])
        ref = A.direct(x, h)
        got = getattr(A, k["op"])(x, h)
        assert len(got) == len(ref)
        assert np.max(np.abs(got - ref)) <= k["tol"], k["src"]
    c = KATS["commutative"]
    assert np.max(np.abs(A.convolve(c["a"], c["b"]) - A.convolve(c["b"], c["a"]))) <= c["tol"]


def check_correlation(A):
    e = KATS["correlate_example"]
    r = A.correlate(e["signal"], e["template"])
    idx, val = A.find_peak(r)
    assert idx == e["peak_index"] and A.lag_from_index(idx, len(e["template"])) == e["lag"]
    assert f"{val:.2f}" == f"{e['peak_value']:.2f}"
    a = KATS["autocorrelate_example"]
    s = KI.gen(a["gen"])
    r = A.auto_correlate_normalized(s)
    n, p = a["gen"]["n"], a["gen"]["period"]
    assert f"{r[n - 1]:.4f}" == f"{a['zero_lag']:.4f}"
    assert f"{r[n - 1 + p]:.4f}" == f"{a['one_period_lag']:.4f}"
    k = KATS["correlate_fft_vs_correlate"]
    assert np.max(np.abs(A.correlate_fft(k["a"], k["b"]) - A.correlate(k["a"], k["b"]))) <= k["tol"]
    k = KATS["correlate_direct_vs_correlate"]
    assert np.max(np.abs(A.correlate_direct(k["a"], k["b"]) - A.correlate(k["a"], k["b"]))) <= k["tol"]
    k = KATS["autocorr_cos_peak"]
    assert A.find_peak(A.auto_correlate(KI.gen(k["gen"])))[0] == k["peak_index"]
    k = KATS["autocorr_normalized"]
    assert abs(A.auto_correlate_normalized(k["a"])[len(k["a"]) - 1] - k["zero_lag"]) <= k["tol"]
    k = KATS["correlate_normalized_peak"]
    assert abs(A.find_peak(A.correlate_normalized(k["a"], k["a"]))[1] - k["peak"]) <= k["tol"]


def check_errors(A):
    for k in KATS["errors"]:
        op = k["op"]
        if op in ("direct", "correlate_fft", "correlate_direct", "direct_circular"):
            _err(lambda: getattr(A, op)(k["a"], k["b"]), k["err"])
        elif op == "new_overlap_save":
            _err(lambda: A.ols_sizes(len(k["kernel"]), k["fft_size"]), k["err"])
        elif op == "process_to_wrong_len":
            sig = (np.arange(k["signal_len"]) % 10).astype(float)
            _err(lambda: A.overlap_save_to(k["kernel"], 0, sig, k["out_len"]), k["err"])
            ok = A.overlap_save_to(k["kernel"], 0, sig, k["signal_len"] + len(k["kernel"]) - 1)
            assert len(ok) == k["signal_len"] + len(k["kernel"]) - 1


def check_partitioned(A):
    P = KATS["partitioned"]
    kern = lambda n: 0.99 ** np.arange(n)
    for o in P["latency"]["orders"]:
        p = A.partitioned(kern(P["latency"]["kernel_len"]), o, o + 4)
        assert A.part_info(p)["latency"] == 1 << o
    for c in P["vs_streaming_ola"]:
        h = kern(c["kernel_len"])
        x = KI.pcg_uniform(c["signal_len"])
        lat = 1 << c["min"]
        p = A.partitioned(h, c["min"], c["max"])
        y = A.part_process(p, np.concatenate([x, np.zeros(lat)]))[lat:]
        ref = np.convolve(x, h)[: len(x)]  # what the streaming OLA reference produces (oracle test pins that)
        assert np.max(np.abs(y[: len(x)] - ref)) <= c["tol"]
    r = P["reset"]
    p = A.partitioned(kern(r["kernel_len"]), r["min"], r["max"])
    x = KI.pcg_uniform(r["signal_len"])
    o1 = A.part_process(p, x)
    A.part_reset(p)
    o2 = A.part_process(p, x)
    assert np.max(np.abs(o1 - o2)) <= r["tol"]
    d = P["dirac"]
    x = KI.pcg_uniform(d["signal_len"])
    lat = 1 << d["min"]
    p = A.partitioned([1.0], d["min"], d["max"])
    out = A.part_process(p, np.concatenate([x, np.zeros(lat)]))
    assert np.max(np.abs(out[lat: lat + len(x)] - x)) <= d["tol"]
    k = P["kernel_len"]
    assert A.part_info(A.partitioned(kern(k["kernel_len"]), k["min"], k["max"]))["kernel_len"] == k["kernel_len"]
    for e in P["errors"]:
        if "in_len" in e:
            p = A.partitioned(e["kernel"], e["min"], e["max"])
            _err(lambda: A.part_process(p, np.zeros(e["in_len"]), e["out_len"]), e["err"])
        else:
            _err(lambda: A.partitioned(e["kernel"], e["min"], e["max"]), e["err"])
    for s in P["stage_layouts"]:
        p = A.partitioned(kern(s["kernel_len"]), s["min"], s["max"])
        info = A.part_info(p)
        assert [tuple(x) for x in info["stages"]] == [(1 << o, c) for o, _, c in s["stages"]]
        n = len(info["stages"])
        _err(lambda: A.part_stage_info(p, -1), "ErrStageIndexOutOfRange")
        _err(lambda: A.part_stage_info(p, n), "ErrStageIndexOutOfRange")


ALL_CHECKS = [check_direct, check_helpers, check_lengths_and_modes, check_cross_impl, check_correlation, check_errors,
              check_partitioned]
