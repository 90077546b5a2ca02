"""Writes tests/golden/reference_kats.json: every known-answer vector the reference's own tests
hold for the dsp/conv hot path, transcribed by hand from /root/reference (file:line cited per
entry).  The Go reference cannot run here (no toolchain), so these literals -- not regenerated
outputs -- are what pins the oracle.  Inputs given by formula in the Go test are described by a
small "gen" record that tests/kat_inputs.py expands.

    python tests/golden/make_reference_kats.py
"""
import json
import os

KATS = {
    "direct": [  # TestDirect, dsp/conv/conv_test.go:9-69 (tol 1e-10)
        {"src": "conv_test.go:16-21", "a": [1, 2, 3], "b": [1, 1, 1], "want": [1, 3, 6, 5, 3], "tol": 1e-10},
        {"src": "conv_test.go:22-27", "a": [1, 2, 3, 4, 5], "b": [1], "want": [1, 2, 3, 4, 5], "tol": 1e-10},
        {"src": "conv_test.go:28-33", "a": [1, 2, 3, 4, 5], "b": [0, 0, 1], "want": [0, 0, 1, 2, 3, 4, 5], "tol": 1e-10},
        {"src": "conv_test.go:34-49", "a": [1, 2, 1], "b": [1, 2, 1], "want": [1, 4, 6, 4, 1], "tol": 1e-10},
    ],
    "direct_example": {  # ExampleDirect, example_test.go:10-27
        "src": "example_test.go:10-27", "a": [1, 2, 3, 4, 5, 4, 3, 2, 1], "b": [0.25, 0.5, 0.25],
        "want_len": 11, "want_head": [0.25, 1.00, 2.00], "print_decimals": 2,
    },
    "direct_circular": {  # TestDirectCircular, conv_test.go:83-98
        "src": "conv_test.go:83-98", "a": [1, 2, 3, 4], "b": [1, 0, 0, 0], "want": [1, 2, 3, 4], "tol": 1e-10,
    },
    "next_power_of_2": {  # TestHelperFunctions, conv_test.go:343-362
        "src": "conv_test.go:343-362", "cases": [[1, 1], [2, 2], [3, 4], [5, 8], [7, 8], [8, 8], [9, 16], [100, 128]],
    },
    "l2_norm": {"src": "conv_test.go:364-370", "x": [3, 4], "want": 5.0, "tol": 1e-10},
    "lengths": {  # ExampleConvolve, example_test.go:29-53
        "src": "example_test.go:29-53", "signal_len": 1000, "short_kernel": [0.2, 0.3, 0.3, 0.2], "short_len": 1003,
        "long_kernel_len": 100, "long_len": 1099,
    },
    "overlap_add_example": {  # ExampleOverlapAdd, example_test.go:55-85
        "src": "example_test.go:55-85", "kernel_len": 64, "block_size_arg": 256, "block_size": 256, "fft_size": 512,
        "signal_len": 500, "result_len": 563,
    },
    "correlate_example": {  # ExampleCorrelate, example_test.go:87-102
        "src": "example_test.go:87-102", "signal": [0, 0, 0, 1, 2, 3, 2, 1, 0, 0, 0], "template": [1, 2, 3, 2, 1],
        "peak_index": 7, "lag": 3, "peak_value": 19.0, "print_decimals": 2,
    },
    "autocorrelate_example": {  # ExampleAutoCorrelate, example_test.go:104-127
        "src": "example_test.go:104-127", "gen": {"kind": "sine_period", "n": 100, "period": 20},
        "zero_lag": 1.0, "one_period_lag": 0.8, "print_decimals": 4,
    },
    "modes": {  # TestConvolveMode conv_test.go:221-242, TestCorrelateMode :536-561
        "src": "conv_test.go:221-242,536-561", "a": [1, 2, 3, 4, 5], "b": [1, 2, 3], "full": 7, "same": 5, "valid": 3,
    },
    "lag_roundtrip": {"src": "conv_test.go:373-385", "len_b": 10, "lags": list(range(-9, 10))},
    "find_peak_empty": {"src": "conv_test.go:668-673", "index": -1, "value": 0.0},
    "cross_impl": [  # FFT paths against the time-domain Direct (tolerance ladder)
        {"src": "conv_test.go:100-130", "op": "overlap_add_convolve", "gen": {"kind": "sine", "n": 1000, "period": 100},
         "kernel": [0.25, 0.5, 0.25], "tol": 1e-10},
        {"src": "conv_test.go:132-169", "op": "overlap_save_convolve", "gen": {"kind": "sine", "n": 500, "period": 50},
         "kernel": [0.2, 0.3, 0.3, 0.2], "tol": 1e-8},
        {"src": "conv_test.go:171-194", "op": "convolve", "gen": {"kind": "mod10", "n": 1000}, "kernel": [1, 2, 1], "tol": 1e-10},
        {"src": "conv_test.go:196-219", "op": "convolve", "gen": {"kind": "mod10", "n": 1000},
         "kernel_gen": {"kind": "exp_decay", "n": 100, "tau": 20}, "tol": 1e-8},
    ],
    "correlate_fft_vs_correlate": {"src": "conv_test.go:463-485", "a": [1, 2, 3, 4, 5], "b": [1, 2, 3], "tol": 1e-8},
    "correlate_direct_vs_correlate": {"src": "conv_test.go:494-511", "a": [1, 2, 3, 4, 5], "b": [1, 2, 3], "tol": 1e-10},
    "autocorr_cos_peak": {"src": "conv_test.go:244-265", "gen": {"kind": "cos", "n": 256, "period": 32}, "peak_index": 255},
    "autocorr_normalized": {"src": "conv_test.go:267-280", "a": [1, 2, 3, 4, 5], "zero_lag": 1.0, "tol": 1e-10},
    "correlate_normalized_peak": {"src": "conv_test.go:513-534", "a": [1, 2, 3, 4, 5], "peak": 1.0, "tol": 0.1},
    "commutative": {"src": "conv_test.go:684-700", "a": [1, 2, 3], "b": [4, 5], "tol": 1e-10},
    "errors": [  # sentinel identity (errors.Is)
        {"src": "conv_test.go:71-81", "op": "direct", "a": [], "b": [1, 2], "err": "ErrEmptyInput"},
        {"src": "conv_test.go:71-81", "op": "direct", "a": [1, 2], "b": [], "err": "ErrEmptyKernel"},
        {"src": "conv_test.go:487-492", "op": "correlate_fft", "a": [], "b": [1, 2], "err": "ErrEmptyInput"},
        {"src": "conv_test.go:513-518", "op": "correlate_direct", "a": [], "b": [1, 2], "err": "ErrEmptyInput"},
        {"src": "conv_test.go:656-666", "op": "direct_circular", "a": [], "b": [1, 2], "err": "ErrEmptyInput"},
        {"src": "conv_test.go:656-666", "op": "direct_circular", "a": [1, 2, 3], "b": [1, 2], "err": "ErrLengthMismatch"},
        {"src": "conv_test.go:675-682", "op": "new_overlap_save", "kernel": [0.25, 0.5, 0.25], "fft_size": 100, "err": "ErrInvalidBlockSize"},
        {"src": "conv_test.go:387-449", "op": "process_to_wrong_len", "kernel": [0.25, 0.5, 0.25], "signal_len": 100, "out_len": 5,
         "err": "ErrLengthMismatch"},
    ],
    "streaming_impulse": {  # TestStreamingOverlapSave / TestStreamingOverlapAdd
        "src": "streaming_overlap_save_test.go:9-49, streaming_overlap_add_test.go:9-56", "kernel": [1.0, 0.5, 0.25], "block_size": 4,
        "block1": [1, 0, 0, 0], "block2": [0, 0, 0, 0], "out1": [1.0, 0.5, 0.25, 0], "tol": 1e-10,
    },
    "streaming_impulse_f32": {  # TestStreamingFloat32ImpulseResponse, streaming_test.go:237-268
        "src": "streaming_test.go:237-268", "kernel": [1.0, 0.5, 0.25], "block_size": 8, "out": [1.0, 0.5, 0.25, 0, 0, 0, 0, 0], "tol": 1e-5,
    },
    "partitioned": {
        "latency": {"src": "partitioned_test.go:103-119", "kernel_len": 64, "orders": [4, 5, 6, 7]},
        "vs_streaming_ola": [  # TestPartitionedConvolutionMatchesSOA, partitioned_test.go:121-165
            {"kernel_len": 64, "signal_len": 512, "min": 4, "max": 10, "tol": 1e-7},
            {"kernel_len": 256, "signal_len": 1024, "min": 5, "max": 12, "tol": 1e-7},
            {"kernel_len": 1024, "signal_len": 4096, "min": 6, "max": 13, "tol": 1e-7},
            {"kernel_len": 8192, "signal_len": 16384, "min": 6, "max": 13, "tol": 1e-7},
        ],
        "reset": {"src": "partitioned_test.go:167-197", "kernel_len": 128, "signal_len": 512, "min": 6, "max": 12, "tol": 1e-12},
        "dirac": {"src": "partitioned_test.go:292-321", "signal_len": 256, "min": 4, "max": 12, "tol": 1e-9},
        "kernel_len": {"src": "partitioned_test.go:279-290", "kernel_len": 300, "min": 6, "max": 13},
        "errors": [
            {"src": "partitioned_test.go:200-205", "kernel": [], "min": 6, "max": 12, "err": "ErrEmptyImpulseResponse"},
            {"src": "partitioned_test.go:207-212", "kernel": [1, 2, 3], "min": 0, "max": 12, "err": "ErrInvalidBlockOrder"},
            {"src": "partitioned_test.go:214-219", "kernel": [1, 2, 3], "min": 8, "max": 5, "err": "ErrInvalidBlockOrder"},
            {"src": "partitioned_test.go:221-234", "kernel": [1, 2, 3, 4], "min": 2, "max": 10, "in_len": 10, "out_len": 8,
             "err": "ErrLengthMismatch"},
        ],
        # stage layouts derived from partitionIR (partitioned.go:269-332) and listed in SURVEY.md 3.3:
        # (kernelLen, minOrder, maxOrder) -> [(order, startPos, count)]
        "stage_layouts": [
            {"kernel_len": 1024, "min": 6, "max": 13, "stages": [[6, 0, 2], [7, 128, 1], [8, 256, 3]]},
            {"kernel_len": 96000, "min": 7, "max": 13,
             "stages": [[7, 0, 2], [8, 256, 2], [9, 768, 2], [10, 1792, 2], [11, 3840, 1], [12, 5888, 2], [13, 14080, 10]]},
            {"kernel_len": 288000, "min": 7, "max": 13,
             "stages": [[7, 0, 2], [8, 256, 2], [9, 768, 1], [10, 1280, 2], [11, 3328, 1], [12, 5376, 1], [13, 9472, 34]]},
        ],
    },
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(out, "w") as f:
        json.dump(KATS, f, indent=1)
    print(out)
