"""ctypes loader and prototypes for libalgodsp_cuda.so (include/algodsp_cuda.h).

The product path has no CPU fallback: if the library is missing, or no CUDA device is
visible, calls fail loudly (ConvError / RuntimeError)."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADSP_LIB_PATH") or os.path.join(HERE, "libalgodsp_cuda.so")   # override: experimental build variants

c_i64 = C.c_int64
c_vp = C.c_void_p
c_dp = C.POINTER(C.c_double)

# status codes (adsp_status)
OK, ERR_EMPTY_INPUT, ERR_EMPTY_KERNEL, ERR_LENGTH_MISMATCH, ERR_INVALID_BLOCK_SIZE, ERR_INVALID_BLOCK_ORDER, \
    ERR_EMPTY_IR, ERR_STAGE_INDEX, ERR_INVALID_ARG, ERR_CUDA, ERR_OOM, ERR_DIVISION_BY_ZERO = range(12)
F64, F32 = 0, 1

# name -> (restype, argtypes); every symbol the header declares
PROTOTYPES = {
    "adsp_version": (C.c_char_p, []),
    "adsp_status_string": (C.c_char_p, [C.c_int]),
    "adsp_last_error": (C.c_size_t, [C.c_char_p, C.c_size_t]),
    "adsp_device_count": (C.c_int, []),
    "adsp_ctx_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "adsp_ctx_destroy": (None, [c_vp]),
    "adsp_ctx_sync": (C.c_int, [c_vp]),
    "adsp_ctx_launch_count": (C.c_uint64, [c_vp]),
    "adsp_ctx_stream": (c_vp, [c_vp]),
    "adsp_ctx_kernel_timing": (None, [c_vp, C.c_int]),
    "adsp_ctx_kernel_time": (C.c_int, [c_vp, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64), C.c_int]),
    "adsp_ctx_host_profile": (None, [c_vp, C.c_int]),
    "adsp_ctx_host_profile_get": (C.c_int, [c_vp, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "adsp_host_ptr_is_pinned": (C.c_int, [c_vp]),
    "adsp_ctx_stage_threads": (C.c_int, [c_vp]),
    "adsp_ctx_measure_pipes": (C.c_int, [c_vp, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "adsp_ctx_copy_ceiling": (C.c_int, [c_vp, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "adsp_gen_uniform_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, c_i64, C.c_int]),
    "adsp_gen_white_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, c_i64, c_i64, c_i64, C.c_int]),
    "adsp_gen_pink_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, c_i64, c_i64, c_i64, C.c_int]),
    "adsp_gen_decaying_ir_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, c_i64, c_i64, C.c_int]),
    "adsp_gen_linear_sweep_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
    "adsp_gen_log_sweep_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int]),
    "adsp_gen_delay_mix_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, C.c_double, c_i64, c_i64, c_i64, c_i64, c_vp, C.c_int]),
    "adsp_normalize_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, c_vp, c_i64, C.c_int]),
    "adsp_remove_dc_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, C.c_int]),
    "adsp_gen_uniform_host": (None, [c_vp, c_i64, c_i64, c_i64]),
    "adsp_gen_white_host": (None, [c_vp, c_i64, C.c_double, c_i64, c_i64]),
    "adsp_gen_pink_host": (None, [c_vp, c_i64, C.c_double, c_i64, c_i64]),
    "adsp_gen_decaying_ir_host": (None, [c_vp, c_i64, C.c_double, c_i64]),
    "adsp_gen_linear_sweep_host": (None, [c_vp, c_i64, c_i64, c_i64, C.c_double, C.c_double, C.c_double, C.c_double]),
    "adsp_gen_log_sweep_host": (None, [c_vp, c_i64, c_i64, c_i64, C.c_double, C.c_double, C.c_double, C.c_double]),
    "adsp_gen_delay_host": (c_i64, [c_i64, c_i64, c_i64]),
    "adsp_ir_schroeder_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64]),
    "adsp_ir_find_impulse_start_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, c_vp]),
    "adsp_ir_find_peak_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp]),
    "adsp_ir_schroeder": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "adsp_ir_find_impulse_start": (C.c_int, [c_vp, c_vp, c_i64, C.c_double, C.POINTER(c_i64)]),
    "adsp_logsweep_samples": (c_i64, [C.c_double, C.c_double]),
    "adsp_logsweep_generate_device": (C.c_int, [c_vp, c_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "adsp_logsweep_inverse_filter_device": (C.c_int, [c_vp, c_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "adsp_logsweep_generate_host": (C.c_int, [c_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "adsp_logsweep_inverse_filter_host": (C.c_int, [c_vp, C.c_double, C.c_double, C.c_double, C.c_double]),
    "adsp_logsweep_deconvolve_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, C.c_double, C.c_double, C.c_double, C.c_double, c_vp, c_i64]),
    "adsp_logsweep_deconvolve": (C.c_int, [c_vp, c_vp, c_i64, C.c_double, C.c_double, C.c_double, C.c_double, c_vp, c_i64]),
    "adsp_fir_create": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, C.POINTER(c_vp)]),
    "adsp_fir_process_block": (C.c_int, [c_vp, c_vp, c_i64, c_i64]),
    "adsp_fir_process_block_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64]),
    "adsp_fir_order": (c_i64, [c_vp]),
    "adsp_fir_reset": (None, [c_vp]),
    "adsp_fir_destroy": (None, [c_vp]),
    "adsp_resample_approximate_ratio": (None, [C.c_double, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "adsp_resampler_create": (C.c_int, [c_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(c_vp)]),
    "adsp_resampler_create_for_rates": (C.c_int, [c_vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "adsp_resampler_ratio": (None, [c_vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "adsp_resampler_taps_per_phase": (C.c_int, [c_vp]),
    "adsp_resampler_prototype": (c_i64, [c_vp, c_vp, c_i64]),
    "adsp_resampler_predict_output_len": (c_i64, [c_vp, c_i64]),
    "adsp_resampler_process": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, C.POINTER(c_i64)]),
    "adsp_resampler_process_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, C.POINTER(c_i64)]),
    "adsp_resampler_reset": (None, [c_vp]),
    "adsp_resampler_destroy": (None, [c_vp]),
    "adsp_host_register": (C.c_int, [c_vp, C.c_size_t]),
    "adsp_host_unregister": (C.c_int, [c_vp]),
    "adsp_host_alloc_pinned": (C.c_int, [C.c_size_t, C.POINTER(c_vp)]),
    "adsp_host_free_pinned": (None, [c_vp]),
    "adsp_device_alloc": (C.c_int, [c_vp, C.c_size_t, C.POINTER(c_vp)]),
    "adsp_device_free": (None, [c_vp, c_vp]),
    "adsp_memcpy_h2d": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "adsp_memcpy_d2h": (C.c_int, [c_vp, c_vp, c_vp, C.c_size_t]),
    "adsp_next_pow2": (c_i64, [c_i64]),
    "adsp_is_pow2": (C.c_int, [c_i64]),
    "adsp_ols_sizes": (C.c_int, [c_i64, c_i64, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "adsp_ola_sizes": (C.c_int, [c_i64, c_i64, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "adsp_trim_mode": (None, [c_i64, c_i64, C.c_int, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "adsp_lag_from_index": (c_i64, [c_i64, c_i64]),
    "adsp_index_from_lag": (c_i64, [c_i64, c_i64]),
    "adsp_direct": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_direct_circular": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_convolve": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_overlap_add_convolve": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_overlap_save_convolve": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_correlate": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_correlate_direct": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_correlate_fft": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_correlate_normalized": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_autocorrelate_normalized": (C.c_int, [c_vp, c_vp, c_i64, c_vp]),
    "adsp_find_peak": (C.c_int, [c_vp, c_vp, c_i64, C.POINTER(c_i64), C.POINTER(C.c_double)]),
    "adsp_direct_f32": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_convolve_f32": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_correlate_f32": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp]),
    "adsp_direct_batch": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64]),
    "adsp_correlate_batch": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp]),
    "adsp_direct_batch_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, C.c_int]),
    "adsp_correlate_batch_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, C.c_int]),
    "adsp_overlap_save_create": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, C.POINTER(c_vp)]),
    "adsp_overlap_add_create": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, C.POINTER(c_vp)]),
    "adsp_plan_destroy": (None, [c_vp]),
    "adsp_plan_reset": (None, [c_vp]),
    "adsp_plan_kernel_len": (c_i64, [c_vp]),
    "adsp_plan_fft_size": (c_i64, [c_vp]),
    "adsp_plan_step_size": (c_i64, [c_vp]),
    "adsp_plan_block_size": (c_i64, [c_vp]),
    "adsp_plan_internal_geometry": (None, [c_vp] + [C.POINTER(c_i64)] * 5),
    "adsp_plan_describe_cover": (C.c_int, [c_vp, c_i64, C.POINTER(c_i64), C.c_int]),
    "adsp_plan_process": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64]),
    "adsp_plan_process_batch": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64]),
    "adsp_plan_process_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64]),
    "adsp_plan_sync": (C.c_int, [c_vp]),
    "adsp_partitioned_create": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "adsp_shard_channel_range": (None, [c_i64, C.c_int, C.c_int, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "adsp_shard_time": (None, [c_i64, c_i64, C.c_int, C.c_int] + [C.POINTER(c_i64)] * 5),
    "adsp_plans_process_batch": (C.c_int, [C.POINTER(c_vp), C.c_int, c_vp, c_i64, c_i64, c_i64, c_vp, c_i64]),
    "adsp_plans_process_long": (C.c_int, [C.POINTER(c_vp), C.c_int, c_vp, c_i64, c_vp, c_i64]),
    "adsp_deconv_out_len": (c_i64, [c_i64, c_i64]),
    "adsp_deconvolve": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64, C.c_int, C.c_double, C.c_double, C.c_double, c_vp, c_i64]),
    "adsp_inverse_filter": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_double, c_vp]),
    "adsp_snr": (C.c_double, [c_vp, c_i64, c_vp, c_i64]),
    "adsp_deconvolve_batch_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, C.c_double, c_vp, c_i64]),
    "adsp_partitioned_process_block": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64]),
    "adsp_partitioned_create_batch": (C.c_int, [c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "adsp_partitioned_process_block_batch": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64]),
    "adsp_partitioned_process_block_batch_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64]),
    "adsp_partitioned_set_wet_dry": (C.c_int, [c_vp, C.c_double, C.c_double]),
    "adsp_partitioned_process_in_place_batch": (C.c_int, [c_vp, c_vp, c_i64, c_i64]),
    "adsp_partitioned_process_in_place_batch_device": (C.c_int, [c_vp, c_vp, c_i64, c_i64]),
    "adsp_partitioned_channels": (C.c_int, [c_vp]),
    "adsp_partitioned_plan_layout": (C.c_int, [c_i64, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(c_i64), C.c_int]),
    "adsp_partitioned_internal_stage_count": (C.c_int, [c_vp]),
    "adsp_partitioned_internal_stage_info": (C.c_int, [c_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(c_i64)]),
    "adsp_partitioned_latency": (C.c_int, [c_vp]),
    "adsp_partitioned_stage_count": (C.c_int, [c_vp]),
    "adsp_partitioned_stage_info": (C.c_int, [c_vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "adsp_streaming_create": (C.c_int, [c_vp, c_vp, c_i64, c_i64, C.c_int, C.c_int, C.POINTER(c_vp)]),
    "adsp_streaming_process_block": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_i64]),
}

_lib = None


def load():
    """Load libalgodsp_cuda.so (building it is the job of algo_dsp_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m algo_dsp_b200.build` "
                "(there is no CPU fallback for the dsp/conv hot path)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    load().adsp_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")
