#!/bin/bash
# Steady-state per-kernel throughput: one huge launch per kernel (no launch tails), scratch aliased onto a few
# L2-resident slots (results wrong, timing valid).  Needs libvar_dbg.so (-DADSP_PHASE_DEBUG).
cd "$(dirname "$0")/.."
export ADSP_LIB_PATH=$PWD/algo_dsp_b200/${LIBV:-libvar_dbg.so}
LABEL="default schedule" python tools/ktimes.py
LABEL="default schedule" python tests/tools/bench_one.py | cut -c1-100
ADSP_STREAMS=1 ADSP_GROUP_PAIRS=100000 ADSP_SCRATCH_ALIAS=3 LABEL="steady alias=3" python tools/ktimes.py
ADSP_STREAMS=1 ADSP_GROUP_PAIRS=100000 ADSP_SCRATCH_ALIAS=3 LABEL="steady alias=3" python tests/tools/bench_one.py | cut -c1-100
