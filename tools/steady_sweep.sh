#!/bin/bash
# Steady-state per-kernel throughput: one huge launch per kernel (no launch tails), scratch aliased onto a few
# L2-resident slots (results wrong, timing valid).  Needs libvar_dbg.so (-DADSP_PHASE_DEBUG).
cd "$(dirname "$0")/.."
export ADSP_LIB_PATH=$PWD/algo_dsp_b200/${LIBV:-libvar_dbg.so}
for odd in 0 1; do
  ADSP_NO_ODD=$odd LABEL="no_odd=$odd default schedule" python tools/ktimes.py
  ADSP_NO_ODD=$odd ADSP_STREAMS=1 ADSP_SCRATCH_MB=100000 ADSP_SCRATCH_ALIAS=3 LABEL="no_odd=$odd steady alias=3" python tools/ktimes.py
done
