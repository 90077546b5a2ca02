#!/bin/bash
# Steady-state per-kernel throughput: one huge launch per kernel (no launch tails), scratch aliased onto a few
# L2-resident slots (results wrong, timing valid).  Needs libvar_dbg.so (-DADSP_PHASE_DEBUG).
cd "$(dirname "$0")/.."
export ADSP_LIB_PATH=$PWD/algo_dsp_b200/${LIBV:-libvar_dbg.so}
for pf in 1 0; do
  ADSP_PF=$pf LABEL="pf=$pf default schedule" python tools/ktimes.py
  ADSP_PF=$pf ADSP_STREAMS=1 ADSP_SCRATCH_MB=100000 ADSP_SCRATCH_ALIAS=3 LABEL="pf=$pf steady alias=3" python tools/ktimes.py
done
for c in 1 2; do
  ADSP_PF_GRID_CTAS=$c ADSP_STREAMS=1 ADSP_SCRATCH_MB=100000 ADSP_SCRATCH_ALIAS=3 LABEL="pf steady alias=3 ctas/sm=$c" python tools/ktimes.py
done
