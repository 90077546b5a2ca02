#!/bin/bash
# A/B of the prefetching persistent kernels against the one-tile-per-CTA kernels, plus stream/group sweeps
cd "$(dirname "$0")/.."
BCHECK=1 LABEL="pf default" python tests/tools/bench_one.py
ADSP_PF=0 BCHECK=1 LABEL="pf off" python tests/tools/bench_one.py
for s in 1 2 3 4; do for mb in 40 79 120; do
  ADSP_STREAMS=$s ADSP_SCRATCH_MB=$mb LABEL="pf streams=$s mb=$mb" python tests/tools/bench_one.py
done; done
LABEL="pf default" python tools/ktimes.py
export ADSP_LIB_PATH=$PWD/algo_dsp_b200/libvar_dbg.so
ADSP_STREAMS=1 ADSP_SCRATCH_MB=100000 ADSP_SCRATCH_ALIAS=3 LABEL="pf steady alias=3" python tools/ktimes.py
ADSP_PF=0 ADSP_STREAMS=1 ADSP_SCRATCH_MB=100000 ADSP_SCRATCH_ALIAS=3 LABEL="old steady alias=3" python tools/ktimes.py
