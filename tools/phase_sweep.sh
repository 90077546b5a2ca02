#!/bin/bash
cd "$(dirname "$0")/.."
export ADSP_LIB_PATH=$PWD/algo_dsp_b200/libvar_dbg.so
for f in 0 128 256 384 127 255 383 511; do ADSP_PHASE_SKIP=$f LABEL="skip=$f" python tests/tools/bench_one.py | cut -c1-75; done
