#!/bin/bash
# A/B build variants on the bench workload: tools/variant_sweep.sh lib1.so lib2.so ...
cd "$(dirname "$0")/.."
for lib in "$@"; do
  ADSP_LIB_PATH=$PWD/algo_dsp_b200/$lib BCHECK=1 LABEL="$lib" python tests/tools/bench_one.py | cut -c1-180
done
