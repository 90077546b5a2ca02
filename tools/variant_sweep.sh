#!/bin/bash
# A/B the CTA-shape build variants on the bench workload
cd "$(dirname "$0")/.."
for lib in libalgodsp_cuda.so libvar_c128.so libvar_r128.so libvar_rc128.so; do
  for n2 in 4096 2048 1024; do
    ADSP_LIB_PATH=$PWD/algo_dsp_b200/$lib ADSP_FFT_N2=$n2 BCHECK=1 LABEL="$lib N2=$n2 mixed" python tools/bench_one.py
  done
done
