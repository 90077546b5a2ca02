#!/bin/bash
# A/B of the persistent TMA-fed column kernels (ADSP_MRP=1) against the one-tile-per-CTA kernels (ADSP_MRP=0) on the
# bench workload: whole step in the default schedule, and each kernel alone on the GPU (exclusive).
cd "$(dirname "$0")/.."
for m in 0 1; do
  ADSP_MRP=$m BCHECK=1 LABEL="MRP=$m default schedule" python tests/tools/bench_one.py
  ADSP_MRP=$m LABEL="MRP=$m default schedule" python tools/ktimes.py
  ADSP_MRP=$m ADSP_STREAMS=1 ADSP_GROUP_PAIRS=100000 LABEL="MRP=$m exclusive (one launch per kernel)" python tools/ktimes.py
done
for c in 1 2; do
  ADSP_MRP=1 ADSP_MRP_GRID_CTAS=$c LABEL="MRP=1 grid ${c} CTA/SM" python tests/tools/bench_one.py
done
