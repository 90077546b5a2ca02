#!/bin/bash
# TMA L2-promotion sweep for the persistent column kernels (see tools/mrp_ab.sh)
cd "$(dirname "$0")/.."
ADSP_MRP=0 LABEL="MRP=0" python tests/tools/bench_one.py
for p in 0 1 2 3; do
  ADSP_TMA_L2PROMO=$p BCHECK=1 LABEL="MRP=1 promo=$p" python tests/tools/bench_one.py
  ADSP_TMA_L2PROMO=$p ADSP_STREAMS=1 ADSP_GROUP_PAIRS=100000 LABEL="MRP=1 promo=$p exclusive" python tools/ktimes.py
done
ADSP_TMA_L2PROMO=2 ADSP_MRP_GRID_CTAS=1 LABEL="MRP=1 promo=2 grid 1 CTA/SM" python tests/tools/bench_one.py
ADSP_TMA_L2PROMO=2 ADSP_MRP_GRID_CTAS=2 LABEL="MRP=1 promo=2 grid 2 CTA/SM" python tests/tools/bench_one.py
