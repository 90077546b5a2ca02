#!/bin/bash
# schedule sweep for the persistent TMA-fed column kernels: streams x pairs per group x tiles per CTA (see tools/mrp_ab.sh)
cd "$(dirname "$0")/.."
ADSP_MRP=0 LABEL="MRP=0 default" python tests/tools/bench_one.py | cut -c1-90
for m in 1 2; do
for st in 2 4; do for gp in 1 2 3; do for k in 1 2 4; do
  ADSP_MRP=$m ADSP_STREAMS=$st ADSP_GROUP_PAIRS=$gp ADSP_MRP_TILES_PER_CTA=$k LABEL="MRP=$m streams=$st pairs/group=$gp tiles/cta=$k" python tests/tools/bench_one.py | cut -c1-90
done; done; done; done
for st in 2 4; do for gp in 2 3; do
  ADSP_MRP=0 ADSP_STREAMS=$st ADSP_GROUP_PAIRS=$gp LABEL="MRP=0 streams=$st pairs/group=$gp" python tests/tools/bench_one.py | cut -c1-90
done; done
