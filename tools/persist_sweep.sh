#!/bin/bash
cd "$(dirname "$0")/.."
for k in 1 2 4 8; do for mb in 80 160 320; do for st in 2 3; do
ADSP_TILES_PER_CTA=$k ADSP_SCRATCH_MB=$mb ADSP_STREAMS=$st LABEL="tiles/cta=$k scratch=$mb streams=$st" python tests/tools/bench_one.py | cut -c1-100
done; done; done
