#!/bin/bash
cd "$(dirname "$0")/.."
for st in 2 3 4 6 8; do for mb in 40 80 160; do ADSP_STREAMS=$st ADSP_SCRATCH_MB=$mb LABEL="streams=$st scratch=$mb" python tools/bench_one.py | cut -c1-75; done; done
