#!/bin/bash
# streams x in-flight scratch budget on three shapes (mixed-radix bench shape, power-of-two bench shape, 2^20 long-signal shape)
cd "$(dirname "$0")/.."
for sm in "2 40" "2 56" "3 40" "3 56" "3 79" "4 40"; do set -- $sm
  export ADSP_STREAMS=$1 ADSP_SCRATCH_MB=$2
  LABEL="odd   streams=$1 mb=$2" python tests/tools/bench_one.py | cut -c1-100
  ADSP_NO_ODD=1 LABEL="pow2  streams=$1 mb=$2" python tests/tools/bench_one.py | cut -c1-100
  BK=288000 BN=4194304 BCH=16 LABEL="2^20  streams=$1 mb=$2" python tests/tools/bench_one.py | cut -c1-100
done
