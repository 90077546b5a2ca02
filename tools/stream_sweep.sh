#!/bin/bash
# worker streams x group size sweep on the bench workload (ADSP_GROUP_PAIRS forces the pairs per launch)
cd "$(dirname "$0")/.."
for s in ${STREAMS:-2 3 4 6 8}; do for g in ${GROUPS_:-1 2 3}; do
  ADSP_STREAMS=$s ADSP_GROUP_PAIRS=$g LABEL="streams=$s pairs/launch=$g" python tests/tools/bench_one.py | cut -c1-100
done; done
