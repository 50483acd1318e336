#!/bin/bash
# r3g: the 64-entry-table exp (exp_table, obm_common.cuh) in the PAR scans, the carbonate solve and the PISCES fast pass
set -e
rm -rf build/variants build/vobj
v() { bash scripts/build_variant.sh "$@" | tail -1; }
v l4_table light -DOBM_LIGHT_EXP=4 &
v c3_table negative_tracers -DOBM_CC_EXP=3 &
v e3_table pisces_tendencies -DOBM_PISCES_EXP=3 &
v e2_horner pisces_tendencies -DOBM_PISCES_EXP=2 &
wait
ls build/variants
