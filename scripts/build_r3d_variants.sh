#!/bin/bash
# r3d: (t) the software-pipelined persistent launch of the PISCES tendency kernel (OBM_PISCES_PIPE) against the direct one;
# (q) the stage prologue: level tables on (copies issued before the tables' barrier) / off, direct loads, exit thresholds.
set -e
rm -rf build/variants build/vobj
v() { bash scripts/build_variant.sh "$@" | tail -1; }
v t_pipe3 pisces_tendencies -DOBM_PISCES_PIPE=1 &
v t_pipe4 pisces_tendencies -DOBM_PISCES_PIPE=1 -DOBM_PISCES_MIN_BLOCKS=4 &
v q_lvl negative_tracers -DOBM_SN_LEVEL=1 &
v q_direct negative_tracers -DOBM_SN_ASYNC=0 &
wait
v q_t5 negative_tracers -DOBM_CC_TOL=5e-6 &
v q_t3 negative_tracers -DOBM_CC_TOL=3e-6 &
v q_t2 negative_tracers -DOBM_CC_TOL=2e-6 &
v q_b7 negative_tracers -DOBM_SN_MIN_BLOCKS=7 &
wait
ls build/variants
