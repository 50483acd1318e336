#!/bin/bash
# r3c: variants of the stage prologue (cp.async staging on / off × resident blocks) and of the PAR scan's exp, timed by scripts/gpu_r3c.sh.
set -e
rm -rf build/variants build/vobj
v() { bash scripts/build_variant.sh "$@" | tail -1; }
v a6_async_b6 negative_tracers -DOBM_SN_MIN_BLOCKS=6 &
v a7_async_b7 negative_tracers -DOBM_SN_MIN_BLOCKS=7 &
v a8_async_b8 negative_tracers -DOBM_SN_MIN_BLOCKS=8 &
v n8_direct_b8 negative_tracers -DOBM_SN_ASYNC=0 -DOBM_SN_MIN_BLOCKS=8 &
wait
v n7_direct_b7 negative_tracers -DOBM_SN_ASYNC=0 -DOBM_SN_MIN_BLOCKS=7 &
v a8_nolevel negative_tracers -DOBM_SN_MIN_BLOCKS=8 -DOBM_CC_LEVEL=0 &
v a8_tol2e6 negative_tracers -DOBM_SN_MIN_BLOCKS=8 -DOBM_CC_TOL=2e-6 &
v l3_exp_clamped light -DOBM_LIGHT_EXP=3 &
wait
ls build/variants
