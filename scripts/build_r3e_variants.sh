#!/bin/bash
# r3e: exit threshold of the Ω solve with the error extrapolation on (and off, for the contrast)
set -e
rm -rf build/variants build/vobj
v() { bash scripts/build_variant.sh "$@" | tail -1; }
v q_t3e5 negative_tracers -DOBM_CC_TOL=3e-5 &
v q_t5e5 negative_tracers -DOBM_CC_TOL=5e-5 &
v q_t1e4 negative_tracers -DOBM_CC_TOL=1e-4 &
v q_noextrap negative_tracers -DOBM_CC_EXTRAP=0 &
wait
ls build/variants
