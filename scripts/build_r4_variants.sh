#!/bin/bash
# r4: the FP32 pre-solve of the carbonate Newton iteration (OBM_CC_F32PRE) on / off and its exit threshold
set -e
rm -rf build/variants build/vobj
v() { bash scripts/build_variant.sh "$@" | tail -1; }
v f0_nopre negative_tracers -DOBM_CC_F32PRE=0 &
v f_t3e5 negative_tracers -DOBM_CC_TOL0=3e-5 &
v f_b7 negative_tracers -DOBM_SN_MIN_BLOCKS=7 &
v f_b6 negative_tracers -DOBM_SN_MIN_BLOCKS=6 &
wait
ls build/variants
# the PAR scan with four levels per lane (OBM_PAR_SCAN4) at 5 / 4 / 3 resident blocks
v s4 light -DOBM_PAR_SCAN4=1 &
v s4_b4 light -DOBM_PAR_SCAN4=1 -DOBM_PAR_DIAG_BLOCKS=4 &
v s4_b3 light -DOBM_PAR_SCAN4=1 -DOBM_PAR_DIAG_BLOCKS=3 &
wait
ls build/variants
