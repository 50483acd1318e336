#!/bin/bash
# r3: the library variants timed by scripts/gpu_r3a.sh (run here; the .so files travel to the GPU box).
# p* = stage prologue (scaling + Ω: csrc/negative_tracers.cu with csrc/carbon_chemistry.cuh), t* = PISCES tendency kernel.
set -e
rm -rf build/variants build/vobj
OLD="-DOBM_CC_EXP=0 -DOBM_CC_LOG=0 -DOBM_CC_INIT=0 -DOBM_CC_TOL=1e-7 -DOBM_CC_POLYEXP=0"
v() { bash scripts/build_variant.sh "$@" | tail -1; }
NEWTON="-DOBM_CC_INIT=1 -DOBM_CC_TOL=1e-5"
v p0_old negative_tracers $OLD -DOBM_SN_MIN_BLOCKS=8 &            # r02 solve: library exp / log, plain quadratic start, 1e-7 exit
v p1_lib_newton_b8 negative_tracers -DOBM_CC_EXP=0 -DOBM_CC_LOG=0 $NEWTON -DOBM_SN_MIN_BLOCKS=8 &   # library exp / log + one refinement, 1e-5 exit, polynomial step
v p2_batch_newton_b8 negative_tracers $NEWTON -DOBM_SN_MIN_BLOCKS=8 &   # lean exp / log, branch-free batch
v p3_batch_newton_b6 negative_tracers $NEWTON -DOBM_SN_MIN_BLOCKS=6 &
wait
v p4_batch_init2_b8 negative_tracers -DOBM_SN_MIN_BLOCKS=8 &        # two refinements, 2e-6 exit
v p5_nobatch_newton_b8 negative_tracers $NEWTON -DOBM_CC_BATCH=0 -DOBM_SN_MIN_BLOCKS=8 &
v p6_batch_newton_b7 negative_tracers $NEWTON -DOBM_SN_MIN_BLOCKS=7 &
v p7_liblog_batchexp_b8 negative_tracers $NEWTON -DOBM_CC_LOG=0 -DOBM_SN_MIN_BLOCKS=8 &
wait
v p8_batch_newton_b8_nolevel negative_tracers $NEWTON -DOBM_SN_MIN_BLOCKS=8 -DOBM_CC_LEVEL=0 &   # without the per-level tables
v p9_lib_newton_b8_nolevel negative_tracers -DOBM_CC_EXP=0 -DOBM_CC_LOG=0 $NEWTON -DOBM_SN_MIN_BLOCKS=8 -DOBM_CC_LEVEL=0 &
wait
v l2_exp_horner light -DOBM_LIGHT_EXP=2 &
v l3_exp_clamped light -DOBM_LIGHT_EXP=3 &
wait
ls build/variants
