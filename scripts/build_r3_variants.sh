#!/bin/bash
# r3: the library variants timed by scripts/gpu_r3a.sh (run here; the .so files travel to the GPU box).
# p* = stage prologue (scaling + Ω: csrc/negative_tracers.cu with csrc/carbon_chemistry.cuh), t* = PISCES tendency kernel.
set -e
rm -rf build/variants build/vobj
OLD="-DOBM_CC_EXP=0 -DOBM_CC_LOG=0 -DOBM_CC_INIT=0 -DOBM_CC_TOL=1e-7 -DOBM_CC_POLYEXP=0"
v() { bash scripts/build_variant.sh "$@" | tail -1; }
v p0_old negative_tracers $OLD -DOBM_SN_MIN_BLOCKS=8 &            # r02 solve: library exp / log, plain quadratic start, 1e-7 exit
v p1_exp negative_tracers -DOBM_CC_EXP=2 -DOBM_CC_LOG=1 -DOBM_CC_INIT=0 -DOBM_CC_TOL=1e-7 -DOBM_CC_POLYEXP=0 -DOBM_SN_MIN_BLOCKS=8 &   # + lean exp / log only
v p2_newton negative_tracers -DOBM_CC_EXP=0 -DOBM_CC_LOG=0 -DOBM_SN_MIN_BLOCKS=8 &   # + refined start, 2e-6 exit, polynomial step only
v p3_all_b8 negative_tracers -DOBM_SN_MIN_BLOCKS=8 &               # everything, 64 registers (48 B spill)
wait
v p4_all_b7 negative_tracers -DOBM_SN_MIN_BLOCKS=7 &               # everything, 72 registers (32 B spill)
v p5_all_b6 negative_tracers -DOBM_SN_MIN_BLOCKS=6 &               # everything, 80 registers, no spill (the default build)
v p6_all_init1 negative_tracers -DOBM_SN_MIN_BLOCKS=6 -DOBM_CC_INIT=1 -DOBM_CC_TOL=1e-5 &
v p7_all_b5 negative_tracers -DOBM_SN_MIN_BLOCKS=5 &
wait
v t1_exp2 pisces_tendencies -DOBM_PISCES_EXP=2 &
v t2_b4 pisces_tendencies -DOBM_PISCES_MIN_BLOCKS=4 &
v t3_exp2_b4 pisces_tendencies -DOBM_PISCES_EXP=2 -DOBM_PISCES_MIN_BLOCKS=4 &
v t4_roll pisces_tendencies -DOBM_PISCES_ROLL=1 &
wait
v t5_roll_b4 pisces_tendencies -DOBM_PISCES_ROLL=1 -DOBM_PISCES_MIN_BLOCKS=4 &
v t6_b2 pisces_tendencies -DOBM_PISCES_MIN_BLOCKS=2 &
v t7_roll_exp2 pisces_tendencies -DOBM_PISCES_ROLL=1 -DOBM_PISCES_EXP=2 &
wait
ls build/variants
