#!/bin/bash
out=$1
REPS=1 bash tools/exp_variants2.sh ${out}_gemv _k
bash tools/exp_skew.sh ${out}_skew _k
QP_SKEW=off QP_LIB_SUFFIX=_prof timeout 200 python tools/phase_profile_step.py 6 > gpurun_out/${out}_prof.log 2>&1
