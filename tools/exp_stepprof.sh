#!/bin/bash
# usage: tools/exp_stepprof.sh out_prefix [modes...]  (needs the _prof library: QP_PROFILE_PHASES=1 QP_FAST_BUILD=1 QP_LIB_SUFFIX=_prof)
out=$1; shift
for mode in ${@:-old}; do
  QP_LIB_SUFFIX=_prof QP_AHEAD_MODE=$mode timeout 200 python tools/phase_profile_step.py 6 >> gpurun_out/${out}.log 2>&1
done
