#!/bin/bash
out=$1
for sk in off "down:7:800,o:19:800,ug:16:940,qkv:18:800"; do
  echo "##### QP_SKEW=$sk" >> gpurun_out/${out}.log
  QP_SKEW=$sk QP_LIB_SUFFIX=_prof timeout 200 python tools/phase_profile_step.py 6 >> gpurun_out/${out}.log 2>&1
done
