#!/bin/bash
# usage: tools/exp_skew.sh out_prefix lib_suffix   -- 8B decode step under several work-split settings (QP_SKEW, decode.py)
out=$1; sfx=$2
for sk in off "down:7:890" "down:7:850,o:19:900" "down:7:890,o:19:900,ug:10:960,qkv:18:900" "down:7:850,o:19:850,ug:12:950,qkv:18:860" "down:7:800,o:19:800,ug:16:940,qkv:18:800" off; do
  echo "=== QP_SKEW=$sk" >> gpurun_out/${out}.log
  QP_SKEW=$sk QP_LIB_SUFFIX=$sfx timeout 300 python bench.py --no-cpu-baseline --no-tp-extra --steps 64 --warmup 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['value'], 'tok/s', d['ms_per_step'], 'ms', 'e2e', d['e2e']['value'])" >> gpurun_out/${out}.log 2>&1
done
