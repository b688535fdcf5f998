#!/bin/bash
# usage: tools/exp_skew.sh out_prefix lib_suffix   -- 8B decode step under several work-split settings (QP_SKEW, decode.py)
out=$1; sfx=$2
run() {
  echo "=== $1 QP_SKEW=$2" >> gpurun_out/${out}.log
  env $1 QP_SKEW=$2 QP_LIB_SUFFIX=$sfx timeout 300 python bench.py --no-cpu-baseline --no-tp-extra --steps 64 --warmup 8 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print(d['value'], 'tok/s', d['ms_per_step'], 'ms', 'e2e', d['e2e']['value'])" >> gpurun_out/${out}.log 2>&1
}
run QP_SPLIT_FLAT=1 off
run X=1 off
run X=1 "down:7:890,o:19:900,ug:10:960,qkv:18:900"
run X=1 "down:7:850,o:19:800,ug:16:930,qkv:18:850"
run X=1 "down:7:850,o:24:750,ug:24:920,qkv:24:850"
run X=1 "down:10:800,o:30:700,ug:30:900,qkv:30:800"
run QP_SPLIT_FLAT=1 off
run X=1 "down:7:850,o:19:800,ug:16:930,qkv:18:850"
