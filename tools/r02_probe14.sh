#!/bin/bash
mkdir -p gpurun_out
for v in parsec_b200/libvariant_m12_cu16_u4.so parsec_b200/libvariant_m20_cu6_u4.so parsec_b200/libvariant_m16_cu8_u4.so parsec_b200/libvariant_m12_cu16_u8.so; do
  echo "== $v"
  PB2_LIB_PATH=$PWD/$v timeout 300 python tools/sweep_hbm.py 0,0,0 2>&1 | tail -1
  PB2_LIB_PATH=$PWD/$v timeout 600 python bench.py --steps 30 --e2e-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.3e ms %.4f' % (d['value'], d['ms_per_step']))
print('e2e_standalone', d.get('e2e_standalone',{}).get('ms_per_step'))
s=d.get('secondary',{})
for k,v in s.items(): print(k, json.dumps(v)[:600])
"
done
