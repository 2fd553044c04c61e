#!/bin/bash
for cfg in ${CFGS:-"0 16 0" "3 16 1" "3 16 0" "3 16 2" "3 16 4" "3 8 0" "3 32 0"}; do
  set -- $cfg
  ADMMB_ND_LEAF=${LEAF:-32} ADMMB_SOLVE_MODE=$1 ADMMB_SOLVE_UNROLL=$2 ADMMB_SOLVE_SPLIT=$3 python bench.py --cube ${CUBE:-55} --steps 10 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; print('mode=$1 unroll=$2 split=$3 value %7.1f local %.3f rhs %.3f solve %.3f' % (d['value'], p['local'], p['rhs'], p['solve']))"
done
