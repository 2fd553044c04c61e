#!/bin/bash
# A/B of local-kernel build variants (variants/*.so): same workload, same frames, steady-state regime
W=${W:-20}; K=${K:-5}
for lib in "" $(ls /root/repo/variants/*.so 2>/dev/null); do
  ADMMB_LIB=$lib python bench.py --cube 55 --steps $K --warmup $W --no-cpu-baseline --no-pairs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; print('%-40s value %7.1f  local %.3f rhs %.3f solve %.3f' % ('$lib'.split('/')[-1] or 'default', d['value'], p['local'], p['rhs'], p['solve']))"
done
