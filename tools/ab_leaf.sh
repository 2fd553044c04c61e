#!/bin/bash
# A/B of the dissection leaf size (levels vs factor bytes) with the prefetching solve kernel
for leaf in ${LEAVES:-32 48 64 96 128 256}; do
  ADMMB_ND_LEAF=$leaf ADMMB_SOLVE_MODE=${MODE:-3} python bench.py --cube ${CUBE:-55} --steps 10 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; s=d['setup']; print('leaf=$leaf value %7.1f local %.3f rhs %.3f solve %.3f  nnzL %.1fM factor %.2fs frac %.2f' % (d['value'], p['local'], p['rhs'], p['solve'], s['nnz_L']/1e6, s['factor_seconds'], d['roofline']['frac']))"
done
