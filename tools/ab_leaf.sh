#!/bin/bash
# nested-dissection leaf size (levels vs fill) with the bulk-copy solve kernel
cd $GRAFT_REPO_ROOT
for N in 30 55; do
for leaf in 64 128 256; do
  ADMMB_ND_LEAF=$leaf python bench.py --cube $N --steps 20 --warmup 5 --no-cpu-baseline --no-pairs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; print('N=$N leaf %4d value %7.1f  local %.3f rhs %.3f solve %.3f  step %.4f  levels %d factor %.2f GB setup %.1f s' % ($leaf, d['value'], p['local'], p['rhs'], p['solve'], p['step'], d['setup']['levels'], d['setup']['factor_bytes']/1e9, d['setup']['seconds']))"
done; done
