#!/bin/bash
python bench.py --cube 55 --steps 10 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; print('value %7.1f (%.3f ms/iter) e2e %7.1f local %.3f rhs %.3f solve %.3f' % (d['value'], d['ms_per_step']/10, d['e2e']['value'], p['local'], p['rhs'], p['solve']))"
