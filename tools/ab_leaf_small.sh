#!/bin/bash
# small shipped scenes: ND leaf size (number of solve levels) vs dense fill
cd $GRAFT_REPO_ROOT
for sc in bunnyexpand windyflag poordillo plinkopony; do
for leaf in 64 256 1024 4096; do
  ADMMB_ND_LEAF=$leaf python bench.py --scene $sc --steps 60 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-12s leaf %5d  value %8.0f it/s  e2e %8.0f  launches/frame %5.0f  levels %d  setup %.2f s' % ('$sc', $leaf, d['value'], d['e2e']['value'], d['gpu_launches']/60.0, d['setup']['levels'], d['setup']['seconds']))"
done; done
