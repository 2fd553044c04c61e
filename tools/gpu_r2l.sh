cd $GRAFT_REPO_ROOT
timeout 1700 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r2l_pytest.log
for sc in bunnyexpand windyflag poordillo plinkopony; do
for cfg in "X=1" "ADMMB_NO_PDL=1" "ADMMB_NO_FUSED_LOCAL=1 ADMMB_NO_PDL=1"; do
  env $cfg python bench.py --scene $sc --steps 60 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%-12s %-40s value %8.0f it/s  e2e %8.0f  launches/frame %5.0f  levels %d' % ('$sc', '$cfg', d['value'], d['e2e']['value'], d['gpu_launches']/60.0, d['setup']['levels']))"
done; done 2>&1 | tee gpurun_out/r2l_small.log
python bench.py --cube 55 --steps 20 --warmup 5 --no-cpu-baseline --no-pairs 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_iteration']; print('cube55 value %7.1f e2e %7.1f local %.3f rhs %.3f solve %.3f  step %.4f launches %d' % (d['value'], d['e2e']['value'], p['local'], p['rhs'], p['solve'], p['step'], d['gpu_launches']))" | tee -a gpurun_out/r2l_small.log
