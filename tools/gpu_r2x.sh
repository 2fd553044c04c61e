cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_dist_gpu.py -x -q -k pcg 2>&1 | tail -3 | tee gpurun_out/r2x_pytest.log
for H in 1 0; do
  ADMMB_PCG_HALO=$H ADMMB_VERBOSE=1 timeout 600 $TR --master-port 2952$H tools/strong_scaling.py --cube 55 --solver pcg --steps 20 2>&1 | grep -E "halo|value" | sed "s/^/halo=$H /" | tee -a gpurun_out/r2x_pcg.log
done
timeout 600 python tools/strong_scaling.py --cube 55 --solver pcg --steps 20 2>&1 | grep value | sed "s/^/1gpu /" | tee -a gpurun_out/r2x_pcg.log
for H in 1 0; do
  ADMMB_PCG_HALO=$H timeout 900 $TR --master-port 2953$H tools/strong_scaling.py --cube 110 --solver pcg 2>&1 | grep -E "value" | sed "s/^/halo=$H /" | tee -a gpurun_out/r2x_pcg.log
done
