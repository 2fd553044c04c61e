cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -k pcg 2>&1 | tail -5 | tee gpurun_out/r2y_pytest.log
for P in 1 0; do
  ADMMB_PCG_P2P=$P ADMMB_VERBOSE=1 timeout 300 $TR --master-port 2952$P tools/strong_scaling.py --cube 55 --solver pcg --steps 20 2>&1 | grep -E "peer|value|rror|timed" | sed "s/^/p2p=$P /" | tee -a gpurun_out/r2y_pcg.log
done
ADMMB_PCG_P2P=1 timeout 600 $TR --master-port 29531 tools/strong_scaling.py --cube 110 --solver pcg 2>&1 | grep -E "value|rror|timed" | sed "s/^/p2p=1 /" | tee -a gpurun_out/r2y_pcg.log
