cd $GRAFT_REPO_ROOT
N=$(nvidia-smi -L | wc -l)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_dist_gpu.py -x -q -s -k "many_ranks or pcg" 2>&1 | grep -E "rank|passed|failed|rror" | tail -12 | tee gpurun_out/r2aa_pytest.log
ADMMB_VERBOSE=1 timeout 300 $TR --master-port 29541 tools/strong_scaling.py --cube 55 --solver pcg --steps 20 2>&1 | grep -E "rank 0.*(halo|peer)|value|rror|timed" | tee gpurun_out/r2aa_pcg.log
timeout 300 $TR --master-port 29542 tools/strong_scaling.py --cube 110 --solver pcg 2>&1 | grep -E "value|rror|timed" | tee -a gpurun_out/r2aa_pcg.log
