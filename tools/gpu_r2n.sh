cd $GRAFT_REPO_ROOT
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_dist_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r2n_pytest.log
ADMMB_FACTOR_VERBOSE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-pairs > gpurun_out/r2n_bench2.json 2> gpurun_out/r2n_bench2.err
grep "sharded" gpurun_out/r2n_bench2.err | head -4
python -c "
import json
d=json.loads(open('gpurun_out/r2n_bench2.json').read().strip().splitlines()[-1])
print('replica value', d['value'], 'partition', json.dumps(d['partition'])[:900])
"
