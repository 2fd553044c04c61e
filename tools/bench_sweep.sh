#!/bin/bash
# default bench line, the reference arm, and the size sweep of DESIGN.md section 7 (one GPU)
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 300 gpurun_out/bench_default.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err
for N in ${SIZES:-30 70 87 110}; do
  python bench.py --cube $N --steps 10 --warmup 20 --no-cpu-baseline > gpurun_out/bench_cube$N.json 2>> gpurun_out/bench_default.err
done
for f in gpurun_out/bench_default.json gpurun_out/bench_reference.json gpurun_out/bench_cube*.json; do
  python - "$f" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
if d.get("impl") == "reference":
    print(sys.argv[1], "reference arm", d.get("value"), d.get("cpu_baseline"))
else:
    p = d["phases_ms_per_iteration"]; s = d["setup"]
    print(sys.argv[1], "value %.1f e2e %.1f | local %.3f rhs %.3f solve %.3f | solve %.0f GB/s frac %.2f | nnzL %.1fM factor %.2f GB setup %.1fs factor %.1fs | cpu %s" % (
        d["value"], d["e2e"]["value"], p["local"], p["rhs"], p["solve"], d["roofline"]["achieved"], d["roofline"]["frac"], s["nnz_L"] / 1e6,
        s["factor_bytes"] / 1e9, s["seconds"], s["factor_seconds"], d.get("cpu_baseline")))
PY
done
