cd $GRAFT_REPO_ROOT
timeout 420 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2_final_pytest.log
timeout 200 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_final_bench.err | tail -1 > gpurun_out/r2_final_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final_bench.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "phases", d["phases_ms_per_iteration"], "roofline", d["roofline"]["frac"], d["roofline_global"]["frac"], d["roofline_global"]["kernel"], "launches", d["gpu_launches"], "clocks", d["clocks"])
print("pairs", {k: (round(v["reference"], 1), round(v["b200"], 1)) for k, v in d.get("same_config_pairs", {}).items()} if isinstance(d.get("same_config_pairs"), dict) else d.get("same_config_pairs"))
PY
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
