"""bench.py's JSON line, dry-run on the CPU: the device layer (admm_b200.System, the FP64 probe, torch.cuda) is replaced by a
stub that returns plausible numbers, everything else -- argument handling, the timing protocol's call sequence, the
arithmetic of value / e2e / roofline / phases, the committed-capture traffic, the cpu_baseline leg on the unmodified
reference -- is bench.py's own code.  Guards the driver contract (keys, types, internal consistency) against edits made
where no GPU is available; the numbers themselves mean nothing here."""
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import pytest

from util import ROOT


class FakeSim:
    def __init__(self, sc, **kw):
        self.sc = sc
        self.frames = 0
        self.launch = 0
        self.cg = 0
        self.timing_on = False
        self.t_iters = 0
        self.n3 = 3 * sc["x"].shape[0]
        self.m_x = np.array(sc["x"], dtype=np.float64).reshape(-1)
        self.solver = kw.get("solver", 0)

    def set_x(self, x): pass
    def upload(self): pass
    def download(self): pass
    def close(self): pass

    def step_resident(self, frames=1, iters=None):
        it = frames * int(self.sc["iters"] if iters is None else iters)
        self.frames += frames
        self.launch += 28 * it + 2 * frames
        self.cg += 7 * it if self.solver == 1 else 0
        self.last_ms = 1.1 * it
        if self.timing_on:
            self.t_iters += it

    def step(self):
        self.step_resident(1)

    def last_region_ms(self): return self.last_ms

    def info(self):
        return dict(launches_total=self.launch, cg_iters_total=self.cg, factor_bytes=1635782656, n_levels=13, n_supernodes=3979,
                    nnz_L=98299397, factor_seconds=2.3)

    def timing(self, on):
        self.timing_on = bool(on)

    def timing_read(self, reset=True):
        it = self.t_iters
        out = dict(local_ms=0.71 * it, rhs_ms=0.07 * it, solve_ms=0.35 * it, step_ms=1.12 * it, iters=it)
        if reset:
            self.t_iters = 0
        return out


@pytest.fixture
def bench(monkeypatch):
    import torch
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    fake = types.ModuleType("admm_b200")
    fake.System = FakeSim
    fake.SOLVER_DIRECT, fake.SOLVER_PCG = 0, 1
    fake.probe_fp64 = lambda dev: dict(dfma_tinst_s=17.0, dadd_tinst_s=18.5, dmul_tinst_s=18.5, dfma_dependent_cycles=8.05, sms=148, sm_max_mhz=1965.0)
    monkeypatch.setitem(sys.modules, "admm_b200", fake)
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    monkeypatch.setattr(b, "algorithmic_flops_per_tet", lambda sim, sc, **kw: dict(
        flops_per_tet_iteration=3200.0, objective_evaluations_per_tet_iteration=17.5, lbfgs_iterations_per_tet_iteration=1.0, sampled_tets=4096))
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    return b


def _run(b, monkeypatch, argv):
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    out = io.StringIO()
    with redirect_stdout(out):
        b.main()
    lines = [ln for ln in out.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.getvalue()
    return json.loads(lines[0])


BASE_KEYS = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
             "config", "e2e", "gpu_launches", "clocks", "roofline"]


def test_default_line_has_the_contract_keys_and_is_consistent(bench, monkeypatch):
    d = _run(bench, monkeypatch, ["--steps", "4", "--warmup", "3", "--no-pairs", "--cpu-cube", "4"])
    for k in BASE_KEYS + ["roofline_global", "cpu_baseline", "phases_ms_per_iteration"]:
        assert k in d, k
    assert d["metric"] == "admm_iterations_per_s_cube_1M_tets" and d["n_gpus"] == 1 and d["steps"] == 4 and d["warmup"] == 3
    assert d["higher_is_better"] is True and d["dtype"] == "f64" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    # value = iterations of the timed region / its device time; ms_per_step the same region per frame
    its = 4 * 10
    assert d["value"] == pytest.approx(its / (d["ms_per_step"] * 4e-3))
    assert set(["value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"]) <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] == 2 * 3 * 175616 * 8
    assert d["gpu_launches"] == 28 * its + 2 * 4
    r, g = d["roofline"], d["roofline_global"]
    for o in (r, g):
        assert set(["bound", "achieved", "peak", "unit", "frac", "traffic"]) <= set(o)
        assert o["frac"] == pytest.approx(o["achieved"] / o["peak"])
    assert r["bound"] == "fp64" and g["bound"] == "hbm"
    # traffic of the default workload comes from the committed captures of the final build and is of the order of the
    # algorithmic bytes (608 B per tet and launch; the packed factor per solve)
    assert 0.5 < r["traffic"] / (608.0 * 998250) < 2.0
    assert 0.9 < g["traffic"] / 1635782656 < 1.3
    ph = d["phases_ms_per_iteration"]
    assert ph["sum"] == pytest.approx(ph["local"] + ph["rhs"] + ph["solve"])
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["one_thread"]["cores"] == 1


def test_other_workloads_and_solvers(bench, monkeypatch):
    d = _run(bench, monkeypatch, ["--steps", "3", "--warmup", "3", "--no-pairs", "--no-cpu-baseline", "--cube", "6", "--solver", "pcg"])
    assert d["roofline"]["traffic"] is None and d["roofline_global"]["traffic"] is None      # no capture of this workload
    assert d["config"]["cg_iterations_per_admm_iteration"] == pytest.approx(7.0)
    assert d["phases_ms_per_iteration"]["cg_iterations_per_admm_iteration"] == pytest.approx(7.0)
    d = _run(bench, monkeypatch, ["--steps", "3", "--warmup", "3", "--no-pairs", "--no-cpu-baseline", "--scene", "windyflag"])
    assert d["metric"] == "admm_iterations_per_s_windyflag" and d["roofline"]["bound"] == "hbm"


def test_reference_arm_line(bench, monkeypatch):
    """`--impl reference` runs for real here (the unmodified reference on the host cores), on a small bounded sample."""
    from oracle import ref
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    d = _run(bench, monkeypatch, ["--impl", "reference", "--steps", "2", "--warmup", "1", "--ref-cube", "6", "--no-pairs"])
    assert d["impl"] == "reference" and d["metric"] == "admm_iterations_per_s_cube_1M_tets" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == pytest.approx(d["value"])
    assert d["e2e"]["value"] == pytest.approx(d["value"]) and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


def test_multi_rank_line_has_the_partition_object(bench, monkeypatch):
    """Rank 0 of a (faked) 2-rank torchrun job: the replica ensemble is `value`, ONE mesh partitioned over the ranks is `partition`."""
    import torch.distributed as dist
    monkeypatch.setenv("WORLD_SIZE", "2"); monkeypatch.setenv("RANK", "0"); monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(dist, "init_process_group", lambda *a, **k: None)
    monkeypatch.setattr(dist, "destroy_process_group", lambda *a, **k: None, raising=False)
    monkeypatch.setattr(dist, "barrier", lambda *a, **k: None)
    monkeypatch.setattr(dist, "broadcast_object_list", lambda obj, src=0: None)
    fake = sys.modules["admm_b200"]
    fake.lib = lambda: types.SimpleNamespace(admmb_set_host_threads=lambda n: 0)
    fake.dist_unique_id = lambda: b"\0" * 128
    d = _run(bench, monkeypatch, ["--gpus", "2", "--steps", "3", "--warmup", "3", "--no-pairs"])
    for k in BASE_KEYS:
        assert k in d, k
    assert d["n_gpus"] == 2 and d["scaling"] == "weak" and "cpu_baseline" not in d
    p = d["partition"]
    assert p["scaling"] == "strong" and p["value"] > 0 and p["unit"] == d["unit"]
    assert set(["local", "rhs_and_allgather", "solve_and_allreduce"]) == set(p["phases_ms_per_iteration_rank0"])
