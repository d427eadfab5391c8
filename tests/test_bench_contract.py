"""bench.py's JSON contract, CPU side: the reference arm (`--impl reference`) really runs here (the reference build when it
is on this machine, else the oracle port) on a tiny sample and must print exactly one line with the agreed keys; the
committed GPU-arm line (profiles/r1_bench_n1.json, written on a B200) is checked against the same contract."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
             "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_reference_arm_prints_one_contract_line():
    # --netgen-nref 0: the reference arm's netgen system without refinement (46 k tets; the default, 3 refinements = 108 M dofs,
    # is for the GPU box); without the reference build the oracle port on the tiny generator sample stands in
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--size", "12", "--steps", "3", "--warmup", "1",
                        "--cpu-sample", "6", "--cpu-sample-t1", "4", "--netgen-nref", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["impl"] == "reference" and d["metric"] == "cg_iterations_per_s" and d["unit"] == "iterations/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["dtype"] == "f64" and d["gpu_launches"] == 0
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True, text=True,
                       timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_round2_line_carries_the_netgen_check():
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_n1.json")))
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    ng = d["config"]["netgen_check"]
    assert ng["ndof"] > 100e6 and ng["reordered"] is True and ng["cg_it_per_s"] > 0 and ng["cpu_reference_it_per_s"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["sample_ndof"] == ng["ndof"]     # measured at full size
    rf = d["roofline"]
    assert abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["traffic"] > 0
    assert d["e2e"]["value"] < d["value"] and d["gpu_launches"] > 0


def test_committed_gpu_line_keeps_the_contract():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1_bench_n1.json")))
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    rf = d["roofline"]
    assert rf["bound"] == "hbm" and rf["unit"] == "GB/s" and abs(rf["frac"] - rf["achieved"] / rf["peak"]) < 1e-12 and rf["traffic"] > 0
    assert d["gpu_launches"] > 0 and d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["e2e"]["value"] < d["value"]                       # the copies are inside the timed region
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert not ({"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"} & set(d["clocks"]["reasons"]))
