"""Contract of `bench.py --impl reference` (the arm the driver runs beside the native one): exactly ONE line on
stdout, a JSON object with the keys of the native line, `impl: reference`, a `cpu_baseline` describing the run and an
`e2e` that repeats the line's own value.  Runs the CPU oracle port for one bounded step (no GPU needed)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    # EG_LIB_PATH points nowhere: importing echoglad_b200 (and so mapping libechoglad_b200.so) on this arm would raise
    env = dict(os.environ, EG_LIB_PATH="/nonexistent/libechoglad_b200.so")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.split("\n") if ln.strip()]
    assert len(lines) == 1, r.stdout[:2000]
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and d["e2e"]["value"] == d["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"]
    # the same `config` as the native arm prints (the bounded sample is described in cpu_baseline.sample)
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(64, 1)
    assert "bounded sample" in d["cpu_baseline"]["sample"]
