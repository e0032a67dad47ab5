"""bench.py contract pieces that run without a GPU: the reference arm prints one JSON line with the agreed keys, and the
algorithmic byte / flop constants the roofline uses are the SURVEY 8(d) figures."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "4"],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "env_steps_per_s" and line["unit"] == "env-steps/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["gpu_launches"] == 0
    have_ref = os.path.exists(os.path.join(ROOT, "baseline", "_ref", "_model", "Burger.py"))
    assert line["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "BASELINE configs[1]" in line["config"]["workload"]


def test_roofline_constants_are_the_survey_figures():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.BYTES_PER_ENV_LAUNCH == 1912          # SURVEY 8(d), Burgers C2
    assert bench.FLOPS_PER_ENV_STEP == 2600
    assert (bench.B_PER_GPU, bench.N, bench.M, bench.NSUB) == (4096, 32, 32, 10)
