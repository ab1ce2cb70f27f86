"""The bench contract on the CPU side: `bench.py --impl reference` (the reference's Reference-platform arithmetic on the
host cores) prints exactly ONE JSON line on stdout carrying the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "fep",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    assert line["metric"] == "sliced-PME force evaluations/s" and line["unit"] == "evals/s"
    assert line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["value"] > 0 and line["value"] == line["cpu_baseline"]["value"] == line["e2e"]["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["workload"].startswith("fep")
