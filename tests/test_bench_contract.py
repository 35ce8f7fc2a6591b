"""bench.py's output contract, checked on the CPU with the reference arm (the arm that needs no GPU): stdout is exactly ONE
line and that line is the JSON object the driver parses, whatever else the process or its libraries print."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(extra_env=None):
    env = dict(os.environ, **(extra_env or {}))
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--size", "2000000"],
                       capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    return p.stdout


def test_reference_arm_prints_one_json_line():
    out = _run()
    lines = out.splitlines()
    assert len(lines) == 1, out[:500]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in line, k
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["chunk_bytes"] % 16 == 0


def test_stdout_is_claimed_at_descriptor_level():
    """A C library writing to file descriptor 1 (NCCL's version banner does) must not reach the driver's pipe."""
    code = ("import sys, os; sys.path.insert(0, %r); import bench; bench.claim_stdout(); os.write(1, b'NCCL version x.y\\n'); "
            "print('python-level noise'); bench.emit({'ok': 1})" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-1000:]
    assert p.stdout == '{"ok": 1}\n'
    assert "NCCL version" in p.stderr and "python-level noise" in p.stderr
