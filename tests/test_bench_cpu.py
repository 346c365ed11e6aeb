"""bench.py contract checks that need no GPU: stdout carries exactly ONE JSON line (native libraries
that print to fd 1 — NCCL's version banner under NCCL_DEBUG — are diverted to stderr), and the
reference arm (CPU restatement, the one arm that runs without a device) emits the agreed keys."""
import json
import os
import subprocess
import sys
import textwrap

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_stdout_guard_diverts_native_writes():
    code = textwrap.dedent(f'''
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import bench
        with bench._CleanStdout() as out:
            os.write(1, b"NCCL version 2.28.9+cuda12.9\\n")      # what a native library does
            print("a python print inside the run")
            out.emit('{{"ok": 1}}')
        print("after")
    ''')
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"ok": 1}\nafter\n'
    assert "NCCL version" in r.stderr and "a python print inside the run" in r.stderr


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "config1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("metric", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "config1",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
