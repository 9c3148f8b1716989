"""bench.py's reference arm runs on the CPU (the oracle on all host cores), so its JSON contract can be checked here."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--seconds", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("encoded audio sec/sec") and d["unit"] == "x realtime"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32"
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"} and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"] == {"value": d["value"], "unit": "x realtime", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "4096" in d["config"]["workload"]


def test_numa_cpulist_parser():
    import importlib
    sys.path.insert(0, ROOT)
    sh = importlib.import_module("swift-mp3_b200.sharding")
    assert sh._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sh._parse_cpulist("") == set()
    assert sh.shard_range(4096, 3, 8) == (1536, 2048)


def test_stdout_carries_only_the_json_line():
    """Libraries print to file descriptor 1 from C (NCCL's version line): bench.py parks it on stderr until emit()."""
    code = ("import os, sys; sys.path.insert(0, %r); import bench; bench.quiet_stdout(); os.write(1, b'NCCL version x\\n'); "
            "print('python noise'); bench.emit({'a': 1})" % ROOT)
    p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    assert p.stdout == '{"a": 1}\n' and "NCCL version x" in p.stderr and "python noise" in p.stderr
