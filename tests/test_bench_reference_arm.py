"""bench.py --impl reference: the CPU arm the driver runs beside the B200 arm (no GPU needed).  One JSON line with the contract's
keys; under torchrun rank 0 alone prints it."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype",
        "data", "config", "cpu_baseline", "e2e"]


def _lines(cmd):
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


def _check(d, n_gpus):
    for key in KEYS:
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "Gbit/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["n_gpus"] == n_gpus and d["vs_baseline"] is None and d["dtype"] == "u8"
    assert "workload" in d["config"] and "n2040_k1530" in d["config"]["workload"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_prints_one_contract_line():
    lines = _lines([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-seconds", "1"])
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_only_rank0_prints():
    lines = _lines([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                    "--master-port", "29541", "bench.py", "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                    "--cpu-seconds", "1"])
    assert len(lines) == 1
    _check(lines[0], 2)
