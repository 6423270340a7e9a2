"""CPU tests of the bench.py contract that needs no GPU: the reference arm (`--impl reference`: the CPU restatement timed
on the host cores) prints exactly one JSON line with the contract's keys -- alone, and under torchrun with two ranks
(rank 0 prints, the other rank exits 0 without work)."""

import json
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMALL = ["--batch", "16", "--prefix-len", "128", "--layers", "2", "--heads", "4", "--kv-heads", "4", "--head-dim", "64", "--steps", "2", "--warmup", "1"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _check_line(out: str, n_gpus: int):
    lines = [ln for ln in out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == n_gpus and d["unit"] == "tokens/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("decode tokens/sec") and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_single_process():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", *SMALL], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    _check_line(r.stdout, 1)


def test_reference_arm_under_torchrun_two_ranks():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", *SMALL]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, OMP_NUM_THREADS="2"))
    assert r.returncode == 0, r.stderr[-2000:]
    _check_line(r.stdout, 2)
