"""GPU, world_size 2+, NCCL: scan -> owner partition -> all-to-all -> group on real devices reproduces
the single-GPU table and the reference pin (skipped when fewer than 2 GPUs are visible)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def n_gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("form,exchange", [("skr", "peer"), ("skr", "nccl"), ("records", "nccl")])
@pytest.mark.parametrize("case", ["cfg5_small", "cfg4_small", "cfg1_reads", "fuzz_polyA"])
def test_nccl_sharded_binner_matches_single_rank(case, form, exchange, tmp_path):
    world = min(n_gpus(), 4)
    if world < 2:
        pytest.skip("needs at least 2 GPUs")
    out = tmp_path / "result.txt"
    port = 29700 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dist_worker.py"), "--backend", "nccl", "--case", case, "--out", str(out), "--form", form,
           "--exchange", exchange, "--repeat", "3" if exchange == "peer" else "1"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert out.read_text().startswith(f"OK world={world}")
