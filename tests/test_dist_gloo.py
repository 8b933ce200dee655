"""CPU, world_size 2, gloo: the multi-rank host logic (even read split, owner partition bookkeeping,
all-to-all-v, per-owner grouping, merge) reproduces the single-rank table and the reference pin."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("world,case", [(2, "cfg5_small"), (3, "cfg4_small")])
def test_sharded_binner_matches_single_rank(world, case, tmp_path):
    out = tmp_path / "result.txt"
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "dist_worker.py"), "--backend", "gloo", "--case", case, "--out", str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert out.read_text().startswith(f"OK world={world}")


def test_split_reads_evenly():
    sys.path.insert(0, os.path.dirname(HERE))
    from genome_assembly_b200.dist import split_reads_evenly
    assert split_reads_evenly(10, 3) == [0, 3, 6, 10]
    assert split_reads_evenly(0, 2) == [0, 0, 0]
    b = split_reads_evenly(1_000_003, 8)
    assert b[0] == 0 and b[-1] == 1_000_003 and all(0 <= y - x - 125000 <= 1 for x, y in zip(b, b[1:]))
