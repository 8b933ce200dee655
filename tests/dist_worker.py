"""Worker for tests/test_dist_gloo.py (world_size-2 gloo on CPU) and tests/test_gpu_dist.py (NCCL).
Runs the multi-rank host logic of genome_assembly_b200.dist on one golden case and checks the merged
owner tables against the single-rank oracle.  With --backend gloo the compute stages are the ORACLE
(test-only stand-ins: the product has no CPU path); with --backend nccl they are the CUDA stages."""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))
import oracle_lib as O  # noqa: E402
from genome_assembly_b200 import binding as B  # noqa: E402
from genome_assembly_b200.dist import GpuStages, ShardedBinner, merge_owner_tables, split_reads_evenly  # noqa: E402


class OracleStages:
    """CPU stand-ins with the same record layout as the CUDA stages ({u64 k[kw]; u32 mmer; u32 arrival})."""

    def __init__(self, K, M, cutoff):
        self.K, self.M, self.cutoff = K, M, cutoff
        self.kw = 1 if K <= 32 else 2
        self.record_bytes = 8 * self.kw + 8
        self.dt = np.dtype([("k", "<u8", (self.kw,)), ("mmer", "<u4"), ("arrival", "<u4")])

    def alloc_records(self, n):
        return torch.empty(max(n, 1) * self.record_bytes, dtype=torch.uint8)

    def _to_tensor(self, rec):
        t = self.alloc_records(len(rec))
        t[: len(rec) * self.record_bytes] = torch.from_numpy(rec.view(np.uint8).reshape(-1).copy())
        return t

    def _from_tensor(self, t, n):
        return t[: n * self.record_bytes].numpy().view(self.dt)

    def scan(self, reads, arrival_base):
        data, starts, lens = reads
        tup, _ = O.scan_all(data, starts, lens, self.K, self.M)
        rec = np.zeros(len(tup), dtype=self.dt)
        rec["k"][:, -1] = tup["klo"]
        if self.kw == 2:
            rec["k"][:, 0] = tup["khi"]
        rec["mmer"] = tup["mmer"]
        rec["arrival"] = tup["arrival"] + arrival_base
        return self._to_tensor(rec), len(rec)

    def partition(self, rec, n, parts):
        r = self._from_tensor(rec, n)
        owner = B.owner_of(r["mmer"], parts)
        order = np.argsort(owner, kind="stable")
        return self._to_tensor(r[order]), [int((owner == p).sum()) for p in range(parts)]

    def group(self, rec, n, id_base):
        r = self._from_tensor(rec, n)
        tup = np.zeros(n, dtype=O.TUPLE_DT)
        tup["mmer"], tup["arrival"], tup["klo"] = r["mmer"], r["arrival"], r["k"][:, -1]
        if self.kw == 2:
            tup["khi"] = r["k"][:, 0]
        t = O.group_tuples(tup, self.K, self.M, self.cutoff, id_base)
        return B.HostTable(K=t.K, M=t.M, cutoff=t.cutoff, kw=t.kw, n_instances=t.n_instances, n_distinct=t.n_distinct,
                           mmer_codes=t.mmer_codes, mmer_kmer_off=t.mmer_kmer_off, kmer_codes=t.kmer_codes,
                           kmer_id_off=t.kmer_id_off, read_ids=t.read_ids)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--backend", default="gloo")
    ap.add_argument("--case", default="cfg5_small")
    ap.add_argument("--out", required=True)
    ap.add_argument("--form", default="skr")
    ap.add_argument("--exchange", default="auto", help="peer (partition kernel stores into the owners' buffers) or nccl")
    ap.add_argument("--repeat", type=int, default=1, help="run the exchange this many times (epochs of the peer exchange)")
    a = ap.parse_args()
    dist.init_process_group(a.backend)
    rank, world = dist.get_rank(), dist.get_world_size()
    case = next(c for c in O.load_pins() if c["name"] == a.case)
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    K, M, cut = case["k"], case["m"], case["cutoff"]
    bounds = split_reads_evenly(len(starts), world)
    lo, hi = bounds[rank], bounds[rank + 1]
    if a.backend == "gloo":
        stages = OracleStages(K, M, cut)
        reads = (data, starts[lo:hi], lens[lo:hi])
        to_host = lambda t: t  # noqa: E731
    else:
        local = int(os.environ.get("LOCAL_RANK", rank))
        torch.cuda.set_device(local)
        binner = B.Binner(K, M, cut, device=local)
        stages = GpuStages(binner, a.form)
        d = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
        s = torch.from_numpy(starts[lo:hi].astype(np.int64)).cuda()
        l = torch.from_numpy(lens[lo:hi].astype(np.int32)).cuda()
        reads = B.Binner._reads(d, d.numel(), hi - lo, starts=s, lens=l)
        to_host = binner.table_to_host
    sb = ShardedBinner(stages, exchange=a.exchange)
    for _ in range(a.repeat):
        table = to_host(sb.run(reads, arrival_base=lo))
    if a.exchange == "peer" and a.form == "skr":
        assert sb.exchange_kind == "peer" and sb._peer_ready, "the peer exchange was not used"
    # every m-mer this rank holds is one it owns
    assert (B.owner_of(table.mmer_codes, world) == rank).all()
    gathered = [None] * world
    dist.gather_object(table, gathered if rank == 0 else None, dst=0)
    if rank == 0:
        merged = merge_owner_tables(gathered)
        want = O.run(data, starts, lens, K, M, cut)
        got = O.Table(merged.K, merged.M, merged.cutoff, merged.n_instances, merged.n_distinct, merged.mmer_codes,
                      merged.mmer_kmer_off, merged.kmer_codes, merged.kmer_id_off, merged.read_ids)
        got.assert_equal(want)
        assert got.md5() == case["md5"]
        with open(a.out, "w") as f:
            f.write(f"OK world={world} kmers={got.n_kmers} sent={sb.stats.sent_records} fallbacks={sb.fallbacks} exchange={sb.exchange_kind}\n")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
