"""Multi-GPU binning: one process per GPU, torch.distributed for the plumbing.

    reads split evenly by contiguous id range (rank g holds arrival indices [base_g, base_g + n_g))
      -> scan stage on every rank (process_read's window/signature work, embarrassingly parallel)
      -> stable partition of the records by owner = gbin_owner_of(mmer_code, world) (a hash of the code)
      -> ONE personalised all-to-all-v of records (count exchange, then payload)
      -> sort / run-length / prune on the owner, which now holds whole m-mer buckets.

Receiving in source-rank order keeps arrival order inside every key (rank ranges are ordered and each
rank's records are in arrival order), so the owner's stable sort yields the reference's list order.
The result stays sharded: every rank returns the table of the m-mer buckets it owns.

The compute stages are C-ABI calls into libgbin.so on device buffers (`GpuStages`).  The class takes
the stage object as a parameter only so that the host-side logic (split sizes, exchange, merge) can be
exercised on CPU with the gloo backend by tests that plug in the oracle; the product path has no
other implementation.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.distributed as dist

from . import binding as B


def split_reads_evenly(n_reads: int, world: int):
    """Contiguous read ranges per rank: rank g gets [bounds[g], bounds[g+1])."""
    return [(n_reads * g) // world for g in range(world + 1)]


class GpuStages:
    """The product stages: scan / partition / group through the C ABI on torch-allocated HBM buffers.

    form="skr" (default) exchanges super-k-mer records (pipeline 2: one 32/48-byte record per signature segment,
    about a fifth of the bytes); form="records" exchanges expanded k-mer instance records (pipeline 1), which is
    also what a batch is re-run with when a shared-memory unit of pipeline 2 overflows."""

    def __init__(self, binner: B.Binner, form: str = "skr"):
        assert form in ("skr", "records")
        self.b = binner
        self.form = form
        self.device = torch.device("cuda", binner.device)
        self.record_bytes = binner.skr_record_bytes if form == "skr" else binner.record_bytes
        self.launches = 0  # kernels launched by the stages since construction

    def fallback(self) -> "GpuStages":
        """The stages a batch is re-run with after an overflow (None if already on the general path)."""
        if self.form == "records":
            return None
        fb = GpuStages(self.b, "records")
        fb.launches = self.launches
        return fb

    def _count(self):
        self.launches += self.b.timings()["kernel_launches"]

    def stream(self):
        return B.stream_handle(torch.cuda.current_stream(self.device))

    def alloc_records(self, n: int) -> torch.Tensor:
        return torch.empty(max(n, 1) * self.record_bytes, dtype=torch.uint8, device=self.device)

    def scan(self, reads: B.CReads, arrival_base: int):
        n = self.b.count_instances_device(reads, self.stream())
        if self.form == "skr":
            # a segment covers about (K-M+2)/2 windows (9-23 at the BASELINE shapes); the call reports the exact need if this is short
            cap = getattr(self, "_skr_cap_seen", 0) or (n // 6 + int(reads.n_reads) + 1024)  # what the last batch needed (+6 %), else an estimate
            while True:
                rec = self.alloc_records(cap)
                try:
                    n_skr, n_inst = self.b.scan_skr_device(reads, arrival_base, rec, cap, self.stream())
                    break
                except B.GbinError as e:
                    if e.code != B.GBIN_E_INVALID_ARG or cap >= n:
                        raise
                    cap = n
            assert n_inst == n
            self._skr_cap_seen = max(getattr(self, "_skr_cap_seen", 0), n_skr + n_skr // 16 + 1024)
            self._count()
            return rec, n_skr
        rec = self.alloc_records(n)
        got = self.b.scan_device(reads, arrival_base, rec, n, self.stream())
        assert got == n
        self._count()
        return rec, n

    def partition(self, rec: torch.Tensor, n: int, parts: int):
        out = self.alloc_records(n)
        if self.form == "skr":
            counts = self.b.partition_skr_device(rec, n, parts, out, self.stream())
        else:
            counts = self.b.partition_device(rec, n, parts, out, self.stream())
        self._count()
        return out, counts

    # ---- partition fused with the exchange: records go straight into the owners' buffers over NVLink (peer stores)
    def peer_exchange_setup(self, rank: int, world: int, capacity_records: int, group=None) -> bool:
        """Collective.  Maps every rank's receive buffer into every other rank (CUDA IPC handles gathered over the process
        group).  False when the form has no peer path (instance records use the NCCL all-to-all)."""
        if self.form != "skr":
            return False
        if getattr(self, "peer_world", 0):  # regrow: all ranks close their mappings before any owner frees what it exported
            self.b.xchg_detach()
            dist.barrier(group=group)
        mine = self.b.xchg_create(rank, world, capacity_records)
        blobs = [None] * world
        dist.all_gather_object(blobs, mine, group=group)
        self.b.xchg_attach(b"".join(blobs))
        dist.barrier(group=group)
        self.peer_world = world
        self.peer_capacity = capacity_records
        return True

    def peer_exchange(self, rec: torch.Tensor, n: int):
        """Collective.  -> (receive buffer address, records received, records sent per owner), or None when some owner's
        buffer was too small (every rank gets None; nothing was stored)."""
        try:
            out = self.b.xchg_exchange_skr(rec, n, self.peer_world, self.stream())
        except B.GbinError as e:
            if e.code != B.GBIN_E_TOO_LARGE:
                raise
            return None
        self._count()
        return out

    def lend_scratch(self, buf: torch.Tensor):
        """The next group() call may use `buf` (device memory, dead data) for its sort buffers."""
        if self.form == "skr":
            self.b.donate_scratch(buf, buf.numel() * buf.element_size())

    def group(self, rec, n: int, id_base: int):
        """Returns the device table, or None when a shared-memory unit overflowed (form "skr" only)."""
        if self.form == "skr":
            try:
                t = self.b.group_skr_device(rec, n, None, id_base, self.stream())
            except B.GbinError as e:
                if e.code != B.GBIN_E_STATE:
                    raise
                return None
        else:
            t = self.b.group_device(rec, n, None, id_base, self.stream())
        self._count()
        return t


@dataclass
class ExchangeStats:
    sent_records: int = 0
    recv_records: int = 0
    sent_bytes_offrank: int = 0
    scan_ms: float = 0.0
    partition_ms: float = 0.0
    exchange_ms: float = 0.0
    group_ms: float = 0.0
    extra: dict = field(default_factory=dict)


class ShardedBinner:
    """Runs the hot path across the ranks of a torch.distributed process group."""

    def __init__(self, stages, group=None, time_stages: bool = False, exchange: str = "auto"):
        """exchange: "peer" = the partition kernel stores into the owners' buffers over NVLink (CUDA IPC peer memory, the
        default when the stages offer it), "nccl" = partition locally, then all_to_all_single."""
        self.stages = stages
        self.group = group
        self.exchange_kind = "nccl" if exchange == "nccl" or not hasattr(stages, "peer_exchange_setup") else "peer"
        self._peer_ready = False
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.time_stages = time_stages
        self.stats = ExchangeStats()
        self.fallbacks = 0

    def _ev(self):
        if not self.time_stages:
            return None
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def exchange(self, part: torch.Tensor, send_counts: list[int]):
        """All-to-all-v of records laid out part-major in `part`. Returns (received buffer, n received)."""
        rb = self.stages.record_bytes
        words = rb // 8
        dev = part.device
        send = torch.tensor(send_counts, dtype=torch.int64, device=dev)
        recv = torch.empty_like(send)
        dist.all_to_all_single(recv, send, group=self.group)
        recv_counts = [int(x) for x in recv.tolist()]
        n_in = sum(recv_counts)
        n_out = sum(send_counts)
        inbuf = self.stages.alloc_records(n_in)
        src = part.view(torch.int64)[: n_out * words]
        dst = inbuf.view(torch.int64)[: n_in * words]
        dist.all_to_all_single(dst, src, output_split_sizes=[c * words for c in recv_counts],
                               input_split_sizes=[c * words for c in send_counts], group=self.group)
        self.stats.sent_records = n_out
        self.stats.recv_records = n_in
        self.stats.sent_bytes_offrank = (n_out - send_counts[self.rank]) * rb
        return inbuf, n_in

    def run(self, reads, arrival_base: int, id_base: int = 0):
        """reads: this rank's shard (device CReads); arrival_base: global index of its first read.
        Returns the device table of the m-mer buckets this rank owns.  If any rank reports an overflow of the
        super-k-mer path, every rank re-runs the batch with the general (instance record) stages."""
        table = self._run_once(reads, arrival_base, id_base)
        ok = 0 if table is None else 1
        if self.world > 1:
            flag = torch.tensor([ok], dtype=torch.int32, device=self._flag_device())
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            ok = int(flag.item())
        if ok:
            return table
        fb = self.stages.fallback() if hasattr(self.stages, "fallback") else None
        if fb is None:
            raise RuntimeError("grouping failed and no fallback stages are available")
        self.fallbacks += 1
        saved, self.stages = self.stages, fb
        try:
            table = self._run_once(reads, arrival_base, id_base)
        finally:
            saved.launches = fb.launches
            self.stages = saved
        return table

    def _peer_path(self, n: int) -> bool:
        """Collective decision (every rank evaluates the same conditions): set the peer exchange up on first use."""
        if self.exchange_kind != "peer" or getattr(self.stages, "form", "") != "skr":
            return False
        if not self._peer_ready:
            # an owner receives about a world-th of every rank's records: the ranks' record counts and the owners' shares differ by a few per cent
            cap = torch.tensor([max(n + n // 3 + 65536, getattr(self, "_peer_capacity_hint", 0))], dtype=torch.int64, device=self._flag_device())
            dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=self.group)
            self._peer_ready = self.stages.peer_exchange_setup(self.rank, self.world, int(cap.item()), self.group)
            if not self._peer_ready:
                self.exchange_kind = "nccl"
        return self._peer_ready

    def _flag_device(self):
        return getattr(self.stages, "device", torch.device("cpu"))

    def _run_once(self, reads, arrival_base: int, id_base: int = 0):
        e0 = self._ev()
        rec, n = self.stages.scan(reads, arrival_base)
        e1 = self._ev()
        if self.world == 1:
            inbuf, n_in = rec, n
            e2 = e3 = e1
        elif self._peer_path(n):
            got = self.stages.peer_exchange(rec, n)
            if got is None:  # an owner's buffer was too small: grow it for the next step, this step goes over NCCL
                self._peer_ready = False
                self._peer_capacity_hint = 2 * getattr(self.stages, "peer_capacity", n)
                part, counts = self.stages.partition(rec, n, self.world)
                e2 = self._ev()
                inbuf, n_in = self.exchange(part, counts)
                del part
            else:
                inbuf, n_in, counts = got
                e2 = e1
                self.stats.sent_records = sum(counts)
                self.stats.recv_records = n_in
                self.stats.sent_bytes_offrank = (sum(counts) - counts[self.rank]) * self.stages.record_bytes
            # the exchange call returned after its stream was synchronised: the records have left.  Their buffer (tens of GB for a
            # config-3-sized shard) is lent to the grouping call for its sort buffers and goes back to the allocator's cache afterwards,
            # where the next step's scan finds it: no cudaFree / cudaMalloc of that size inside the step.
            lent = rec if (got is not None and hasattr(self.stages, "lend_scratch")) else None
            del rec
            if lent is not None:
                self.stages.lend_scratch(lent)
            elif n * self.stages.record_bytes > (12 << 30):  # nothing to lend it to: a block of that size must not idle in the cache
                torch.cuda.empty_cache()
            e3 = self._ev()
        else:
            part, counts = self.stages.partition(rec, n, self.world)
            del rec
            e2 = self._ev()
            inbuf, n_in = self.exchange(part, counts)
            del part
            e3 = self._ev()
        table = self.stages.group(inbuf, n_in, id_base)
        lent = None  # (the group call has returned: its stream was synchronised)
        e4 = self._ev()
        if self.time_stages:
            torch.cuda.synchronize()
            self.stats.scan_ms = e0.elapsed_time(e1)
            self.stats.partition_ms = e1.elapsed_time(e2)
            self.stats.exchange_ms = e2.elapsed_time(e3)
            self.stats.group_ms = e3.elapsed_time(e4)
        self._keep = inbuf  # the group stage's input must outlive the call (it is sort scratch)
        return table


def merge_owner_tables(tables: list[B.HostTable]) -> B.HostTable:
    """Union of per-owner tables (disjoint m-mer buckets) in canonical order — for parity checks."""
    t0 = tables[0]
    kw = t0.kw
    entries = []
    for t in tables:
        kc = t.kmer_codes.reshape(-1, kw)
        for b in range(t.n_buckets):
            entries.append((int(t.mmer_codes[b]), t, kc, int(t.mmer_kmer_off[b]), int(t.mmer_kmer_off[b + 1])))
    entries.sort(key=lambda e: e[0])
    mmer_codes, mmer_off, kcs, id_lens, ids = [], [0], [], [], []
    for code, t, kc, s0, s1 in entries:
        mmer_codes.append(code)
        kcs.append(kc[s0:s1])
        lo, hi = int(t.kmer_id_off[s0]), int(t.kmer_id_off[s1])
        id_lens.append(np.diff(t.kmer_id_off[s0:s1 + 1].astype(np.int64)))
        ids.append(t.read_ids[lo:hi])
        mmer_off.append(mmer_off[-1] + (s1 - s0))
    if entries:
        kmer_codes = np.concatenate(kcs).reshape(-1)
        id_off = np.concatenate([[0], np.cumsum(np.concatenate(id_lens))]).astype(np.uint64)
        read_ids = np.concatenate(ids).astype(np.int32)
    else:
        kmer_codes = np.zeros(0, np.uint64)
        id_off = np.zeros(1, np.uint64)
        read_ids = np.zeros(0, np.int32)
    return B.HostTable(K=t0.K, M=t0.M, cutoff=t0.cutoff, kw=kw,
                       n_instances=sum(t.n_instances for t in tables), n_distinct=sum(t.n_distinct for t in tables),
                       mmer_codes=np.array(mmer_codes, dtype=np.uint32), mmer_kmer_off=np.array(mmer_off, dtype=np.uint64),
                       kmer_codes=kmer_codes, kmer_id_off=id_off, read_ids=read_ids)
