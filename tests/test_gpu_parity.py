"""GPU parity tests proper: the CUDA path, called through the C ABI (libgbin.so via ctypes), against
the oracle on the same inputs and against the reference-execution pins in tests/golden/pins.json.
Bit-exact everywhere (integer / byte / index work)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
from genome_assembly_b200 import binding as B
from genome_assembly_b200 import synth

pytestmark = pytest.mark.gpu

CASES = O.load_pins()
GPU_CASES = [c for c in CASES if c["acgt_only"] and c["k"] >= 2 * c["m"]]


def torch_cuda():
    import torch
    assert torch.cuda.is_available(), "these tests need a CUDA device"
    return torch


def assert_tables_equal(got: B.HostTable, want: O.Table):
    assert (got.K, got.M, got.kw) == (want.K, want.M, want.kw)
    assert got.n_instances == want.n_instances
    assert got.n_distinct == want.n_distinct
    np.testing.assert_array_equal(got.mmer_codes, want.mmer_codes)
    np.testing.assert_array_equal(got.mmer_kmer_off, want.mmer_kmer_off)
    np.testing.assert_array_equal(got.kmer_codes, want.kmer_codes.reshape(-1))
    np.testing.assert_array_equal(got.kmer_id_off, want.kmer_id_off)
    np.testing.assert_array_equal(got.read_ids, want.read_ids)


def c_table_of(t: B.HostTable):
    keep = dict(mc=np.ascontiguousarray(t.mmer_codes, np.uint32), mo=np.ascontiguousarray(t.mmer_kmer_off, np.uint64),
                kc=np.ascontiguousarray(t.kmer_codes, np.uint64), ko=np.ascontiguousarray(t.kmer_id_off, np.uint64),
                ids=np.ascontiguousarray(t.read_ids, np.int32))
    ct = B.CTable(t.K, t.M, t.cutoff, t.kw, 0, 1, t.n_instances, t.n_distinct, t.n_buckets, t.n_kmers, len(t.read_ids),
                  keep["mc"].ctypes.data, keep["mo"].ctypes.data, keep["kc"].ctypes.data, keep["ko"].ctypes.data, keep["ids"].ctypes.data)
    return ct, keep


def as_oracle_table(t: B.HostTable) -> O.Table:
    return O.Table(t.K, t.M, t.cutoff, t.n_instances, t.n_distinct, t.mmer_codes, t.mmer_kmer_off, t.kmer_codes, t.kmer_id_off,
                   t.read_ids)


def records_to_numpy(torch, buf, n, kw):
    raw = buf[: n * (8 * kw + 8)].cpu().numpy()
    dt = np.dtype([("k", "<u8", (kw,)), ("mmer", "<u4"), ("arrival", "<u4")])
    return raw.view(dt)


def dev_reads(torch, data, starts, lens, ids=None):
    d = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    s = torch.from_numpy(starts.astype(np.int64)).cuda()
    l = torch.from_numpy(lens.astype(np.int32)).cuda()
    i = torch.from_numpy(ids.astype(np.int32)).cuda() if ids is not None else None
    rd = B.Binner._reads(d, d.numel(), len(starts), starts=s, lens=l, read_ids=i)
    return rd, (d, s, l, i)


@pytest.mark.parametrize("case", GPU_CASES, ids=lambda c: c["name"])
def test_scan_stage_matches_process_read(case):
    """Per-window (m-mer, oriented k-mer, arrival) records == the oracle's process_read restatement."""
    torch = torch_cuda()
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    K, M = case["k"], case["m"]
    tup, _ = O.scan_all(data, starts, lens, K, M)
    b = B.Binner(K, M, case["cutoff"])
    rd, keep = dev_reads(torch, data, starts, lens)
    assert b.count_instances_device(rd) == len(tup) == case["instances"]
    kw = 1 if K <= 32 else 2
    buf = torch.zeros(max(len(tup), 1) * b.record_bytes, dtype=torch.uint8, device="cuda")
    n = b.scan_device(rd, 0, buf, len(tup))
    assert n == len(tup)
    rec = records_to_numpy(torch, buf, n, kw)
    np.testing.assert_array_equal(rec["mmer"], tup["mmer"])
    np.testing.assert_array_equal(rec["arrival"], tup["arrival"])
    np.testing.assert_array_equal(rec["k"][:, -1], tup["klo"])
    if kw == 2:
        np.testing.assert_array_equal(rec["k"][:, 0], tup["khi"])
    b.close()


def expand_skr(rec_words: np.ndarray, K: int):
    """Host expansion of super-k-mer records (include/gbin.h layout) into per-window (mmer, arrival, khi, klo)."""
    nw = rec_words.shape[1]
    out = []
    mask = (1 << (2 * K)) - 1
    pbits = 32 * (nw - 4)
    for r in rec_words:
        arrival, mmer, meta, start = int(r[0]), int(r[1]), int(r[2]), int(r[3])
        n, rev = meta & 0xff, (meta >> 8) & 1
        payload = 0
        for w in r[4:]:
            payload = (payload << 32) | int(w)
        for t in range(n):
            k = (payload >> (pbits - 2 * K - 2 * t)) & mask
            if rev:
                k = ~k & mask
            out.append((mmer, arrival, k >> 64, k & ((1 << 64) - 1), start + t))
    return out


@pytest.mark.parametrize("case", GPU_CASES, ids=lambda c: c["name"])
def test_super_kmer_scan_matches_process_read(case):
    """Pipeline-2 scan stage: expanding every super-k-mer record window by window reproduces the oracle's
    per-window tuples in arrival order (so segments, orientation and m-mer codes are all right)."""
    torch = torch_cuda()
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    K, M = case["k"], case["m"]
    if case["instances"] > 400000:
        starts, lens = starts[:1200], lens[:1200]
    tup, win = O.scan_all(data, starts, lens, K, M)
    b = B.Binner(K, M, case["cutoff"])
    rd, keep = dev_reads(torch, data, starts, lens)
    rb = b.skr_record_bytes
    assert rb == (32 if K <= 32 else 48)
    buf = torch.zeros(max(len(tup), 1) * rb, dtype=torch.uint8, device="cuda")
    n_skr, n_inst = b.scan_skr_device(rd, 0, buf, len(tup))
    assert n_inst == len(tup)
    words = buf[: n_skr * rb].cpu().numpy().view(np.uint32).reshape(n_skr, rb // 4)
    got = expand_skr(words, K)
    assert len(got) == len(tup)
    np.testing.assert_array_equal(np.array([g[0] for g in got], dtype=np.uint32), tup["mmer"])
    np.testing.assert_array_equal(np.array([g[1] for g in got], dtype=np.uint32), tup["arrival"])
    np.testing.assert_array_equal(np.array([g[3] for g in got], dtype=np.uint64), tup["klo"])
    np.testing.assert_array_equal(np.array([g[2] for g in got], dtype=np.uint64), tup["khi"])
    # a segment ends exactly where the oracle's signature position changes (or the read ends)
    seg_first = np.array([True] + [bool(win["sig_pos"][i] != win["sig_pos"][i - 1] or tup["arrival"][i] != tup["arrival"][i - 1])
                                   for i in range(1, len(tup))]) if len(tup) else np.zeros(0, bool)
    assert int(seg_first.sum()) == n_skr
    b.close()


@pytest.mark.parametrize("pipeline", [3, 2, 1])
@pytest.mark.parametrize("case", GPU_CASES, ids=lambda c: c["name"])
def test_table_matches_oracle_and_reference_pin(case, pipeline, tmp_path):
    """Whole path through gbin_read_file_fgets + gbin_bin_reads_host: identical arrays to the oracle and
    the same md5 as the reference binary's sorted dump."""
    torch_cuda()
    data = O.load_case_bytes(case)
    p = tmp_path / "reads.txt"
    p.write_bytes(data)
    d, starts, lens = B.read_file_fgets(str(p), case["read_length_define"])
    b = B.Binner(case["k"], case["m"], case["cutoff"], pipeline=pipeline)
    got = b.bin_host(d, len(starts), starts=starts, lens=lens)
    want = O.run(data, starts, lens, case["k"], case["m"], case["cutoff"])
    assert_tables_equal(got, want)
    assert got.n_kmers == case["surviving_kmers"] and got.n_buckets == case["surviving_buckets"]
    assert as_oracle_table(got).md5() == case["md5"]
    if "digest" in case:  # gbin_table_digest of the host table == the pinned digest (anchored on the reference's dump)
        dg = C.c_uint64()
        ct, keep = c_table_of(got)
        assert B.load_library().gbin_table_digest(None, C.byref(ct), None, C.byref(dg)) == 0
        assert "%016x" % dg.value == case["digest"]
    tm = b.timings()
    assert tm["kernel_launches"] > 0
    info = b.pipeline_info()
    assert info["configured"] == pipeline
    if pipeline == 1 or case["instances"] == 0:
        assert info["last_used"] == 1
    elif case["name"] not in ("fuzz_polyA", "fuzz_ragged_k15"):
        # homopolymers and 2-letter-alphabet reads hold single k-mers with more instances than a shared-memory unit:
        # pipelines 3 and 2 detect the overflow and the batch is redone by the next pipeline down
        if pipeline == 3 and case["m"] <= 5:
            # few, huge m-mer buckets: a (bucket, d) class can exceed a unit of pipeline 3, the batch then goes to pipeline 2
            assert info["last_used"] in (3, 2), info
        else:
            assert info["last_used"] == pipeline and info["fallbacks"] == 0, info
    b.close()


def test_device_path_and_staged_path_agree_with_host_path():
    torch = torch_cuda()
    case = next(c for c in CASES if c["name"] == "cfg2_small")
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    K, M, cut = case["k"], case["m"], case["cutoff"]
    want = O.run(data, starts, lens, K, M, cut)
    b = B.Binner(K, M, cut)
    rd, keep = dev_reads(torch, data, starts, lens)
    dev = b.bin_device_raw(rd, B.stream_handle(torch.cuda.current_stream()))
    assert dev.on_device == 1
    assert "%016x" % b.table_digest(dev) == case["digest"]  # digest kernel on the device table == pinned digest
    assert_tables_equal(b.table_to_host(dev), want)
    # staged: scan -> group
    n = b.count_instances_device(rd)
    buf = torch.empty(n * b.record_bytes, dtype=torch.uint8, device="cuda")
    assert b.scan_device(rd, 0, buf, n) == n
    dev2 = b.group_device(buf, n)
    assert_tables_equal(b.table_to_host(dev2), want)
    b.close()


def test_fixed_stride_form_and_explicit_ids():
    torch_cuda()
    rs = synth.generate(4000, 100, error_rate=0.02, seed=5, starts="uniform")
    K, M = 31, 11
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    ids = (np.arange(rs.n_reads, dtype=np.int32) * 7 + 100)[::-1].copy()  # arbitrary caller ids, not sorted
    b = B.Binner(K, M, 1)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    want = O.run(rs.as_bytes(), starts, lens, K, M, 1)
    assert_tables_equal(got, want)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len, read_ids=ids)
    want = O.run(rs.as_bytes(), starts, lens, K, M, 1, ids=ids)
    assert_tables_equal(got, want)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len, id_base=1000)
    np.testing.assert_array_equal(got.read_ids, O.run(rs.as_bytes(), starts, lens, K, M, 1).read_ids + 1000)
    b.close()


@pytest.mark.parametrize("pipeline", [3, 2, 1])
@pytest.mark.parametrize("K,M,cutoff,L", [(32, 15, 1, 80), (33, 8, 0, 90), (64, 15, 1, 200), (8, 4, 2, 40), (4, 2, 5, 30),
                                          (31, 4, -1, 60), (40, 13, 3, 150), (63, 2, 1, 100)])
def test_key_width_and_parameter_edges(K, M, cutoff, L, pipeline):
    """K = 32/33/64 (64-bit and 128-bit code boundaries), smallest/largest M, cutoff 0 / none."""
    torch_cuda()
    rs = synth.generate(1500, L, genome_len=2000, error_rate=0.01, seed=K * 100 + M, starts="uniform")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    b = B.Binner(K, M, cutoff, pipeline=pipeline)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    want = O.run(rs.as_bytes(), starts, lens, K, M, cutoff)
    assert want.n_kmers > 0
    assert_tables_equal(got, want)
    b.close()


def test_host_path_streams_reads_in_and_table_out():
    """gbin_bin_reads_host overlaps PCIe with the kernels: fixed-stride reads are copied and scanned in chunks, and from the
    second call on a context the finished part of the table is copied out while the grouping continues.  Every call must give
    the oracle's table, whatever the chunking, also when a larger batch follows a smaller one (arena regrowth)."""
    torch_cuda()
    rs = synth.generate(40_000, 100, error_rate=0.01, seed=11, starts="uniform")
    K, M = 31, 11
    n_small = 9000
    want = {}
    for n in (rs.n_reads, n_small):
        starts = np.arange(n, dtype=np.uint64) * rs.stride
        lens = np.full(n, rs.read_len, dtype=np.uint32)
        want[n] = O.run(rs.buf[: n * rs.stride].tobytes(), starts, lens, K, M, 1)
    b = B.Binner(K, M, 1)
    for n in (n_small, rs.n_reads, rs.n_reads, n_small, rs.n_reads):
        got = b.bin_host(rs.buf[: n * rs.stride], n, stride=rs.stride, read_len=rs.read_len)
        assert_tables_equal(got, want[n])
        assert b.pipeline_info()["last_used"] == 3
    # device table -> pinned arena (what the multi-GPU end-to-end path uses)
    torch = torch_cuda()
    d = torch.from_numpy(rs.buf).cuda()
    dev = b.bin_device_raw(B.Binner._reads(d, d.numel(), rs.n_reads, stride=rs.stride, read_len=rs.read_len),
                           B.stream_handle(torch.cuda.current_stream()))
    assert_tables_equal(B.host_table_from_c(b.table_to_pinned_raw(dev)), want[rs.n_reads])
    b.close()


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_device_side_fgets_split_matches_mains_read_loop(case):
    """main's read loop (binning.c:1154-1166) on the device: same starts / lens as the host replay of fgets, for every
    fixture (bundled reads.txt with READ_LENGTH 101: a 99-base read and an empty read per line; ragged lines, long lines,
    missing trailing newline) and for a range of READ_LENGTH values around the line length."""
    torch = torch_cuda()
    data = O.load_case_bytes(case)
    d = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    b = B.Binner(31, 11, 1)
    for R in sorted({case["read_length_define"], 2, 3, 17, 100, 101, 102, 103, 4096}):
        want_s, want_l = O.fgets_split(data, R)
        rd = b.split_reads_device(d, d.numel(), R, B.stream_handle(torch.cuda.current_stream()))
        n = int(rd.n_reads)
        assert n == len(want_s), (R, n, len(want_s))
        got_s = b.device_array(rd.starts, n, np.uint64)
        got_l = b.device_array(rd.lens, n, np.uint32)
        np.testing.assert_array_equal(got_s, want_s)
        np.testing.assert_array_equal(got_l, want_l)
    b.close()


@pytest.mark.parametrize("case", [c for c in GPU_CASES if c["name"] in ("cfg1_reads", "cfg2_small", "fuzz_ragged_k25", "short_reads")], ids=lambda c: c["name"])
def test_file_to_table_in_one_call(case, tmp_path):
    """gbin_bin_file_host: file -> device -> split there -> hot path -> host table; equal to the oracle and the reference pin."""
    torch_cuda()
    data = O.load_case_bytes(case)
    path = tmp_path / "reads.txt"
    path.write_bytes(data)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    want = O.run(data, starts, lens, case["k"], case["m"], case["cutoff"])
    b = B.Binner(case["k"], case["m"], case["cutoff"])
    got = b.bin_file_host(str(path), case["read_length_define"])
    assert_tables_equal(got, want)
    assert as_oracle_table(got).md5() == case["md5"]
    b.close()


@pytest.mark.parametrize("L", [1700, 2600, 4000])
def test_long_reads_pick_a_scan_kernel_that_fits(L):
    """Both scan kernels keep a read in shared memory: above about 1 800 bases the super-k-mer scan runs with fewer warps per block
    or hands the batch to pipeline 1's scan, which holds reads up to gbin_max_read_len().  Same table as the oracle either way."""
    torch_cuda()
    assert B.load_library().gbin_max_read_len() >= 4000
    rs = synth.generate(300, L, genome_len=3 * L, error_rate=0.01, seed=L, starts="uniform")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    for K, M in ((31, 11), (63, 15)):
        b = B.Binner(K, M, 1)
        got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
        assert_tables_equal(got, O.run(rs.as_bytes(), starts, lens, K, M, 1))
        b.close()


def test_file_call_with_a_large_read_length_define(tmp_path):
    """gbin_bin_file_host(path, 4096) on 100-base reads: READ_LENGTH bounds the line length only; the kernels are sized by the
    real maximum read length, found on the device."""
    torch_cuda()
    case = next(c for c in CASES if c["name"] == "cfg2_small")
    data = O.load_case_bytes(case)
    path = tmp_path / "reads.txt"
    path.write_bytes(data)
    starts, lens = O.fgets_split(data, 4096)
    b = B.Binner(case["k"], case["m"], case["cutoff"])
    got = b.bin_file_host(str(path), 4096)
    assert_tables_equal(got, O.run(data, starts, lens, case["k"], case["m"], case["cutoff"]))
    assert b.pipeline_info()["last_used"] == 3
    b.close()


def test_empty_and_degenerate_batches():
    torch_cuda()
    b = B.Binner(31, 4, 1)
    t = b.bin_host(np.zeros(1, np.uint8), 0)
    assert (t.n_instances, t.n_kmers, t.n_buckets, t.n_ids) == (0, 0, 0, 0)
    # every read shorter than K
    data = b"ACGT\nACGTACGT\n\n"
    starts, lens = O.fgets_split(data, 101)
    t = b.bin_host(data, len(starts), starts=starts, lens=lens)
    assert (t.n_instances, t.n_kmers) == (0, 0)
    np.testing.assert_array_equal(t.kmer_id_off, [0])
    # one read exactly K long: one instance, pruned away at cutoff 1, kept with the prune disabled
    one = b"A" * 31 + b"\n"
    s1, l1 = np.array([0], np.uint64), np.array([31], np.uint32)
    t = b.bin_host(one, 1, starts=s1, lens=l1)
    assert (t.n_instances, t.n_distinct, t.n_kmers) == (1, 1, 0)
    b2 = B.Binner(31, 4, -1)
    t = b2.bin_host(one, 1, starts=s1, lens=l1)
    assert_tables_equal(t, O.run(one, s1, l1, 31, 4, -1))
    # a homopolymer batch: one giant group (every window identical)
    poly = (b"A" * 100 + b"\n") * 3000
    sp, lp = np.arange(3000, dtype=np.uint64) * 101, np.full(3000, 100, np.uint32)
    t = b.bin_host(poly, 3000, stride=101, read_len=100)
    w = O.run(poly, sp, lp, 31, 4, 1)
    assert w.n_kmers == 1 and w.n_instances == 210000
    assert_tables_equal(t, w)
    b.close()
    b2.close()


def test_non_acgt_input_is_rejected():
    torch_cuda()
    case = next(c for c in CASES if c["name"] == "fuzz_nonacgt")
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    b = B.Binner(case["k"], case["m"], case["cutoff"])
    with pytest.raises(B.GbinError) as e:
        b.bin_host(data, len(starts), starts=starts, lens=lens)
    assert e.value.code == B.GBIN_E_NON_ACGT
    b.close()


def test_owner_partition_is_stable_and_complete():
    torch = torch_cuda()
    case = next(c for c in CASES if c["name"] == "cfg5_small")
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    K, M = case["k"], case["m"]
    tup, _ = O.scan_all(data, starts, lens, K, M)
    b = B.Binner(K, M, 1)
    rd, keep = dev_reads(torch, data, starts, lens)
    n = len(tup)
    buf = torch.empty(n * b.record_bytes, dtype=torch.uint8, device="cuda")
    out = torch.empty_like(buf)
    assert b.scan_device(rd, 0, buf, n) == n
    for parts in (1, 2, 3, 8):
        counts = b.partition_device(buf, n, parts, out)
        owner = B.owner_of(tup["mmer"], parts)
        assert counts == [int((owner == p).sum()) for p in range(parts)]
        rec = records_to_numpy(torch, out, n, 1)
        order = np.argsort(owner, kind="stable")
        np.testing.assert_array_equal(rec["mmer"], tup["mmer"][order])
        np.testing.assert_array_equal(rec["arrival"], tup["arrival"][order])
        np.testing.assert_array_equal(rec["k"][:, 0], tup["klo"][order])
    b.close()


def test_reference_entry_points_process_read_prune_data():
    """The reference-named shims: process_read per read, prune_data as the flush, result walked through
    the reference's own struct layouts (cfg1 replayed exactly as main does, binning.c:1150-1169)."""
    torch_cuda()
    from test_capi_cpu import ZTable, walk_zhash
    import hashlib
    case = next(c for c in CASES if c["name"] == "cfg1_reads")
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, 101)
    L = B.load_library()
    assert L.gbin_ref_configure(31, 4, 1, 0) == 0
    root = ZTable(0, 0, None)
    for i in range(len(starts)):  # read_id++ for every fgets return, empty reads included
        read = data[int(starts[i]):int(starts[i]) + int(lens[i])]
        r = L.process_read(C.addressof(root), read, i)
        assert r == C.addressof(root)
    assert L.gbin_ref_last_status() == 0
    assert L.prune_data(C.addressof(root)) == C.addressof(root)
    assert L.gbin_ref_last_status() == 0
    got = walk_zhash(root)
    assert len(got) == case["surviving_kmers"]
    assert hashlib.md5(b"".join(x + b"\n" for x in sorted(got))).hexdigest() == case["md5"]
    # inserting after the flush is a state error, reported through gbin_ref_last_status
    L.process_read(C.addressof(root), b"ACGT", 0)
    assert L.gbin_ref_last_status() == B.GBIN_E_STATE
    L.gbin_zhash_release(C.addressof(root))
    L.gbin_ref_reset(C.addressof(root))


@pytest.mark.parametrize("pipeline", [3, 2, 1])
def test_medium_synthetic_cfg2_shape(pipeline):
    """60 000 reads x 100 bp (4.2 M instances, 1026 sort tiles): multi-tile sort, multi-level scans."""
    torch_cuda()
    rs = synth.generate(60000, 100, error_rate=0.01, seed=20, starts="triangular")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    b = B.Binner(31, 11, 1, pipeline=pipeline)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    want = O.run(rs.as_bytes(), starts, lens, 31, 11, 1)
    assert_tables_equal(got, want)
    assert b.pipeline_info()["last_used"] == pipeline
    b.close()


@pytest.mark.parametrize("pipeline", [3, 2])
def test_homopolymer_batch_falls_back_to_the_hbm_pipeline(pipeline):
    """One k-mer with 210 000 instances cannot fit a shared-memory unit: pipelines 3 and 2 detect it and the batch is
    redone by pipeline 1 — same table either way."""
    torch_cuda()
    poly = (b"A" * 100 + b"\n") * 3000
    sp, lp = np.arange(3000, dtype=np.uint64) * 101, np.full(3000, 100, np.uint32)
    b = B.Binner(31, 4, 1, pipeline=pipeline)
    t = b.bin_host(poly, 3000, stride=101, read_len=100)
    assert_tables_equal(t, O.run(poly, sp, lp, 31, 4, 1))
    info = b.pipeline_info()
    assert info["last_used"] == 1 and info["fallbacks"] == pipeline - 1
    b.close()


@pytest.mark.parametrize("cap", [1024, 512])
@pytest.mark.parametrize("n_reads", [2500, 12000])
def test_deep_coverage_rounds_and_spans(n_reads, cap):
    """Pipeline 3 on a tiny genome at ~125x / ~600x coverage: every m-mer bucket is larger than a unit, so it is grouped in
    d-rounds (ranges of the signature's offset inside the k-mer) by several warps and put back in k-mer order by the span
    pass; id lists hold hundreds of entries."""
    torch_cuda()
    rs = synth.generate(n_reads, 100, genome_len=2000, error_rate=0.002, seed=77, starts="uniform")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    for K, M in ((31, 11), (31, 4), (63, 15)):
        b = B.Binner(K, M, 1, pipeline=3)
        b.set_tuning("v3_cap", cap)
        got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
        want = O.run(rs.as_bytes(), starts, lens, K, M, 1)
        assert_tables_equal(got, want)
        info = b.pipeline_info()
        if n_reads == 2500 and M > 4:  # no (bucket, d) class exceeds a unit: no fallback (M = 4 has a few huge buckets)
            assert info["last_used"] == 3 and info["fallbacks"] == 0, (K, M, cap, info)
        b.close()


@pytest.mark.parametrize("n_reads,expect_v2", [(2500, True), (12000, False)])
def test_deep_coverage_long_id_lists(n_reads, expect_v2):
    """A tiny genome at ~125x / ~600x coverage: id lists of hundreds of entries and m-mer buckets larger than a
    shared-memory unit (split by k-mer prefix).  At 600x a handful of k-mers with ~500 instances each no longer
    spread evenly over the prefix sub-units, pipeline 2 reports the overflow and pipeline 1 redoes the batch."""
    torch_cuda()
    rs = synth.generate(n_reads, 100, genome_len=2000, error_rate=0.002, seed=77, starts="uniform")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    for K, M in ((31, 11), (31, 4), (63, 15)):
        b = B.Binner(K, M, 1, pipeline=2)
        got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
        want = O.run(rs.as_bytes(), starts, lens, K, M, 1)
        assert np.diff(want.kmer_id_off.astype(np.int64)).max() > (60 if expect_v2 else 200)
        assert_tables_equal(got, want)
        if expect_v2:
            assert b.pipeline_info()["last_used"] == 2, (K, M, b.pipeline_info())
        b.close()


def test_large_buckets_are_sliced_not_redone():
    """19 M instances over a few thousand m-mer buckets (M=7): most buckets are several times larger than a shared-memory
    unit and are range-partitioned on the k-mer prefix with sampled splitters; no slice may overflow (no fallback)."""
    torch_cuda()
    rs = synth.generate(150_000, 150, error_rate=0.01, seed=31, starts="uniform")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    b = B.Binner(25, 7, 1, pipeline=2)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    info, st = b.pipeline_info(), b.run_stats()
    assert info["last_used"] == 2 and info["fallbacks"] == 0, (info, st)
    assert st["n_units"] > 3 * st["n_mmer_runs"], st
    assert_tables_equal(got, O.run(rs.as_bytes(), starts, lens, 25, 7, 1))
    b.close()


@pytest.mark.parametrize("cap", [1024, 512])
def test_v3_large_buckets_in_rounds(cap):
    """Pipeline 3 on the same 19 M instances: buckets of ~15 000 instances are cut into d-rounds; a (bucket, d) class that
    exceeds a unit sends the batch to pipeline 2.  Either way the table is the oracle's."""
    torch_cuda()
    rs = synth.generate(150_000, 150, error_rate=0.01, seed=31, starts="uniform")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    b = B.Binner(25, 7, 1, pipeline=3)
    b.set_tuning("v3_cap", cap)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    assert b.pipeline_info()["last_used"] in (3, 2)
    assert_tables_equal(got, O.run(rs.as_bytes(), starts, lens, 25, 7, 1))
    b.close()


@pytest.mark.parametrize("nc,h", [(0, 0), (2, 2)])
def test_v3_passes_append_to_one_table(nc, h):
    """A batch with more k-mer instances than 32-bit coordinates hold is grouped in passes over ranges of the sorted entries, cut
    between m-mer buckets, every pass appending to the table.  Forced here on 4.2 M instances with a pass limit of 300 000."""
    torch_cuda()
    rs = synth.generate(60000, 100, error_rate=0.01, seed=20, starts="triangular")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    b = B.Binner(31, 11, 1, pipeline=3)
    b.set_tuning("v3_pass_max", 300_000)
    b.set_tuning("v3_nc", nc)
    b.set_tuning("v3_h", h)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    st = b.run_stats()
    assert b.pipeline_info()["last_used"] == 3 and st["n_passes"] >= 3, st
    assert_tables_equal(got, O.run(rs.as_bytes(), starts, lens, 31, 11, 1))
    b.close()


def check_table_invariants(t: B.HostTable, n_reads, W):
    """Size-independent properties of a pruned table (used at BASELINE's full sizes)."""
    assert t.n_instances == n_reads * W
    assert (np.diff(t.mmer_codes.astype(np.int64)) > 0).all(), "buckets ascend strictly by m-mer code"
    assert (t.mmer_codes >= (1 << (2 * t.M - 1))).all(), "stored m-mer is the larger of m-mer / complement"
    assert t.mmer_kmer_off[0] == 0 and t.mmer_kmer_off[-1] == t.n_kmers and (np.diff(t.mmer_kmer_off.astype(np.int64)) > 0).all()
    assert t.kmer_id_off[0] == 0 and t.kmer_id_off[-1] == t.n_ids
    cnt = np.diff(t.kmer_id_off.astype(np.int64))
    assert (cnt > t.cutoff).all(), "every surviving k-mer has more than cutoff occurrences"
    if t.kw == 1:  # k-mers ascend strictly inside a bucket
        kc = t.kmer_codes
        asc = kc[1:] > kc[:-1]
        bucket_start = np.zeros(t.n_kmers, dtype=bool)
        bucket_start[t.mmer_kmer_off[:-1].astype(np.int64)] = True
        assert (asc | bucket_start[1:]).all()
    # ids newest-first inside every list
    ids = t.read_ids.astype(np.int64)
    desc = ids[1:] <= ids[:-1]
    list_start = np.zeros(t.n_ids, dtype=bool)
    list_start[t.kmer_id_off[:-1].astype(np.int64)] = True
    assert (desc | list_start[1:]).all()
    assert ids.min() >= 0 and ids.max() < n_reads
    assert t.n_ids <= t.n_instances and t.n_kmers <= t.n_distinct <= t.n_instances


def test_full_size_cfg2_properties_and_prefix_parity():
    """BASELINE config 2 at full size (1 M reads x 100 bp, K=31, M=11): invariants on the whole table,
    checksum agreement between two runs, and exact parity with the oracle on a 30 000-read prefix."""
    torch_cuda()
    rs = synth.generate(1_000_000, 100, error_rate=0.01, seed=20, starts="triangular")
    b = B.Binner(31, 11, 1)
    t = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    check_table_invariants(t, rs.n_reads, 70)
    t2 = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    for a in ("mmer_codes", "mmer_kmer_off", "kmer_codes", "kmer_id_off", "read_ids"):
        np.testing.assert_array_equal(getattr(t, a), getattr(t2, a))  # idempotent / deterministic
    n = 30000
    starts = np.arange(n, dtype=np.uint64) * rs.stride
    lens = np.full(n, rs.read_len, dtype=np.uint32)
    got = b.bin_host(rs.buf[: n * rs.stride], n, stride=rs.stride, read_len=rs.read_len)
    assert_tables_equal(got, O.run(rs.buf[: n * rs.stride].tobytes(), starts, lens, 31, 11, 1))
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["cfg1_reads", "cfg2_small", "cfg4_small", "kat_twice"])
def test_expand_read_id_list_on_the_device(name, tmp_path):
    """expand_read_id_list (binning.c:857-888) as a device CSR replicate.  The expanded lists, copied back and
    printed in print_kmer_read_ids's layout, give byte for byte the text of gbin_table_dump_expanded_format (which the CPU suite
    pins against the unmodified reference's own expand + print), and — where the reference harness is built — the same
    {(m-mer, k-mer): K id lines} as the reference's --expanded output."""
    import os
    import subprocess
    torch = torch_cuda()
    case = next(c for c in CASES if c["name"] == name)
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    K = case["k"]
    b = B.Binner(K, case["m"], case["cutoff"])
    rd, keep = dev_reads(torch, data, starts, lens)
    dev = b.bin_device_raw(rd)
    xd = b.expand_read_ids_device(dev)
    assert (xd.n_lists, xd.n_ids, xd.on_device) == (dev.n_kmers * K, dev.n_ids * K, 1)
    xh = b.expanded_to_host(xd)
    ht = B.CTable()
    b._check(b.lib.gbin_table_to_host(b.h, C.byref(dev), C.byref(ht)))
    try:
        off = np.ctypeslib.as_array(C.cast(xh.list_off, C.POINTER(C.c_uint64)), shape=(xh.n_lists + 1,))
        ids = np.ctypeslib.as_array(C.cast(xh.ids, C.POINTER(C.c_int32)), shape=(max(xh.n_ids, 1),))[: xh.n_ids]
        t_off = np.ctypeslib.as_array(C.cast(ht.kmer_id_off, C.POINTER(C.c_uint64)), shape=(ht.n_kmers + 1,))
        t_ids = np.ctypeslib.as_array(C.cast(ht.read_ids, C.POINTER(C.c_int32)), shape=(max(ht.n_ids, 1),))[: ht.n_ids]
        # CSR replicate: list (j, b) is a copy of list j, the K lists of a k-mer back to back
        cnt = np.diff(t_off)
        np.testing.assert_array_equal(np.diff(off), np.repeat(cnt, K))
        assert off[0] == 0 and off[-1] == K * ht.n_ids
        want = np.concatenate([np.tile(t_ids[t_off[j]:t_off[j + 1]], K) for j in range(int(ht.n_kmers))]) if ht.n_kmers else np.zeros(0, np.int32)
        np.testing.assert_array_equal(ids, want)
        p1, p2 = tmp_path / "lists.txt", tmp_path / "format.txt"
        assert b.lib.gbin_table_dump_expanded_lists(C.byref(ht), C.byref(xh), str(p1).encode()) == 0
        assert b.lib.gbin_table_dump_expanded_format(C.byref(ht), str(p2).encode()) == 0
        assert p1.read_bytes() == p2.read_bytes()
        exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                           f"ref_K{K}_M{case['m']}_C{case['cutoff']}_R{case['read_length_define']}")
        if os.path.exists(exe) and name in ("cfg1_reads", "cfg2_small"):
            src = tmp_path / "reads.txt"
            src.write_bytes(data)
            ref = subprocess.run([exe, str(src), "--expanded"], capture_output=True, check=True).stdout
            from test_capi_cpu import _parse_expanded
            assert _parse_expanded(p1.read_bytes(), K) == _parse_expanded(ref, K)
    finally:
        b.lib.gbin_expanded_free(C.byref(xh))
        b.lib.gbin_table_free(C.byref(ht))
    b.close()


@pytest.mark.gpu
@pytest.mark.parametrize("M", [4, 5, 7])
def test_v3_long_spans_are_sorted_span_by_span(M):
    """Few, huge m-mer buckets (M = 4: at most 136 of them for 4.2 M instances): every bucket is a long span whose surviving k-mers
    are put in k-mer order by the per-span merge sort in shared memory (2 048- and 8 192-k-mer launches) or, for spans above that,
    by the global radix sort.  The table must be the oracle's whichever path a span took."""
    torch_cuda()
    rs = synth.generate(60000, 100, error_rate=0.01, seed=20, starts="triangular")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    b = B.Binner(25, M, 1, pipeline=3)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    want = O.run(rs.as_bytes(), starts, lens, 25, M, 1)
    assert_tables_equal(got, want)
    info, st = b.pipeline_info(), b.run_stats()
    if info["last_used"] == 3:
        assert st["n_lsd_kmers"] > 0, st
        sizes = np.diff(want.mmer_kmer_off.astype(np.int64))
        print("M", M, "buckets", len(sizes), "largest", int(sizes.max()), "lsd k-mers", st["n_lsd_kmers"])
    b.close()


_SWEEP = (
    # (read length, K, M): every (positions per lane, (K-M+1 - positions per lane) mod positions per lane) pair of the
    # lane-per-position scan kernel, 32- and 64-bit keys, windows narrower than a lane's positions + 1 and wider than a warp
    [(100, K, 11) for K in (31, 32, 33)] + [(100, 4, 2), (100, 5, 2), (100, 6, 2), (100, 9, 4), (100, 31, 13), (100, 40, 15), (100, 63, 15),
                                             (100, 63, 11), (100, 64, 12), (60, 31, 11), (96, 24, 12)]
    + [(150, K, 11) for K in (30, 31, 32, 33, 34)] + [(150, 31, 14), (150, 63, 15), (150, 10, 5), (160, 64, 11)]
    + [(250, K, 11) for K in (31, 32, 33, 34, 35, 36, 37, 38)] + [(250, 63, 15), (250, 64, 13), (250, 16, 8), (266, 22, 11), (267, 31, 11)]
)


@pytest.mark.gpu
@pytest.mark.parametrize("L,K,M", _SWEEP)
def test_scan_kernel_window_shapes(L, K, M):
    """The sliding arg-max of the scan stage is composed from per-lane prefix / suffix maxima and a doubling over lane maxima whose
    shape depends on (read length, K - M + 1); every shape must give the oracle's table (267-base reads go to the hop-chain kernel)."""
    torch_cuda()
    rs = synth.generate(400, L, genome_len=6000, error_rate=0.01, seed=1000 * L + 10 * K + M, starts="uniform")
    starts = np.arange(rs.n_reads, dtype=np.uint64) * rs.stride
    lens = np.full(rs.n_reads, rs.read_len, dtype=np.uint32)
    b = B.Binner(K, M, 1)
    got = b.bin_host(rs.buf, rs.n_reads, stride=rs.stride, read_len=rs.read_len)
    assert_tables_equal(got, O.run(rs.as_bytes(), starts, lens, K, M, 1))
    b.close()
