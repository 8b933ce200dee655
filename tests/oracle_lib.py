"""ctypes view of oracle/_build/libgbin_oracle.so — TEST INFRASTRUCTURE ONLY (the checker, never the product)."""
import ctypes as C
import gzip
import hashlib
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
GOLDEN = os.path.join(ROOT, "tests", "golden")


class OrcTuple(C.Structure):
    _fields_ = [("mmer", C.c_uint32), ("arrival", C.c_uint32), ("khi", C.c_uint64), ("klo", C.c_uint64)]


class OrcWindow(C.Structure):
    _fields_ = [("sig_pos", C.c_int32), ("is_rev", C.c_int32), ("mmer", C.c_uint32)]


class OrcResult(C.Structure):
    _fields_ = [("K", C.c_int), ("M", C.c_int), ("cutoff", C.c_int), ("kw", C.c_int),
                ("n_instances", C.c_uint64), ("n_distinct", C.c_uint64), ("n_buckets", C.c_uint64),
                ("n_kmers", C.c_uint64), ("n_ids", C.c_uint64),
                ("mmer_codes", C.POINTER(C.c_uint32)), ("mmer_kmer_off", C.POINTER(C.c_uint64)),
                ("kmer_codes", C.POINTER(C.c_uint64)), ("kmer_id_off", C.POINTER(C.c_uint64)),
                ("read_ids", C.POINTER(C.c_int32))]


TUPLE_DT = np.dtype([("mmer", "<u4"), ("arrival", "<u4"), ("khi", "<u8"), ("klo", "<u8")])
WINDOW_DT = np.dtype([("sig_pos", "<i4"), ("is_rev", "<i4"), ("mmer", "<u4")])

_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(ORACLE_DIR, "_build", "libgbin_oracle.so")
        if not os.path.exists(so):
            subprocess.run(["make", "-s", "-C", ORACLE_DIR, "_build/libgbin_oracle.so"], check=True)
        L = C.CDLL(so)
        L.orc_getval.argtypes = [C.c_char]
        L.orc_getval.restype = C.c_int
        L.orc_getbp.argtypes = [C.c_int]
        L.orc_getbp.restype = C.c_char
        L.orc_getscore.argtypes = [C.c_char_p]
        L.orc_getscore.restype = C.c_int
        L.orc_fgets_split.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.POINTER(C.POINTER(C.c_uint64)),
                                      C.POINTER(C.POINTER(C.c_uint32))]
        L.orc_fgets_split.restype = C.c_size_t
        L.orc_run.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int,
                              C.POINTER(OrcResult)]
        L.orc_run.restype = C.c_int
        L.orc_result_free.argtypes = [C.POINTER(OrcResult)]
        L.orc_group_tuples.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32, C.c_int, C.c_int, C.c_int, C.POINTER(OrcResult)]
        L.orc_group_tuples.restype = C.c_int
        L.orc_scan_all.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_scan_all.restype = C.c_size_t
        L.orc_dump_strings.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                       C.c_void_p]
        L.orc_dump_strings.restype = C.c_int
        L.orc_dump.argtypes = [C.POINTER(OrcResult), C.c_void_p]
        L.orc_dump.restype = C.c_int
        L.free = C.CDLL(None).free
        L.free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


_libc = C.CDLL(None)
_libc.fopen.restype = C.c_void_p
_libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
_libc.fclose.argtypes = [C.c_void_p]


def fgets_split(data: bytes, read_length_define: int):
    """main's read loop (binning.c:1154-1166) -> (starts u64[n], lens u32[n])."""
    L = lib()
    sp = C.POINTER(C.c_uint64)()
    lp = C.POINTER(C.c_uint32)()
    n = L.orc_fgets_split(data, len(data), read_length_define, C.byref(sp), C.byref(lp))
    starts = np.ctypeslib.as_array(sp, shape=(max(n, 1),))[:n].copy()
    lens = np.ctypeslib.as_array(lp, shape=(max(n, 1),))[:n].copy()
    L.free(sp)
    L.free(lp)
    return starts, lens


class Table:
    """Flat pruned table in canonical order (host numpy arrays)."""

    def __init__(self, K, M, cutoff, n_instances, n_distinct, mmer_codes, mmer_kmer_off, kmer_codes, kmer_id_off, read_ids):
        self.K, self.M, self.cutoff = K, M, cutoff
        self.kw = 1 if K <= 32 else 2
        self.n_instances, self.n_distinct = n_instances, n_distinct
        self.mmer_codes, self.mmer_kmer_off = mmer_codes, mmer_kmer_off
        self.kmer_codes, self.kmer_id_off, self.read_ids = kmer_codes, kmer_id_off, read_ids

    @property
    def n_buckets(self):
        return len(self.mmer_codes)

    @property
    def n_kmers(self):
        return len(self.kmer_id_off) - 1

    def assert_equal(self, other):
        assert (self.K, self.M) == (other.K, other.M)
        assert self.n_instances == other.n_instances, (self.n_instances, other.n_instances)
        assert self.n_distinct == other.n_distinct, (self.n_distinct, other.n_distinct)
        np.testing.assert_array_equal(self.mmer_codes, other.mmer_codes)
        np.testing.assert_array_equal(self.mmer_kmer_off, other.mmer_kmer_off)
        np.testing.assert_array_equal(self.kmer_codes.reshape(-1), other.kmer_codes.reshape(-1))
        np.testing.assert_array_equal(self.kmer_id_off, other.kmer_id_off)
        np.testing.assert_array_equal(self.read_ids, other.read_ids)

    def dump_lines(self):
        """'<mmer> <kmer> <ids...>' lines, like oracle/ref_harness_main.c prints them."""
        lut = np.frombuffer(b"TGCA", dtype=np.uint8)

        def dec(code_hi, code_lo, n):
            v = (int(code_hi) << 64) | int(code_lo)
            return bytes(lut[[(v >> (2 * (n - 1 - t))) & 3 for t in range(n)]])

        out = []
        kc = self.kmer_codes.reshape(-1, self.kw)
        for b in range(self.n_buckets):
            mm = dec(0, self.mmer_codes[b], self.M)
            for s in range(int(self.mmer_kmer_off[b]), int(self.mmer_kmer_off[b + 1])):
                km = dec(kc[s, 0] if self.kw == 2 else 0, kc[s, -1], self.K)
                ids = self.read_ids[int(self.kmer_id_off[s]):int(self.kmer_id_off[s + 1])]
                out.append(mm + b" " + km + b"".join(b" %d" % i for i in ids))
        return out

    def md5(self):
        return hashlib.md5(b"".join(x + b"\n" for x in sorted(self.dump_lines()))).hexdigest()


def _table_from_result(L, r, K, M, cutoff) -> "Table":
    kw = r.kw

    def arr(p, n, dt):
        if n == 0:
            return np.zeros(0, dtype=dt)
        return np.ctypeslib.as_array(p, shape=(n,)).astype(dt, copy=True)

    t = Table(K, M, cutoff, r.n_instances, r.n_distinct,
              arr(r.mmer_codes, r.n_buckets, np.uint32), arr(r.mmer_kmer_off, r.n_buckets + 1, np.uint64),
              arr(r.kmer_codes, r.n_kmers * kw, np.uint64), arr(r.kmer_id_off, r.n_kmers + 1, np.uint64),
              arr(r.read_ids, r.n_ids, np.int32))
    L.orc_result_free(C.byref(r))
    return t


def group_tuples(tuples: np.ndarray, K, M, cutoff, id_base=0) -> "Table":
    """Grouping + prune of tuples given in arrival order (the owner-side work of the multi-rank path)."""
    L = lib()
    t = np.ascontiguousarray(tuples, dtype=TUPLE_DT).copy()
    r = OrcResult()
    rc = L.orc_group_tuples(t.ctypes.data, len(t), None, id_base, K, M, cutoff, C.byref(r))
    assert rc == 0
    return _table_from_result(L, r, K, M, cutoff)


def run(data: bytes, starts, lens, K, M, cutoff, ids=None) -> Table:
    L = lib()
    starts = np.ascontiguousarray(starts, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    idp = None
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        idp = ids.ctypes.data
    r = OrcResult()
    rc = L.orc_run(data, starts.ctypes.data, lens.ctypes.data, len(starts), idp, K, M, cutoff, C.byref(r))
    assert rc == 0, rc
    return _table_from_result(L, r, K, M, cutoff)


def scan_all(data: bytes, starts, lens, K, M):
    """Per-window tuples and signature trace in arrival order (pre-sort)."""
    L = lib()
    starts = np.ascontiguousarray(starts, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    total = int(np.maximum(lens.astype(np.int64) - K + 1, 0).sum())
    tup = np.zeros(max(total, 1), dtype=TUPLE_DT)
    win = np.zeros(max(total, 1), dtype=WINDOW_DT)
    n = L.orc_scan_all(data, starts.ctypes.data, lens.ctypes.data, len(starts), K, M, tup.ctypes.data, win.ctypes.data)
    assert n == total
    return tup[:n], win[:n]


def dump_strings_md5(data: bytes, starts, lens, K, M, cutoff):
    L = lib()
    starts = np.ascontiguousarray(starts, dtype=np.uint64)
    lens = np.ascontiguousarray(lens, dtype=np.uint32)
    with tempfile.NamedTemporaryFile(suffix=".dump") as tf:
        f = _libc.fopen(tf.name.encode(), b"wb")
        rc = L.orc_dump_strings(data, starts.ctypes.data, lens.ctypes.data, len(starts), None, K, M, cutoff, f)
        _libc.fclose(f)
        assert rc == 0
        lines = open(tf.name, "rb").read().split(b"\n")[:-1]
    return hashlib.md5(b"".join(x + b"\n" for x in sorted(lines))).hexdigest(), lines


def load_pins():
    with open(os.path.join(GOLDEN, "pins.json")) as f:
        return json.load(f)["cases"]


def load_case_bytes(case) -> bytes:
    with gzip.open(os.path.join(GOLDEN, case["file"]), "rb") as f:
        return f.read()


def _mix64(x: int) -> int:
    m = (1 << 64) - 1
    x ^= x >> 33
    x = (x * 0xff51afd7ed558ccd) & m
    x ^= x >> 33
    x = (x * 0xc4ceb9fe1a85ec53) & m
    x ^= x >> 33
    return x


def table_digest(t) -> int:
    """Python restatement of gbin_table_digest (csrc/table_digest.cu): sum over k-mers of a hash of (m-mer code, k-mer code,
    read ids in list order), modulo 2^64.  For small tables only (pure-Python loop)."""
    m = (1 << 64) - 1
    kw = t.kw
    kc = np.asarray(t.kmer_codes).reshape(-1, kw)
    acc = 0
    for b in range(len(t.mmer_codes)):
        mm = int(t.mmer_codes[b])
        for s in range(int(t.mmer_kmer_off[b]), int(t.mmer_kmer_off[b + 1])):
            a, e = int(t.kmer_id_off[s]), int(t.kmer_id_off[s + 1])
            lh = 0x9E3779B97F4A7C15
            for i in t.read_ids[a:e]:
                lh = _mix64(lh ^ (int(i) & 0xffffffff))
            h = ((mm * 0x9E3779B97F4A7C15) & m) ^ _mix64(int(kc[s, 0]))
            if kw == 2:
                h ^= _mix64((int(kc[s, 1]) + 0x632BE59BD9B4E019) & m)
            acc = (acc + _mix64(h ^ lh ^ (((e - a) << 40) & m))) & m
    return acc
