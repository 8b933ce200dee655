"""CPU: the C-ABI library loads, exports every symbol include/gbin.h declares, and its host-only
pieces (scalar helpers, main's fgets reader, the ZHashTable adapter, dumps) agree with the oracle.
No compute entry point is exercised here (they need a GPU and have no fallback)."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

import oracle_lib as O
from genome_assembly_b200 import binding as B

ROOT = O.ROOT
CASES = O.load_pins()


def declared_functions():
    src = open(os.path.join(ROOT, "include", "gbin.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src))
    return sorted(n for n in names if n.startswith("gbin_") or n in ("getval", "getbp", "getscore", "process_read", "prune_data"))


def test_library_exports_every_declared_symbol():
    L = B.load_library()
    decl = declared_functions()
    assert len(decl) >= 30
    for name in decl:
        assert hasattr(L, name), f"libgbin.so does not export {name}"
    assert sorted(B.EXPORTS) == decl, "binding.EXPORTS out of sync with include/gbin.h"


def test_scalar_helpers_match_reference_kats():
    """binning.c:69-124 known answers."""
    L = B.load_library()
    for s, v in [(b"AAGTCC", 3914), (b"TTCAGG", 181), (b"AAAA", 255), (b"CTTT", 128), (b"AACA", 251), (b"CAGA", 183)]:
        assert L.getscore(s) == v
    OL = O.lib()
    for c in range(1, 256):
        ch = bytes([c])
        assert L.getval(ch) == OL.orc_getval(ch)
    for v in range(-3, 8):
        assert L.getbp(v) == OL.orc_getbp(v)


def test_invalid_configs_are_rejected_without_touching_cuda():
    L = B.load_library()
    for k, m in [(6, 4), (31, 16), (65, 4), (31, 1), (20, 11)]:
        h = C.c_void_p()
        assert L.gbin_create(C.byref(B.Config(k, m, 1, 0)), C.byref(h)) == B.GBIN_E_INVALID_CONFIG
        assert L.gbin_ref_configure(k, m, 1, 0) == B.GBIN_E_INVALID_CONFIG


def test_no_cpu_fallback_without_a_gpu():
    """On a machine without a CUDA device the product refuses to run instead of falling back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B.GbinError) as e:
        B.Binner(31, 4, 1, 0)
    assert e.value.code == B.GBIN_E_CUDA


@pytest.mark.parametrize("case", [c for c in CASES if c["name"] in ("cfg1_reads", "fuzz_ragged_k15", "fuzz_nonacgt", "short_reads")],
                         ids=lambda c: c["name"])
def test_fgets_reader_matches_oracle(case, tmp_path):
    data = O.load_case_bytes(case)
    p = tmp_path / "reads.txt"
    p.write_bytes(data)
    d, starts, lens = B.read_file_fgets(str(p), case["read_length_define"])
    os_, ol = O.fgets_split(data, case["read_length_define"])
    assert d.tobytes() == data
    np.testing.assert_array_equal(starts, os_)
    np.testing.assert_array_equal(lens, ol)
    assert len(starts) == case["read_ids"]


def c_table_from_oracle(t: O.Table):
    """A host gbin_table whose arrays are the oracle's (stand-in for a GPU result in CPU-only tests)."""
    keep = dict(mc=np.ascontiguousarray(t.mmer_codes, np.uint32), mo=np.ascontiguousarray(t.mmer_kmer_off, np.uint64),
                kc=np.ascontiguousarray(t.kmer_codes, np.uint64), ko=np.ascontiguousarray(t.kmer_id_off, np.uint64),
                ids=np.ascontiguousarray(t.read_ids, np.int32))
    ct = B.CTable(t.K, t.M, t.cutoff, t.kw, 0, 1, t.n_instances, t.n_distinct, t.n_buckets, t.n_kmers, len(t.read_ids),
                  keep["mc"].ctypes.data, keep["mo"].ctypes.data, keep["kc"].ctypes.data, keep["ko"].ctypes.data,
                  keep["ids"].ctypes.data)
    return ct, keep


class ZEntry(C.Structure):
    pass


ZEntry._fields_ = [("key", C.c_char_p), ("val", C.c_void_p), ("next", C.POINTER(ZEntry))]


class ZTable(C.Structure):
    _fields_ = [("size_index", C.c_size_t), ("entry_count", C.c_size_t), ("entries", C.POINTER(C.POINTER(ZEntry)))]


class LNode(C.Structure):
    pass


LNode._fields_ = [("next", C.POINTER(LNode)), ("read_id", C.c_int), ("_pad", C.c_int)]

ZSIZES = [53, 101, 211, 503, 1553, 3407, 6803, 12503, 25013, 50261, 104729, 250007, 500009, 1000003]


def walk_zhash(root: ZTable):
    """Walks the pointer graph the way iterate_level_one/two_hash do (binning.c:298-460)."""
    assert C.sizeof(ZEntry) == 24 and C.sizeof(ZTable) == 24 and C.sizeof(LNode) == 16  # zhash.h:14-26, llist.h:7-13
    lines = []
    for i in range(ZSIZES[root.size_index]):
        e = root.entries[i]
        while e:
            kt = C.cast(e.contents.val, C.POINTER(ZTable)).contents
            for j in range(ZSIZES[kt.size_index]):
                ke = kt.entries[j]
                while ke:
                    ids = []
                    n = C.cast(ke.contents.val, C.POINTER(LNode))
                    while n:
                        ids.append(n.contents.read_id)
                        n = n.contents.next
                    lines.append(e.contents.key + b" " + ke.contents.key + b"".join(b" %d" % x for x in ids))
                    ke = ke.contents.next
            e = e.contents.next
    return lines


def ref_hash(key: bytes, size: int) -> int:
    h = 0
    for ch in key:
        h = (17 * h + ch) % size
    return h


@pytest.mark.parametrize("name", ["cfg1_reads", "cfg2_small", "cfg4_small"])
def test_zhash_adapter_and_dumps(name, tmp_path):
    case = next(c for c in CASES if c["name"] == name)
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    t = O.run(data, starts, lens, case["k"], case["m"], case["cutoff"])
    ct, keep = c_table_from_oracle(t)
    L = B.load_library()
    # flat dump
    p = tmp_path / "dump.txt"
    assert L.gbin_table_dump(C.byref(ct), str(p).encode()) == 0
    lines = p.read_bytes().split(b"\n")[:-1]
    assert hashlib.md5(b"".join(x + b"\n" for x in sorted(lines))).hexdigest() == case["md5"]
    # reference print_kmer_read_ids layout
    p2 = tmp_path / "dump_ref.txt"
    assert L.gbin_table_dump_reference_format(C.byref(ct), str(p2).encode()) == 0
    blocks = p2.read_bytes().split(b"\n\n")
    assert len([b for b in blocks if b.strip()]) == case["surviving_buckets"]
    # pointer graph with the reference's struct layouts
    root = ZTable(0, 0, None)
    assert L.gbin_table_to_zhash(C.byref(ct), C.byref(root)) == 0
    assert root.entry_count == case["surviving_buckets"]
    got = walk_zhash(root)
    assert hashlib.md5(b"".join(x + b"\n" for x in sorted(got))).hexdigest() == case["md5"]
    # entries sit in the chain the reference's zgenerate_hash (zhash.c:171-182) would look in
    size = ZSIZES[root.size_index]
    assert root.entry_count <= size // 2 or root.size_index == len(ZSIZES) - 1
    for i in range(size):
        e = root.entries[i]
        while e:
            assert ref_hash(e.contents.key, size) == i
            e = e.contents.next
    L.gbin_zhash_release(C.byref(root))
    assert not root.entries


def test_owner_function_host_and_numpy_agree_and_balance():
    """gbin_owner_of (C, host) == binding.owner_of (numpy) == what the partition kernel computes (tests/test_gpu_parity.py
    checks the kernel against the numpy form); every owner is in range and a uniform code space is spread evenly."""
    import numpy as np
    from genome_assembly_b200 import binding as B
    L = B.load_library()
    rng = np.random.default_rng(7)
    codes = np.concatenate([rng.integers(0, 4 ** 15, size=4000, dtype=np.int64), np.arange(0, 64), np.array([4 ** 11 - 1, 4 ** 15 - 1, 2 ** 32 - 1])])
    for parts in (1, 2, 3, 4, 7, 8, 16, 255):
        want = B.owner_of(codes, parts)
        got = np.array([L.gbin_owner_of(int(c), parts) for c in codes])
        np.testing.assert_array_equal(got, want)
        assert got.min() >= 0 and got.max() < parts
    assert L.gbin_owner_of(12345, 0) == 0
    dense = B.owner_of(np.arange(4 ** 9), 8)
    share = np.bincount(dense, minlength=8) / dense.size
    assert abs(share - 0.125).max() < 0.01


def _parse_expanded(text: bytes, K: int):
    """{(mmer, kmer): K id lines} from the print_kmer_read_ids layout after expand_read_id_list."""
    out = {}
    for block in text.split(b"\n\n"):
        lines = block.split(b"\n")
        if not block.strip():
            continue
        mmer, rest = lines[0], lines[1:]
        assert len(rest) % (K + 1) == 0, (mmer, len(rest))
        for j in range(0, len(rest), K + 1):
            key = (mmer, rest[j])
            assert key not in out
            out[key] = tuple(rest[j + 1: j + 1 + K])
    return out


@pytest.mark.parametrize("name", ["cfg1_reads", "cfg2_small"])
def test_expanded_dump_equals_the_references_own_expand_and_print(name, tmp_path):
    """gbin_table_dump_expanded_format against the UNMODIFIED reference running its own expand_read_id_list (binning.c:857-888)
    and print_kmer_read_ids (binning.c:792-823) on its own pruned table (oracle/_ref harness, --expanded): same buckets, same
    k-mers, the same K id lines per k-mer (the reference prints in hash-iteration order, so blocks are compared as sets)."""
    import os
    import subprocess
    case = next(c for c in CASES if c["name"] == name)
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref",
                       f"ref_K{case['k']}_M{case['m']}_C{case['cutoff']}_R{case['read_length_define']}")
    if not os.path.exists(exe):
        pytest.skip("reference harness not built")
    data = O.load_case_bytes(case)
    src = tmp_path / "reads.txt"
    src.write_bytes(data)
    ref = subprocess.run([exe, str(src), "--expanded"], capture_output=True, check=True).stdout
    starts, lens = O.fgets_split(data, case["read_length_define"])
    t = O.run(data, starts, lens, case["k"], case["m"], case["cutoff"])
    ct, keep = c_table_from_oracle(t)
    p = tmp_path / "expanded.txt"
    assert B.load_library().gbin_table_dump_expanded_format(C.byref(ct), str(p).encode()) == 0
    want, got = _parse_expanded(ref, case["k"]), _parse_expanded(p.read_bytes(), case["k"])
    assert len(want) == case["surviving_kmers"]
    assert got == want


@pytest.mark.parametrize("name", ["cfg2_small", "cfg4_small", "kat_twice"])
def test_table_digest_host_matches_its_python_restatement(name):
    """gbin_table_digest on a HOST table needs no GPU: same value as the Python restatement over the oracle's table, invariant
    under reordering of the buckets, and additive over a split of the buckets (what the multi-GPU owners' digests rely on)."""
    case = next(c for c in O.load_pins() if c["name"] == name)
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    t = O.run(data, starts, lens, case["k"], case["m"], case["cutoff"])
    L = B.load_library()
    ct, keep = c_table_from_oracle(t)
    d = C.c_uint64()
    assert L.gbin_table_digest(None, C.byref(ct), None, C.byref(d)) == 0
    assert d.value == O.table_digest(t)
    # split the buckets in two tables: the digests add up
    nb = t.n_buckets
    if nb >= 2:
        h = nb // 2
        s_h, n_h = int(t.mmer_kmer_off[h]), int(t.kmer_id_off[int(t.mmer_kmer_off[h])])
        kw = t.kw
        lo = O.Table(t.K, t.M, t.cutoff, 0, 0, t.mmer_codes[:h], t.mmer_kmer_off[:h + 1], t.kmer_codes[:s_h * kw], t.kmer_id_off[:s_h + 1], t.read_ids[:n_h])
        hi = O.Table(t.K, t.M, t.cutoff, 0, 0, t.mmer_codes[h:], t.mmer_kmer_off[h:] - np.uint64(s_h), t.kmer_codes[s_h * kw:],
                     t.kmer_id_off[s_h:] - np.uint64(n_h), t.read_ids[n_h:])
        tot = 0
        for part in (lo, hi):
            cp, keep2 = c_table_from_oracle(part)
            assert L.gbin_table_digest(None, C.byref(cp), None, C.byref(d)) == 0
            tot = (tot + d.value) & ((1 << 64) - 1)
        assert tot == O.table_digest(t)


@pytest.mark.parametrize("name", ["cfg1_reads", "kat_twice"])
def test_dump_from_expanded_lists_equals_the_expanded_dump(name, tmp_path):
    """Host half of row f3 (no GPU needed): gbin_table_dump_expanded_lists prints a list-of-lists object (what
    gbin_expand_read_ids_device + gbin_expanded_to_host deliver; built here with numpy as the CSR replicate of the oracle's table)
    byte for byte like gbin_table_dump_expanded_format, which the test above pins against the reference's own expand + print.
    A list-of-lists object of another batch (wrong list count) is refused."""
    case = next(c for c in CASES if c["name"] == name)
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    t = O.run(data, starts, lens, case["k"], case["m"], case["cutoff"])
    ct, keep = c_table_from_oracle(t)
    K = case["k"]
    off = np.asarray(t.kmer_id_off, np.int64)
    cnt = np.diff(off)
    ids = np.asarray(t.read_ids, np.int32)
    x_ids = np.concatenate([np.tile(ids[off[j]:off[j + 1]], K) for j in range(t.n_kmers)]).astype(np.int32) if t.n_kmers else np.zeros(1, np.int32)
    x_off = np.concatenate([[0], np.cumsum(np.repeat(cnt, K))]).astype(np.uint64)
    x = B.CExpanded(K, 0, t.n_kmers * K, int(x_off[-1]), x_off.ctypes.data, x_ids.ctypes.data)
    L = B.load_library()
    p1, p2 = tmp_path / "lists.txt", tmp_path / "format.txt"
    assert L.gbin_table_dump_expanded_lists(C.byref(ct), C.byref(x), str(p1).encode()) == 0
    assert L.gbin_table_dump_expanded_format(C.byref(ct), str(p2).encode()) == 0
    assert p1.read_bytes() == p2.read_bytes() and p1.stat().st_size > 0
    bad = B.CExpanded(K, 0, t.n_kmers * K + 1, int(x_off[-1]), x_off.ctypes.data, x_ids.ctypes.data)
    assert L.gbin_table_dump_expanded_lists(C.byref(ct), C.byref(bad), str(p1).encode()) == B.GBIN_E_INVALID_ARG
