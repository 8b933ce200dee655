"""CPU: the restated oracle against the reference-execution pins in tests/golden/pins.json
(each pin = md5 of the sorted table dump of the UNMODIFIED reference binary, oracle/_ref)."""
import gzip
import os

import numpy as np
import pytest

import oracle_lib as O

CASES = O.load_pins()


def test_scalar_kats():
    """SURVEY §4.3 scalar known answers (getscore / getval / getbp, binning.c:69-124)."""
    L = O.lib()
    for s, v in [(b"AAGTCC", 3914), (b"TTCAGG", 181), (b"AAAA", 255), (b"CTTT", 128), (b"AACA", 251), (b"CAGA", 183)]:
        assert L.orc_getscore(s) == v
    for c in b"Na\n":
        assert L.orc_getval(bytes([c])) == 3
    assert [L.orc_getval(c) for c in (b"T", b"G", b"C", b"A")] == [0, 1, 2, 3]
    assert L.orc_getbp(4) == b"A" and L.orc_getbp(-1) == b"A"
    assert b"".join(L.orc_getbp(i) for i in range(4)) == b"TGCA"


def test_fgets_quirk_on_bundled_fixture():
    """binning.c:1154-1166 with READ_LENGTH 101 on 100-char lines: even ids = 99-bp reads, odd ids empty."""
    case = next(c for c in CASES if c["name"] == "cfg1_reads")
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, 101)
    assert len(starts) == 10000 == case["read_ids"]
    assert (lens[0::2] == 99).all() and (lens[1::2] == 0).all()


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference_pin(case):
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    assert len(starts) == case["read_ids"]
    md5s, lines = O.dump_strings_md5(data, starts, lens, case["k"], case["m"], case["cutoff"])
    assert md5s == case["md5"], "string-faithful oracle differs from the reference pin"
    assert len(lines) == case["surviving_kmers"]
    if case["acgt_only"]:
        t = O.run(data, starts, lens, case["k"], case["m"], case["cutoff"])
        assert t.n_instances == case["instances"]
        assert t.n_kmers == case["surviving_kmers"]
        assert t.n_buckets == case["surviving_buckets"]
        assert t.md5() == case["md5"], "code-space oracle differs from the reference pin"


def test_kat_dump_lines():
    case = next(c for c in CASES if c["name"] == "kat_twice")
    with gzip.open(os.path.join(O.GOLDEN, case["dump"]), "rb") as f:
        want = f.read().split(b"\n")[:-1]
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, 101)
    t = O.run(data, starts, lens, 31, 4, 1)
    assert sorted(t.dump_lines()) == want
    assert b"AACA GTGTCCTCCCTCGGCTAATCATGAACACCGG 1 0" in want
    assert all(l.startswith(b"AACA ") and l.endswith(b" 1 0") for l in want) and len(want) == 10


def test_scan_trace_is_consistent():
    """orc_scan_all's per-window trace satisfies the closed form of SURVEY §8(a) (K >= 2M):
    sig(i) = sig(i-1) if sig(i-1) >= i else leftmost argmax of w over [i, i+K-M]."""
    case = next(c for c in CASES if c["name"] == "cfg2_small")
    data = O.load_case_bytes(case)
    starts, lens = O.fgets_split(data, case["read_length_define"])
    K, M = case["k"], case["m"]
    tup, win = O.scan_all(data, starts[:50], lens[:50], K, M)
    lut = np.full(256, 3, dtype=np.int64)
    lut[ord("T")], lut[ord("G")], lut[ord("C")] = 0, 1, 2
    pos = 0
    for r in range(50):
        v = lut[np.frombuffer(data[int(starts[r]):int(starts[r]) + int(lens[r])], dtype=np.uint8)]
        L = len(v)
        s = np.array([sum(int(v[p + t]) << (2 * (M - 1 - t)) for t in range(M)) for p in range(L - M + 1)])
        w = np.maximum(s, 4 ** M - 1 - s)
        sig = -1
        for i in range(L - K + 1):
            if sig < i:
                sig = i + int(np.argmax(w[i:i + K - M + 1]))
            assert win["sig_pos"][pos] == sig and win["mmer"][pos] == w[sig] == tup["mmer"][pos]
            assert win["is_rev"][pos] == int(4 ** M - 1 - s[sig] > s[sig])
            pos += 1
    assert pos == len(tup)


def _split_per_line(data: bytes, read_length_define: int):
    """The per-line form of main's fgets loop that csrc/split_reads.cu evaluates in parallel: a line of l bytes (newline
    included, or up to the end of the file) makes ceil(l / (R-1)) reads, piece k covering bytes [k(R-1), min((k+1)(R-1), l))
    of the line minus its last byte."""
    cap = read_length_define - 1
    a = np.frombuffer(data, dtype=np.uint8)
    nl = np.flatnonzero(a == 10)
    ends = nl + 1
    if len(a) and (len(nl) == 0 or nl[-1] != len(a) - 1):
        ends = np.append(ends, len(a))
    begins = np.concatenate([[0], ends[:-1]]) if len(ends) else np.zeros(0, np.int64)
    starts, lens = [], []
    for s0, e0 in zip(begins.tolist(), ends.tolist()):
        for off in range(0, e0 - s0, cap):
            starts.append(s0 + off)
            lens.append(min(cap, e0 - s0 - off) - 1)
    return np.array(starts, dtype=np.uint64), np.array(lens, dtype=np.uint32)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_per_line_split_equals_sequential_fgets(case):
    """The decomposition the device-side splitter relies on (independent lines) gives exactly what the sequential replay of
    fgets gives — on every fixture, for READ_LENGTH around and far from the line length, and on random bytes."""
    data = O.load_case_bytes(case)[:200_000]
    rng = np.random.default_rng(len(data))
    noise = rng.choice(np.frombuffer(b"ACGT\nN", dtype=np.uint8), size=5000, p=[0.24, 0.24, 0.24, 0.24, 0.03, 0.01]).tobytes()
    for blob in (data, noise, noise + b"ACGT", b"", b"\n", b"A", b"\n\n\nAC"):
        for R in sorted({case["read_length_define"], 2, 3, 5, 64, 100, 101, 102, 4096}):
            want_s, want_l = O.fgets_split(blob, R)
            got_s, got_l = _split_per_line(blob, R)
            np.testing.assert_array_equal(got_s, want_s)
            np.testing.assert_array_equal(got_l, want_l)
