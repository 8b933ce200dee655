"""GPU: the link-time substitution of INTEGRATION.md, end to end.  oracle/_ref/dropin_* is the reference's main() flow
(binning.c:1147-1181) with process_read / prune_data bound to libgbin.so and every downstream function
(expand_read_id_list, find_kmer_extensions, print_kmers, iterators, zhash, llist) being unmodified reference code;
oracle/_ref/stock_* is the reference program as shipped.  Both are prebuilt by oracle/build_dropin.sh."""
import gzip
import os
import subprocess

import pytest

import oracle_lib as O

pytestmark = pytest.mark.gpu
REFDIR = os.path.join(O.ROOT, "oracle", "_ref")


def run(exe, path, *extra):
    p = subprocess.run([os.path.join(REFDIR, exe), path, *extra], capture_output=True, timeout=300)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    return sorted(p.stdout.split(b"\n")[:-1])


def need(*names):
    for n in names:
        if not os.path.exists(os.path.join(REFDIR, n)):
            pytest.skip(f"{n} not built (oracle/build_dropin.sh needs /root/reference)")


def test_dropin_program_prints_what_the_reference_program_prints(tmp_path):
    """K=31, M=11: a.out's stdout (print_kmers after expand_read_id_list + find_kmer_extensions) from the stock binary and
    from the binary whose hot path runs on the GPU are the same multiset of lines."""
    need("stock_K31_M11_C1_R102", "dropin_K31_M11_C1_R102")
    case = next(c for c in O.load_pins() if c["name"] == "cfg2_small")
    path = tmp_path / "reads.txt"
    path.write_bytes(O.load_case_bytes(case))
    stock = run("stock_K31_M11_C1_R102", str(path))
    dropin = run("dropin_K31_M11_C1_R102", str(path))
    assert len(stock) == case["surviving_kmers"]
    assert dropin == stock


def test_dropin_table_feeds_unmodified_downstream_code_on_the_bundled_fixture(tmp_path):
    """Config 1 (bundled reads.txt, K=31, M=4): right after prune_data the reference's own print_kmers walks the table the
    GPU built and prints exactly the k-mers of the reference's table; the unitig phase (order dependent on hash layout for
    M=4) then runs to completion on it."""
    need("stock_K31_M4_C1_R101", "dropin_K31_M4_C1_R101")
    case = next(c for c in O.load_pins() if c["name"] == "cfg1_reads")
    data = O.load_case_bytes(case)
    path = tmp_path / "reads.txt"
    path.write_bytes(data)
    table_only = run("dropin_K31_M4_C1_R101", str(path), "--table-only")
    starts, lens = O.fgets_split(data, 101)
    want = sorted(l.split(b" ")[1] for l in O.run(data, starts, lens, 31, 4, 1).dump_lines())
    assert table_only == want and len(want) == case["surviving_kmers"]
    full = run("dropin_K31_M4_C1_R101", str(path))
    stock = run("stock_K31_M4_C1_R101", str(path))
    assert len(full) > 0 and len(stock) > 0
    # every printed line is a string over ACGT at least K long (k-mers or merged unitigs)
    assert all(len(l) >= 31 and set(l) <= set(b"ACGT") for l in full)
