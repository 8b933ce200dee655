"""Stage-by-stage check of pipeline 3 against the oracle on small inputs (debug aid, GPU box)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle_lib as O
from genome_assembly_b200 import binding as B, synth


def compare(got, want, tag):
    ok = True
    for name in ("mmer_codes", "mmer_kmer_off", "kmer_codes", "kmer_id_off", "read_ids"):
        a, b = getattr(got, name), getattr(want, name).reshape(-1)
        if a.shape != b.shape or not np.array_equal(a, b):
            ok = False
            n = min(len(a), len(b))
            bad = np.nonzero(a[:n] != b[:n])[0]
            print(f"  [{tag}] {name}: shapes {a.shape} vs {b.shape}, first mismatch at {bad[:5] if len(bad) else 'tail'}")
            if len(bad):
                i = int(bad[0])
                print("     got ", a[max(0, i - 2): i + 6])
                print("     want", b[max(0, i - 2): i + 6])
                print("     bad idx", bad[:40])
    print(f"  [{tag}] instances {got.n_instances}/{want.n_instances} distinct {got.n_distinct}/{want.n_distinct} kmers {got.n_kmers}/{want.n_kmers} -> {'OK' if ok else 'MISMATCH'}")
    return ok


def run_case(tag, buf, n, stride, L, K, M, cutoff, cap=1024, nc=0, h=0, pipeline=3):
    starts = np.arange(n, dtype=np.uint64) * stride
    lens = np.full(n, L, dtype=np.uint32)
    want = O.run(bytes(buf[: n * stride]), starts, lens, K, M, cutoff)
    b = B.Binner(K, M, cutoff, pipeline=pipeline)
    b.set_tuning("v3_cap", cap); b.set_tuning("v3_nc", nc); b.set_tuning("v3_h", h)
    t0 = time.time()
    got = b.bin_host(buf[: n * stride], n, stride=stride, read_len=L)
    dt = time.time() - t0
    info = b.pipeline_info(); st = b.run_stats()
    print(f"{tag}: K={K} M={M} cap={cap} nc={nc} h={h} n={n}: {dt*1e3:.1f} ms info={info} stats={st} err='{b.lib.gbin_last_error(b.h).decode()}'")
    ok = compare(got, want, tag)
    b.close()
    return ok


if __name__ == "__main__":
    allok = True
    if "--k64" in sys.argv:
        rs = synth.generate(1500, 200, genome_len=2000, error_rate=0.01, seed=6415, starts="uniform")
        for pl in (3, 2):
            run_case("k64", rs.buf, 1500, rs.stride, 200, 64, 15, 1, pipeline=pl)
        run_case("k64_512", rs.buf, 1500, rs.stride, 200, 64, 15, 1, cap=512)
        sys.exit(0)
    rs = synth.generate(3000, 100, error_rate=0.01, seed=3, starts="uniform")
    allok &= run_case("tiny", rs.buf, 200, rs.stride, 100, 31, 11, 1)
    allok &= run_case("small", rs.buf, 3000, rs.stride, 100, 31, 11, 1)
    allok &= run_case("small512", rs.buf, 3000, rs.stride, 100, 31, 11, 1, cap=512)
    allok &= run_case("m4", rs.buf, 3000, rs.stride, 100, 31, 4, 1)
    allok &= run_case("m4_512", rs.buf, 3000, rs.stride, 100, 31, 4, 1, cap=512)
    allok &= run_case("cut-1", rs.buf, 3000, rs.stride, 100, 31, 11, -1)
    rs2 = synth.generate(2500, 100, genome_len=2000, error_rate=0.002, seed=77, starts="uniform")
    allok &= run_case("deep", rs2.buf, 2500, rs2.stride, 100, 31, 11, 1)
    allok &= run_case("deep63", rs2.buf, 2500, rs2.stride, 100, 63, 15, 1)
    rs3 = synth.generate(60000, 100, error_rate=0.01, seed=20, starts="triangular")
    allok &= run_case("medium", rs3.buf, 60000, rs3.stride, 100, 31, 11, 1)
    allok &= run_case("medium512", rs3.buf, 60000, rs3.stride, 100, 31, 11, 1, cap=512)
    allok &= run_case("nc2", rs.buf, 3000, rs.stride, 100, 31, 11, 1, nc=2, h=0)
    allok &= run_case("nc2h2", rs3.buf, 60000, rs3.stride, 100, 31, 11, 1, nc=2, h=2)
    allok &= run_case("nc2h4_512", rs3.buf, 60000, rs3.stride, 100, 31, 11, 1, nc=2, h=4, cap=512)
    allok &= run_case("deep_nc2", rs2.buf, 2500, rs2.stride, 100, 31, 11, 1, nc=2, h=1)
    allok &= run_case("deep63_nc2", rs2.buf, 2500, rs2.stride, 100, 63, 15, 1, nc=2, h=0)
    allok &= run_case("m4_nc2", rs.buf, 3000, rs.stride, 100, 31, 4, 1, nc=2, h=3)
    rs4 = synth.generate(150000, 150, error_rate=0.01, seed=31, starts="uniform")
    allok &= run_case("m7_auto", rs4.buf, 150000, rs4.stride, 150, 25, 7, 1, nc=0)
    allok &= run_case("m7_auto512", rs4.buf, 150000, rs4.stride, 150, 25, 7, 1, nc=0, cap=512)
    print("ALL OK" if allok else "FAILURES")
    sys.exit(0 if allok else 1)
