#!/usr/bin/env python
"""Generates tests/golden/*: small input files and pins.json.

Every pin is the md5 of the LC_ALL=C-sorted table dump ("<mmer> <kmer> <ids...>" per line) printed
by the UNMODIFIED reference hot path (oracle/_ref/ref_K*_M*_C*_R*, built by oracle/build_ref.sh from
/root/reference).  Run in the authoring container only (needs /root/reference):
    make -C oracle && python tests/golden/make_golden.py
"""
import gzip
import hashlib
import json
import os
import random
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from genome_assembly_b200 import synth  # noqa: E402

REFDIR = os.path.join(ROOT, "oracle", "_ref")


def ref_dump(path, k, m, c, r):
    exe = os.path.join(REFDIR, f"ref_K{k}_M{m}_C{c}_R{r}")
    p = subprocess.run([exe, path], capture_output=True, check=True)
    lines = sorted(p.stdout.split(b"\n")[:-1])
    stats = json.loads(p.stderr.decode().strip().splitlines()[-1])
    blob = b"".join(x + b"\n" for x in lines)
    return hashlib.md5(blob).hexdigest(), stats, blob


def write_gz(name, data: bytes):
    with open(os.path.join(HERE, name), "wb") as raw:
        with gzip.GzipFile(fileobj=raw, mode="wb", mtime=0, filename="") as f:
            f.write(data)


def ragged(rng, n, maxlen, alphabet, newline_at_end=True, genome_len=300):
    """Lines are substrings of a small random genome (so k-mers repeat and survive the prune),
    with lengths that straddle K, the fgets buffer and zero."""
    genome = "".join(rng.choice(alphabet) for _ in range(genome_len))
    out = []
    for _ in range(n):
        ln = rng.choice([0, 1, 5, 14, 15, 16, 24, 25, 26, 30, 31, 40, 62, 63, 64, 65, 90, 130, rng.randint(0, maxlen)])
        st = rng.randint(0, genome_len - 1)
        out.append((genome + genome)[st:st + ln])
    s = "\n".join(out)
    if newline_at_end:
        s += "\n"
    return s.encode()


def main():
    cases = []

    def add(name, data, k, m, c, r, note, acgt_only=True, keep_dump=False):
        fn = f"{name}.txt.gz"
        write_gz(fn, data)
        tmp = os.path.join("/tmp", f"golden_{name}.txt")
        with open(tmp, "wb") as f:
            f.write(data)
        md5, stats, blob = ref_dump(tmp, k, m, c, r)
        case = dict(name=name, file=fn, k=k, m=m, cutoff=c, read_length_define=r, md5=md5, acgt_only=acgt_only,
                    note=note, **{q: stats[q] for q in ("read_ids", "instances", "surviving_kmers", "surviving_buckets")})
        if keep_dump:
            write_gz(f"{name}.dump.gz", blob)
            case["dump"] = f"{name}.dump.gz"
        cases.append(case)
        print(case)

    # config 1: the bundled fixture, exactly as main reads it (READ_LENGTH 101 => last base chopped, empty odd ids)
    with open("/root/reference/reads.txt", "rb") as f:
        bundled = f.read()
    add("cfg1_reads", bundled, 31, 4, 1, 101, "BASELINE config 1: bundled reads.txt, makefile defaults")
    cases[-1]["file_is_reference_fixture"] = True
    # same input, higher cutoff
    write = cases  # noqa
    tmp = "/tmp/golden_cfg1_reads.txt"
    md5, stats, _ = ref_dump(tmp, 31, 4, 3, 101)
    cases.append(dict(name="cfg1_reads_cut3", file="cfg1_reads.txt.gz", k=31, m=4, cutoff=3, read_length_define=101, md5=md5,
                      acgt_only=True, note="config 1 input with ABUNDANCE_CUTOFF 3",
                      **{q: stats[q] for q in ("read_ids", "instances", "surviving_kmers", "surviving_buckets")}))

    kat = b"GTGTCCTCCCTCGGCTAATCATGAACACCGGTCAGGCATG\n" * 2
    add("kat_twice", kat, 31, 4, 1, 101, "SURVEY §4.3 KAT: one 40-bp read fed twice, all 10 k-mers in bucket AACA with list '1 0'",
        keep_dump=True)

    # scaled-down versions of BASELINE configs 2-5 (same K/M/cutoff/L/error model)
    rs = synth.generate(3000, 100, error_rate=0.01, seed=20, starts="triangular")
    add("cfg2_small", rs.as_bytes(), 31, 11, 1, 102, "config 2 shape: 3000 x 100 bp, 1% subs, triangular starts")
    rs = synth.generate(2000, 150, error_rate=0.01, seed=21, starts="uniform")
    add("cfg3_small", rs.as_bytes(), 31, 11, 1, 152, "config 3 shape: 2000 x 150 bp, 1% subs")
    rs = synth.generate(1000, 250, error_rate=0.01, seed=22, starts="uniform")
    add("cfg4_small", rs.as_bytes(), 63, 15, 1, 252, "config 4 shape: 1000 x 250 bp, K=63 (128-bit codes), M=15")
    rs = synth.generate(2000, 150, error_rate=0.05, seed=23, starts="uniform")
    add("cfg5_small", rs.as_bytes(), 25, 9, 1, 152, "config 5 shape: 2000 x 150 bp, 5% subs, K=25 M=9")

    # ragged / edge inputs through main's fgets loop (READ_LENGTH 64: long lines are split, chops happen)
    rng = random.Random(7)
    add("fuzz_ragged_k25", ragged(rng, 1500, 140, "ACGT"), 25, 9, 2, 64, "ragged lengths incl. <K, =K, >buffer; cutoff 2")
    add("fuzz_ragged_k15", ragged(rng, 1500, 140, "AC", newline_at_end=False), 15, 5, 1, 64,
        "2-letter alphabet => heavy duplication and ties; file lacks the trailing newline")
    add("fuzz_polyA", (b"A" * 63 + b"\n") * 40 + (b"T" * 50 + b"\n") * 30 + ragged(rng, 200, 100, "AT"), 15, 5, 1, 64,
        "homopolymers: every window ties, leftmost-max rule and huge single groups")
    add("fuzz_nonacgt", ragged(rng, 800, 140, "ACGTNacgt"), 15, 5, 1, 64,
        "non-ACGT bytes: oracle string mode only (outside the 2-bit contract of the GPU path)", acgt_only=False)
    add("short_reads", ragged(rng, 300, 30, "ACGT"), 31, 4, 1, 40, "lines split into <=38-char reads by a 40-byte fgets buffer; many reads shorter than K")
    add("k_lt_2m", ragged(rng, 500, 60, "ACGT"), 6, 4, 1, 64,
        "K < 2M: the reference's else-branch loop (binning.c:997) is live; oracle only, GPU path rejects K < 2M")

    with open(os.path.join(HERE, "pins.json"), "w") as f:
        json.dump(dict(generated_by="tests/golden/make_golden.py", reference="twitu/genome-assembly binning.c via oracle/_ref",
                       sort="LC_ALL=C sort of dump lines, md5 of the result", cases=cases), f, indent=1)


if __name__ == "__main__":
    main()
