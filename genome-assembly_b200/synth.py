"""Synthetic read sets for the binning hot path.

Restates the *distribution* of the reference's generate_reads.py (generate_reads.py:93-112):
an i.i.d. uniform genome over ACGT, fixed-length forward-strand substrings, one read per
``\\n``-terminated line.  Two things the script lacks are added, as SURVEY.md §8(d) prescribes:
the seed is applied *before* the genome is drawn (the script seeds after, generate_reads.py:96-97,
so its genome is not reproducible), and i.i.d. per-base substitution to one of the other three
bases at rate ``error_rate`` (the script has no error model).

``starts="triangular"`` replays the script's chained ``random.triangular(0, G-1-L, mode)`` walk
(generate_reads.py:98-105) with Python's ``random`` module; it is inherently sequential, so it is
used for the 1 M-read configuration only.  ``starts="uniform"`` draws i.i.d. uniform starts from a
counter-based numpy generator (Philox) and is used for the larger configurations.
"""
from __future__ import annotations

import random
from dataclasses import dataclass

import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class ReadSet:
    """Fixed-stride ASCII reads: row i is ``buf[i*stride : i*stride+read_len]`` followed by ``\\n``."""

    buf: np.ndarray  # uint8 [n_reads * stride]
    n_reads: int
    read_len: int
    stride: int
    genome_len: int
    seed: int
    error_rate: float
    starts_kind: str

    def as_bytes(self) -> bytes:
        return self.buf.tobytes()

    def rows(self) -> np.ndarray:
        return self.buf.reshape(self.n_reads, self.stride)[:, : self.read_len]


def default_genome_len(n_reads: int, read_len: int, coverage: float = 30.0) -> int:
    """genome length = n_reads * L / 30 (about 30x coverage)."""
    return max(read_len + 2, int(n_reads * read_len / coverage))


def generate(n_reads: int, read_len: int, *, genome_len: int | None = None, error_rate: float = 0.01,
             seed: int = 20, starts: str = "uniform", chunk: int = 1 << 18, read_seed: int | None = None) -> ReadSet:
    """``seed`` fixes the genome; ``read_seed`` (default: ``seed``) fixes starts and substitutions, so that
    several ranks can draw disjoint read shards from one shared genome."""
    if genome_len is None:
        genome_len = default_genome_len(n_reads, read_len)
    if genome_len < read_len + 2:
        raise ValueError("genome shorter than a read")
    rng = np.random.Generator(np.random.Philox(seed))
    genome = _ACGT[rng.integers(0, 4, size=genome_len, dtype=np.uint8)]
    hi = genome_len - 1 - read_len  # generate_reads.py:98
    if read_seed is not None and read_seed != seed:
        rng = np.random.Generator(np.random.Philox(key=read_seed + (1 << 40)))
    else:
        read_seed = seed
    if starts == "triangular":
        r = random.Random(read_seed)
        mode = r.randint(0, hi)
        pos = np.empty(n_reads, dtype=np.int64)
        for i in range(n_reads):
            mode = int(r.triangular(0, hi, mode))
            pos[i] = mode
    elif starts == "uniform":
        pos = rng.integers(0, hi + 1, size=n_reads, dtype=np.int64)
    else:
        raise ValueError(starts)
    stride = read_len + 1
    buf = np.empty(n_reads * stride, dtype=np.uint8)
    view = buf.reshape(n_reads, stride)
    view[:, read_len] = ord("\n")
    ar = np.arange(read_len, dtype=np.int64)
    for lo in range(0, n_reads, chunk):
        hi_i = min(n_reads, lo + chunk)
        idx = pos[lo:hi_i, None] + ar[None, :]
        bases = genome[idx]
        if error_rate > 0:
            hit = rng.random(bases.shape, dtype=np.float32) < error_rate
            nhit = int(hit.sum())
            if nhit:
                # substitute with one of the three *other* bases: rotate the ACGT index by 1..3
                lut = np.zeros(256, dtype=np.uint8)
                lut[_ACGT] = np.arange(4, dtype=np.uint8)
                cur = lut[bases[hit]]
                rot = rng.integers(1, 4, size=nhit, dtype=np.uint8)
                bases[hit] = _ACGT[(cur + rot) & 3]
        view[lo:hi_i, :read_len] = bases
    return ReadSet(buf=buf, n_reads=n_reads, read_len=read_len, stride=stride, genome_len=genome_len,
                   seed=seed, error_rate=error_rate, starts_kind=starts)


def make_genome(genome_len: int, seed: int = 20) -> np.ndarray:
    """The shared genome of generate(seed=seed, genome_len=genome_len)."""
    rng = np.random.Generator(np.random.Philox(seed))
    return _ACGT[rng.integers(0, 4, size=genome_len, dtype=np.uint8)]


def reads_from_genome(genome: np.ndarray, n_reads: int, read_len: int, *, error_rate: float, read_seed: int, out: np.ndarray | None = None,
                      chunk: int = 1 << 18) -> np.ndarray:
    """n_reads fixed-stride reads (uniform starts, i.i.d. substitutions) drawn from `genome` with a generator that depends on
    read_seed alone — so a read set cut into blocks with read_seed = block index is the same whatever process draws which block."""
    rng = np.random.Generator(np.random.Philox(key=read_seed + (1 << 40)))
    hi = len(genome) - 1 - read_len
    pos = rng.integers(0, hi + 1, size=n_reads, dtype=np.int64)
    stride = read_len + 1
    buf = out if out is not None else np.empty(n_reads * stride, dtype=np.uint8)
    view = buf.reshape(n_reads, stride)
    view[:, read_len] = ord("\n")
    ar = np.arange(read_len, dtype=np.int64)
    lut = np.zeros(256, dtype=np.uint8)
    lut[_ACGT] = np.arange(4, dtype=np.uint8)
    for lo in range(0, n_reads, chunk):
        hi_i = min(n_reads, lo + chunk)
        bases = genome[pos[lo:hi_i, None] + ar[None, :]]
        if error_rate > 0:
            hit = rng.random(bases.shape, dtype=np.float32) < error_rate
            nhit = int(hit.sum())
            if nhit:
                cur = lut[bases[hit]]
                rot = rng.integers(1, 4, size=nhit, dtype=np.uint8)
                bases[hit] = _ACGT[(cur + rot) & 3]
        view[lo:hi_i, :read_len] = bases
    return buf


def blocked_reads_torch(genome_len: int, n_blocks: int, first_block: int, block_reads: int, read_len: int, *, error_rate: float, seed: int, device):
    """The same kind of read set as reads_from_genome, drawn on the GPU with torch generators (one per block, seeded by the block
    index; the genome with `seed`): a 100 M-read set takes seconds instead of minutes.  Returns a uint8 device tensor
    [n_blocks * block_reads * (read_len + 1)].  The values differ from the numpy generators' (another RNG) — what matters is that
    block b holds the same reads whichever process draws it."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    genome = torch.randint(0, 4, (genome_len,), generator=g, device=device, dtype=torch.uint8)  # codes 0..3 = A, C, G, T
    stride = read_len + 1
    out = torch.empty(n_blocks * block_reads * stride, dtype=torch.uint8, device=device)
    ar = torch.arange(read_len, device=device, dtype=torch.int64)
    hi = genome_len - 1 - read_len
    for k in range(n_blocks):
        g.manual_seed((1 << 40) + first_block + k)
        pos = torch.randint(0, hi + 1, (block_reads,), generator=g, device=device, dtype=torch.int64)
        codes = genome[pos[:, None] + ar[None, :]]
        if error_rate > 0:
            hit = torch.rand(codes.shape, generator=g, device=device) < error_rate
            rot = torch.randint(1, 4, codes.shape, generator=g, device=device, dtype=torch.uint8)
            codes = torch.where(hit, (codes + rot) & 3, codes)
        view = out[k * block_reads * stride:(k + 1) * block_reads * stride].view(block_reads, stride)
        view[:, :read_len] = acgt[codes.long()]
        view[:, read_len] = ord("\n")
    return out


# BASELINE.json configs 2-5 (config 1 is the bundled reads.txt)
WORKLOADS = {
    "cfg2": dict(n_reads=1_000_000, read_len=100, k=31, m=11, cutoff=1, error_rate=0.01, starts="triangular"),
    "cfg3": dict(n_reads=100_000_000, read_len=150, k=31, m=11, cutoff=1, error_rate=0.01, starts="uniform"),
    "cfg4": dict(n_reads=10_000_000, read_len=250, k=63, m=15, cutoff=1, error_rate=0.01, starts="uniform"),
    "cfg5": dict(n_reads=50_000_000, read_len=150, k=25, m=9, cutoff=1, error_rate=0.05, starts="uniform"),
}
