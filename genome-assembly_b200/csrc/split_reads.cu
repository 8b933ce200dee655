// split_reads.cu — main's read loop (binning.c:1154-1166) on the device: the file image is split into reads exactly the way
//   while (fgets(read, READ_LENGTH, file)) { len = strlen(read); read[--len] = '\0'; process_read(table, read, read_id++); }
// does it.  fgets stores at most READ_LENGTH-1 bytes and stops after a newline, the loop then drops the last byte whatever
// it is, and every fgets return owns a read id.  Per line (its bytes up to and including the newline, or up to the end of
// the file) of length l that makes ceil(l / (READ_LENGTH-1)) reads: piece k covers bytes [k*(R-1), min((k+1)*(R-1), l)) of the
// line and loses its last byte — on the bundled reads.txt (100 bases + newline, READ_LENGTH 101) every line becomes a
// 99-base read and an empty read with an id of its own.
// Kernels: newline count per 8 KB tile -> scan of the tile counts -> newline positions (stream compaction, positions in file
// order) -> reads per line -> scan -> starts / lens.
#include "gbin_internal.h"
#include "prefix_scan.cuh"

namespace gbin {

constexpr int SPL_THREADS = 256;
constexpr int SPL_BYTES = 32;                       // bytes per thread
constexpr int SPL_TILE = SPL_THREADS * SPL_BYTES;   // bytes per block

// 0x80 in every byte of w that equals '\n' (exact: no carries between bytes)
__device__ __forceinline__ uint32_t newline_mask(uint32_t w) {
    const uint32_t x = w ^ 0x0a0a0a0au;
    const uint32_t t = (x & 0x7f7f7f7fu) + 0x7f7f7f7fu;
    return ~(t | x | 0x7f7f7f7fu);
}

// The 32 bytes of thread `tid` of tile `tile` as 8 words (zero past the end of the data: zero is not a newline).
__device__ __forceinline__ void load_span(const uint8_t *__restrict__ data, uint64_t n, uint64_t first, uint32_t (&w)[8]) {
    if (first + SPL_BYTES <= n && ((reinterpret_cast<uintptr_t>(data) + first) & 15u) == 0) {
        const uint4 a = *reinterpret_cast<const uint4 *>(data + first), b = *reinterpret_cast<const uint4 *>(data + first + 16);
        w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
        w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            uint32_t v = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const uint64_t p = first + 4 * i + j;
                if (p < n) v |= (uint32_t)data[p] << (8 * j);
            }
            w[i] = v;
        }
    }
}

__global__ void __launch_bounds__(SPL_THREADS)
    count_newlines_kernel(const uint8_t *__restrict__ data, uint64_t n, uint32_t *__restrict__ tile_counts) {
    uint32_t w[8];
    load_span(data, n, (uint64_t)blockIdx.x * SPL_TILE + (uint64_t)threadIdx.x * SPL_BYTES, w);
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) c += __popc(newline_mask(w[i]));
    uint32_t total;
    (void)block_exclusive_scan<uint32_t>(c, &total);
    if (threadIdx.x == 0) tile_counts[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SPL_THREADS)
    emit_newlines_kernel(const uint8_t *__restrict__ data, uint64_t n, const uint32_t *__restrict__ tile_base, uint64_t *__restrict__ nl) {
    uint32_t w[8];
    const uint64_t first = (uint64_t)blockIdx.x * SPL_TILE + (uint64_t)threadIdx.x * SPL_BYTES;
    load_span(data, n, first, w);
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) c += __popc(newline_mask(w[i]));
    uint32_t total;
    uint64_t at = (uint64_t)tile_base[blockIdx.x] + block_exclusive_scan<uint32_t>(c, &total);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t m = newline_mask(w[i]);
        while (m) {
            const int bit = __ffs(m) - 1;  // 7, 15, 23 or 31: byte bit / 8 of word i (little endian: lowest address first)
            nl[at++] = first + 4 * i + (bit >> 3);
            m &= m - 1;
        }
    }
}

struct LineView {
    const uint64_t *nl;  // newline positions, ascending
    uint64_t n_nl, size;
    uint32_t cap;        // READ_LENGTH - 1: bytes one fgets call can take
    __device__ __forceinline__ void span(uint64_t j, uint64_t *start, uint64_t *len) const {
        const uint64_t s = j ? nl[j - 1] + 1 : 0;
        const uint64_t e = j < n_nl ? nl[j] + 1 : size;  // one past the line's last byte (its newline, or the end of the file)
        *start = s;
        *len = e - s;
    }
};
struct ReadsOfLine {
    LineView lv;
    __device__ __forceinline__ uint32_t operator()(uint64_t j) const {
        uint64_t s, l;
        lv.span(j, &s, &l);
        return (uint32_t)((l + lv.cap - 1) / lv.cap);
    }
};

__global__ void emit_reads_kernel(LineView lv, uint64_t n_lines, const uint32_t *__restrict__ read_base, uint64_t *__restrict__ starts,
                                  uint32_t *__restrict__ lens) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_lines) return;
    uint64_t s, l;
    lv.span(j, &s, &l);
    uint64_t r = read_base[j];
    for (uint64_t off = 0; off < l; off += lv.cap, r++) {
        const uint64_t take = l - off < lv.cap ? l - off : lv.cap;
        starts[r] = s + off;
        lens[r] = (uint32_t)(take - 1);  // read[--len] = '\0'
    }
}

uint32_t split_tiles(uint64_t n) { return (uint32_t)((n + SPL_TILE - 1) / SPL_TILE); }

// Stage 1: newline positions.  tile_counts: [tiles + scan scratch] u32; *n_nl_dev receives the number of newlines.
int split_find_newlines(const uint8_t *data, uint64_t n, uint32_t *tile_counts, uint32_t *scan_scratch, uint64_t *nl, uint32_t *n_nl_dev,
                        bool emit, cudaStream_t st) {
    const uint32_t tiles = split_tiles(n);
    if (tiles == 0) return 0;
    if (!emit) {
        count_newlines_kernel<<<tiles, SPL_THREADS, 0, st>>>(data, n, tile_counts);
        return 1 + exclusive_scan<uint32_t, PtrIn<uint32_t>>(PtrIn<uint32_t>{tile_counts}, tile_counts, tiles, scan_scratch, n_nl_dev, st);
    }
    emit_newlines_kernel<<<tiles, SPL_THREADS, 0, st>>>(data, n, tile_counts, nl);
    return 1;
}

// Stage 2: reads per line -> read_base (exclusive), *n_reads_dev; then starts / lens.
int split_count_reads(const uint64_t *nl, uint64_t n_nl, uint64_t size, uint32_t cap, uint64_t n_lines, uint32_t *read_base, uint32_t *scan_scratch,
                      uint32_t *n_reads_dev, cudaStream_t st) {
    return exclusive_scan<uint32_t, ReadsOfLine>(ReadsOfLine{LineView{nl, n_nl, size, cap}}, read_base, n_lines, scan_scratch, n_reads_dev, st);
}
int split_emit_reads(const uint64_t *nl, uint64_t n_nl, uint64_t size, uint32_t cap, uint64_t n_lines, const uint32_t *read_base, uint64_t *starts,
                     uint32_t *lens, cudaStream_t st) {
    if (n_lines == 0) return 0;
    emit_reads_kernel<<<(unsigned)((n_lines + 255) / 256), 256, 0, st>>>(LineView{nl, n_nl, size, cap}, n_lines, read_base, starts, lens);
    return 1;
}

}  // namespace gbin
