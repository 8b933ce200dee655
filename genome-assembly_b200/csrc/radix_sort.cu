// radix_sort.cu — stable LSD radix sort of k-mer instance records by (mmer, kmer), 8 bits per pass.
//
// This replaces the reference's per-k-mer hash chaining (zhash_get/zhash_set + strcmp along the
// chain, binning.c:1044-1058, zhash.c:53-93): equal (m-mer, k-mer) keys become adjacent, and because
// every pass is stable and the scan stage emits records in arrival order, the records of one key
// stay in arrival order — which is what the linked list's head-insert order encodes
// (binning.c:1059-1069), read backwards.
//
// Per pass: (1) per-tile digit histogram, (2) exclusive scan of the [256][tiles] table,
// (3) stable scatter: warp-level match ranking, tile-local reorder through shared memory so that
// every digit run leaves the SM as one contiguous, coalesced store.
// The same machinery with digit = mmer % n_parts is the stable owner partition of the multi-GPU path.
#include "gbin_device.cuh"
#include "gbin_internal.h"
#include "prefix_scan.cuh"

namespace gbin {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_THREADS * RS_ITEMS;  // 4096 records per tile
constexpr int RS_RADIX = 256;

struct DigitSel {
    int mode;  // 0: k-mer word `word` >> shift ; 1: mmer >> shift ; 2: mmer % nparts
    int word;
    int shift;
    uint32_t nparts;
};

template <int KW>
__device__ __forceinline__ uint32_t digit_of(const Rec<KW> &r, const DigitSel &s) {
    if (s.mode == 0) return (uint32_t)((s.word == 0 ? r.k[0] : r.k[KW - 1]) >> s.shift) & 0xffu;  // no dynamic indexing: keeps records in registers
    if (s.mode == 1) return (r.mmer >> s.shift) & 0xffu;
    return r.mmer % s.nparts;
}

template <int KW>
__global__ void __launch_bounds__(RS_THREADS)
    radix_hist_kernel(const Rec<KW> *__restrict__ in, uint64_t n, DigitSel sel, uint32_t *__restrict__ tile_hist, uint32_t ntiles) {
    __shared__ uint32_t h[RS_RADIX];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int i = 0; i < RS_ITEMS; i++) {
        const uint64_t j = base + (uint64_t)i * RS_THREADS + threadIdx.x;
        if (j < n) {
            const Rec<KW> r = load_rec<KW>(in + j);
            atomicAdd(&h[digit_of<KW>(r, sel)], 1u);
        }
    }
    __syncthreads();
    tile_hist[(uint64_t)threadIdx.x * ntiles + blockIdx.x] = h[threadIdx.x];
}

template <int KW>
__global__ void __launch_bounds__(RS_THREADS)
    radix_scatter_kernel(const Rec<KW> *__restrict__ in, Rec<KW> *__restrict__ out, uint64_t n, DigitSel sel,
                         const uint32_t *__restrict__ tile_off, uint32_t ntiles) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    Rec<KW> *exch = reinterpret_cast<Rec<KW> *>(smem_raw);
    uint32_t *wc = reinterpret_cast<uint32_t *>(smem_raw + (size_t)RS_TILE * sizeof(Rec<KW>));  // [RS_WARPS][256]
    uint32_t *dstart = wc + RS_WARPS * RS_RADIX;
    uint32_t *goff = dstart + RS_RADIX;

    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
    const uint32_t count = (uint32_t)min((uint64_t)RS_TILE, n - base);
    for (uint32_t i = tid; i < RS_WARPS * RS_RADIX; i += RS_THREADS) wc[i] = 0;
    __syncthreads();

    // warp `warp` owns the contiguous slice [warp*512, warp*512+512) of the tile, 16 rounds of 32
    Rec<KW> items[RS_ITEMS];
    uint32_t rank[RS_ITEMS];
    uint32_t *mywc = wc + warp * RS_RADIX;
    const uint32_t lt_mask = (1u << lane) - 1;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const uint32_t idx = warp * (32 * RS_ITEMS) + r * 32 + lane;
        const bool valid = idx < count;
        const unsigned act = __ballot_sync(0xffffffffu, valid);
        rank[r] = 0;
        if (valid) {
            items[r] = load_rec<KW>(in + base + idx);
            const uint32_t d = digit_of<KW>(items[r], sel);
            const unsigned peers = __match_any_sync(act, d);
            const int leader = __ffs(peers) - 1;
            uint32_t c = 0;
            if ((int)lane == leader) {
                c = mywc[d];
                mywc[d] = c + __popc(peers);
            }
            c = __shfl_sync(peers, c, leader);
            rank[r] = c + __popc(peers & lt_mask);
        }
        __syncwarp();
    }
    __syncthreads();
    {  // thread = digit: exclusive scan over warps, then over digits
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const uint32_t t = wc[w * RS_RADIX + tid];
            wc[w * RS_RADIX + tid] = run;
            run += t;
        }
        uint32_t tot;
        const uint32_t excl = block_exclusive_scan<uint32_t>(run, &tot);
        dstart[tid] = excl;
        goff[tid] = tile_off[(uint64_t)tid * ntiles + blockIdx.x] - excl;  // mod 2^32, n < 2^32
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; r++) {
        const uint32_t idx = warp * (32 * RS_ITEMS) + r * 32 + lane;
        if (idx < count) {
            const uint32_t d = digit_of<KW>(items[r], sel);
            exch[dstart[d] + mywc[d] + rank[r]] = items[r];
        }
    }
    __syncthreads();
    for (uint32_t i = tid; i < count; i += RS_THREADS) {
        const Rec<KW> r = exch[i];
        const uint32_t d = digit_of<KW>(r, sel);
        store_rec<KW>(out + (uint32_t)(goff[d] + i), r);
    }
}

__global__ void part_counts_kernel(const uint32_t *__restrict__ tile_off, uint32_t ntiles, uint32_t nparts, uint64_t n,
                                   uint64_t *__restrict__ counts) {
    const uint32_t p = threadIdx.x;
    if (p >= nparts) return;
    const uint64_t lo = tile_off[(uint64_t)p * ntiles];
    const uint64_t hi = (p + 1 < RS_RADIX) ? tile_off[(uint64_t)(p + 1) * ntiles] : n;
    counts[p] = hi - lo;
}

static inline uint32_t ntiles_of(uint64_t n) { return (uint32_t)((n + RS_TILE - 1) / RS_TILE); }

size_t radix_scratch_bytes(uint64_t n) {
    const uint64_t table = (uint64_t)RS_RADIX * ntiles_of(n ? n : 1);
    return (table + scan_scratch_elems(table) + 16) * sizeof(uint32_t);
}

template <int KW>
static int one_pass(const Rec<KW> *in, Rec<KW> *out, uint64_t n, const DigitSel &sel, uint32_t *scratch, KernelProf *prof,
                    cudaStream_t st) {
    const uint32_t nt = ntiles_of(n);
    const uint64_t table = (uint64_t)RS_RADIX * nt;
    uint32_t *tile_hist = scratch;
    uint32_t *scan_tmp = scratch + table;
    int launches = 0;
    bool on = prof && prof->begin(KK_RADIX_HIST, st);
    radix_hist_kernel<KW><<<nt, RS_THREADS, 0, st>>>(in, n, sel, tile_hist, nt);
    if (prof) prof->end(on, 1, st);
    launches++;
    on = prof && prof->begin(KK_RADIX_TILESCAN, st);
    const int ls = exclusive_scan<uint32_t, PtrIn<uint32_t>>(PtrIn<uint32_t>{tile_hist}, tile_hist, table, scan_tmp, nullptr, st);
    if (prof) prof->end(on, ls, st);
    launches += ls;
    const size_t smem = (size_t)RS_TILE * sizeof(Rec<KW>) + (RS_WARPS * RS_RADIX + 2 * RS_RADIX) * sizeof(uint32_t);
    static bool attr_set[3] = {false, false, false};
    if (!attr_set[KW]) {
        cudaFuncSetAttribute(radix_scatter_kernel<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr_set[KW] = true;
    }
    on = prof && prof->begin(KK_RADIX_SCATTER, st);
    radix_scatter_kernel<KW><<<nt, RS_THREADS, smem, st>>>(in, out, n, sel, tile_hist, nt);
    if (prof) prof->end(on, 1, st);
    return launches + 1;
}

template <int KW>
static int sort_impl(Rec<KW> *a, Rec<KW> *b, uint64_t n, int K, int M, uint32_t *scratch, bool *result_in_b, int *passes_out,
                     KernelProf *prof, cudaStream_t st) {
    int launches = 0, passes = 0;
    Rec<KW> *src = a, *dst = b;
    auto run = [&](DigitSel sel) {
        launches += one_pass<KW>(src, dst, n, sel, scratch, prof, st);
        Rec<KW> *t = src;
        src = dst;
        dst = t;
        passes++;
    };
    // least significant digits first: k-mer low word ... k-mer high word, then the m-mer code
    int kbits = 2 * K;
    for (int w = KW - 1; w >= 0; w--) {
        const int bits = kbits > 64 ? 64 : kbits;
        for (int s = 0; s < bits; s += 8) run(DigitSel{0, w, s, 0});
        kbits -= bits;
    }
    for (int s = 0; s < 2 * M; s += 8) run(DigitSel{1, 0, s, 0});
    *result_in_b = (src == b);
    *passes_out = passes;
    return launches;
}

int radix_sort_records(void *a, void *b, uint64_t n, int KW, int K, int M, void *scratch, bool *result_in_b, int *passes_out,
                       KernelProf *prof, cudaStream_t st) {
    *result_in_b = false;
    *passes_out = 0;
    if (n == 0) return 0;
    if (KW == 1)
        return sort_impl<1>(static_cast<Rec<1> *>(a), static_cast<Rec<1> *>(b), n, K, M, static_cast<uint32_t *>(scratch), result_in_b,
                            passes_out, prof, st);
    return sort_impl<2>(static_cast<Rec<2> *>(a), static_cast<Rec<2> *>(b), n, K, M, static_cast<uint32_t *>(scratch), result_in_b,
                        passes_out, prof, st);
}

int radix_partition_by_owner(const void *in, void *out, uint64_t n, int KW, uint32_t n_parts, void *scratch, uint64_t *d_counts,
                             cudaStream_t st) {
    if (n == 0) {
        cudaMemsetAsync(d_counts, 0, sizeof(uint64_t) * n_parts, st);
        return 0;
    }
    const DigitSel sel{2, 0, 0, n_parts};
    int launches;
    if (KW == 1)
        launches = one_pass<1>(static_cast<const Rec<1> *>(in), static_cast<Rec<1> *>(out), n, sel, static_cast<uint32_t *>(scratch), nullptr, st);
    else
        launches = one_pass<2>(static_cast<const Rec<2> *>(in), static_cast<Rec<2> *>(out), n, sel, static_cast<uint32_t *>(scratch), nullptr, st);
    part_counts_kernel<<<1, RS_RADIX, 0, st>>>(static_cast<uint32_t *>(scratch), ntiles_of(n), n_parts, n, d_counts);
    return launches + 1;
}

}  // namespace gbin
