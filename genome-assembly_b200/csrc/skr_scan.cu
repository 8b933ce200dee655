// skr_scan.cu — pipeline v2 scan stage: process_read's window/signature work (binning.c:918-1040)
// emitting one super-k-mer record per signature segment instead of one record per window.
//
// One warp per read as in scan_reads.cu (pack in shared memory, signature chain by hops).  A CTA
// works on tiles of 8 x RPW consecutive reads; every warp stages the records of its reads in shared
// memory, the CTA obtains the tile's global output offset from a single-pass chained scan over
// tiles (decoupled look-back on a packed {flag, count} word per tile, tiles handed out in order by an
// atomic ticket so a predecessor is always running), and the staged records leave the SM as one
// contiguous run of 16-byte stores.  The output is therefore in arrival order (read-major,
// segment-minor) and deterministic — the later stable sort by m-mer keeps arrival order per bucket.
#include "gbin_device.cuh"
#include "gbin_internal.h"
#include "read_pack.cuh"
#include "skr.cuh"

namespace gbin {

constexpr int SKR_WARPS = 8;
constexpr unsigned long long LB_FLAG_AGG = 1ull << 62, LB_FLAG_PREFIX = 2ull << 62, LB_VALUE_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// Chained scan: publishes this tile's count and returns the sum of all earlier tiles' counts.
__device__ __forceinline__ unsigned long long lookback_exclusive(unsigned long long *state, uint32_t tile, unsigned long long count) {
    if (tile == 0) {
        atomicExch(&state[0], LB_FLAG_PREFIX | count);
        return 0;
    }
    atomicExch(&state[tile], LB_FLAG_AGG | count);
    unsigned long long sum = 0;
    for (int64_t j = (int64_t)tile - 1;; j--) {
        unsigned long long v;
        while (((v = ld_volatile_u64(&state[j])) >> 62) == 0) __nanosleep(40);
        sum += v & LB_VALUE_MASK;
        if ((v >> 62) == 2) break;
    }
    atomicExch(&state[tile], LB_FLAG_PREFIX | (sum + count));
    return sum;
}

// per warp: packed words, m-mer scores, is_rev bit mask, segment list of the current read (2 words per segment), staged records;
// every region is a multiple of 16 bytes
__host__ __device__ inline uint32_t skr_len4(uint32_t max_len) { return (max_len + 3u) & ~3u; }
__host__ __device__ inline uint32_t skr_mask_words(uint32_t max_len) { return ((max_len >> 5) + 1u + 3u) & ~3u; }
__host__ __device__ inline uint32_t skr_warp_smem(uint32_t max_len, uint32_t seg_cap, int words) {
    return 4 * scan_pk_words(max_len) + 4 * skr_len4(max_len) + 4 * skr_mask_words(max_len) + 8 * skr_len4(max_len) + 4 * seg_cap * words;
}

template <int PW>
__global__ void __launch_bounds__(SKR_WARPS * 32)
    skr_scan_kernel(ReadsView rv, int K, int M, uint32_t arrival_base, uint32_t max_len, uint32_t rpw, uint32_t seg_cap, uint32_t ntiles,
                    uint32_t *__restrict__ out, unsigned long long capacity, unsigned long long *__restrict__ tile_state, uint32_t *__restrict__ ticket,
                    unsigned long long *__restrict__ counters /* [0] bad bases, [1] records, [2] instances */) {
    constexpr int NW = SkrLayout<PW>::WORDS;
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ uint32_t s_tile, s_warp_cnt[SKR_WARPS], s_warp_off[SKR_WARPS];
    __shared__ unsigned long long s_base;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t *wbase = smem + (size_t)warp * skr_warp_smem(max_len, seg_cap, NW);
    uint32_t *pk = reinterpret_cast<uint32_t *>(wbase);
    uint32_t *wv = pk + scan_pk_words(max_len);
    uint32_t *revmask = wv + skr_len4(max_len);
    uint32_t *segl = revmask + skr_mask_words(max_len);  // per segment of the current read: start | n << 16 | rev << 24, m-mer code
    uint32_t *stage = segl + 2 * skr_len4(max_len);
    const uint32_t FULL = (1u << (2 * M)) - 1;
    const uint32_t C = K - M + 1;
    uint32_t nbad = 0;
    unsigned long long ninst = 0;

    for (;;) {
        if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const uint32_t tile = s_tile;
        if (tile >= ntiles) break;
        const uint64_t first = (uint64_t)tile * (SKR_WARPS * rpw) + (uint64_t)warp * rpw;
        uint32_t nseg = 0;
        for (uint32_t rr = 0; rr < rpw; rr++) {
            const uint64_t r = first + rr;
            if (r >= rv.n_reads) break;
            const uint32_t L = rv.len(r);
            if (L < (uint32_t)K) continue;
            const uint32_t W = L - K + 1;
            nbad += warp_pack_read(pk, rv.data + rv.start(r), L, lane);
            warp_mmer_scores_rev(pk, wv, revmask, L, M, FULL, lane);
            const uint32_t arrival = arrival_base + (uint32_t)r;
            // phase 1: the signature chain, one hop per segment; lane 0 notes the segments of this read
            uint32_t i = 0, nsr = 0;
            while (i < W) {
                uint32_t mx;
                const uint32_t sig = warp_signature_hop(wv, i, C, lane, &mx);
                const uint32_t rev = (revmask[sig >> 5] >> (sig & 31)) & 1u;  // is_rev of the signature (binning.c:943,948)
                const uint32_t next = min(sig + 1, W);
                if (lane == 0) {
                    segl[2 * nsr] = i | ((next - i) << 16) | (rev << 24);
                    segl[2 * nsr + 1] = mx;
                }
                nsr++;
                i = next;
            }
            __syncwarp();
            // phase 2: all lanes build record words, one (segment, word) pair per lane and step — no divergence
            for (uint32_t e = lane; e < nsr * NW; e += 32) {
                const uint32_t sg = e / NW, j = e - sg * NW;
                const uint32_t info = segl[2 * sg], mx = segl[2 * sg + 1];
                const uint32_t st0 = info & 0xffffu, n = (info >> 16) & 0xffu, rev = (info >> 24) & 1u;
                const uint32_t jp = j >= 4 ? j - 4 : 0;
                const uint32_t bit = 2 * st0, wi = (bit >> 5) + jp, sh = bit & 31;
                uint32_t word = __funnelshift_l(pk[wi + 1], pk[wi], sh);
                const int keep = 2 * (int)(K + n - 1) - 32 * (int)jp;  // valid payload bits in this word
                word = keep >= 32 ? word : (keep <= 0 ? 0u : (word & (0xffffffffu << (32 - keep))));
                word = j == 0 ? arrival : (j == 1 ? mx : (j == 2 ? (n | (rev << 8)) : (j == 3 ? st0 : word)));
                stage[(nseg + sg) * NW + j] = word;
            }
            nseg += nsr;
            ninst += (lane == 0) ? W : 0;
            __syncwarp();
        }
        if (lane == 0) s_warp_cnt[warp] = nseg;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < SKR_WARPS; w++) {
                s_warp_off[w] = tot;
                tot += s_warp_cnt[w];
            }
            s_base = lookback_exclusive(tile_state, tile, tot);
            if (tile == ntiles - 1) counters[1] = s_base + tot;
        }
        __syncthreads();
        if (s_base + s_warp_off[warp] + nseg <= capacity) {  // past the capacity only the total is produced (the caller re-runs)
            uint4 *dst = reinterpret_cast<uint4 *>(out + (s_base + s_warp_off[warp]) * NW);
            const uint4 *src = reinterpret_cast<const uint4 *>(stage);
            const uint32_t nvec = nseg * (NW / 4);
            for (uint32_t v = lane; v < nvec; v += 32) dst[v] = src[v];
        }
        __syncthreads();
    }
    if (nbad) atomicAdd(&counters[0], (unsigned long long)nbad);
    if (ninst) atomicAdd(&counters[2], ninst);
}

// Host launcher.  tile_state must hold ntiles u64 (zeroed here), ticket one u32 (zeroed here).
// Returns kernels launched; *ntiles_out = number of tiles.
int launch_skr_scan(const ReadsView &rv, int K, int M, uint32_t arrival_base, uint32_t max_len, void *out, uint64_t capacity,
                    unsigned long long *tile_state, uint32_t *ticket, unsigned long long *counters, int sm_count, cudaStream_t st) {
    if (rv.n_reads == 0) return 0;
    const int PW = skr_payload_units(K);
    const int NW = 4 + 2 * PW;
    const uint32_t W = max_len >= (uint32_t)K ? max_len - K + 1 : 1;
    // reads per warp per tile: as many as fit a ~4.5 KB staging area in the worst case (one segment per window);
    // a small staging area keeps 6+ CTAs resident per SM
    uint32_t rpw = (4608u) / (W * NW * 4u);
    if (rpw < 1) rpw = 1;
    if (rpw > 8) rpw = 8;
    const uint32_t seg_cap = rpw * W;
    const uint32_t per_tile = SKR_WARPS * rpw;
    const uint32_t ntiles = (uint32_t)((rv.n_reads + per_tile - 1) / per_tile);
    const size_t smem = (size_t)SKR_WARPS * skr_warp_smem(max_len, seg_cap, NW);
    cudaMemsetAsync(tile_state, 0, sizeof(unsigned long long) * ntiles, st);
    cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st);
    uint32_t blocks = (uint32_t)sm_count * 8;
    if (blocks > ntiles) blocks = ntiles;
    if (PW == 2) {
        cudaFuncSetAttribute(skr_scan_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        skr_scan_kernel<2><<<blocks, SKR_WARPS * 32, smem, st>>>(rv, K, M, arrival_base, max_len, rpw, seg_cap, ntiles,
                                                               static_cast<uint32_t *>(out), capacity, tile_state, ticket, counters);
    } else {
        cudaFuncSetAttribute(skr_scan_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        skr_scan_kernel<4><<<blocks, SKR_WARPS * 32, smem, st>>>(rv, K, M, arrival_base, max_len, rpw, seg_cap, ntiles,
                                                               static_cast<uint32_t *>(out), capacity, tile_state, ticket, counters);
    }
    return 1;
}

uint32_t skr_scan_tiles(uint64_t n_reads, int K, uint32_t max_len) {
    const int NW = skr_words(K);
    const uint32_t W = max_len >= (uint32_t)K ? max_len - K + 1 : 1;
    uint32_t rpw = (4608u) / (W * NW * 4u);
    if (rpw < 1) rpw = 1;
    if (rpw > 8) rpw = 8;
    const uint32_t per_tile = SKR_WARPS * rpw;
    return (uint32_t)((n_reads + per_tile - 1) / per_tile);
}

}  // namespace gbin
