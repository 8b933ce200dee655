// expand_ids.cu — expand_read_id_list (binning.c:857-888) on the device.
//
// The reference replaces the id list of every surviving k-mer by a list of K lists, one per base, each a copy of the k-mer's
// list (head first, then K-1 duplicates appended: binning.c:871-884); print_kmer_read_ids then prints K identical lines per
// k-mer.  In the table's CSR form this is a replicate: list (j, b) of k-mer j and base b is ids[off[j] .. off[j+1]).  The
// expanded lists of k-mer j are laid out back to back at K * off[j], so the output is again CSR:
//   list_off[j * K + b] = K * off[j] + b * (off[j+1] - off[j]),   exp_ids[list_off[j*K+b] + i] = ids[off[j] + i].
// One warp per k-mer: its K * c output elements are consecutive, so the stores coalesce; the c source ids are re-read from L1.
#include "gbin_device.cuh"
#include "gbin_internal.h"

namespace gbin {

__global__ void __launch_bounds__(256)
    expand_ids_kernel(const uint64_t *__restrict__ id_off, const int32_t *__restrict__ ids, uint64_t n_kmers, uint32_t K, uint64_t *__restrict__ list_off,
                      int32_t *__restrict__ exp_ids) {
    const uint32_t lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    for (uint64_t j = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n_kmers; j += warps) {
        const uint64_t o = id_off[j], c = id_off[j + 1] - o;
        const uint64_t dst0 = (uint64_t)K * o;
        for (uint32_t b = lane; b < K; b += 32) list_off[j * K + b] = dst0 + (uint64_t)b * c;
        if (j + 1 == n_kmers && lane == 0) list_off[n_kmers * K] = dst0 + (uint64_t)K * c;
        const uint64_t total = (uint64_t)K * c;
        uint64_t src = lane % c;  // c >= 1: a surviving k-mer has at least one id
        const uint64_t step = 32 % c;
        for (uint64_t x = lane; x < total; x += 32) {
            exp_ids[dst0 + x] = ids[o + src];
            src += step;
            if (src >= c) src -= c;
        }
    }
}

int expand_ids_device(const uint64_t *id_off, const int32_t *ids, uint64_t n_kmers, uint32_t K, uint64_t *list_off, int32_t *exp_ids, int sm_count,
                      cudaStream_t st) {
    if (n_kmers == 0) {
        cudaMemsetAsync(list_off, 0, sizeof(uint64_t), st);
        return 0;
    }
    uint64_t blocks = (n_kmers + 7) / 8;
    if (blocks > (uint64_t)sm_count * 16) blocks = (uint64_t)sm_count * 16;
    expand_ids_kernel<<<(unsigned)blocks, 256, 0, st>>>(id_off, ids, n_kmers, K, list_off, exp_ids);
    return 1;
}

}  // namespace gbin
