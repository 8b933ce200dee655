// table_digest.cu — order-independent 64-bit digest of a pruned table.
//
// Every surviving k-mer contributes h(m-mer code, k-mer code, its read-id list in list order); the digest is the sum of the
// contributions modulo 2^64, so it does not depend on the order of buckets or k-mers and the digests of tables over disjoint
// bucket sets add up to the digest of their union (multi-GPU: sum over owners; multi-pass: sum over passes).  The id list
// enters in order: the newest-first order of the reference's linked lists (binning.c:1059-1069) is part of what is checked.
#include "../../include/gbin.h"
#include "gbin_internal.h"

namespace gbin {

__host__ __device__ inline uint64_t dg_mix(uint64_t x) {  // murmur3 finaliser
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

__host__ __device__ inline uint64_t dg_kmer(uint32_t mmer, const uint64_t *code, int kw, const int32_t *ids, uint64_t cnt) {
    uint64_t lh = 0x9E3779B97F4A7C15ull;
    for (uint64_t i = 0; i < cnt; i++) lh = dg_mix(lh ^ (uint64_t)(uint32_t)ids[i]);
    uint64_t h = (uint64_t)mmer * 0x9E3779B97F4A7C15ull ^ dg_mix(code[0]);
    if (kw == 2) h ^= dg_mix(code[1] + 0x632BE59BD9B4E019ull);
    return dg_mix(h ^ lh ^ (cnt << 40));
}

__global__ void table_digest_kernel(const uint32_t *__restrict__ mmer_codes, const uint64_t *__restrict__ mmer_kmer_off, uint64_t n_buckets,
                                    const uint64_t *__restrict__ kmer_codes, const uint64_t *__restrict__ kmer_id_off, const int32_t *__restrict__ read_ids,
                                    uint64_t n_kmers, int kw, unsigned long long *__restrict__ out) {
    unsigned long long acc = 0;
    for (uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_kmers; s += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t lo = 0, hi = n_buckets;  // last bucket whose first k-mer is <= s
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (mmer_kmer_off[mid] <= s) lo = mid;
            else hi = mid;
        }
        const uint64_t a = kmer_id_off[s], b = kmer_id_off[s + 1];
        acc += dg_kmer(mmer_codes[lo], kmer_codes + s * kw, kw, read_ids + a, b - a);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

int table_digest_device(const gbin_table *t, unsigned long long *out_dev, cudaStream_t st) {
    cudaMemsetAsync(out_dev, 0, sizeof(unsigned long long), st);
    if (t->n_kmers == 0) return 0;
    table_digest_kernel<<<148 * 8, 256, 0, st>>>(t->mmer_codes, t->mmer_kmer_off, t->n_buckets, t->kmer_codes, t->kmer_id_off, t->read_ids, t->n_kmers,
                                                 t->kmer_words, out_dev);
    return 1;
}

uint64_t table_digest_host(const gbin_table *t) {
    uint64_t acc = 0;
    for (uint64_t b = 0; b < t->n_buckets; b++)
        for (uint64_t s = t->mmer_kmer_off[b]; s < t->mmer_kmer_off[b + 1]; s++) {
            const uint64_t a = t->kmer_id_off[s], e = t->kmer_id_off[s + 1];
            acc += dg_kmer(t->mmer_codes[b], t->kmer_codes + s * t->kmer_words, t->kmer_words, t->read_ids + a, e - a);
        }
    return acc;
}

}  // namespace gbin
