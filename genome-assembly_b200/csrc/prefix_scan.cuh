// prefix_scan.cuh — device-wide exclusive prefix sums (reduce-then-scan, recursive over block sums).
//
// Used by the record sort (per-tile digit tables), the run-length stage (head flags -> group index)
// and the prune stage (survivor index, id offsets).  HBM-bound: the input functor is evaluated twice
// (reduce pass + scan pass) and the output written once.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gbin {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;  // 4096 items per block

template <typename T>
__device__ __forceinline__ T warp_inclusive_scan(T v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// Block-wide exclusive scan of one value per thread (SCAN_THREADS threads); returns the exclusive
// prefix and writes the block total to *total (same value in every thread).
template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T *total) {
    __shared__ T warp_sums[SCAN_THREADS / 32];
    __shared__ T block_total;
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = warp_inclusive_scan(v);
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T s = lane < SCAN_THREADS / 32 ? warp_sums[lane] : T(0);
        T si = warp_inclusive_scan(s);
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = si - s;
        if (lane == SCAN_THREADS / 32 - 1) block_total = si;
    }
    __syncthreads();
    T res = inc - v + warp_sums[warp];
    *total = block_total;
    __syncthreads();  // shared scratch is reused by the next call
    return res;
}

// Both kernels read their tile striped (thread t takes elements t, t + 256, ...): consecutive threads touch consecutive
// elements, whatever the input functor reads.  The scan kernel needs each thread's 16 elements to be consecutive for the
// sequential part, so it passes the tile through shared memory (one pad element per 16 / 32 keeps the blocked accesses
// spread over the banks).
template <typename TOut>
__device__ __forceinline__ uint32_t scan_pad(uint32_t p) {
    return p + (sizeof(TOut) == 8 ? (p >> 4) : (p >> 5));
}

template <typename TOut, typename InOp>
__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(InOp in, TOut *__restrict__ block_sums, uint64_t n) {
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
    TOut s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint64_t j = base + (uint64_t)i * SCAN_THREADS + threadIdx.x;
        if (j < n) s += (TOut)in(j);
    }
    TOut total;
    (void)block_exclusive_scan<TOut>(s, &total);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

template <typename TOut, typename InOp>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_apply_kernel(InOp in, TOut *__restrict__ out, const TOut *__restrict__ block_offsets, uint64_t n, TOut *__restrict__ total_out) {
    __shared__ TOut ex[SCAN_TILE + SCAN_TILE / 16 + 1];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint32_t p = (uint32_t)i * SCAN_THREADS + threadIdx.x;
        const uint64_t j = base + p;
        ex[scan_pad<TOut>(p)] = j < n ? (TOut)in(j) : TOut(0);
    }
    __syncthreads();
    TOut v[SCAN_ITEMS];
    TOut s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = ex[scan_pad<TOut>(threadIdx.x * SCAN_ITEMS + i)];
        s += v[i];
    }
    TOut total;
    TOut run = block_exclusive_scan<TOut>(s, &total) + (block_offsets ? block_offsets[blockIdx.x] : TOut(0));
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        ex[scan_pad<TOut>(threadIdx.x * SCAN_ITEMS + i)] = run;
        run += v[i];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total_out = run;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint32_t p = (uint32_t)i * SCAN_THREADS + threadIdx.x;
        const uint64_t j = base + p;
        if (j < n) out[j] = ex[scan_pad<TOut>(p)];
    }
}

// Same as scan_apply_kernel, but the exclusive prefix of element j is handed to out(j, prefix) instead of being stored in an array
// (the consumer can write several derived arrays in one go and the prefix array itself never exists).
template <typename TOut, typename InOp, typename OutOp>
__global__ void __launch_bounds__(SCAN_THREADS)
    scan_apply_to_kernel(InOp in, OutOp out, const TOut *__restrict__ block_offsets, uint64_t n, TOut *__restrict__ total_out) {
    __shared__ TOut ex[SCAN_TILE + SCAN_TILE / 16 + 1];
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_TILE;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint32_t p = (uint32_t)i * SCAN_THREADS + threadIdx.x;
        const uint64_t j = base + p;
        ex[scan_pad<TOut>(p)] = j < n ? (TOut)in(j) : TOut(0);
    }
    __syncthreads();
    TOut v[SCAN_ITEMS];
    TOut s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        v[i] = ex[scan_pad<TOut>(threadIdx.x * SCAN_ITEMS + i)];
        s += v[i];
    }
    TOut total;
    TOut run = block_exclusive_scan<TOut>(s, &total) + (block_offsets ? block_offsets[blockIdx.x] : TOut(0));
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        ex[scan_pad<TOut>(threadIdx.x * SCAN_ITEMS + i)] = run;
        run += v[i];
    }
    if (total_out && blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_THREADS - 1) *total_out = run;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; i++) {
        const uint32_t p = (uint32_t)i * SCAN_THREADS + threadIdx.x;
        const uint64_t j = base + p;
        if (j < n) out(j, ex[scan_pad<TOut>(p)]);
    }
}

template <typename T>
struct PtrIn {
    const T *p;
    __device__ __forceinline__ T operator()(uint64_t j) const { return p[j]; }
};

// scratch must hold scan_scratch_elems(n) elements of TOut.
inline uint64_t scan_scratch_elems(uint64_t n) {
    uint64_t tot = 0;
    while (n > (uint64_t)SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        tot += n;
    }
    return tot + 1;
}

// out[j] = sum_{i<j} in(i); *total_out (device, optional) = sum of all. in/out may alias when InOp reads `out`.
// Returns the number of kernels launched.
template <typename TOut, typename InOp>
int exclusive_scan(InOp in, TOut *out, uint64_t n, TOut *scratch, TOut *total_out, cudaStream_t st) {
    if (n == 0) {
        if (total_out) cudaMemsetAsync(total_out, 0, sizeof(TOut), st);
        return 0;
    }
    const uint64_t blocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (blocks == 1) {
        scan_apply_kernel<TOut, InOp><<<1, SCAN_THREADS, 0, st>>>(in, out, nullptr, n, total_out);
        return 1;
    }
    int launches = 0;
    scan_reduce_kernel<TOut, InOp><<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(in, scratch, n);
    launches++;
    launches += exclusive_scan<TOut, PtrIn<TOut>>(PtrIn<TOut>{scratch}, scratch, blocks, scratch + blocks, nullptr, st);
    scan_apply_kernel<TOut, InOp><<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(in, out, scratch, n, total_out);
    return launches + 1;
}

// exclusive_scan whose result goes to out(j, prefix).  scratch as for exclusive_scan.
template <typename TOut, typename InOp, typename OutOp>
int exclusive_scan_to(InOp in, OutOp out, uint64_t n, TOut *scratch, TOut *total_out, cudaStream_t st) {
    if (n == 0) {
        if (total_out) cudaMemsetAsync(total_out, 0, sizeof(TOut), st);
        return 0;
    }
    const uint64_t blocks = (n + SCAN_TILE - 1) / SCAN_TILE;
    if (blocks == 1) {
        scan_apply_to_kernel<TOut, InOp, OutOp><<<1, SCAN_THREADS, 0, st>>>(in, out, nullptr, n, total_out);
        return 1;
    }
    int launches = 0;
    scan_reduce_kernel<TOut, InOp><<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(in, scratch, n);
    launches++;
    launches += exclusive_scan<TOut, PtrIn<TOut>>(PtrIn<TOut>{scratch}, scratch, blocks, scratch + blocks, nullptr, st);
    scan_apply_to_kernel<TOut, InOp, OutOp><<<(unsigned)blocks, SCAN_THREADS, 0, st>>>(in, out, scratch, n, total_out);
    return launches + 1;
}

}  // namespace gbin
