// lookback.cuh — single-pass chained scan ("decoupled look-back") over tiles handed out in order by an atomic ticket.
//
// state[t] packs {flag: 2 bits, value: 62 bits}: flag 1 = the tile's own aggregate is known, flag 2 = its inclusive
// prefix is known.  A tile publishes its aggregate as early as it can and later resolves its exclusive prefix by walking
// back over its predecessors 32 at a time (one predecessor per lane) until it meets an inclusive prefix.  Because tiles
// are taken in ticket order a predecessor is always already running, so the wait is bounded.
#pragma once
#include <cstdint>

namespace gbin {

constexpr unsigned long long LKB_VALUE_MASK = (1ull << 62) - 1;

__device__ __forceinline__ unsigned long long lkb_load(const unsigned long long *p) {
    return *reinterpret_cast<const volatile unsigned long long *>(p);
}

// one thread
__device__ __forceinline__ void lkb_publish_aggregate(unsigned long long *state, uint32_t tile, unsigned long long mine) {
    atomicExch(&state[tile], ((tile == 0 ? 2ull : 1ull) << 62) | mine);
}

// all 32 lanes of one warp; returns (in every lane) the sum of the aggregates of tiles [0, tile) and publishes the
// inclusive prefix of `tile`.  The warp reads a window of 32*LKB_DEPTH predecessors with independent loads (one
// round trip to L2), because with hundreds of long-running tiles in flight the nearest inclusive prefix is
// typically hundreds of tiles back and a 32-at-a-time walk would pay one dependent round trip per step.  Short tiles
// (the scan stage) use LKB_DEPTH = 1: their predecessors resolve quickly and a deeper window only adds L2 traffic.
template <int LKB_DEPTH>
__device__ __forceinline__ unsigned long long lkb_resolve_warp(unsigned long long *state, uint32_t tile, unsigned long long mine,
                                                               uint32_t lane, uint32_t sleep_ns = 100) {
    if (tile == 0) return 0ull;
    // Cheap wait first: one lane polls the nearest predecessor until it has published anything at all (tiles finish in
    // roughly ticket order, so this is where almost all of the waiting happens), then the warp evaluates the window.
    if (lane == 0) {
        while ((lkb_load(&state[tile - 1]) >> 62) == 0) __nanosleep(sleep_ns);
    }
    __syncwarp();
    unsigned long long sum = 0;
    int64_t j = (int64_t)tile - 1;
    for (;;) {
        unsigned long long v[LKB_DEPTH];
#pragma unroll
        for (int k = 0; k < LKB_DEPTH; k++) {
            const int64_t idx = j - (int64_t)(k * 32 + lane);  // lane 0 of step 0 looks at the closest predecessor
            v[k] = idx >= 0 ? lkb_load(&state[idx]) : (2ull << 62);
        }
        bool done = false, retry = false;
        unsigned long long part = 0;
        int consumed = 0;
#pragma unroll
        for (int k = 0; k < LKB_DEPTH; k++) {
            if (done || retry) continue;
            const unsigned flag = (unsigned)(v[k] >> 62);
            const unsigned pmask = __ballot_sync(0xffffffffu, flag == 2u);
            const unsigned zmask = __ballot_sync(0xffffffffu, flag == 0u);
            const int first_p = pmask ? __ffs(pmask) - 1 : 32;
            const unsigned need = first_p < 31 ? ((1u << (first_p + 1)) - 1u) : 0xffffffffu;
            if (zmask & need) {  // a predecessor in this step has not published yet: keep what was summed, re-read from here
                retry = true;
            } else {
                part += ((int)lane <= first_p) ? (v[k] & LKB_VALUE_MASK) : 0ull;
                consumed = k + 1;
                if (first_p < 32) done = true;
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
        sum += part;
        if (done) break;
        j -= 32 * consumed;
        if (retry) __nanosleep(sleep_ns);
    }
    if (lane == 0) atomicExch(&state[tile], (2ull << 62) | (sum + mine));
    return sum;
}

}  // namespace gbin
