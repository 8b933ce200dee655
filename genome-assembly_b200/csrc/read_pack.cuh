// read_pack.cuh — per-warp staging of one read: ASCII -> 2-bit codes (getval, binning.c:91-111),
// packed MSB-first 16 bases per u32 in shared memory, and extraction helpers on the packed form.
#pragma once
#include "gbin_device.cuh"

namespace gbin {

constexpr int PK_PAD_WORDS = 10;  // zero words after the read so window extraction never reads garbage

// Packed words a warp writes for a read of up to max_len bases: whole 128-base chunks (8 words each) plus
// the zero padding, rounded to a multiple of 4 so that what follows stays 16-byte aligned.
__host__ __device__ inline uint32_t scan_pk_words(uint32_t max_len) { return (8u * ((max_len + 127u) / 128u) + PK_PAD_WORDS + 3u) & ~3u; }

// Whole warp: reads the L bytes at src with aligned 32-bit loads (one coalesced load per 128 bases, realigned
// with a funnel shift), converts four bases at a time with byte-parallel arithmetic, and stores the packed words.
//   x = (c >> 1) & 3 maps A,C,T,G to 0,1,2,3;  code = 3 ^ x ^ (x >> 1) gives A3 C2 G1 T0 (binning.c:91-111);
//   a byte is valid iff re-encoding the code gives the byte back (PRMT table lookup), so non-ACGT bytes are
//   counted exactly.  Returns (per lane) the number of non-ACGT bytes it saw.
// The loads of one 128-base chunk: lane l gets the aligned word that holds bytes 4l.. of the chunk (w0), lane 31 also the
// word after its own (w31).  Split from the conversion so that a caller can issue the loads of the NEXT read early.
__device__ __forceinline__ void warp_pack_load(const uint8_t *__restrict__ src, uint32_t L, uint32_t c0, uint32_t lane, uint32_t &w0, uint32_t &w31) {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(src);
    const uint32_t mis = (uint32_t)(addr & 3u);
    const uint32_t *__restrict__ base = reinterpret_cast<const uint32_t *>(addr - mis);
    const uint32_t nwords_in = (mis + L + 3u) >> 2;  // aligned words that overlap the read
    const uint32_t idx = (c0 >> 2) + lane;
    w0 = idx < nwords_in ? base[idx] : 0u;
    w31 = 0u;
    if (lane == 31) w31 = (idx + 1 < nwords_in) ? base[idx + 1] : 0u;
}

// Converts one loaded chunk (bases c0 .. c0+127 of the read) and stores its 8 packed words; returns the lane's count of
// non-ACGT bytes.
__device__ __forceinline__ uint32_t warp_pack_chunk(uint32_t *pk, uint32_t w0, uint32_t w31, uint32_t mis, uint32_t L, uint32_t c0, uint32_t lane) {
    uint32_t w1 = __shfl_down_sync(0xffffffffu, w0, 1);
    if (lane == 31) w1 = w31;
    uint32_t chars = __funnelshift_r(w0, w1, mis * 8);  // bytes src[c0+4*lane .. +3], first byte lowest
    const int rem = (int)L - (int)(c0 + 4 * lane);      // valid bytes in this group of four
    if (rem <= 0) chars = 0x54545454u;                  // past the end: 'T' = code 0
    else if (rem < 4) {
        const uint32_t m = (1u << (8 * rem)) - 1u;
        chars = (chars & m) | (0x54545454u & ~m);
    }
    const uint32_t x = (chars >> 1) & 0x03030303u;
    const uint32_t y = x ^ ((x >> 1) & 0x01010101u) ^ 0x03030303u;  // per byte: A3 C2 G1 T0
    const uint32_t sel = (y & 0x3u) | ((y >> 4) & 0x30u) | ((y >> 8) & 0x300u) | ((y >> 12) & 0x3000u);
    const uint32_t recon = __byte_perm(0x41434754u, 0u, sel);  // bytes T,G,C,A indexed by code
    const uint32_t diff = recon ^ chars;
    uint32_t nbad = 0;
    if (diff) nbad = __popc((((diff & 0x7f7f7f7fu) + 0x7f7f7f7fu) | diff) & 0x80808080u);
    const uint32_t v8 = (y * 0x40100401u) >> 24;  // four 2-bit codes, first base most significant
    uint32_t part = v8 << (24 - 8 * (lane & 3));
    part |= __shfl_xor_sync(0xffffffffu, part, 1);
    part |= __shfl_xor_sync(0xffffffffu, part, 2);
    if ((lane & 3) == 0) pk[(c0 >> 4) + (lane >> 2)] = part;
    return nbad;
}

__device__ __forceinline__ uint32_t warp_pack_read(uint32_t *pk, const uint8_t *__restrict__ src, uint32_t L, uint32_t lane) {
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(src) & 3u);
    uint32_t nbad = 0, nchunks = 0;
    for (uint32_t c0 = 0; c0 < L; c0 += 128, nchunks++) {
        uint32_t w0, w31;
        warp_pack_load(src, L, c0, lane, w0, w31);
        nbad += warp_pack_chunk(pk, w0, w31, mis, L, c0, lane);
    }
    if (lane < PK_PAD_WORDS) pk[8 * nchunks + lane] = 0u;
    __syncwarp();
    return nbad;
}

// 2M-bit code of the m-mer starting at base p.
__device__ __forceinline__ uint32_t mmer_at(const uint32_t *pk, uint32_t p, int M) {
    const uint32_t bit = 2 * p, wi = bit >> 5, sh = bit & 31;
    const uint64_t x = ((uint64_t)pk[wi] << 32) | pk[wi + 1];
    return (uint32_t)((x << sh) >> (64 - 2 * M));
}

// Whole warp: wv[p] = w(p) = max(s(p), 4^M-1-s(p)) for every m-mer start p in [0, L-M].
__device__ __forceinline__ void warp_mmer_scores(const uint32_t *pk, uint32_t *wv, uint32_t L, int M, uint32_t FULL, uint32_t lane) {
    for (uint32_t p = lane; p + M <= L; p += 32) {
        const uint32_t s = mmer_at(pk, p, M);
        wv[p] = max(s, FULL - s);
    }
    __syncwarp();
}
// One hop of the signature chain (binning.c:931-988 at a restart window i): the LEFTMOST position
// p in [i, i+C) maximising w(p).  Whole warp; returns sig, and w(sig) in *wmax.
__device__ __forceinline__ uint32_t warp_signature_hop(const uint32_t *wv, uint32_t i, uint32_t C, uint32_t lane, uint32_t *wmax) {
    uint32_t best_w = 0, best_p = 0xffffffffu;
    for (uint32_t b = 0; b < C; b += 32) {
        const uint32_t off = b + lane;
        if (off < C) {
            const uint32_t w = wv[i + off];
            if (w > best_w) {  // strict: the earlier (smaller p) candidate survives ties
                best_w = w;
                best_p = i + off;
            }
        }
    }
    const uint32_t mx = __reduce_max_sync(0xffffffffu, best_w);
    *wmax = mx;
    return __reduce_min_sync(0xffffffffu, best_w == mx ? best_p : 0xffffffffu);
}


// ---- pipeline v2 scan: scores carry the strand bit, one hop = one REDUX

// Whole warp: wr[p] = w(p) << 1 | is_rev(p) for every m-mer start p in [0, L-M]
// (w(p) = max(s, 4^M-1-s) < 2^30, is_rev = the complement scored higher, binning.c:943,948).
__device__ __forceinline__ void warp_mmer_scores_strand(const uint32_t *pk, uint32_t *wr, uint32_t L, int M, uint32_t FULL, uint32_t lane) {
    const uint32_t down = 32u - 2u * (uint32_t)M;
    for (uint32_t p = lane; p + M <= L; p += 32) {
        const uint32_t bit = 2 * p, wi = bit >> 5, sh = bit & 31;
        const uint32_t s = __funnelshift_l(pk[wi + 1], pk[wi], sh) >> down;
        const uint32_t c = FULL - s;
        wr[p] = c > s ? ((c << 1) | 1u) : (s << 1);
    }
    __syncwarp();
}

// One hop of the signature chain (binning.c:931-988 at a restart window i): the LEFTMOST position p in [i, i+C)
// maximising w(p).  Whole warp.  Returns the offset of the signature from i; *w_rev = w(sig) << 1 | is_rev(sig).
// PACKED: score, (inverted) offset and strand bit fit one 32-bit key, so the hop is a single REDUX.MAX
// (needs 2M + 1 + OB <= 32 with OB = 5 offset bits when C <= 32, else 6).
template <bool PACKED>
__device__ __forceinline__ uint32_t warp_signature_hop_strand(const uint32_t *wr, uint32_t i, uint32_t C, uint32_t lane, uint32_t *w_rev) {
    if (PACKED) {
        if (C <= 32) {
            const uint32_t v = lane < C ? wr[i + lane] : 0u;
            const uint32_t key = ((v & ~1u) << 5) | ((31u - lane) << 1) | (v & 1u);
            const uint32_t mx = __reduce_max_sync(0xffffffffu, key);
            *w_rev = ((mx >> 6) << 1) | (mx & 1u);
            return 31u - ((mx >> 1) & 31u);
        }
        const uint32_t v0 = wr[i + lane];  // C > 32: lanes cover offsets lane and lane + 32 (C <= 63)
        const uint32_t v1 = lane + 32 < C ? wr[i + lane + 32] : 0u;
        const uint32_t k0 = ((v0 & ~1u) << 6) | ((63u - lane) << 1) | (v0 & 1u);
        const uint32_t k1 = ((v1 & ~1u) << 6) | ((31u - lane) << 1) | (v1 & 1u);
        const uint32_t mx = __reduce_max_sync(0xffffffffu, max(k0, k1));
        *w_rev = ((mx >> 7) << 1) | (mx & 1u);
        return 63u - ((mx >> 1) & 63u);
    }
    uint32_t best = 0, best_off = 0;
    for (uint32_t b = 0; b < C; b += 32) {
        const uint32_t o = b + lane;
        const uint32_t v = o < C ? wr[i + o] : 0u;
        if ((v >> 1) > (best >> 1)) {  // strict: the earlier candidate survives ties
            best = v;
            best_off = o;
        }
    }
    const uint32_t mx = __reduce_max_sync(0xffffffffu, best >> 1);
    const uint32_t sel = __reduce_min_sync(0xffffffffu, (best >> 1) == mx ? ((best_off << 1) | (best & 1u)) : 0xffffffffu);
    *w_rev = (mx << 1) | (sel & 1u);
    return sel >> 1;
}

}  // namespace gbin
