// skr.cuh — super-k-mer records (pipeline v2).
//
// Consecutive windows of a read that share one signature (binning.c:922: the signature is kept until
// the window start passes it) share m-mer bucket and orientation, so they are one substring of
// K+n-1 bases.  A super-k-mer record stores that substring once instead of n expanded k-mer records:
//
//   word 0        arrival index of the read
//   word 1        m-mer bucket code  w(sig) = max(s, 4^M-1-s)
//   word 2        n (bits 0-7: windows in the segment, 1..K-M+1 <= 63) | is_rev << 8 | so << 16 (bits 16-21: offset of the
//                 signature m-mer from the segment's first base; window t holds it at offset d = so - t, 0 <= d <= K-M)
//   word 3        start: index of the segment's first window in the read (diagnostic / parity)
//   word 4..      the K+n-1 bases, 2 bits each, MSB first, 16 bases per word, zero padded
//
// PW = payload size in 64-bit units: 2 (K <= 32: at most 2K-M <= 64 bases) or 4 (K <= 64).
#pragma once
#include <cstdint>

namespace gbin {

template <int PW>
struct SkrLayout {
    static constexpr int PAYLOAD_WORDS = 2 * PW;       // u32 words of packed bases
    static constexpr int WORDS = 4 + PAYLOAD_WORDS;    // 8 (32 B) or 12 (48 B)
    static constexpr int BYTES = 4 * WORDS;
};

__host__ __device__ inline int skr_payload_units(int K) { return K <= 32 ? 2 : 4; }
__host__ __device__ inline int skr_words(int K) { return 4 + 2 * skr_payload_units(K); }

}  // namespace gbin
