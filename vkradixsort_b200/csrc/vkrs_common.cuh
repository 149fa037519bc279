// vkrs_common.cuh -- small device helpers shared by every kernel of the sort.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vkrs {

constexpr int RADIX_BITS = 8;        // 8 bits per pass (multi_radixsort.comp:12, RADIX_SORT_BINS 256)
constexpr int RADIX = 1 << RADIX_BITS;

// Chained-scan tile status word: [31:30] flag, [29:0] count.  30 bits of count is the
// reference's own limit (its byte sizes are uint32, MultiRadixSort.h:29-31 => N < 2^30).
constexpr uint32_t STATUS_FLAG_AGGREGATE = 1u << 30; // this tile's own digit count
constexpr uint32_t STATUS_FLAG_INCLUSIVE = 2u << 30; // count of this tile and all tiles before it
constexpr uint32_t STATUS_FLAG_MASK = 3u << 30;
constexpr uint32_t STATUS_VALUE_MASK = ~STATUS_FLAG_MASK;
constexpr uint32_t LOOKBACK_SPIN_LIMIT = 1u << 22;
constexpr int LOOKBACK_WINDOW = 8;                  // earlier tiles polled together by one digit thread // polls before a tile gives up and raises the error flag

enum DeviceError : uint32_t { DEVERR_NONE = 0, DEVERR_LOOKBACK_TIMEOUT = 1 };

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still draining; it must not touch the
// predecessor's output before this returns.  A no-op for ordinary launches.
__device__ __forceinline__ void grid_dependency_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// L2-coherent (never L1-cached) accesses for the tile status words that other CTAs poll.
__device__ __forceinline__ uint32_t ld_relaxed_gpu(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_gpu(uint32_t *p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Streaming loads/stores for key data that is touched once per pass: keep it out of L1.
template <typename T>
__device__ __forceinline__ T ld_stream(const T *p) {
    return __ldcs(p);
}

template <typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT key, uint32_t shift) {
    return static_cast<uint32_t>(key >> shift) & (RADIX - 1);
}

__device__ __forceinline__ uint32_t lanemask_gt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
    return m;
}

// Peer mask of the lanes holding the same 8-bit digit.
//   MATCH_BALLOT: 8 ballots + LOP3s as the compiler schedules them
//   MATCH_HW    : the match.any instruction (measured slower than ballots on B200)
//   MATCH_PTX   : 8 ballots, hand-scheduled: per bit one predicate-producing AND on the key
//                 itself, one vote, one predicated fold (3 ALU-pipe + 1 vote instruction per bit;
//                 the compiler's version needs 5 + 1) -- see match_key_ptx()
//   MATCH_TABLE : 8 ballots, then two 16-entry nibble tables spread over the lanes and two
//                 shuffles -- see match_key_table()
enum MatchMode { MATCH_BALLOT = 0, MATCH_HW = 1, MATCH_PTX = 2, MATCH_TABLE = 3 };

// The eight single-bit masks of the current digit, 1 << (shift + b).  Uniform per kernel.
struct DigitBitMasks {
    uint32_t m[RADIX_BITS];
    __device__ __forceinline__ explicit DigitBitMasks(uint32_t shift) {
#pragma unroll
        for (int b = 0; b < RADIX_BITS; ++b) m[b] = 1u << (shift + b);
    }
};

#define VKRS_MATCH_BIT(N)                                            \
    "and.b32 t, %1, %" #N ";\n\t"                                    \
    "setp.ne.u32 p, t, 0;\n\t"                                       \
    "vote.sync.ballot.b32 v, p, 0xffffffff;\n\t"                     \
    "@p and.b32 m, m, v;\n\t"                                        \
    "@!p lop3.b32 m, m, v, 0, 0x30;\n\t" /* m & ~v */

// Lanes of the warp whose key has the same digit (the 8 bits selected by `bm`).  All 32 lanes
// must call it (vote.sync over the full warp).  Works on 32-bit key words; 64-bit keys pass the
// word that holds the digit.
__device__ __forceinline__ uint32_t match_key_ptx(uint32_t key_word, const DigitBitMasks &bm) {
    uint32_t peers;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t, v, m;\n\t"
        "mov.b32 m, 0xffffffff;\n\t"
        VKRS_MATCH_BIT(2) VKRS_MATCH_BIT(3) VKRS_MATCH_BIT(4) VKRS_MATCH_BIT(5)
        VKRS_MATCH_BIT(6) VKRS_MATCH_BIT(7) VKRS_MATCH_BIT(8) VKRS_MATCH_BIT(9)
        "mov.b32 %0, m;\n\t"
        "}"
        : "=r"(peers)
        : "r"(key_word), "r"(bm.m[0]), "r"(bm.m[1]), "r"(bm.m[2]), "r"(bm.m[3]), "r"(bm.m[4]), "r"(bm.m[5]),
          "r"(bm.m[6]), "r"(bm.m[7]));
    return peers;
}
#undef VKRS_MATCH_BIT

// Ballot of "key has this bit set": one predicate-producing AND + one vote (the compiler's own
// rendering of the same C expression is shift + mask + compare + vote).
__device__ __forceinline__ uint32_t ballot_bit(uint32_t key_word, uint32_t bit_mask) {
    uint32_t v;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        ".reg .b32 t;\n\t"
        "and.b32 t, %1, %2;\n\t"
        "setp.ne.u32 p, t, 0;\n\t"
        "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
        "}"
        : "=r"(v)
        : "r"(key_word), "r"(bit_mask));
    return v;
}

// Per-lane constants of match_key_table(): c[b] = all-ones iff bit b of (lane & 15) is CLEAR.
struct LaneNibbleConsts {
    uint32_t c[4];
    __device__ __forceinline__ explicit LaneNibbleConsts(int lane) {
#pragma unroll
        for (int b = 0; b < 4; ++b) c[b] = ((lane >> b) & 1) ? 0u : 0xffffffffu;
    }
};

// Peer mask through nibble tables.  The eight ballots B_b (lanes whose key has digit bit b set)
// are warp-uniform.  Lane i turns them into two table entries with 8 logic ops whose selectors
// are lane constants, not data:  lo = lanes whose low digit nibble == (i & 15), hi = lanes whose
// high nibble == (i & 15).  A lane then fetches the two entries that belong to ITS digit with
// two shuffles: peers = lo[d & 15] & hi[d >> 4].  ~19 ALU-pipe instructions + 8 votes +
// 2 shuffles per key, against ~32 + 8 for folding the ballots with per-lane selects.
__device__ __forceinline__ uint32_t match_key_table(uint32_t key_word, uint32_t digit, const DigitBitMasks &bm,
                                                    const LaneNibbleConsts &lc) {
    uint32_t B[RADIX_BITS];
#pragma unroll
    for (int b = 0; b < RADIX_BITS; ++b) B[b] = ballot_bit(key_word, bm.m[b]);
    const uint32_t lo = (B[0] ^ lc.c[0]) & (B[1] ^ lc.c[1]) & (B[2] ^ lc.c[2]) & (B[3] ^ lc.c[3]);
    const uint32_t hi = (B[4] ^ lc.c[0]) & (B[5] ^ lc.c[1]) & (B[6] ^ lc.c[2]) & (B[7] ^ lc.c[3]);
    // shfl takes the source lane modulo 32 and lane i and i+16 hold the same `lo` entry: no mask needed
    return __shfl_sync(0xffffffffu, lo, (int) digit) & __shfl_sync(0xffffffffu, hi, (int) (digit >> 4));
}

// Byte extraction of the current digit (shift is a multiple of 8): one PRMT instead of shift + mask.
__device__ __forceinline__ uint32_t digit_selector(uint32_t shift) { return 0x4440u + ((shift & 31u) >> 3); }
__device__ __forceinline__ uint32_t digit_prmt(uint32_t key_word, uint32_t selector) {
    return __byte_perm(key_word, 0u, selector);
}

template <int MODE>
__device__ __forceinline__ uint32_t match_digit(uint32_t digit) {
    if (MODE == MATCH_HW) {
        return __match_any_sync(0xffffffffu, digit);
    } else {
        uint32_t mask = 0xffffffffu;
#pragma unroll
        for (int b = 0; b < RADIX_BITS; ++b) {
            const bool bit = (digit >> b) & 1u;
            const uint32_t vote = __ballot_sync(0xffffffffu, bit);
            mask &= bit ? vote : ~vote;
        }
        return mask;
    }
}

// Order-preserving key transforms fused into the first / last pass (the reference leaves signed and
// floating-point keys to caller-side preprocessing, README.md:98-99,154-155).  fwd maps the key type's
// order onto unsigned order, inv undoes it.  XF: 0 = unsigned (identity), 1 = two's-complement signed,
// 2 = IEEE-754 binary32/binary64 (negative values reversed, -0 < +0, NaNs at the two ends by sign).
template <typename KeyT, int XF>
struct KeyXform {
    static constexpr KeyT SIGN = KeyT(1) << (8 * sizeof(KeyT) - 1);
    static __device__ __forceinline__ KeyT fwd(KeyT k) {
        if (XF == 1) return k ^ SIGN;
        if (XF == 2) return k ^ ((k & SIGN) ? ~KeyT(0) : SIGN);
        return k;
    }
    static __device__ __forceinline__ KeyT inv(KeyT k) {
        if (XF == 1) return k ^ SIGN;
        if (XF == 2) return k ^ ((k & SIGN) ? SIGN : ~KeyT(0));
        return k;
    }
};

__device__ __forceinline__ uint32_t warp_inclusive_scan(uint32_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += t;
    }
    return v;
}

// Exclusive scan of one value per thread over the first 256 threads of the block (one per
// digit).  Every thread of the block must call it (it contains __syncthreads); threads
// >= 256 pass 0 and ignore the result.  `scratch` = 8 uint32 of shared memory.
__device__ __forceinline__ uint32_t block_exclusive_scan_256(uint32_t v, uint32_t *scratch, uint32_t *total_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t incl = warp_inclusive_scan(v, lane);
    if (warp < 8 && lane == 31) scratch[warp] = incl;
    __syncthreads();
    uint32_t warp_prefix = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        const uint32_t s = scratch[w];
        if (w < warp) warp_prefix += s;
        total += s;
    }
    if (total_out) *total_out = total;
    return warp_prefix + incl - v;
}

} // namespace vkrs
