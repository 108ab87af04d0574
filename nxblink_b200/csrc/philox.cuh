// Philox4x32-10 counter-based generator (Salmon et al., SC'11), written out so
// that the same function compiles for host (known-answer test through the
// C-ABI) and device.  Every random number in a sweep is a pure function of
// (seed, unit, sweep, stream, block), so results do not depend on the launch
// geometry, the shared-memory capacity tier or the number of GPUs.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define NXB_HD __host__ __device__ __forceinline__
#else
#define NXB_HD inline
#endif

struct NxbU4 { uint32_t x, y, z, w; };

NXB_HD uint32_t nxb_mulhi(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

NXB_HD NxbU4 nxb_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                               uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 10; ++i) {
        uint32_t hi0 = nxb_mulhi(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = nxb_mulhi(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    NxbU4 r; r.x = c0; r.y = c1; r.z = c2; r.w = c3;
    return r;
}

// 53-bit uniform in the open interval (0, 1)
NXB_HD double nxb_u53(uint32_t hi, uint32_t lo) {
    uint64_t v = ((uint64_t)hi << 21) ^ (uint64_t)(lo >> 11);
    return ((double)v + 0.5) * (1.0 / 9007199254740992.0);
}

// 32-bit uniform in the open interval (0, 1)
NXB_HD double nxb_u32(uint32_t x) {
    return ((double)x + 0.5) * (1.0 / 4294967296.0);
}

// xoshiro128++ (Blackman & Vigna): the per-track sequential stream of the tolerance
// update (thinning and backward-sampling decisions, consumed in the canonical
// (edge, time) walk order).  Seeded from one Philox block per (unit, sweep, track).
struct NxbXo { uint32_t s0, s1, s2, s3; };
NXB_HD uint32_t nxb_rotl(uint32_t x, int k) { return (x << k) | (x >> (32 - k)); }
NXB_HD uint32_t nxb_xo_next(NxbXo& g) {
    uint32_t result = nxb_rotl(g.s0 + g.s3, 7) + g.s0;
    uint32_t t = g.s1 << 9;
    g.s2 ^= g.s0; g.s3 ^= g.s1; g.s1 ^= g.s2; g.s0 ^= g.s3;
    g.s2 ^= t;
    g.s3 = nxb_rotl(g.s3, 11);
    return result;
}

// RNG stream identifiers (counter word c2)
#define NXB_STREAM_PRI_EDGE   0x01000000u   // | edge        : candidate count, times, thinning
#define NXB_STREAM_PRI_FFBS   0x02000000u   //               : block = chunk index
#define NXB_STREAM_BLK_PAIR   0x03000000u   // | k<<16 | edge: candidate times (2 per block)
#define NXB_STREAM_BLK_COUNT  0x06000000u   // | pair group  : 4 candidate counts per block
#define NXB_STREAM_BLK_FFBS   0x04000000u   // | k<<16       : block 0 seeds the track's xoshiro stream
#define NXB_STREAM_INIT_BLK   0x05000000u   // | k<<16 | edge
#define NXB_STREAM_TOL_SLICE  0x07000000u   // | slice       : block 0 seeds the xoshiro stream of one slice of the candidate axis
#define NXB_STREAM_TOL_RET    0x08000000u   //               : backward-sampling words of the retained events (block = index / 4)
#define NXB_STREAM_TOL_ROOT   0x09000000u   //               : root draws of the tracks (block = track / 4)
#define NXB_SWEEP_INIT        0xFFFFFFFFu   // counter word c1 used by the initialisation pass
