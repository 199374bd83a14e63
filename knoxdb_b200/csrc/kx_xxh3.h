// kx_xxh3.h — XXH3-64 (seed 0) for host and device: hash.Hash / hash.Uint64 … of the reference
// (internal/hash/hash.go:26,67-92, internal/hash/xxh3.go:22-58; byte strings go through
// github.com/zeebo/xxh3 v1.1.0, i.e. canonical XXH3_64bits).  The host uses it for probe hashes,
// the device for building bloom filters from column values (stats.BuildBloomFilter).
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "kx_types.h"

namespace kx {
namespace xxh3 {

constexpr uint64_t P32_1 = 0x9E3779B1ull, P32_2 = 0x85EBCA77ull, P32_3 = 0xC2B2AE3Dull;
constexpr uint64_t P64_1 = 0x9E3779B185EBCA87ull, P64_2 = 0xC2B2AE3D27D4EB4Full, P64_3 = 0x165667B19E3779F9ull,
                   P64_4 = 0x85EBCA77C2B2AE63ull, P64_5 = 0x27D4EB2F165667C5ull;

// default XXH3 secret (xxHash v0.8) as little-endian 64-bit words at byte offsets 0, 8, 16 …; the
// reference's key64_008 / key64_016 / key32_* constants (internal/hash/xxh3.go:11-20) are words of it
KX_HD inline uint64_t secret64(int i) {
    constexpr uint64_t S[24] = {
        0xbe4ba423396cfeb8ull, 0x1cad21f72c81017cull, 0xdb979083e96dd4deull, 0x1f67b3b7a4a44072ull,
        0x78e5c0cc4ee679cbull, 0x2172ffcc7dd05a82ull, 0x8e2443f7744608b8ull, 0x4c263a81e69035e0ull,
        0xcb00c391bb52283cull, 0xa32e531b8b65d088ull, 0x4ef90da297486471ull, 0xd8acdea946ef1938ull,
        0x3f349ce33f76faa8ull, 0x1d4f0bc7c7bbdcf9ull, 0x3159b4cd4be0518aull, 0x647378d9c97e9fc8ull,
        0xc3ebd33483acc5eaull, 0xeb6313faffa081c5ull, 0x49daf0b751dd0d17ull, 0x9e68d429265516d3ull,
        0xfca1477d58be162bull, 0xce31d07ad1b8f88full, 0x280416958f3acb45ull, 0x7e404bbbcafbd7afull,
    };
    return S[i];
}
// unaligned little-endian 64-bit read of the secret at byte offset `off` (0 <= off <= 184)
KX_HD inline uint64_t sec(int off) {
    int i = off >> 3, sh = (off & 7) * 8;
    uint64_t lo = secret64(i);
    if (sh == 0) return lo;
    return (lo >> sh) | (secret64(i + 1) << (64 - sh));
}
KX_HD inline uint64_t r64(const uint8_t* p) {
    uint64_t v = 0;
    for (int i = 7; i >= 0; --i) v = (v << 8) | p[i];
    return v;
}
KX_HD inline uint32_t r32(const uint8_t* p) { return uint32_t(p[0]) | uint32_t(p[1]) << 8 | uint32_t(p[2]) << 16 | uint32_t(p[3]) << 24; }
KX_HD inline uint64_t rotl(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
KX_HD inline uint64_t bswap(uint64_t x) {
    x = ((x & 0x00ff00ff00ff00ffull) << 8) | ((x >> 8) & 0x00ff00ff00ff00ffull);
    x = ((x & 0x0000ffff0000ffffull) << 16) | ((x >> 16) & 0x0000ffff0000ffffull);
    return (x << 32) | (x >> 32);
}
KX_HD inline uint64_t fold(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
    return (a * b) ^ __umul64hi(a, b);
#else
    unsigned __int128 m = (unsigned __int128)a * b;
    return uint64_t(m) ^ uint64_t(m >> 64);
#endif
}
KX_HD inline uint64_t aval64(uint64_t h) { h ^= h >> 33; h *= P64_2; h ^= h >> 29; h *= P64_3; h ^= h >> 32; return h; }
KX_HD inline uint64_t aval3(uint64_t h) { h ^= h >> 37; h *= 0x165667919E3779F9ull; h ^= h >> 32; return h; }
KX_HD inline uint64_t rrmxmx(uint64_t h, uint64_t len) {
    h ^= rotl(h, 49) ^ rotl(h, 24); h *= 0x9FB21C651E98DF25ull; h ^= (h >> 35) + len; h *= 0x9FB21C651E98DF25ull;
    return h ^ (h >> 28);
}
KX_HD inline uint64_t mix(const uint8_t* in, int soff) { return fold(r64(in) ^ sec(soff), r64(in + 8) ^ sec(soff + 8)); }
KX_HD inline void stripe(uint64_t* acc, const uint8_t* in, int soff) {
    for (int i = 0; i < 8; i++) {
        uint64_t v = r64(in + 8 * i), k = v ^ sec(soff + 8 * i);
        acc[i ^ 1] += v; acc[i] += uint64_t(uint32_t(k)) * (k >> 32);
    }
}

// fixed-width specialisations: hash.Uint64/Uint32/Uint16/Uint8 (internal/hash/xxh3.go:22-58)
KX_HD inline uint64_t u64(uint64_t v) { return rrmxmx(((v >> 32) + (v << 32)) ^ (sec(8) ^ sec(16)), 8); }
KX_HD inline uint64_t u32(uint32_t v) { return rrmxmx((uint64_t(v) + (uint64_t(v) << 32)) ^ (sec(8) ^ sec(16)), 4); }
KX_HD inline uint64_t u16(uint16_t v) {
    uint32_t c = (uint32_t(v & 0xff) << 16) | (uint32_t(v >> 8) << 24) | uint32_t(v >> 8) | (2u << 8);
    return aval64(uint64_t(c) ^ uint64_t(uint32_t(sec(0)) ^ uint32_t(sec(0) >> 32)));
}
KX_HD inline uint64_t u8(uint8_t v) {
    uint32_t c = (uint32_t(v) << 16) | (uint32_t(v) << 24) | uint32_t(v) | (1u << 8);
    return aval64(uint64_t(c) ^ uint64_t(uint32_t(sec(0)) ^ uint32_t(sec(0) >> 32)));
}

// arbitrary byte strings: hash.Hash = xxh3.Hash (internal/hash/hash.go:26)
KX_HD inline uint64_t bytes(const uint8_t* in, size_t len) {
    if (len == 0) return aval64(sec(56) ^ sec(64));
    if (len < 4) {
        uint32_t c = (uint32_t(in[0]) << 16) | (uint32_t(in[len >> 1]) << 24) | in[len - 1] | (uint32_t(len) << 8);
        return aval64(uint64_t(c) ^ uint64_t(uint32_t(sec(0)) ^ uint32_t(sec(0) >> 32)));
    }
    if (len <= 8) return rrmxmx((uint64_t(r32(in + len - 4)) + (uint64_t(r32(in)) << 32)) ^ (sec(8) ^ sec(16)), len);
    if (len <= 16) {
        uint64_t lo = r64(in) ^ (sec(24) ^ sec(32)), hi = r64(in + len - 8) ^ (sec(40) ^ sec(48));
        return aval3(len + bswap(lo) + hi + fold(lo, hi));
    }
    if (len <= 128) {
        uint64_t acc = len * P64_1;
        int pairs = int((len - 1) / 32);   // 0..3 extra (front, back) pairs beyond the outermost one
        for (int i = pairs; i > 0; i--) { acc += mix(in + 16 * i, 32 * i); acc += mix(in + len - 16 * (i + 1), 32 * i + 16); }
        acc += mix(in, 0); acc += mix(in + len - 16, 16);
        return aval3(acc);
    }
    if (len <= 240) {
        uint64_t acc = len * P64_1;
        for (int i = 0; i < 8; i++) acc += mix(in + 16 * i, 16 * i);
        acc = aval3(acc);
        for (int i = 8; i < int(len / 16); i++) acc += mix(in + 16 * i, 16 * (i - 8) + 3);
        acc += mix(in + len - 16, 119);
        return aval3(acc);
    }
    uint64_t acc[8] = {P32_3, P64_1, P64_2, P64_3, P64_4, P32_2, P64_5, P32_1};
    const size_t per_block = 16, block = per_block * 64;
    size_t nblocks = (len - 1) / block;
    for (size_t b = 0; b < nblocks; b++) {
        for (size_t k = 0; k < per_block; k++) stripe(acc, in + b * block + k * 64, int(8 * k));
        for (int i = 0; i < 8; i++) { uint64_t x = acc[i]; x ^= x >> 47; x ^= sec(128 + 8 * i); acc[i] = x * P32_1; }
    }
    size_t tail = ((len - 1) - nblocks * block) / 64;
    for (size_t k = 0; k < tail; k++) stripe(acc, in + nblocks * block + k * 64, int(8 * k));
    stripe(acc, in + len - 64, 121);
    uint64_t h = len * P64_1;
    for (int i = 0; i < 4; i++) h += fold(acc[2 * i] ^ sec(11 + 16 * i), acc[2 * i + 1] ^ sec(19 + 16 * i));
    return aval3(h);
}

}  // namespace xxh3
}  // namespace kx
