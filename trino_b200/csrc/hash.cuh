// hash.cuh — the reference's hash arithmetic as __host__ __device__ inlines.
// Fused into every consumer kernel; never a pass of its own.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define TG_HD __host__ __device__ __forceinline__
#else
#define TG_HD inline
#endif

namespace tg {

constexpr uint64_t XXP1 = 0x9E3779B185EBCA87ULL, XXP2 = 0xC2B2AE3D27D4EB4FULL, XXP3 = 0x165667B19E3779F9ULL,
                   XXP4 = 0x85EBCA77C2B2AE63ULL, XXP5 = 0x27D4EB2F165667C5ULL;

TG_HD uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

// BIGINT / INTEGER / DATE / SMALLINT / TINYINT hash code (S/type/AbstractLongType.java:121-125;
// narrower integers are sign-extended first, S/type/AbstractIntType.java:183-187)
TG_HD uint64_t hash_long(int64_t v) { return rotl64((uint64_t)v * XXP2, 31) * XXP1; }

// DOUBLE hash code (S/type/DoubleType.java:199-206): -0.0 -> +0.0, every NaN -> canonical NaN
// (Double.doubleToLongBits), then hash_long of the bits
TG_HD uint64_t hash_double_bits(int64_t bits)
{
    uint64_t u = (uint64_t)bits;
    if ((u << 1) == 0) u = 0;                                            // +-0.0
    if ((u & 0x7FFFFFFFFFFFFFFFULL) > 0x7FF0000000000000ULL) u = 0x7FF8000000000000ULL;  // NaN
    return hash_long((int64_t)u);
}

// REAL hash code (S/type/RealType.java:151-159): -0.0 -> +0.0, every NaN -> 0x7fc00000 (Float.floatToIntBits), the int widened with its
// sign, then hash_long.  `bits`: the float's raw bits (any upper half is ignored)
TG_HD uint64_t hash_real_bits(int64_t bits)
{
    uint32_t u = (uint32_t)bits;
    if ((u << 1) == 0) u = 0;
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) u = 0x7FC00000u;
    return hash_long((int64_t)(int32_t)u);
}

// murmur3 fmix64 (M/operator/join/PagesHash.java:44-50 == fastutil HashCommon.murmurHash3 used at
// M/operator/BigintGroupByHash.java:297-300)
TG_HD uint64_t murmur3_mix(uint64_t x)
{
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

// CombineHashFunction.getHash (M/operator/scalar/CombineHashFunction.java:29-32)
TG_HD uint64_t combine_hash(uint64_t prev, uint64_t v) { return 31 * prev + v; }

TG_HD uint64_t xxh64_long(int64_t v);

TG_HD uint64_t reverse_bits64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    x = ((x >> 8) & 0x00FF00FF00FF00FFULL) | ((x & 0x00FF00FF00FF00FFULL) << 8);
    x = ((x >> 16) & 0x0000FFFF0000FFFFULL) | ((x & 0x0000FFFF0000FFFFULL) << 16);
    return (x >> 32) | (x << 32);
#endif
}

// Partition of a raw row hash.
//   count > 0: HashGenerator.processRawHash (M/operator/HashGenerator.java:41-46) - the bucket of the inter-stage hash distribution.
//   count < 0: LocalPartitionGenerator.getPartition (M/operator/exchange/LocalPartitionGenerator.java:45-77) over -count partitions
//              (a power of two): (int) XxHash64.hash(Long.reverse(rawHash)) & (partitionCount - 1) - the local exchange re-mixes the
//              bits so that it does not reuse the hash the stages were distributed by.
TG_HD int32_t process_raw_hash(uint64_t raw, int32_t count)
{
    if (count < 0) return (int32_t)(uint32_t)xxh64_long((int64_t)reverse_bits64(raw)) & (-count - 1);
    uint32_t x = (uint32_t)(raw ^ (raw >> 32));
    return (int32_t)(((uint64_t)x * (uint64_t)(uint32_t)count) >> 32);
}

// XXH64 (io.airlift.slice.XxHash64, public algorithm) of a byte range; unaligned-safe.  The reference always hashes with seed 0; other
// seeds give the independent hash functions the string dictionary rehashes colliding strings with (strdict.cuh)
TG_HD uint64_t xxh64_bytes(const uint8_t* p, int64_t len, uint64_t seed = 0)
{
    const uint8_t* end = p + len;
    uint64_t h;
    auto rd64 = [](const uint8_t* q) { uint64_t v = 0; for (int i = 7; i >= 0; i--) v = (v << 8) | q[i]; return v; };
    auto rd32 = [](const uint8_t* q) { uint32_t v = 0; for (int i = 3; i >= 0; i--) v = (v << 8) | q[i]; return v; };
    if (len >= 32) {
        uint64_t v1 = seed + XXP1 + XXP2, v2 = seed + XXP2, v3 = seed, v4 = seed - XXP1;
        const uint8_t* limit = end - 32;
        do {
            v1 = rotl64(v1 + rd64(p) * XXP2, 31) * XXP1;
            v2 = rotl64(v2 + rd64(p + 8) * XXP2, 31) * XXP1;
            v3 = rotl64(v3 + rd64(p + 16) * XXP2, 31) * XXP1;
            v4 = rotl64(v4 + rd64(p + 24) * XXP2, 31) * XXP1;
            p += 32;
        } while (p <= limit);
        h = rotl64(v1, 1) + rotl64(v2, 7) + rotl64(v3, 12) + rotl64(v4, 18);
        h ^= rotl64(v1 * XXP2, 31) * XXP1; h = h * XXP1 + XXP4;
        h ^= rotl64(v2 * XXP2, 31) * XXP1; h = h * XXP1 + XXP4;
        h ^= rotl64(v3 * XXP2, 31) * XXP1; h = h * XXP1 + XXP4;
        h ^= rotl64(v4 * XXP2, 31) * XXP1; h = h * XXP1 + XXP4;
    }
    else {
        h = seed + XXP5;
    }
    h += (uint64_t)len;
    while (p + 8 <= end) { h ^= rotl64(rd64(p) * XXP2, 31) * XXP1; h = rotl64(h, 27) * XXP1 + XXP4; p += 8; }
    if (p + 4 <= end) { h ^= (uint64_t)rd32(p) * XXP1; h = rotl64(h, 23) * XXP2 + XXP3; p += 4; }
    while (p < end) { h ^= (uint64_t)(*p) * XXP5; h = rotl64(h, 11) * XXP1; p++; }
    h ^= h >> 33; h *= XXP2; h ^= h >> 29; h *= XXP3; h ^= h >> 32;
    return h;
}

// io.airlift.slice.XxHash64.hash(long) - XXH64 (seed 0) of the value's eight little-endian bytes, written out: the 8-byte input takes
// the short-input branch and one 8-byte tail round.  (airlift slice is a dependency of the reference, not vendored in it; the
// algorithm is the published XXH64, pinned on its test vectors in tests/test_oracle_hashes.py.)
TG_HD uint64_t xxh64_long(int64_t v)
{
    uint64_t h = XXP5 + 8;
    h ^= rotl64((uint64_t)v * XXP2, 31) * XXP1;
    h = rotl64(h, 27) * XXP1 + XXP4;
    h ^= h >> 33; h *= XXP2; h ^= h >> 29; h *= XXP3; h ^= h >> 32;
    return h;
}

// long DECIMAL hash code (S/type/LongDecimalType.java:203-229): XxHash64.hash(high) ^ XxHash64.hash(low)
TG_HD uint64_t hash_int128(int64_t high, int64_t low) { return xxh64_long(high) ^ xxh64_long(low); }

// splitmix64: counter-based generator of the synthetic TPC-H-shaped columns
TG_HD uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

}  // namespace tg
