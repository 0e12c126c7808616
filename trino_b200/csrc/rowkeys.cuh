// rowkeys.cuh — row-level key helpers over a set of key columns: the reference's row hash
// (M/operator/InterpretedHashGenerator.java:102-110: h = 31*h + typeHash(col), NULL -> 0) and the two notions of
// key equality the operators need: EQUAL (joins; M/operator/SimplePagesHashStrategy.java:194-260: NaN never matches,
// -0.0 == +0.0) and IDENTICAL (group by; S/type/DoubleType.java:218-229: NaN identical to NaN, NULL identical to NULL).
#pragma once
#include "common.cuh"

namespace tg {

constexpr int MAX_KEY_COLS = 8;

struct KeyCols {
    int32_t count;
    ColRef cols[MAX_KEY_COLS];
    const int32_t* offsets[MAX_KEY_COLS];   // UTF8 only
    int32_t is_utf8[MAX_KEY_COLS];
    int32_t is_double[MAX_KEY_COLS];        // 1 = DOUBLE (raw bits in 8 bytes), 2 = REAL (raw bits in 4 bytes)
    int32_t is_const[MAX_KEY_COLS];         // partitionConstants (M/operator/output/PagePartitioner.java:436-451): no column, one hash for every row
    uint64_t const_hash[MAX_KEY_COLS];
};

static inline void key_cols_set(KeyCols* k, int c, const DevColumn& col)
{
    k->cols[c] = tg_colref(col);
    k->offsets[c] = col.offsets;
    k->is_utf8[c] = col.type == TGPU_UTF8;
    k->is_double[c] = col.type == TGPU_FLOAT64 ? 1 : col.type == TGPU_FLOAT32 ? 2 : 0;
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t type_hash(const KeyCols& k, int c, int64_t row)
{
    if (k.is_const[c]) return k.const_hash[c];
    const ColRef& col = k.cols[c];
    if (!tg_valid(col.validity, row)) return 0;   // NULL_HASH_CODE (S/type/TypeUtils.java:34)
    if (k.is_utf8[c]) {
        int32_t a = k.offsets[c][row], b = k.offsets[c][row + 1];
        return xxh64_bytes((const uint8_t*)col.data + a, b - a);
    }
    if (col.elem == 16) {      // Int128ArrayBlock: high word first
        const int64_t* w = (const int64_t*)col.data + row * 2;
        return hash_int128(w[0], w[1]);
    }
    int64_t v = tg_load_i64(col, row);
    return k.is_double[c] == 1 ? hash_double_bits(v) : k.is_double[c] == 2 ? hash_real_bits(v) : hash_long(v);
}

__device__ __forceinline__ uint64_t row_hash(const KeyCols& k, int64_t row)
{
    uint64_t h = 0;
    for (int c = 0; c < k.count; c++) h = combine_hash(h, type_hash(k, c, row));
    return h;
}

// The family of row hashes the fingerprint tables are keyed by.  attempt 0 is the reference's row hash; attempt a >= 1 is an independent
// 64-bit hash of the same key values (strings are re-hashed from their bytes with seed a, fixed-width values are re-mixed), used only for
// key tuples whose lower-attempt hash is already owned by a different tuple (join.cu: build / probe retries).  -0.0 hashes like +0.0.
__device__ __forceinline__ uint64_t row_hash_attempt(const KeyCols& k, int64_t row, int attempt)
{
    if (attempt == 0) return row_hash(k, row);
    uint64_t h = 0x9E3779B97F4A7C15ULL * (uint64_t)attempt;
    for (int c = 0; c < k.count; c++) {
        const ColRef& col = k.cols[c];
        uint64_t t;
        if (!tg_valid(col.validity, row)) t = 0x5851F42D4C957F2DULL;
        else if (k.is_utf8[c]) {
            int32_t a = k.offsets[c][row], b = k.offsets[c][row + 1];
            t = xxh64_bytes((const uint8_t*)col.data + a, b - a, (uint64_t)attempt);
        }
        else if (col.elem == 16) {
            const uint64_t* w = (const uint64_t*)col.data + row * 2;
            t = murmur3_mix(w[0] + 0xD1B54A32D192ED03ULL * (uint64_t)attempt) ^ murmur3_mix(w[1] + 0x9E3779B97F4A7C15ULL * (uint64_t)(attempt + 1));
        }
        else {
            uint64_t u = (uint64_t)tg_load_i64(col, row);
            if (k.is_double[c] == 2) u = (uint32_t)u;
            if (k.is_double[c] == 1 && (u << 1) == 0) u = 0;
            if (k.is_double[c] == 2 && (u << 33) == 0) u = 0;
            t = murmur3_mix(u + 0xD1B54A32D192ED03ULL * (uint64_t)attempt);
        }
        h = murmur3_mix(h ^ t) * 31 + (uint64_t)(c + 1);
    }
    return h;
}

// true when the row can take part in an equi-join: no NULL key and no NaN key
__device__ __forceinline__ bool row_joinable(const KeyCols& k, int64_t row)
{
    for (int c = 0; c < k.count; c++) {
        const ColRef& col = k.cols[c];
        if (!tg_valid(col.validity, row)) return false;
        if (k.is_double[c] == 1) {
            unsigned long long u = (unsigned long long)tg_load_i64(col, row);
            if ((u & 0x7FFFFFFFFFFFFFFFULL) > 0x7FF0000000000000ULL) return false;
        }
        else if (k.is_double[c] == 2) {
            unsigned int u = (unsigned int)tg_load_i64(col, row);
            if ((u & 0x7FFFFFFFu) > 0x7F800000u) return false;
        }
    }
    return true;
}

// value equality of non-NULL keys column by column (EQUAL for joinable rows; with nan_equal also IDENTICAL on values)
__device__ __forceinline__ bool col_value_equal(const KeyCols& a, int c, int64_t ra, const KeyCols& b, int64_t rb, bool nan_equal)
{
    if (a.is_utf8[c]) {
        int32_t a0 = a.offsets[c][ra], la = a.offsets[c][ra + 1] - a0;
        int32_t b0 = b.offsets[c][rb], lb = b.offsets[c][rb + 1] - b0;
        if (la != lb) return false;
        const uint8_t* pa = (const uint8_t*)a.cols[c].data + a0;
        const uint8_t* pb = (const uint8_t*)b.cols[c].data + b0;
        for (int32_t i = 0; i < la; i++)
            if (pa[i] != pb[i]) return false;
        return true;
    }
    if (a.cols[c].elem == 16) {
        const int64_t* wa = (const int64_t*)a.cols[c].data + ra * 2;
        const int64_t* wb = (const int64_t*)b.cols[c].data + rb * 2;
        return wa[0] == wb[0] && wa[1] == wb[1];
    }
    int64_t va = tg_load_i64(a.cols[c], ra), vb = tg_load_i64(b.cols[c], rb);
    if (a.is_double[c] == 1) {
        double x = __longlong_as_double(va), y = __longlong_as_double(vb);
        if (nan_equal && x != x && y != y) return true;
        return x == y;
    }
    if (a.is_double[c] == 2) {
        float x = __int_as_float((int)va), y = __int_as_float((int)vb);
        if (nan_equal && x != x && y != y) return true;
        return x == y;
    }
    return va == vb;
}

__device__ __forceinline__ bool rows_equal_for_join(const KeyCols& a, int64_t ra, const KeyCols& b, int64_t rb)
{
    for (int c = 0; c < a.count; c++)
        if (!col_value_equal(a, c, ra, b, rb, false)) return false;
    return true;
}
#endif

}  // namespace tg
