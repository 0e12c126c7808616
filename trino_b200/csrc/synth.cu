// synth.cu — counter-based synthetic TPC-H-shaped columns generated directly in HBM.
// The CPU checker restates the same generators; tests compare them element by element.
#include "common.cuh"

namespace {

__host__ __device__ inline uint64_t feistel_perm(uint64_t i, uint64_t n, uint64_t seed)
{
    int bits = 1;
    while ((1ULL << bits) < n) bits++;
    int hb = (bits + 1) / 2;
    uint64_t hm = (1ULL << hb) - 1;
    uint64_t x = i;
    do {
        uint64_t l = x >> hb, r = x & hm;
        for (int round = 0; round < 4; round++) {
            uint64_t f = tg::splitmix64(r ^ (seed + 0x1000003ULL * (uint64_t)round)) & hm;
            uint64_t nl = r, nr = l ^ f;
            l = nl; r = nr;
        }
        x = (l << hb) | r;
    } while (x >= n);
    return x;
}

__host__ __device__ inline int64_t order_key(int64_t i) { return (i / 8) * 32 + (i % 8) + 1; }

__host__ __device__ inline int64_t lineitem_order_index(int64_t r)
{
    int64_t b = r / 28, w = r % 28;
    int64_t acc = 0;
    for (int64_t jj = 0; jj < 7; jj++) {
        int64_t c = 1 + ((jj + b) % 7);
        if (w < acc + c) return b * 7 + jj;
        acc += c;
    }
    return b * 7 + 6;
}

__global__ void synth_orders_kernel(int64_t n_total, int64_t first, int64_t count, uint64_t seed, int shuffle, int64_t* __restrict__ out)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < count; j += stride) {
        int64_t i = shuffle ? (int64_t)feistel_perm((uint64_t)(first + j), (uint64_t)n_total, seed) : first + j;
        out[j] = order_key(i);
    }
}

__global__ void synth_lineitem_keys_kernel(int64_t rows, int64_t first, int64_t count, uint64_t seed, int shuffle, int64_t* __restrict__ out)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < count; k += stride) {
        int64_t r = shuffle ? (int64_t)feistel_perm((uint64_t)(first + k), (uint64_t)rows, seed) : first + k;
        out[k] = order_key(lineitem_order_index(r));
    }
}

__global__ void synth_q1_kernel(int64_t n, int64_t first, uint64_t seed, int32_t* __restrict__ shipdate, int8_t* __restrict__ returnflag,
                                int8_t* __restrict__ linestatus, double* __restrict__ quantity, double* __restrict__ extendedprice,
                                double* __restrict__ discount, double* __restrict__ tax)
{
    int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; k < n; k += stride) {
        uint64_t x = tg::splitmix64(seed ^ (uint64_t)(first + k));
        uint64_t y = tg::splitmix64(x);
        int32_t sd = 8036 + (int32_t)(x % 2556);
        int32_t receipt = sd + 1 + (int32_t)((y >> 40) % 30);
        int64_t qty = 1 + (int64_t)((x >> 12) % 50);
        int64_t retail_cents = 90000 + (int64_t)(y % 20001);
        shipdate[k] = sd;
        linestatus[k] = sd > 9298 ? 'O' : 'F';
        returnflag[k] = receipt <= 9298 ? (((y >> 50) & 1) ? 'R' : 'A') : 'N';
        quantity[k] = (double)qty;
        extendedprice[k] = __ddiv_rn((double)(qty * retail_cents), 100.0);
        discount[k] = __ddiv_rn((double)((y >> 20) % 11), 100.0);
        tax[k] = __ddiv_rn((double)((y >> 30) % 9), 100.0);
    }
}

// o_custkey of order index i: uniform over the customers that have orders (TPC-H: custkey % 3 != 0, two thirds of them)
__host__ __device__ inline int64_t order_custkey(int64_t i, int64_t n_customers, uint64_t seed)
{
    uint64_t with_orders = (uint64_t)(n_customers - n_customers / 3);
    uint64_t j = tg::splitmix64(seed ^ (0x9E3779B97F4A7C15ULL * (uint64_t)(i + 1))) % with_orders;
    return (int64_t)((j / 2) * 3 + (j % 2) + 1);
}

__global__ void synth_orders_custkeys_kernel(int64_t n_total, int64_t first, int64_t count, uint64_t seed, int shuffle, int64_t n_customers, uint64_t cust_seed,
                                             int64_t* __restrict__ out)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < count; j += stride) {
        int64_t i = shuffle ? (int64_t)feistel_perm((uint64_t)(first + j), (uint64_t)n_total, seed) : first + j;
        out[j] = order_custkey(i, n_customers, cust_seed);
    }
}

__global__ void synth_sequence32_kernel(int32_t first_value, int64_t count, int32_t* __restrict__ out)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < count; j += stride) out[j] = first_value + (int32_t)j;
}

__global__ void synth_sequence_kernel(int64_t first_value, int64_t count, int64_t* __restrict__ out)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; j < count; j += stride) out[j] = first_value + j;
}

// TPC-DS store_sales shape: one thread makes 8 consecutive rows = one byte of each validity bitmap
struct StoreSalesRow { int64_t date_sk, item_sk, customer_sk, store_sk; double net_paid; bool customer_null, store_null; };
__host__ __device__ inline StoreSalesRow store_sales_row(int64_t r, uint64_t seed)
{
    uint64_t x = tg::splitmix64(seed ^ (uint64_t)r);
    uint64_t y = tg::splitmix64(x);
    StoreSalesRow o;
    o.date_sk = 2450816 + (int64_t)(x % 1823);                   // a 1 823-day window of date_dim's 73 049 days (2415022 ..)
    o.item_sk = 1 + (int64_t)((x >> 16) % 300000);
    o.customer_sk = 1 + (int64_t)(y % 12000000);
    o.store_sk = 1 + (int64_t)((y >> 32) % 1002);
    o.customer_null = ((x >> 40) % 1000) < 45;                   // 4.5 % NULL
    o.store_null = ((y >> 48) % 1000) < 45;
    o.net_paid = (double)((x >> 8) % 2000000) / 100.0;
    return o;
}

__global__ void synth_store_sales_kernel(int64_t n, int64_t first, uint64_t seed, int64_t* __restrict__ date_sk, int64_t* __restrict__ item_sk,
                                         int64_t* __restrict__ customer_sk, uint8_t* __restrict__ customer_valid, int64_t* __restrict__ store_sk,
                                         uint8_t* __restrict__ store_valid, double* __restrict__ net_paid, unsigned long long* __restrict__ both_valid)
{
    int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t nbytes = (n + 7) / 8;
    unsigned int both = 0;
    for (; b < nbytes; b += stride) {
        unsigned int cv = 0, sv = 0;
        for (int k = 0; k < 8; k++) {
            int64_t j = b * 8 + k;
            if (j >= n) break;
            StoreSalesRow r = store_sales_row(first + j, seed);
            date_sk[j] = r.date_sk;
            item_sk[j] = r.item_sk;
            customer_sk[j] = r.customer_null ? 0 : r.customer_sk;
            store_sk[j] = r.store_null ? 0 : r.store_sk;
            net_paid[j] = r.net_paid;
            if (!r.customer_null) cv |= 1u << k;
            if (!r.store_null) sv |= 1u << k;
            both += (!r.customer_null && !r.store_null) ? 1 : 0;
        }
        customer_valid[b] = (uint8_t)cv;
        store_valid[b] = (uint8_t)sv;
    }
    for (int off = 16; off > 0; off >>= 1) both += __shfl_xor_sync(0xffffffffu, both, off);
    if ((threadIdx.x & 31) == 0 && both) atomicAdd(both_valid, (unsigned long long)both);
}

// wrapping 64-bit sum of the values of a fixed-width column (value % mod when mod > 0); NULL rows are skipped
__global__ void column_sum_kernel(ColRef col, int64_t n, long long mod, unsigned long long* __restrict__ out)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (; i < n; i += stride) {
        if (!tg_valid(col.validity, i)) continue;
        long long v = tg_load_i64(col, i);
        acc += (unsigned long long)(mod > 0 ? v % mod : v);
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

}  // namespace

extern "C" int tgpu_column_sum(tgpu_ctx* ctx, const tgpu_column* device_column, int64_t mod, int64_t* out_sum)
{
    if (!ctx || !device_column || !out_sum) return TGPU_ERR_INVALID_ARGUMENT;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    DevColumn c;
    TG_TRY(tg_ingest_column(ctx, device_column, true, &c));
    if (c.elem_size() == 0) return tg_fail(ctx, TGPU_ERR_NOT_SUPPORTED, "tgpu_column_sum needs a fixed-width column");
    DevBuf acc;
    TG_TRY(acc.alloc(ctx, 8));
    TG_CUDA(ctx, cudaMemsetAsync(acc.p, 0, 8, ctx->stream));
    if (c.length > 0) TG_LAUNCH(ctx, column_sum_kernel, tg_grid(ctx, c.length, 1024, 8), 256, 0, tg_colref(c), c.length, (long long)mod, acc.as<unsigned long long>());
    return tg_read_i64(ctx, acc.p, out_sum);
}

extern "C" int64_t tgpu_synth_lineitem_rows(int64_t n_orders)
{
    int64_t full = n_orders / 7, rem = n_orders % 7;
    int64_t rows = full * 28;
    for (int64_t jj = 0; jj < rem; jj++) rows += 1 + ((jj + full) % 7);
    return rows;
}

extern "C" int tgpu_synth_orders_keys(tgpu_ctx* ctx, int64_t n_total, int64_t first, int64_t count, uint64_t seed, int shuffle, int64_t* out)
{
    if (!ctx || !out) return TGPU_ERR_INVALID_ARGUMENT;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return TGPU_OK;
    TG_LAUNCH(ctx, synth_orders_kernel, tg_grid(ctx, count, 256, 8), 256, 0, n_total, first, count, seed, shuffle, out);
    return TGPU_OK;
}

extern "C" int tgpu_synth_lineitem_keys(tgpu_ctx* ctx, int64_t n_orders, int64_t first, int64_t count, uint64_t seed, int shuffle, int64_t* out)
{
    if (!ctx || !out) return TGPU_ERR_INVALID_ARGUMENT;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return TGPU_OK;
    int64_t rows = tgpu_synth_lineitem_rows(n_orders);
    TG_LAUNCH(ctx, synth_lineitem_keys_kernel, tg_grid(ctx, count, 256, 8), 256, 0, rows, first, count, seed, shuffle, out);
    return TGPU_OK;
}

extern "C" int tgpu_synth_lineitem_q1(tgpu_ctx* ctx, int64_t n, int64_t first, uint64_t seed, int32_t* shipdate, int8_t* returnflag, int8_t* linestatus,
                                      double* quantity, double* extendedprice, double* discount, double* tax)
{
    if (!ctx) return TGPU_ERR_INVALID_ARGUMENT;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (n == 0) return TGPU_OK;
    TG_LAUNCH(ctx, synth_q1_kernel, tg_grid(ctx, n, 256, 8), 256, 0, n, first, seed, shipdate, returnflag, linestatus, quantity, extendedprice, discount, tax);
    return TGPU_OK;
}

extern "C" int tgpu_synth_orders_custkeys(tgpu_ctx* ctx, int64_t n_total, int64_t first, int64_t count, uint64_t seed, int shuffle, int64_t n_customers,
                                          uint64_t cust_seed, int64_t* out)
{
    if (!ctx || !out || n_customers < 3) return TGPU_ERR_INVALID_ARGUMENT;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return TGPU_OK;
    TG_LAUNCH(ctx, synth_orders_custkeys_kernel, tg_grid(ctx, count, 256, 8), 256, 0, n_total, first, count, seed, shuffle, n_customers, cust_seed, out);
    return TGPU_OK;
}

extern "C" int tgpu_synth_sequence(tgpu_ctx* ctx, int64_t first_value, int64_t count, int64_t* out)
{
    if (!ctx || !out) return TGPU_ERR_INVALID_ARGUMENT;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return TGPU_OK;
    TG_LAUNCH(ctx, synth_sequence_kernel, tg_grid(ctx, count, 1024, 8), 256, 0, first_value, count, out);
    return TGPU_OK;
}

extern "C" int tgpu_synth_store_sales(tgpu_ctx* ctx, int64_t n, int64_t first, uint64_t seed, int64_t* date_sk, int64_t* item_sk, int64_t* customer_sk,
                                      uint8_t* customer_valid, int64_t* store_sk, uint8_t* store_valid, double* net_paid, int64_t* rows_with_both_keys)
{
    if (!ctx || !date_sk || !item_sk || !customer_sk || !customer_valid || !store_sk || !store_valid || !net_paid) return TGPU_ERR_INVALID_ARGUMENT;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (rows_with_both_keys) *rows_with_both_keys = 0;
    if (n == 0) return TGPU_OK;
    DevBuf cnt;
    TG_TRY(cnt.alloc(ctx, 8));
    TG_CUDA(ctx, cudaMemsetAsync(cnt.p, 0, 8, ctx->stream));
    TG_LAUNCH(ctx, synth_store_sales_kernel, tg_grid(ctx, (n + 7) / 8, 256, 8), 256, 0, n, first, seed, date_sk, item_sk, customer_sk, customer_valid, store_sk,
              store_valid, net_paid, cnt.as<unsigned long long>());
    int64_t v = 0;
    TG_TRY(tg_read_i64(ctx, cnt.p, &v));
    if (rows_with_both_keys) *rows_with_both_keys = v;
    return TGPU_OK;
}

extern "C" int tgpu_synth_sequence32(tgpu_ctx* ctx, int32_t first_value, int64_t count, int32_t* out)
{
    if (!ctx || !out) return TGPU_ERR_INVALID_ARGUMENT;
    TG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (count == 0) return TGPU_OK;
    TG_LAUNCH(ctx, synth_sequence32_kernel, tg_grid(ctx, count, 1024, 8), 256, 0, first_value, count, out);
    return TGPU_OK;
}
