// driver_loop.cpp — a Driver-loop harness over the C ABI: drives libtrino_gpu.so the way the Java side would.
//
// TEST INFRASTRUCTURE.  It plays the roles of Driver.processInternal (M/operator/Driver.java:391-424: needsInput / addInput / getOutput /
// finish / isFinished) and of the Java shim's PageMarshaller (java/io/trino/spi/block/PageMarshaller.java): the input are "Java pages" -
// 8192 positions (PageProcessor.java:58) of heap arrays with one-byte-per-position boolean[] null maps - which are copied batch by batch
// (>= 1 M rows) into ONE pinned staging region and handed over as a single host tgpu_page with TGPU_COL_NULLS_BYTEMAP columns; outputs are
// copied back with tgpu_page_copy_to_host and cut into <= 8192-row pages again.  Results are checked against the oracle (liboracle.so),
// and the wall-clock rows/s of the whole loop - marshalling copies, H2D, kernels, D2H, un-marshalling - is printed as one JSON line:
// the end-to-end number a JVM-hosted operator can expect, next to bench.py's `e2e` (which feeds 32 M-row pinned pages directly).
//
//   driver_loop q1   <rows> [drivers]     fused filter + project + GROUP BY (TPC-H Q1) through tgpu_agg_create with a pre-program
//   driver_loop join <orders> [drivers]   hash build from pages, then LookupJoinOperator probes by `drivers` threads sharing the lookup
#include <dlfcn.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/trino_gpu.h"
#include "../../oracle/oracle.h"

namespace {

constexpr int64_t PAGE_ROWS = 8192;
constexpr int64_t BATCH_ROWS = 1 << 20;

#define CHECK(ctx, call)                                                                                  \
    do {                                                                                                  \
        int _s = (call);                                                                                  \
        if (_s != 0) { fprintf(stderr, "%s failed: %d %s\n", #call, _s, tgpu_last_error(ctx)); exit(2); } \
    } while (0)

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// a "Java page": one heap array + one boolean[] per channel
struct JavaColumn {
    int32_t type;
    std::vector<uint8_t> values;      // positions x width bytes
    std::vector<uint8_t> is_null;     // empty = the block has no null array (mayHaveNull() == false)
};
struct JavaPage {
    int64_t rows = 0;
    std::vector<JavaColumn> cols;
};

int width_of(int32_t type) { return type == TGPU_INT64 || type == TGPU_FLOAT64 ? 8 : type == TGPU_INT32 ? 4 : type == TGPU_INT16 ? 2 : 1; }

// PageMarshaller.flush: concatenate the batch column by column into the pinned staging region, describe it as one host tgpu_page
struct Marshaller {
    uint8_t* staging = nullptr;
    size_t staging_bytes = 0, used = 0;
    std::vector<tgpu_column> cols;
    tgpu_page page;

    uint8_t* reserve(size_t bytes)
    {
        size_t at = (used + 63) & ~(size_t)63;
        if (at + bytes > staging_bytes) { fprintf(stderr, "staging region too small\n"); exit(2); }
        used = at + bytes;
        return staging + at;
    }

    const tgpu_page* flush(const std::vector<const JavaPage*>& batch)
    {
        used = 0;
        int64_t rows = 0;
        for (auto* p : batch) rows += p->rows;
        size_t channels = batch[0]->cols.size();
        cols.assign(channels, tgpu_column());
        for (size_t c = 0; c < channels; c++) {
            int32_t type = batch[0]->cols[c].type;
            int w = width_of(type);
            bool any_nulls = false;
            for (auto* p : batch) any_nulls |= !p->cols[c].is_null.empty();
            uint8_t* data = reserve((size_t)rows * w);
            uint8_t* nulls = any_nulls ? reserve((size_t)rows) : nullptr;
            int64_t at = 0;
            for (auto* p : batch) {
                memcpy(data + at * w, p->cols[c].values.data(), (size_t)p->rows * w);
                if (nulls) {
                    if (p->cols[c].is_null.empty()) memset(nulls + at, 0, (size_t)p->rows);
                    else memcpy(nulls + at, p->cols[c].is_null.data(), (size_t)p->rows);
                }
                at += p->rows;
            }
            tgpu_column& col = cols[c];
            memset(&col, 0, sizeof(col));
            col.type = type;
            col.flags = any_nulls ? TGPU_COL_NULLS_BYTEMAP : 0;
            col.length = rows;
            col.data = data;
            col.validity = nulls;
        }
        page.num_columns = (int32_t)channels;
        page.flags = 0;
        page.num_rows = rows;
        page.columns = cols.data();
        return &page;
    }
};

// the relevant slice of Driver.processInternal for one operator: returns the device pages through `sink`
template <class Sink>
void drive(tgpu_ctx* ctx, tgpu_op* op, Marshaller& m, const std::vector<JavaPage>& pages, size_t first, size_t step, Sink sink)
{
    std::vector<const JavaPage*> batch;
    int64_t batch_rows = 0;
    auto drain = [&]() {
        while (true) {
            tgpu_page* out = nullptr;
            CHECK(ctx, tgpu_op_get_output(op, &out));
            if (!out) break;
            sink(out, batch);
            tgpu_page_release(ctx, out);
        }
    };
    auto flush = [&]() {
        if (batch.empty()) return;
        int needs = 0;
        CHECK(ctx, tgpu_op_needs_input(op, &needs));
        if (!needs) { fprintf(stderr, "operator refuses input with nothing pending\n"); exit(2); }
        CHECK(ctx, tgpu_op_add_input(op, m.flush(batch)));
        drain();
        batch.clear();
        batch_rows = 0;
    };
    for (size_t i = first; i < pages.size(); i += step) {
        batch.push_back(&pages[i]);
        batch_rows += pages[i].rows;
        if (batch_rows >= BATCH_ROWS) flush();
    }
    flush();
    CHECK(ctx, tgpu_op_finish(op));
    CHECK(ctx, tgpu_op_finish(op));        // finish() is re-entrant (Driver.java:380-388)
    int finished = 0;
    while (true) {
        drain();
        CHECK(ctx, tgpu_op_is_finished(op, &finished));
        if (finished) break;
    }
}

tgpu_operand col(int ch) { tgpu_operand o; memset(&o, 0, sizeof(o)); o.kind = TGPU_OPND_COLUMN; o.index = ch; return o; }
tgpu_operand tmp(int t) { tgpu_operand o; memset(&o, 0, sizeof(o)); o.kind = TGPU_OPND_TEMP; o.index = t; return o; }
tgpu_operand cd(double v) { tgpu_operand o; memset(&o, 0, sizeof(o)); o.kind = TGPU_OPND_CONST; o.imm.f64 = v; return o; }
tgpu_operand ci(int64_t v) { tgpu_operand o; memset(&o, 0, sizeof(o)); o.kind = TGPU_OPND_CONST; o.imm.i64 = v; return o; }
tgpu_expr_insn insn(int op, int vtype, int dst, tgpu_operand a, tgpu_operand b)
{
    tgpu_expr_insn i;
    memset(&i, 0, sizeof(i));
    i.op = op; i.vtype = vtype; i.dst = dst; i.a = a; i.b = b;
    return i;
}

// ------------------------------------------------------------------------------------------------ Q1
int run_q1(int64_t n, int drivers)
{
    // synthetic lineitem columns (the oracle's generator), cut into Java pages; l_discount carries a boolean[] (no NULLs set) so the
    // byte-map path is exercised, the rest have no null array
    std::vector<int32_t> shipdate(n);
    std::vector<int8_t> returnflag(n), linestatus(n);
    std::vector<double> quantity(n), extendedprice(n), discount(n), tax(n);
    orc_synth_lineitem_q1(n, 0, 0x7C01, shipdate.data(), returnflag.data(), linestatus.data(), quantity.data(), extendedprice.data(), discount.data(), tax.data());
    std::vector<JavaPage> pages;
    for (int64_t lo = 0; lo < n; lo += PAGE_ROWS) {
        int64_t m = std::min(PAGE_ROWS, n - lo);
        JavaPage p;
        p.rows = m;
        auto add = [&](int32_t type, const void* src, int w, bool nullmap) {
            JavaColumn c;
            c.type = type;
            c.values.assign((const uint8_t*)src + lo * w, (const uint8_t*)src + (lo + m) * w);
            if (nullmap) c.is_null.assign((size_t)m, 0);
            p.cols.push_back(std::move(c));
        };
        add(TGPU_INT32, shipdate.data(), 4, false);
        add(TGPU_INT8, returnflag.data(), 1, false);
        add(TGPU_INT8, linestatus.data(), 1, false);
        add(TGPU_FLOAT64, quantity.data(), 8, false);
        add(TGPU_FLOAT64, extendedprice.data(), 8, false);
        add(TGPU_FLOAT64, discount.data(), 8, true);
        add(TGPU_FLOAT64, tax.data(), 8, false);
        pages.push_back(std::move(p));
    }
    // filter l_shipdate <= 10471; projections returnflag, linestatus, quantity, extendedprice, ep*(1-d), ep*(1-d)*(1+t), discount (q01.sql)
    tgpu_expr_insn insns[5] = {
        insn(TGPU_EX_LE, TGPU_V_BIGINT, 0, col(0), ci(10471)),
        insn(TGPU_EX_SUB, TGPU_V_DOUBLE, 1, cd(1.0), col(5)),
        insn(TGPU_EX_MUL, TGPU_V_DOUBLE, 2, col(4), tmp(1)),
        insn(TGPU_EX_ADD, TGPU_V_DOUBLE, 3, cd(1.0), col(6)),
        insn(TGPU_EX_MUL, TGPU_V_DOUBLE, 4, tmp(2), tmp(3)),
    };
    tgpu_projection proj[7] = {{0, 1, 0}, {0, 2, 0}, {0, 3, 0}, {0, 4, 0}, {1, 2, TGPU_V_DOUBLE}, {1, 4, TGPU_V_DOUBLE}, {0, 5, 0}};
    tgpu_expr_program prog;
    memset(&prog, 0, sizeof(prog));
    prog.num_insns = 5; prog.insns = insns; prog.filter_temp = 0; prog.num_filter_insns = 1; prog.num_projections = 7; prog.projections = proj;
    int32_t keys[2] = {0, 1};
    tgpu_agg_fn fns[8] = {{TGPU_AGG_SUM, 2, -1, 0}, {TGPU_AGG_SUM, 3, -1, 0}, {TGPU_AGG_SUM, 4, -1, 0}, {TGPU_AGG_SUM, 5, -1, 0},
                          {TGPU_AGG_AVG, 2, -1, 0}, {TGPU_AGG_AVG, 3, -1, 0}, {TGPU_AGG_AVG, 6, -1, 0}, {TGPU_AGG_COUNT_STAR, -1, -1, 0}};
    tgpu_agg_spec spec;
    memset(&spec, 0, sizeof(spec));
    spec.num_keys = 2; spec.key_channels = keys; spec.step = TGPU_STEP_PARTIAL; spec.num_aggs = 8; spec.aggs = fns; spec.expected_groups = 16; spec.pre = &prog;
    spec.group_id_key = -1;
    // `drivers` PARTIAL aggregations in parallel (one per driver thread, as task.concurrency drivers would), then one FINAL over their outputs
    struct Partial { std::vector<int8_t> k0, k1; std::vector<std::vector<int64_t>> cols; };
    std::vector<Partial> partials(drivers);
    std::vector<tgpu_ctx*> ctxs(drivers);
    std::vector<Marshaller> ms(drivers);
    for (int d = 0; d < drivers; d++) {
        CHECK(nullptr, tgpu_ctx_create(0, &ctxs[d]));
        ms[d].staging_bytes = (size_t)(BATCH_ROWS + PAGE_ROWS) * 48 + (1 << 20);
        void* p = nullptr;
        CHECK(ctxs[d], tgpu_host_alloc_pinned(ms[d].staging_bytes, &p));
        ms[d].staging = (uint8_t*)p;
    }
    auto one_pass = [&]() {
        std::vector<std::thread> threads;
        for (int d = 0; d < drivers; d++) {
            threads.emplace_back([&, d]() {
                tgpu_op* op = nullptr;
                CHECK(ctxs[d], tgpu_agg_create(ctxs[d], &spec, &op));
                Partial& out = partials[d];
                out = Partial();
                drive(ctxs[d], op, ms[d], pages, (size_t)d, (size_t)drivers, [&](tgpu_page* dev, const std::vector<const JavaPage*>&) {
                    // intermediate page: 2 INT8 keys, then 12 state columns of 8 bytes (include/trino_gpu.h: intermediate state layout)
                    int64_t g = dev->num_rows;
                    std::vector<tgpu_column> hc(dev->num_columns);
                    std::vector<std::vector<uint8_t>> bufs(dev->num_columns), valid(dev->num_columns);
                    for (int c = 0; c < dev->num_columns; c++) {
                        memset(&hc[c], 0, sizeof(tgpu_column));
                        hc[c].type = dev->columns[c].type;
                        hc[c].length = g;
                        bufs[c].assign((size_t)g * 8, 0);
                        valid[c].assign((size_t)g / 8 + 8, 0);
                        hc[c].data = bufs[c].data();
                        hc[c].validity = valid[c].data();
                    }
                    tgpu_page host = {dev->num_columns, 0, g, hc.data()};
                    CHECK(ctxs[d], tgpu_page_copy_to_host(ctxs[d], dev, &host));
                    out.cols.resize(dev->num_columns - 2);
                    for (int64_t r = 0; r < g; r++) {
                        out.k0.push_back((int8_t)bufs[0][r]);
                        out.k1.push_back((int8_t)bufs[1][r]);
                        for (int c = 2; c < dev->num_columns; c++) out.cols[c - 2].push_back(((const int64_t*)bufs[c].data())[r]);
                    }
                });
                tgpu_op_close(op);
            });
        }
        for (auto& t : threads) t.join();
    };
    one_pass();     // warm-up (NVRTC specialisation, pool growth)
    double t0 = now_s();
    one_pass();
    double secs = now_s() - t0;
    // FINAL step over the partial rows (tiny): sum / avg(count, sum) / count per (returnflag, linestatus), compared with the oracle's Q1
    orc_q1_result want;
    orc_q1_run(n, shipdate.data(), returnflag.data(), linestatus.data(), quantity.data(), extendedprice.data(), discount.data(), tax.data(), 10471, 1, &want);
    int bad = 0;
    for (int g = 0; g < want.num_groups; g++) {
        double sum_qty = 0, cnt_qty = 0;
        int64_t count = 0;
        for (auto& p : partials)
            for (size_t r = 0; r < p.k0.size(); r++)
                if (p.k0[r] == want.returnflag[g] && p.k1[r] == want.linestatus[g]) {
                    double v;
                    memcpy(&v, &p.cols[0][r], 8);
                    sum_qty += v;
                    cnt_qty += (double)p.cols[4][r];     // avg(quantity) state: count
                    count += p.cols[10][r];              // count(*) state
                }
        if (count != want.count_order[g]) bad++;
        if (std::fabs(sum_qty - want.sum_qty[g]) > 1e-6 * std::fabs(want.sum_qty[g])) bad++;      // north_star tolerance for DOUBLE aggregates
        if ((int64_t)cnt_qty != want.count_order[g]) bad++;
    }
    printf("{\"harness\": \"driver_loop\", \"mode\": \"q1\", \"rows\": %lld, \"drivers\": %d, \"java_page_rows\": %lld, \"batch_rows\": %lld, \"seconds\": %.6f, "
           "\"rows_per_s\": %.1f, \"groups\": %d, \"mismatches\": %d}\n",
           (long long)n, drivers, (long long)PAGE_ROWS, (long long)BATCH_ROWS, secs, (double)n / secs, want.num_groups, bad);
    for (int d = 0; d < drivers; d++) { tgpu_host_free_pinned(ms[d].staging); tgpu_ctx_destroy(ctxs[d]); }
    return bad ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------ join
int run_join(int64_t n_orders, int drivers)
{
    int64_t rows = orc_synth_lineitem_rows(n_orders);
    std::vector<int64_t> okeys(n_orders), lkeys(rows);
    orc_synth_orders_keys(n_orders, 0, n_orders, 0x7C02, 1, okeys.data());
    orc_synth_lineitem_keys(n_orders, 0, rows, 0x7C01, 0, lkeys.data());
    auto make_pages = [&](const std::vector<int64_t>& keys, bool build) {
        std::vector<JavaPage> pages;
        int64_t n = (int64_t)keys.size();
        for (int64_t lo = 0; lo < n; lo += PAGE_ROWS) {
            int64_t m = std::min(PAGE_ROWS, n - lo);
            JavaPage p;
            p.rows = m;
            JavaColumn k, v;
            k.type = TGPU_INT64;
            k.values.assign((const uint8_t*)(keys.data() + lo), (const uint8_t*)(keys.data() + lo + m));
            if (!build) k.is_null.assign((size_t)m, 0);       // the probe key block carries a boolean[] without NULLs
            v.type = build ? TGPU_INT64 : TGPU_FLOAT64;
            v.values.resize((size_t)m * 8);
            for (int64_t i = 0; i < m; i++) {
                if (build) { int64_t pay = keys[lo + i] % 2557; memcpy(&v.values[i * 8], &pay, 8); }
                else { double price = (double)keys[lo + i] * 0.5; memcpy(&v.values[i * 8], &price, 8); }
            }
            p.cols.push_back(std::move(k));
            p.cols.push_back(std::move(v));
            pages.push_back(std::move(p));
        }
        return pages;
    };
    std::vector<JavaPage> build_pages = make_pages(okeys, true), probe_pages = make_pages(lkeys, false);
    std::vector<tgpu_ctx*> ctxs(drivers);
    std::vector<Marshaller> ms(drivers);
    for (int d = 0; d < drivers; d++) {
        CHECK(nullptr, tgpu_ctx_create(0, &ctxs[d]));
        ms[d].staging_bytes = (size_t)(BATCH_ROWS + PAGE_ROWS) * 24 + (1 << 20);
        void* p = nullptr;
        CHECK(ctxs[d], tgpu_host_alloc_pinned(ms[d].staging_bytes, &p));
        ms[d].staging = (uint8_t*)p;
    }
    // HashBuilderOperator: addInput page batches, finish() builds the table
    int32_t bkey[1] = {0}, bout[1] = {1};
    tgpu_join_build_spec bspec = {1, bkey, 1, bout, n_orders};
    tgpu_op* builder = nullptr;
    CHECK(ctxs[0], tgpu_join_build_create(ctxs[0], &bspec, &builder));
    double tb = now_s();
    drive(ctxs[0], builder, ms[0], build_pages, 0, 1, [](tgpu_page*, const std::vector<const JavaPage*>&) {});
    double build_secs = now_s() - tb;
    tgpu_lookup* lookup = nullptr;
    CHECK(ctxs[0], tgpu_join_build_get_lookup(builder, &lookup));
    int32_t pkey[1] = {0}, pout[2] = {0, 1};
    tgpu_join_probe_spec pspec = {TGPU_JOIN_INNER, 0, 1, pkey, 2, pout};
    std::atomic<long long> out_rows{0}, wrong{0};
    auto one_pass = [&]() {
        out_rows = 0;
        std::vector<std::thread> threads;
        for (int d = 0; d < drivers; d++) {
            threads.emplace_back([&, d]() {
                tgpu_op* probe = nullptr;
                CHECK(ctxs[d], tgpu_join_probe_create(ctxs[d], &pspec, lookup, &probe));
                CHECK(ctxs[d], tgpu_join_probe_set_passthrough_by_reference(probe, 1));     // probe blocks stay on the "heap"
                std::vector<int64_t> payload((size_t)BATCH_ROWS + PAGE_ROWS);
                std::vector<uint8_t> valid((size_t)(BATCH_ROWS + PAGE_ROWS) / 8 + 8);
                long long mine = 0, bad = 0;
                drive(ctxs[d], probe, ms[d], probe_pages, (size_t)d, (size_t)drivers, [&](tgpu_page* dev, const std::vector<const JavaPage*>& batch) {
                    // output = probe key, probe price (views of the caller's blocks: not copied back), build payload
                    tgpu_column hc[3];
                    memset(hc, 0, sizeof(hc));
                    for (int c = 0; c < 3; c++) {
                        int32_t src = -1;
                        CHECK(ctxs[d], tgpu_page_passthrough_channel(dev, c, &src));
                        hc[c].type = dev->columns[c].type;
                        hc[c].length = dev->num_rows;
                        hc[c].validity = valid.data();
                        hc[c].data = src >= 0 ? nullptr : payload.data();
                        if ((c < 2) != (src >= 0)) bad++;        // the two probe channels pass through, the payload does not
                    }
                    tgpu_page host = {3, 0, dev->num_rows, hc};
                    CHECK(ctxs[d], tgpu_page_copy_to_host(ctxs[d], dev, &host));
                    // un-marshal: cut into <= 8192-row pages along the input page boundaries; check payload == key % 2557 against the heap keys
                    int64_t at = 0;
                    for (auto* p : batch) {
                        const int64_t* keys = (const int64_t*)p->cols[0].values.data();
                        for (int64_t i = 0; i < p->rows; i += 97) bad += payload[at + i] != keys[i] % 2557;
                        at += p->rows;
                    }
                    if (at != dev->num_rows) bad++;
                    mine += dev->num_rows;
                });
                tgpu_op_close(probe);
                out_rows += mine;
                wrong += bad;
            });
        }
        for (auto& t : threads) t.join();
    };
    one_pass();
    double t0 = now_s();
    one_pass();
    double secs = now_s() - t0;
    printf("{\"harness\": \"driver_loop\", \"mode\": \"join\", \"build_rows\": %lld, \"probe_rows\": %lld, \"drivers\": %d, \"java_page_rows\": %lld, \"batch_rows\": %lld, "
           "\"build_seconds\": %.6f, \"seconds\": %.6f, \"rows_per_s\": %.1f, \"output_rows\": %lld, \"mismatches\": %lld}\n",
           (long long)n_orders, (long long)rows, drivers, (long long)PAGE_ROWS, (long long)BATCH_ROWS, build_secs, secs, (double)rows / secs, out_rows.load(), wrong.load());
    tgpu_lookup_release(lookup);
    tgpu_op_close(builder);
    for (int d = 0; d < drivers; d++) { tgpu_host_free_pinned(ms[d].staging); tgpu_ctx_destroy(ctxs[d]); }
    return (wrong.load() || out_rows.load() != rows) ? 1 : 0;
}

}  // namespace

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: driver_loop q1|join <rows> [drivers]\n"); return 2; }
    std::string mode = argv[1];
    int64_t n = atoll(argv[2]);
    int drivers = argc > 3 ? atoi(argv[3]) : 4;
    if (mode == "q1") return run_q1(n, drivers);
    if (mode == "join") return run_join(n, drivers);
    fprintf(stderr, "unknown mode %s\n", mode.c_str());
    return 2;
}
