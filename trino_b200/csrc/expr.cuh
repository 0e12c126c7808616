// expr.cuh — device-side evaluator of tgpu_expr_program (the GPU stand-in for the bytecode that
// ExpressionCompiler.compilePageProcessor emits, M/sql/gen/ExpressionCompiler.java:50-85).
//
// Semantics reproduced:
//   - SQL three-valued logic, NULL-propagating arithmetic/comparison, Kleene AND/OR
//     (M/sql/gen/columnar/AndFilterEvaluator.java, OrFilterEvaluator.java; filters reject NULL:
//      M/sql/gen/columnar/ColumnarFilter.java:27-30)
//   - BIGINT arithmetic is checked (Math.addExact/subtractExact/multiplyExact/negateExact,
//     M/type/BigintOperators.java:52-110) -> NUMERIC_VALUE_OUT_OF_RANGE / DIVISION_BY_ZERO
//   - DOUBLE arithmetic is IEEE-754 binary64 with NO fused multiply-add (M/type/DoubleOperators.java:66-86):
//     every operation goes through __dadd_rn/__dmul_rn/__ddiv_rn which the compiler never contracts.
//
// Execution model: one thread evaluates one row; the <= TGPU_MAX_TEMPS temporaries of a row live in shared
// memory ([temp][thread], conflict-free 64-bit accesses) so the instruction stream can index them
// dynamically; null flags of the temporaries are one register bitmask.
#pragma once
#include "common.cuh"

namespace tg {

struct DOperand {
    int32_t kind;
    int32_t index;
    int64_t imm;
};

struct DInsn {
    int32_t op, vtype, dst, pad;
    DOperand a, b, c;
};

struct DProgram {
    int32_t num_insns;
    int32_t num_filter_insns;   // instructions [0, num_filter_insns) compute the filter
    int32_t filter_temp;        // -1: no filter
    int32_t num_in_lists;
    int32_t in_offset[8];
    int32_t in_count[8];
    int64_t in_values[128];
    DInsn insns[TGPU_MAX_INSNS];
};

struct DColumns {
    ColRef cols[TGPU_MAX_CHANNELS];
};

enum { TG_ERR_BIT_OVERFLOW = 1, TG_ERR_BIT_DIV_ZERO = 2 };

#if defined(__CUDACC__)

struct Value {
    int64_t bits;
    bool is_null;
};

__device__ __forceinline__ Value vm_fetch(const DOperand& o, const DColumns& cols, int64_t row, const int64_t* temps, int tstride, uint32_t nullbits)
{
    Value v;
    switch (o.kind) {
        case TGPU_OPND_COLUMN: {
            const ColRef& c = cols.cols[o.index];
            v.is_null = !tg_valid(c.validity, row);
            v.bits = tg_load_i64(c, row);
            break;
        }
        case TGPU_OPND_TEMP:
            v.bits = temps[o.index * tstride];
            v.is_null = (nullbits >> o.index) & 1;
            break;
        case TGPU_OPND_CONST:
            v.bits = o.imm;
            v.is_null = false;
            break;
        default:   // TGPU_OPND_NULL / NONE
            v.bits = 0;
            v.is_null = true;
            break;
    }
    return v;
}

__device__ __forceinline__ bool vm_cmp(int op, int vtype, int64_t a, int64_t b)
{
    if (vtype == TGPU_V_DOUBLE) {
        double x = __longlong_as_double(a), y = __longlong_as_double(b);
        switch (op) {
            case TGPU_EX_EQ: return x == y;
            case TGPU_EX_NE: return !(x == y);
            case TGPU_EX_LT: return x < y;
            case TGPU_EX_LE: return x <= y;
            case TGPU_EX_GT: return x > y;
            default: return x >= y;
        }
    }
    switch (op) {
        case TGPU_EX_EQ: return a == b;
        case TGPU_EX_NE: return a != b;
        case TGPU_EX_LT: return a < b;
        case TGPU_EX_LE: return a <= b;
        case TGPU_EX_GT: return a > b;
        default: return a >= b;
    }
}

// Runs instructions [first, last) for one row.  `temps` points at this thread's column of the shared
// [temp][thread] array (stride tstride).  Returns the updated null bitmask; *err accumulates TG_ERR_BIT_*.
__device__ __forceinline__ uint32_t vm_run(const DProgram* __restrict__ prog, int first, int last, const DColumns& cols, int64_t row,
                                           int64_t* temps, int tstride, uint32_t nullbits, uint32_t* err)
{
    for (int pc = first; pc < last; pc++) {
        const DInsn& in = prog->insns[pc];
        Value a = vm_fetch(in.a, cols, row, temps, tstride, nullbits);
        Value b = vm_fetch(in.b, cols, row, temps, tstride, nullbits);
        int64_t r = 0;
        bool rn = false;
        const int op = in.op;
        const bool dbl = in.vtype == TGPU_V_DOUBLE;
        switch (op) {
            case TGPU_EX_MOV: r = a.bits; rn = a.is_null; break;
            case TGPU_EX_ADD: case TGPU_EX_SUB: case TGPU_EX_MUL: case TGPU_EX_DIV: case TGPU_EX_MOD: {
                rn = a.is_null || b.is_null;
                if (rn) break;
                if (dbl) {
                    double x = __longlong_as_double(a.bits), y = __longlong_as_double(b.bits), z;
                    if (op == TGPU_EX_ADD) z = __dadd_rn(x, y);
                    else if (op == TGPU_EX_SUB) z = __dsub_rn(x, y);
                    else if (op == TGPU_EX_MUL) z = __dmul_rn(x, y);
                    else if (op == TGPU_EX_DIV) z = __ddiv_rn(x, y);
                    else z = fmod(x, y);
                    r = __double_as_longlong(z);
                }
                else {
                    long long x = a.bits, y = b.bits, z = 0;
                    if (op == TGPU_EX_ADD) {
                        z = (long long)((unsigned long long)x + (unsigned long long)y);
                        if (((x ^ z) & (y ^ z)) < 0) *err |= TG_ERR_BIT_OVERFLOW;
                    }
                    else if (op == TGPU_EX_SUB) {
                        z = (long long)((unsigned long long)x - (unsigned long long)y);
                        if (((x ^ y) & (x ^ z)) < 0) *err |= TG_ERR_BIT_OVERFLOW;
                    }
                    else if (op == TGPU_EX_MUL) {
                        z = (long long)((unsigned long long)x * (unsigned long long)y);
                        long long hi = __mul64hi(x, y);
                        if (hi != (z >> 63)) *err |= TG_ERR_BIT_OVERFLOW;
                    }
                    else {
                        if (y == 0) { *err |= TG_ERR_BIT_DIV_ZERO; }
                        else if (y == -1) {
                            if (op == TGPU_EX_DIV) {
                                if (x == LLONG_MIN) *err |= TG_ERR_BIT_OVERFLOW;
                                else z = -x;
                            }
                            else z = 0;
                        }
                        else z = op == TGPU_EX_DIV ? x / y : x % y;
                    }
                    r = z;
                }
                break;
            }
            case TGPU_EX_NEG:
                rn = a.is_null;
                if (rn) break;
                if (dbl) r = a.bits ^ (long long)0x8000000000000000ULL;
                else {
                    if (a.bits == LLONG_MIN) *err |= TG_ERR_BIT_OVERFLOW;
                    r = (long long)(0ULL - (unsigned long long)a.bits);
                }
                break;
            case TGPU_EX_EQ: case TGPU_EX_NE: case TGPU_EX_LT: case TGPU_EX_LE: case TGPU_EX_GT: case TGPU_EX_GE:
                rn = a.is_null || b.is_null;
                if (!rn) r = vm_cmp(op, in.vtype, a.bits, b.bits) ? 1 : 0;
                break;
            case TGPU_EX_AND: {
                bool af = !a.is_null && a.bits == 0, bf = !b.is_null && b.bits == 0;
                if (af || bf) { r = 0; rn = false; }
                else if (a.is_null || b.is_null) rn = true;
                else r = 1;
                break;
            }
            case TGPU_EX_OR: {
                bool at = !a.is_null && a.bits != 0, bt = !b.is_null && b.bits != 0;
                if (at || bt) { r = 1; rn = false; }
                else if (a.is_null || b.is_null) rn = true;
                else r = 0;
                break;
            }
            case TGPU_EX_NOT: rn = a.is_null; r = a.bits == 0 ? 1 : 0; break;
            case TGPU_EX_IS_NULL: r = a.is_null ? 1 : 0; break;
            case TGPU_EX_IS_NOT_NULL: r = a.is_null ? 0 : 1; break;
            case TGPU_EX_BETWEEN: {
                // value BETWEEN min AND max  ==  value >= min AND value <= max (Kleene AND)
                Value c = vm_fetch(in.c, cols, row, temps, tstride, nullbits);
                bool n1 = a.is_null || b.is_null, n2 = a.is_null || c.is_null;
                bool v1 = !n1 && vm_cmp(TGPU_EX_GE, in.vtype, a.bits, b.bits);
                bool v2 = !n2 && vm_cmp(TGPU_EX_LE, in.vtype, a.bits, c.bits);
                bool f1 = !n1 && !v1, f2 = !n2 && !v2;
                if (f1 || f2) r = 0;
                else if (n1 || n2) rn = true;
                else r = 1;
                break;
            }
            case TGPU_EX_CAST_BIGINT_TO_DOUBLE:
                rn = a.is_null;
                r = __double_as_longlong((double)a.bits);
                break;
            case TGPU_EX_CAST_DOUBLE_TO_BIGINT: {
                rn = a.is_null;
                if (rn) break;
                double x = __longlong_as_double(a.bits);
                // DoubleMath.roundToLong(x, HALF_UP): NaN / out of range is an error
                if (!(x >= -9.2233720368547758e18 && x < 9.2233720368547758e18)) *err |= TG_ERR_BIT_OVERFLOW;
                else r = llround(x);
                break;
            }
            case TGPU_EX_IN: {
                rn = a.is_null;
                if (rn) break;
                int li = (int)in.b.imm;
                int off = prog->in_offset[li], cnt = prog->in_count[li];
                bool hit = false;
                for (int k = 0; k < cnt; k++) {
                    int64_t c = prog->in_values[off + k];
                    hit |= dbl ? (__longlong_as_double(a.bits) == __longlong_as_double(c)) : (a.bits == c);
                }
                r = hit ? 1 : 0;
                break;
            }
            default: break;
        }
        temps[in.dst * tstride] = r;
        nullbits = (nullbits & ~(1u << in.dst)) | ((rn ? 1u : 0u) << in.dst);
    }
    return nullbits;
}

#endif  // __CUDACC__

// host side: validate + flatten a tgpu_expr_program into a DProgram (defined in expr.cu)
int expr_compile(tgpu_ctx* ctx, const tgpu_expr_program* program, DProgram* out, int32_t* max_channel);

}  // namespace tg
