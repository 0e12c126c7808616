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


#if defined(__CUDACC__)

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

// Runs instructions [first, last) for one row.  `temps` points at this thread's column of the shared
// [temp][thread] array (stride tstride).  Returns the updated null bitmask; *err accumulates TG_ERR_BIT_*.
__device__ __forceinline__ uint32_t vm_run(const DProgram* __restrict__ prog, int first, int last, const DColumns& cols, int64_t row,
                                           int64_t* temps, int tstride, uint32_t nullbits, uint32_t* err)
{
    for (int pc = first; pc < last; pc++) {
        const DInsn& in = prog->insns[pc];
        Value a = vm_fetch(in.a, cols, row, temps, tstride, nullbits);
        Value b = vm_fetch(in.b, cols, row, temps, tstride, nullbits);
        int64_t r;
        bool rn;
        if (in.op == TGPU_EX_IN) {
            rn = a.is_null;
            r = 0;
            if (!rn) {
                int li = (int)in.b.imm;
                int off = prog->in_offset[li], cnt = prog->in_count[li];
                bool hit = false;
                for (int k = 0; k < cnt; k++) {
                    int64_t c = prog->in_values[off + k];
                    hit |= in.vtype == TGPU_V_DOUBLE ? (__longlong_as_double(a.bits) == __longlong_as_double(c)) : (a.bits == c);
                }
                r = hit ? 1 : 0;
            }
        }
        else {
            Value c;
            c.bits = 0; c.is_null = true;
            if (in.op == TGPU_EX_BETWEEN) c = vm_fetch(in.c, cols, row, temps, tstride, nullbits);
            Value res = vm_apply(in.op, in.vtype, a, b, c, err);
            r = res.bits;
            rn = res.is_null;
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
