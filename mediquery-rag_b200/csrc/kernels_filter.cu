// kernels_filter.cu -- K5: Chroma `where` predicates evaluated on the device into the packed allow-bitmask the scan
// kernels consume.  Metadata columns live in HBM as typed arrays (int64 / float64; strings are
// dictionary codes assigned by the host, ordered like the strings) with one presence bit per row; the filter arrives as
// a postfix program of leaf comparisons and AND / OR.  One thread per row, one ballot per warp = one mask word.
// HBM-bound: 8 bytes per row per referenced column (+ 1 bit presence, + 1 bit live) -> 4 bytes per 32 rows out.
// `where_document` ($contains over the stored documents) is a leaf of the same program: the documents' UTF-8 bytes are a
// CSR text column and a warp searches its 32 rows one after another with coalesced byte loads (warp_text_contains).
#include "common.cuh"
#include "filter.h"
#include "launch.h"

namespace mqvs {
namespace {

constexpr int kFilterThreads = 256;

__device__ __forceinline__ bool leaf_i64(int op, long long v, long long c) {
    switch (op) {
        case MQVS_OP_EQ: return v == c;
        case MQVS_OP_NE: return v != c;
        case MQVS_OP_GT: return v > c;
        case MQVS_OP_GTE: return v >= c;
        case MQVS_OP_LT: return v < c;
        default: return v <= c;
    }
}
__device__ __forceinline__ bool leaf_f64(int op, double v, double c) {
    switch (op) {
        case MQVS_OP_EQ: return v == c;
        case MQVS_OP_NE: return v != c;
        case MQVS_OP_GT: return v > c;
        case MQVS_OP_GTE: return v >= c;
        case MQVS_OP_LT: return v < c;
        default: return v <= c;
    }
}

// where_document leaf: does row `row`'s document contain the needle?  The WARP searches one row at a time (lane l tests the
// start positions b + l, b + l + 32, ...: byte loads of neighbouring lanes coalesce) and takes its 32 rows in turn; the
// answer for lane l's own row comes back as bit l.  Every argument but the lane id is warp-uniform.
__device__ __forceinline__ uint32_t warp_text_contains(const FilterColumn& col, uint64_t warp_row0, uint64_t n_rows,
                                                       const uint8_t* __restrict__ needle, uint32_t m, int lane) {
    const uint8_t* __restrict__ text = reinterpret_cast<const uint8_t*>(col.values);
    const uint8_t n0 = __ldg(needle);
    uint32_t hits = 0;
    for (int rr = 0; rr < 32; ++rr) {
        const uint64_t row = warp_row0 + rr;
        if (row >= n_rows || row >= col.n) break;
        const uint64_t b = col.offsets[row], e = col.offsets[row + 1];
        if (e - b < m) continue;
        const uint64_t last = e - m;  // last start position
        for (uint64_t base = b; base <= last; base += 32) {
            const uint64_t pos = base + lane;
            bool found = pos <= last && text[pos] == n0;
            for (uint32_t j = 1; found && j < m; ++j) found = text[pos + j] == __ldg(needle + j);
            if (__any_sync(0xffffffffu, found)) {
                hits |= 1u << rr;
                break;
            }
        }
    }
    return hits;
}

__global__ void __launch_bounds__(kFilterThreads)
filter_eval_kernel(const FilterArgs a) {
    __shared__ mqvs_filter_op prog[kFilterMaxOps];
    for (int i = threadIdx.x; i < a.n_ops; i += kFilterThreads) prog[i] = a.prog[i];
    __syncthreads();
    const uint64_t r = (uint64_t)blockIdx.x * kFilterThreads + threadIdx.x;
    const bool in_range = r < a.n_rows;
    uint32_t stk = 0;  // bit stack, top = bit 0
    for (int i = 0; i < a.n_ops; ++i) {
        const mqvs_filter_op op = prog[i];
        if (op.op == MQVS_OP_AND || op.op == MQVS_OP_OR) {
            const uint32_t b2 = stk & 1u;
            stk >>= 1;
            const uint32_t b1 = stk & 1u;
            stk = (stk & ~1u) | (op.op == MQVS_OP_AND ? (b1 & b2) : (b1 | b2));
            continue;
        }
        bool hit = false;
        if (op.op == MQVS_OP_TEXT_CONTAINS || op.op == MQVS_OP_TEXT_NOT_CONTAINS) {  // warp-uniform branch (one program)
            const FilterColumn col = a.cols[op.col];
            const int lane = threadIdx.x & 31;
            const uint32_t hits = warp_text_contains(col, r - lane, a.n_rows, reinterpret_cast<const uint8_t*>(a.pool) + 8ull * op.set_off,
                                                     op.set_len, lane);
            const bool present = r < col.n && (col.present == nullptr || ((col.present[r >> 5] >> (r & 31)) & 1u));
            hit = in_range && ((present && ((hits >> lane) & 1u)) != (op.op == MQVS_OP_TEXT_NOT_CONTAINS));
        } else if (in_range) {
            const FilterColumn col = a.cols[op.col];
            // rows beyond the column's length (added since it was uploaded) count as "field missing"
            const bool present = r < col.n && (col.present == nullptr || ((col.present[r >> 5] >> (r & 31)) & 1u));
            const bool negated = op.op == MQVS_OP_NE || op.op == MQVS_OP_NIN || op.op == MQVS_OP_NOT_CONTAINS;
            if (col.type == MQVS_COL_TOKENS) {
                // token-set membership (comma-joined `tags`): a handful of token ids per row, CSR
                bool found = false;
                if (r < col.n) {
                    const long long* tok = reinterpret_cast<const long long*>(col.values);
                    for (uint64_t j = col.offsets[r], e = col.offsets[r + 1]; j < e; ++j) found |= tok[j] == (long long)op.ival;
                }
                hit = found != negated;
            } else if (!present) {
                hit = negated;  // a missing field matches only $ne / $nin
            } else if (op.op == MQVS_OP_IN || op.op == MQVS_OP_NIN) {
                bool found = false;
                if (col.type == MQVS_COL_I64) {
                    const long long v = reinterpret_cast<const long long*>(col.values)[r];
                    for (uint32_t j = 0; j < op.set_len; ++j) found |= v == reinterpret_cast<const long long*>(a.pool)[op.set_off + j];
                } else {
                    const double v = reinterpret_cast<const double*>(col.values)[r];
                    for (uint32_t j = 0; j < op.set_len; ++j) found |= v == reinterpret_cast<const double*>(a.pool)[op.set_off + j];
                }
                hit = found != negated;
            } else if (col.type == MQVS_COL_I64) {
                hit = leaf_i64(op.op, reinterpret_cast<const long long*>(col.values)[r], (long long)op.ival);
            } else {
                hit = leaf_f64(op.op, reinterpret_cast<const double*>(col.values)[r], op.fval);
            }
        }
        stk = (stk << 1) | (hit ? 1u : 0u);
    }
    const bool allowed = in_range && (a.n_ops == 0 || (stk & 1u));
    uint32_t word = __ballot_sync(0xffffffffu, allowed);
    if ((threadIdx.x & 31) == 0 && (r >> 5) < a.n_words) {
        if (a.live != nullptr) word &= a.live[r >> 5];
        a.mask[r >> 5] = word;
        if (word) atomicAdd(a.allowed, (unsigned long long)__popc(word));
    }
}

}  // namespace

void launch_filter_eval(const FilterArgs& a, cudaStream_t st) {
    const uint64_t rows_padded = a.n_words * 32ull;
    if (rows_padded == 0) return;
    const unsigned grid = (unsigned)((rows_padded + kFilterThreads - 1) / kFilterThreads);
    filter_eval_kernel<<<grid, kFilterThreads, 0, st>>>(a);
}

}  // namespace mqvs
