// filter.h -- device-side `where` evaluation (kernels_filter.cu): argument block shared with api.cu.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mqvs.h"

namespace mqvs {

constexpr int kFilterMaxOps = 64;
constexpr int kFilterMaxCols = 64;

struct FilterColumn {
    const void* values;       // int64 or float64 [n]; MQVS_COL_TOKENS: int64 token ids, all rows concatenated;
                              // MQVS_COL_TEXT: the documents' UTF-8 bytes, all rows concatenated
    const uint32_t* present;  // presence bits, or nullptr = every row has the field
    uint64_t n;
    int type;                 // MQVS_COL_I64 / MQVS_COL_F64 / MQVS_COL_TOKENS / MQVS_COL_TEXT
    const uint64_t* offsets;  // MQVS_COL_TOKENS / MQVS_COL_TEXT: [n + 1] CSR offsets into values
};

struct FilterArgs {
    const mqvs_filter_op* prog;  // device, postfix
    int n_ops;
    const FilterColumn* cols;    // device table indexed by column id
    const void* pool;            // device, 8-byte constants of the $in / $nin sets
    const uint32_t* live;        // optional live-row bitmap to AND in (device), or nullptr
    uint64_t n_rows, n_words;
    uint32_t* mask;              // out: ceil(n_rows / 32) words
    unsigned long long* allowed; // out: number of allowed rows (zeroed by the caller)
};

void launch_filter_eval(const FilterArgs& a, cudaStream_t st);

}  // namespace mqvs
