// kernels_scan.cu -- K2 (SIMT) + K3: bandwidth-bound exact fp32 scan with fused per-warp top-k.
//
// Does the arithmetic of vectorstore.similarity_search(query, k) (/root/reference/src/agents/nodes.py:93,
// src/medical_engine.py:70) for small query batches: every stored row is streamed from HBM exactly once per
// pass (128-bit / 64-bit coalesced ld.global.nc, L1 no-allocate), scored against QB register-resident
// queries in the canonical fp32 order (common.cuh), and offered to a warp-private top-k list that lives in
// registers.  Roofline: HBM; algorithmic bytes per pass = n_rows * 768 * sizeof(elem).
//
// Top-k (K3): per (warp, query) an unsorted list of 64-bit keys spread over the lanes (slot j on lane j%32),
// the smallest key cached as a score threshold.  A row is inserted only when it beats the threshold, which
// after warm-up happens with probability ~ k / rows_seen.  At the end the warps' lists are sorted together in
// shared memory and the CTA writes its best k keys, sorted, to cand[q][cta][k].
// The optional allow-mask (Chroma `where`, compiled on the host) is tested BEFORE the row is loaded, so
// filtered-out rows cost no HBM traffic.
#include "common.cuh"
#include "launch.h"

namespace mqvs {

namespace {

constexpr int kScanWarps = 8;
constexpr int kScanThreads = kScanWarps * 32;

template <int QB, int KPL>
struct TopK {
    uint64_t keys[QB][KPL];
    uint64_t minkey[QB];
    float thr[QB];

    __device__ __forceinline__ void init(int k, int lane) {
#pragma unroll
        for (int b = 0; b < QB; ++b) {
#pragma unroll
            for (int j = 0; j < KPL; ++j) keys[b][j] = (j * 32 + lane) < k ? 0ull : ~0ull;
            minkey[b] = 0ull;
            thr[b] = -INFINITY;
        }
    }
    // warp-uniform: replace the smallest key of list b with `key` (caller checked key > minkey[b])
    __device__ __forceinline__ void insert(int b, uint64_t key, int lane) {
        bool done = false;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            unsigned hit = __ballot_sync(0xffffffffu, !done && keys[b][j] == minkey[b]);
            if (hit) {
                if (lane == __ffs(hit) - 1) keys[b][j] = key;
                done = true;
            }
        }
        uint64_t m = keys[b][0];
#pragma unroll
        for (int j = 1; j < KPL; ++j) m = keys[b][j] < m ? keys[b][j] : m;
        m = warp_min_u64(m);
        minkey[b] = m;
        thr[b] = m ? key_score(m) : -INFINITY;
    }
};

template <typename T, int QB, int KPL, int U>
__device__ __forceinline__ void scan_body(const T* __restrict__ rows, uint32_t n_rows, const float* __restrict__ qn, int nq,
                                          const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ cand,
                                          int cand_stride, uint64_t* smem_keys) {
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // queries -> registers (lane's canonical share); missing queries are zero and never written out
    float4 q[QB][kChunks];
#pragma unroll
    for (int b = 0; b < QB; ++b)
#pragma unroll
        for (int c = 0; c < kChunks; ++c)
            q[b][c] = b < nq ? *reinterpret_cast<const float4*>(qn + (size_t)b * kDim + c * 128 + lane * 4)
                             : make_float4(0.f, 0.f, 0.f, 0.f);

    TopK<QB, KPL> top;
    top.init(k, lane);

    const uint32_t warps_total = gridDim.x * kScanWarps;
    const uint32_t gw = blockIdx.x * kScanWarps + warp;
    // each warp walks groups of U consecutive rows, ascending
    for (uint64_t r0 = (uint64_t)gw * U; r0 < n_rows; r0 += (uint64_t)warps_total * U) {
        RowFrag<T> x[U];
        bool live[U];
        uint32_t mw = 0xffffffffu;
        if (mask != nullptr) mw = __ldg(mask + (r0 >> 5)) >> (r0 & 31);  // U divides 32: one word covers the group
#pragma unroll
        for (int u = 0; u < U; ++u) {
            live[u] = (r0 + u < n_rows) && ((mw >> u) & 1u);
            if (live[u]) x[u].load(rows + (r0 + u) * (uint64_t)kDim, lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (!live[u]) continue;
            float4 xv[kChunks];
#pragma unroll
            for (int c = 0; c < kChunks; ++c) xv[c] = x[u].get(c);
#pragma unroll
            for (int b = 0; b < QB; ++b) {
                float s = warp_butterfly_sum(lane_dot(q[b], xv));
                if (s >= top.thr[b]) {
                    uint64_t key = pack_key(s, (uint32_t)(r0 + u));
                    if (key > top.minkey[b]) top.insert(b, key, lane);
                }
            }
        }
    }

    // CTA-level reduction: all warps' lists of query b -> sort -> best k
    const int sort_n = next_pow2(kScanWarps * k);
    for (int i = threadIdx.x; i < QB * sort_n; i += kScanThreads) smem_keys[i] = 0ull;
    __syncthreads();
#pragma unroll
    for (int b = 0; b < QB; ++b)
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
            int slot = j * 32 + lane;
            if (slot < k) smem_keys[b * sort_n + warp * k + slot] = top.keys[b][j];
        }
    __syncthreads();
    for (int b = 0; b < nq; ++b) {
        uint64_t* s = smem_keys + b * sort_n;
        block_bitonic_sort_desc(s, sort_n);
        uint64_t* out = cand + ((size_t)b * cand_stride + blockIdx.x) * k;
        for (int j = threadIdx.x; j < k; j += kScanThreads) out[j] = s[j];
    }
}

template <typename T, int QB, int KPL, int U>
__global__ void __launch_bounds__(kScanThreads)
scan_simt_kernel(const T* __restrict__ rows, uint32_t n_rows, const float* __restrict__ qn, int nq,
                 const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ cand, int cand_stride) {
    extern __shared__ __align__(16) uint64_t smem_keys[];  // [QB][sort_n]
    scan_body<T, QB, KPL, U>(rows, n_rows, qn, nq, mask, k, cand, cand_stride, smem_keys);
}

// Device-driven repair pass of the tensor-core path: re-scans the corpus exactly for the (few) queries whose
// shortlist could not be proven exact.  The list lives in device memory, so no host round trip is needed; with an
// empty list every CTA returns at once.
template <typename T, int KPL, int U>
__global__ void __launch_bounds__(kScanThreads)
scan_fix_kernel(const T* __restrict__ rows, uint32_t n_rows, const float* __restrict__ qn_all,
                const uint32_t* __restrict__ fix_list, const uint32_t* __restrict__ fix_n,
                const uint32_t* __restrict__ mask, int k, uint64_t* __restrict__ cand, int cand_stride) {
    extern __shared__ __align__(16) uint64_t smem_keys[];
    const uint32_t n = *fix_n;
    for (uint32_t f = 0; f < n; ++f) {
        scan_body<T, 1, KPL, U>(rows, n_rows, qn_all + (size_t)fix_list[f] * kDim, 1, mask, k,
                                cand + (size_t)f * cand_stride * k, cand_stride, smem_keys);
        __syncthreads();
    }
}

struct ScanCfg {
    int qb, kpl;
};

ScanCfg pick_cfg(int nq, int k) {
    ScanCfg c;
    c.kpl = k <= 32 ? 1 : (k <= 128 ? 4 : 32);
    int qmax = c.kpl == 1 ? 4 : (c.kpl == 4 ? 2 : 1);
    c.qb = nq >= 4 && qmax >= 4 ? 4 : (nq >= 2 && qmax >= 2 ? 2 : 1);
    return c;
}

using ScanKernel = void (*)(const void*, uint32_t, const float*, int, const uint32_t*, int, uint64_t*, int);

template <typename T, int QB, int KPL, int U>
const void* kernel_ptr() {
    return reinterpret_cast<const void*>(&scan_simt_kernel<T, QB, KPL, U>);
}

const void* find_kernel(int dtype, const ScanCfg& c) {
#define MQVS_CASE(QB_, KPL_)                                                                    \
    if (c.qb == QB_ && c.kpl == KPL_)                                                           \
        return dtype == MQVS_F32 ? kernel_ptr<float, QB_, KPL_, 2>() : kernel_ptr<__nv_bfloat16, QB_, KPL_, 4>();
    MQVS_CASE(1, 1)
    MQVS_CASE(2, 1)
    MQVS_CASE(4, 1)
    MQVS_CASE(1, 4)
    MQVS_CASE(2, 4)
    MQVS_CASE(1, 32)
#undef MQVS_CASE
    return nullptr;
}

size_t scan_smem(const ScanCfg& c, int k) { return (size_t)c.qb * next_pow2(kScanWarps * k) * sizeof(uint64_t); }

constexpr int kMaxCtasPerSm = 3;

}  // namespace

int scan_fix_grid(int sm_count) { return sm_count * 2; }

void launch_scan_fix(const void* rows, int dtype, uint32_t n_rows, const float* qn_all, const uint32_t* fix_list,
                     const uint32_t* fix_n, const uint32_t* mask, int k, uint64_t* cand, int sm_count, cudaStream_t st) {
    const int kpl = k <= 32 ? 1 : 4;  // the tensor path serves k <= 128
    const size_t smem = (size_t)next_pow2(kScanWarps * k) * sizeof(uint64_t);
    const int grid = scan_fix_grid(sm_count), stride = grid;  // every slot of a listed query is written: no memset
#define MQVS_FIX(T_, KPL_, U_)                                                                                   \
    scan_fix_kernel<T_, KPL_, U_><<<grid, kScanThreads, smem, st>>>((const T_*)rows, n_rows, qn_all, fix_list, fix_n, \
                                                                     mask, k, cand, stride)
    if (dtype == MQVS_F32) {
        if (kpl == 1) MQVS_FIX(float, 1, 2); else MQVS_FIX(float, 4, 2);
    } else {
        if (kpl == 1) MQVS_FIX(__nv_bfloat16, 1, 4); else MQVS_FIX(__nv_bfloat16, 4, 4);
    }
#undef MQVS_FIX
}

void scan_plan(int dtype, int nq_remaining, int k, int sm_count, int* nq_pass, int* grid) {
    ScanCfg c = pick_cfg(nq_remaining, k);
    const void* kern = find_kernel(dtype, c);
    size_t smem = scan_smem(c, k);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int occ = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kScanThreads, smem) != cudaSuccess || occ < 1) {
        cudaGetLastError();
        occ = 1;
    }
    *nq_pass = nq_remaining < c.qb ? nq_remaining : c.qb;
    *grid = sm_count * (occ < kMaxCtasPerSm ? occ : kMaxCtasPerSm);
}

int scan_max_grid(int sm_count) { return sm_count * kMaxCtasPerSm; }

void launch_scan_simt(const void* rows, int dtype, uint32_t n_rows, const float* qn, int nq, const uint32_t* mask,
                      int k, uint64_t* cand, int cand_stride, int grid, cudaStream_t st) {
    ScanCfg c = pick_cfg(nq, k);
    const void* kern = find_kernel(dtype, c);
    void* args[] = {(void*)&rows, (void*)&n_rows, (void*)&qn, (void*)&nq, (void*)&mask, (void*)&k, (void*)&cand,
                    (void*)&cand_stride};
    cudaLaunchKernel(kern, dim3(grid), dim3(kScanThreads), args, scan_smem(c, k), st);
}

}  // namespace mqvs
