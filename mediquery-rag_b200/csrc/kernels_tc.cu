// kernels_tc.cu -- K2' (tensor-core scoring) + K3 (fused shortlist epilogue) + K4 (exact fp32 rescoring).
//
// For query batches where the contraction [B, 768] x [768, N] is real GEMM work, scoring runs on the 5th-gen
// tensor cores: tcgen05.mma (kind::f16 for bf16 rows with bf16-rounded queries, kind::tf32 for fp32 rows),
// operands staged by TMA into 128B-swizzled shared memory through a 4-stage mbarrier ring, fp32 accumulators in
// TMEM, double-buffered so the epilogue of tile t overlaps the MMAs of tile t+1.
//
// Orientation: QUERIES on the MMA M axis (TMEM lanes), CORPUS rows on N (TMEM columns).  After tcgen05.ld each
// epilogue thread owns one query and walks that query's scores for 256 corpus rows in registers: one compare per
// score against a per-query threshold held in a register.  Survivors are appended to a thread-private list
// in global memory ([cta][query][CAP] 64-bit keys); when a list passes CAP-32 entries the warp sorts it in
// registers (bitonic, CAP/32 keys per lane), keeps the best k' and raises the threshold, which is shared across
// CTAs through an atomicMax on tau[query] -- every CTA filters with the tightest bound any CTA has proven.
// The score matrix (B x N; 164 GB at 4096 x 10M) is never materialised.
//
// The tensor-core scores are approximate (bf16-rounded queries / tf32 operands), so they only SHORTLIST.
// K4 recomputes the k' best candidates per query in the canonical fp32 order (common.cuh) and the final top-k
// is taken from those exact scores.  Exactness is checked, not assumed: every discarded row had an approximate
// score <= tau~, hence an exact score <= tau~ + eps (eps = rigorous bound on |approx - exact|); if the k-th exact
// score does not exceed tau~ + eps the query is flagged and re-run on the exact SIMT scan (kernels_scan.cu).
//
// Roofline: HBM while B is small (corpus streamed once per batch), tensor pipe for large B:
// 2 * N * 768 * B flop per batch.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <vector>

#include "common.cuh"
#include "launch.h"
#include "tc.h"

namespace mqvs {
namespace tc {

// ------------------------------------------------------------------------------------------ geometry
constexpr int BM = 128;           // queries per tile  (TMEM lanes)
constexpr int BN = 256;           // corpus rows per tile (TMEM columns)
constexpr int kStages = 4;
constexpr int kRowBytes = 128;    // one swizzle-128B row of a K-chunk
constexpr int kABytes = BM * kRowBytes;
constexpr int kBBytes = BN * kRowBytes;
constexpr int kStageBytes = kABytes + kBBytes;  // 48 KB
constexpr int kThreads = 192;     // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, warps 2-5: epilogue
constexpr int kTmemCols = 512;    // two 128 x 256 fp32 accumulators
constexpr int kStageChunkBytes = 4 * 32 * 32 * 4;  // 4 epilogue warps x (32 scores x 32 lanes) staging for the hit path
constexpr size_t kSmemBytes = (size_t)kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/ + kStageChunkBytes;
constexpr int kMaxQueriesPerLaunch = 4096;
constexpr uint32_t kFirstSegmentTiles = 256;  // first filtered segment; its thresholds come from the cold pass below
// Corpus segments end at 16, 256, 4096, ... tiles: each is 16x everything before it.  What matters is hits per
// warp-chunk (32 queries x 32 rows = 1024 scores against thresholds learnt on S rows: ~1024 k'/S), not appends per
// list.  Measured alternatives on B200 (r01, 1M and 10M rows): growth 8 / 32 / 64, a two-segment schedule for a single
// query tile, and "second segment = one tile per CTA" were all equal or slower.
// (also measured and rejected: a cold segment of one tile on EVERY CTA + 32x growth for B <= 256 -- two segments up to
// 1.2M rows -- 0.61 -> 0.68 ms at config 2, B=64: the cold tiles and the higher hit rate of the big segment cost more
// than the launches saved.)
constexpr uint32_t kSegmentGrowth = 16;
__host__ inline uint64_t first_segment_end(uint32_t /*m_tiles*/, int /*grid*/) { return kFirstSegmentTiles; }
__host__ inline uint64_t next_segment_end(uint64_t end, uint32_t /*m_tiles*/) { return end * kSegmentGrowth; }
// a short remainder (< half of what is already scored) is not worth a launch + selection of its own
__host__ inline uint64_t clamp_segment_end(uint64_t end, uint32_t n_tiles) {
    return (end >= n_tiles || n_tiles - end < end / 2) ? n_tiles : end;
}

template <typename T>
struct Kind;
template <>
struct Kind<__nv_bfloat16> {  // kind::f16, bf16 x bf16 -> fp32, K = 16 per instruction
    static constexpr int kElemsPerChunk = 64, kFmt = 1;
    static constexpr bool kTf32 = false;
};
template <>
struct Kind<float> {  // kind::tf32, fp32 data read as tf32, K = 8 per instruction
    static constexpr int kElemsPerChunk = 32, kFmt = 2;
    static constexpr bool kTf32 = true;
};

// ------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void tcgen05_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if constexpr (kTf32) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
// max of 32 fp32 values held as raw bits: 16 three-input max instructions, depth 4
__device__ __forceinline__ float max32(const uint32_t (&v)[32], float (&a)[11]) {
#pragma unroll
    for (int i = 0; i < 10; ++i)
        a[i] = fmax3(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
    a[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
    const float b0 = fmax3(a[0], a[1], a[2]), b1 = fmax3(a[3], a[4], a[5]), b2 = fmax3(a[6], a[7], a[8]);
    const float b3 = fmaxf(a[9], a[10]);
    return fmaxf(fmax3(b0, b1, b2), b3);
}

// K-major, 128B-swizzled operand tile: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3FFFFu) >> 4);   // start address, 16-byte units
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                    // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                    // SWIZZLE_128B
    return d;
}

__device__ __forceinline__ uint32_t ord_u32(float s) {
    if (s == 0.0f) s = 0.0f;
    uint32_t u = __float_as_uint(s);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_to_float(uint32_t u) {
    u = (u & 0x80000000u) ? (u & 0x7fffffffu) : ~u;
    return __uint_as_float(u);
}
constexpr uint32_t kOrdNegInf = 0x007FFFFFu;  // ord_u32(-inf)

// ------------------------------------------------------------------------------------------ warp bitonic
// Sort 32*KPL keys held as element e = r*32 + lane, DESCENDING in e.
// Deliberately NOT fully unrolled: the 36-45 compare-exchange stages, straight-line, are 50-150 KB of code that runs
// once per call -- measured on B200, fetching it cold cost ~10 us per sort (instruction-cache misses), several times
// the arithmetic.  Stages that exchange between lanes share one loop body; only the register-to-register stages
// (partner = another register of the same lane) are unrolled, because their register indices must be constants.
template <int KPL>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&key)[KPL], int lane) {
#pragma unroll 1
    for (int size = 2; size <= 32 * KPL; size <<= 1) {
        if (size > 32) {
#pragma unroll
            for (int rs = KPL >> 1; rs >= 1; rs >>= 1) {
                if ((rs << 5) < size) {  // stride = 32 * rs <= size / 2
#pragma unroll
                    for (int r = 0; r < KPL; ++r) {
                        if ((r & rs) == 0) {
                            const bool desc = (((r << 5) | lane) & size) == 0;
                            const uint64_t a = key[r], b = key[r | rs];
                            const bool swap = desc ? (a < b) : (a > b);
                            key[r] = swap ? b : a;
                            key[r | rs] = swap ? a : b;
                        }
                    }
                }
            }
        }
#pragma unroll 1
        for (int stride = size > 32 ? 16 : (size >> 1); stride > 0; stride >>= 1) {
            const bool lower = (lane & stride) == 0;
#pragma unroll
            for (int r = 0; r < KPL; ++r) {
                const uint64_t other = __shfl_xor_sync(0xffffffffu, key[r], stride);
                const bool desc = (((r << 5) | lane) & size) == 0;
                const bool take_max = (lower == desc);
                const uint64_t mx = key[r] > other ? key[r] : other;
                const uint64_t mn = key[r] > other ? other : key[r];
                key[r] = take_max ? mx : mn;
            }
        }
    }
}

// One 128 x 256 accumulator tile, seen by the epilogue thread that owns TMEM lane (= query) `taddr >> 16`: threshold
// test of this query's 256 scores, survivors appended to the thread's private list, lists about to overflow compacted.
// The 8 largest chunk maxima (order-preserving bits; one chunk = 32 consecutive rows, masked rows excluded) this
// (CTA, query) has seen, unsorted, mirrored in global memory slot by slot: tc_fused_kernel's threshold refresh reads them.
struct ChunkTop {
    uint32_t v[8];
    uint32_t mn;     // min of v
    uint32_t* out;   // this thread's 8 slots of cmax
    bool on;         // record chunks of the tile being processed
};

template <int CAP, int KPRIME, bool TRACK = false>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, uint32_t tile_row0, uint32_t n_rows,
                                              const uint32_t* __restrict__ mask, uint32_t* st, uint32_t* tau_glob_q, int lane,
                                              float& tau, uint32_t& cnt, uint64_t* lst, ChunkTop* trk = nullptr) {
    constexpr int KPL = CAP / 32;
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
        const uint32_t row0 = tile_row0 + ch * 32;
        if (row0 >= n_rows) break;  // uniform
        // allow-mask word of this chunk (uniform over the warp).  Masked and out-of-range rows become -inf BEFORE the
        // fast-reject test: thresholds are learnt on allowed rows only, so at 10 % selectivity unmasked scores would send
        // ten times as many chunks through the hit path only to be dropped there (measured: +17 % on config 5); chunks
        // without any allowed row are not even read from TMEM.
        uint32_t mw = mask ? __ldg(mask + (row0 >> 5)) : 0xffffffffu;
        if (n_rows - row0 < 32) mw &= (1u << (n_rows - row0)) - 1u;
        if (mw == 0u) continue;
        uint32_t v[32];
        __syncwarp();
        tmem_ld32(taddr + ch * 32, v);
        if (mw != 0xffffffffu) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = (mw >> j) & 1u ? v[j] : 0xff800000u;
        }
        // fast reject: one 3-input max tree per 32 scores, one compare
        float grp[11];
        const float mx = max32(v, grp);
        const unsigned hit = __ballot_sync(0xffffffffu, mx >= tau);
        if (hit == 0u) continue;
        {
            // hit path (some lane of the warp has a survivor): branch-free pass mask per lane, the chunk
            // staged in shared memory so survivors can be fetched with a dynamic index -- cost is
            // ~100 straight-line instructions + a few per survivor, instead of 32 divergent tests
            uint32_t pm = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                st[j * 32 + lane] = v[j];
                pm |= __uint_as_float(v[j]) >= tau ? (1u << j) : 0u;
            }
            pm &= mw;  // tau may still be -inf
            if constexpr (TRACK) {
                if (pm != 0u && trk->on) {  // this chunk's best allowed row clears tau: is it among this thread's 8 best chunks?
                    const uint32_t o = ord_u32(mx);
                    if (o > trk->mn) {
                        bool done = false;
                        uint32_t nm = 0xffffffffu;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (!done && trk->v[i] == trk->mn) {  // replace ONE slot: a reader never sees a chunk twice
                                trk->v[i] = o;
                                trk->out[i] = o;
                                done = true;
                            }
                            nm = min(nm, trk->v[i]);
                        }
                        trk->mn = nm;
                    }
                }
            }
            __syncwarp();
            while (pm) {
                const int j = __ffs(pm) - 1;
                pm &= pm - 1;
                if (cnt < CAP) lst[cnt] = pack_key(__uint_as_float(st[j * 32 + lane]), row0 + j);
                ++cnt;
            }
            __syncwarp();
        }
        // lists that could overflow on the next chunk are compacted now (warp-cooperative)
        unsigned todo = __ballot_sync(0xffffffffu, cnt > CAP - 32);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            __syncwarp();
            const uint64_t* L = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, (unsigned long long)lst, src));
            const uint32_t c = min(__shfl_sync(0xffffffffu, cnt, src), (uint32_t)CAP);
            uint64_t key[KPL];
#pragma unroll
            for (int r = 0; r < KPL; ++r) key[r] = (uint32_t)(r * 32 + lane) < c ? __ldcg(L + r * 32 + lane) : 0ull;
            warp_sort_desc<KPL>(key, lane);
#pragma unroll
            for (int r = 0; r < KPL; ++r)
                if (r * 32 + lane < KPRIME) const_cast<uint64_t*>(L)[r * 32 + lane] = key[r];
            const uint64_t kth = __shfl_sync(0xffffffffu, key[(KPRIME - 1) >> 5], (KPRIME - 1) & 31);
            if (lane == src) {
                cnt = KPRIME;  // c > CAP - 32 >= KPRIME, so the list is full down to its k'-th entry
                const uint32_t t = (uint32_t)(kth >> 32);
                atomicMax(tau_glob_q, t);
                tau = fmaxf(tau, ord_to_float(t));
            }
            __syncwarp();
        }
    }
}

// Maximum of each 32-row chunk of one accumulator tile for this thread's query, masked and out-of-range rows excluded,
// as order-preserving bits (0 = the chunk has no allowed row).  Distinct chunks are distinct rows, so the k'-th largest
// chunk maximum of a set of tiles is a valid shortlist threshold for them: at least k' rows score that much.
__device__ __forceinline__ void tile_chunk_maxima(uint32_t taddr, uint32_t tile_row0, uint32_t n_rows,
                                                  const uint32_t* __restrict__ mask, uint32_t (&out)[BN / 32]) {
#pragma unroll
    for (int ch = 0; ch < BN / 32; ++ch) {
        const uint32_t row0 = tile_row0 + ch * 32;
        uint32_t ord = 0u;
        if (row0 < n_rows) {  // uniform
            uint32_t v[32];
            __syncwarp();
            tmem_ld32(taddr + ch * 32, v);
            uint32_t mw = mask ? __ldg(mask + (row0 >> 5)) : 0xffffffffu;
            if (n_rows - row0 < 32) mw &= (1u << (n_rows - row0)) - 1u;
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, (mw >> j) & 1u ? __uint_as_float(v[j]) : -INFINITY);
            ord = mw ? ord_u32(m) : 0u;
        }
        out[ch] = ord;
    }
}

// ------------------------------------------------------------------------------------------ main kernel
struct ScoreArgs {
    uint32_t n_rows;        // valid corpus rows
    uint32_t n_queries;     // valid queries (B); lanes beyond are padding
    uint32_t m_tiles;       // ceil(B / 128)
    uint32_t tile_begin;    // this launch scores corpus tiles [tile_begin, tile_end) (256 rows each)
    uint32_t tile_end;
    const uint32_t* mask;   // optional allow-mask over rows (device), or nullptr
    uint32_t* tau;          // [B_pad] shared per-query threshold, order-preserving uint, atomicMax
    uint32_t* cnt;          // [grid][B_pad] list lengths
    uint64_t* lists;        // [grid][B_pad][CAP]
    uint32_t b_pad;
    uint32_t* cmax;         // cold pass only (else nullptr): [B_pad][ncols] chunk maxima, no list appends
    uint32_t ncols;         // 8 * CTAs of the cold pass
};

template <typename T, int CAP, int KPRIME>
__global__ void __launch_bounds__(kThreads, 1)
tc_score_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, ScoreArgs a) {
    using K = Kind<T>;
    constexpr int kChunks768 = kDim / K::kElemsPerChunk;  // 12 (bf16) or 24 (tf32) K-chunks per tile
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)K::kFmt << 7) | ((uint32_t)K::kFmt << 10) |
                                ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + (size_t)kStages * kStageBytes);
    // bars[0..S): full, [S..2S): empty, [2S..2S+2): tmem_full, [2S+2..2S+4): tmem_empty
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    uint32_t* hit_stage = reinterpret_cast<uint32_t*>(smem_al + (size_t)kStages * kStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);  // one arrive per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const uint32_t first_tile = a.tile_begin + blockIdx.x;
    const uint32_t total_work = a.tile_end > first_tile ? ((a.tile_end - first_tile + gridDim.x - 1) / gridDim.x) * a.m_tiles : 0;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t nt = first_tile; nt < a.tile_end; nt += gridDim.x)
                for (uint32_t mt = 0; mt < a.m_tiles; ++mt)
                    for (int kc = 0; kc < kChunks768; ++kc) {
                        mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                        const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                        mbar_expect_tx(bar_full + 8 * stage, kStageBytes);
                        tma_load_2d(sa, &map_q, bar_full + 8 * stage, kc * K::kElemsPerChunk, (int)(mt * BM));
                        tma_load_2d(sb, &map_x, bar_full + 8 * stage, kc * K::kElemsPerChunk, (int)(nt * BN));
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t w = 0; w < total_work; ++w) {
                const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);  // epilogue has drained this accumulator
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kc = 0; kc < kChunks768; ++kc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                    const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)  // 4 x 32-byte K-slices inside the 128-byte swizzle row
                        tcgen05_mma<K::kTf32>(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, kIdesc, (kc | kk) != 0);
                    tcgen05_commit(bar_empty + 8 * stage);  // frees the smem slot when these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(bar_tfull + 8 * acc);  // accumulator complete -> epilogue
            }
        }
    } else {
        // ===================== epilogue: threshold + shortlist =====================
        const int quad = warp & 3;  // TMEM lanes 32*quad .. 32*quad+31 are this warp's
        uint32_t w = 0;
        for (uint32_t nt = first_tile; nt < a.tile_end; nt += gridDim.x)
            for (uint32_t mt = 0; mt < a.m_tiles; ++mt, ++w) {
                const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                const uint32_t q = mt * BM + quad * 32 + lane;
                const bool q_ok = q < a.n_queries;
                const size_t slot = (size_t)blockIdx.x * a.b_pad + q;
                uint64_t* lst = a.lists + slot * CAP;
                uint32_t cnt = q_ok ? a.cnt[slot] : 0;
                float tau = q_ok ? ord_to_float(__ldcg(a.tau + q)) : INFINITY;

                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tcgen05_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * BN;
                if (a.cmax != nullptr) {  // cold pass (uniform): only the chunk maxima that seed the thresholds
                    uint32_t cm[BN / 32];
                    tile_chunk_maxima(taddr, nt * BN, a.n_rows, a.mask, cm);
                    if (q_ok) {
                        uint4* dst = reinterpret_cast<uint4*>(a.cmax + (size_t)q * a.ncols + (size_t)(nt - a.tile_begin) * 8);
                        dst[0] = make_uint4(cm[0], cm[1], cm[2], cm[3]);
                        dst[1] = make_uint4(cm[4], cm[5], cm[6], cm[7]);
                    }
                } else
                epilogue_tile<CAP, KPRIME>(taddr, nt * BN, a.n_rows, a.mask, hit_stage + (warp - 2) * 1024, a.tau + q, lane, tau, cnt, lst);
                // accumulator fully read -> hand TMEM back to the MMA warp
                tcgen05_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
                if (q_ok) a.cnt[slot] = cnt;
            }
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ prep
// Normalised fp32 queries -> operand matrix [b_pad, 768] (bf16-rounded or fp32 copy), zero rows for padding,
// and eps[q], the rigorous bound on |tensor-core score - exact fp32 score| for unit-norm rows.
template <typename T>
__global__ void prep_queries_kernel(const float* __restrict__ qn, int B, int b_pad, T* __restrict__ q_op,
                                    float* __restrict__ eps, uint32_t* __restrict__ tau) {
    const int q = blockIdx.x, lane = threadIdx.x;  // one warp per query row
    float err2 = 0.f;
    for (int j = lane; j < kDim; j += 32) {
        float v = q < B ? qn[(size_t)q * kDim + j] : 0.f;
        if constexpr (sizeof(T) == 2) {
            __nv_bfloat16 h = __float2bfloat16_rn(v);
            q_op[(size_t)q * kDim + j] = h;
            float d = v - __bfloat162float(h);
            err2 += d * d;
        } else {
            q_op[(size_t)q * kDim + j] = v;
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, m);
    if (lane == 0) {
        // bf16 rows: stored rows are exact operands, only the query is rounded: |err| <= ||dq|| * ||x||, with
        // ||x|| <= 1.004 after bf16 rounding of a unit row.  tf32: both operands lose 13 mantissa bits (relative
        // error < 2^-10 each): |err| <= 2^-9 * sum|q_i x_i| <= 2^-9 * 1.001.  Both get 1e-4 for the tensor core's fp32
        // accumulation (768 additions, each off by at most 2^-23 of a partial sum bounded by ||q|| ||x||).
        float e = sizeof(T) == 2 ? sqrtf(err2) * 1.004f * 1.0001f + 1e-4f : 2.0f * 0.0009765625f * 1.001f + 1e-4f;
        if (q < B) eps[q] = e;
        tau[q] = kOrdNegInf;
    }
}

// ------------------------------------------------------------------------------------------ select + rescore
constexpr int kSelThreads = 256;

// General streaming selection, correct for any number of candidates: (previous shortlist) U (per-CTA lists of query q)
// -> the kp (= k') best keys, sorted descending, in buf[0, kp).  Streams the lists through a 2048-key buffer in batches
// that fit even if every entry survives the filter, with a block bitonic sort per batch.  This is the FALLBACK of
// select_lists_kernel (NT = 256) and of tc_fused_kernel's tail (NT = 192) when the candidates overflow their fast
// paths (adversarial orderings, thousands of ties).  buf: 2048 keys; pre_s: NT + 1; warp_tot: NT / 32.
template <int NT>
__device__ __forceinline__ void select_lists_body(const uint64_t* __restrict__ lists, const uint32_t* __restrict__ cnt,
                                                  int grid_ctas, int b_pad, int cap, int kp, const uint64_t* __restrict__ prev_keys,
                                                  int q, uint64_t* buf, int* pre_s, int* warp_tot, int* fill_sp) {
    constexpr int kBuf = 2048;
    int& fill_s = *fill_sp;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // list lengths -> block-wide exclusive scan (one list per thread; grid_ctas <= NT)
    const int mine = tid < grid_ctas ? min((int)__ldcg(cnt + (size_t)tid * b_pad + q), cap) : 0;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    const int n_sort0 = prev_keys ? kp : 0;
    for (int i = tid; i < kBuf; i += NT) buf[i] = i < n_sort0 ? prev_keys[(size_t)q * kp + i] : 0ull;
    if (tid == 0) fill_s = kp;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    pre_s[tid + 1] = base + incl;
    if (tid == 0) pre_s[0] = 0;
    __syncthreads();
    uint64_t floor_key = buf[kp - 1];  // sorted by the previous call; 0 = shortlist not full yet
    const int grand_total = pre_s[grid_ctas];
    int g0 = 0;
    if (floor_key == 0ull && grand_total > kBuf - kp) {
        // Cold call with far more candidates than the buffer holds (one whole tile from every CTA): sort the first
        // buffer-full to get a pivot (its k'-th key), then ONE filtered pass over everything else.  A pivot from 1792
        // keys lets ~k'/1792 of the rest through; should they overflow the buffer anyway, the batch loop below redoes
        // the pass safely (correct for any data, just slower).
        int g1 = 0;
        while (g1 < grid_ctas && pre_s[g1 + 1] <= kBuf - kp) ++g1;
        for (int idx = tid; idx < pre_s[g1]; idx += NT) {
            int lo = 0, hi = g1;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pre_s[mid] <= idx) lo = mid; else hi = mid;
            }
            buf[kp + idx] = __ldcg(lists + ((size_t)lo * b_pad + q) * cap + (idx - pre_s[lo]));
        }
        __syncthreads();
        block_bitonic_sort_desc(buf, kBuf);
        const uint64_t pivot = buf[kp - 1];
        __syncthreads();
        for (int i = kp + tid; i < kBuf; i += NT) buf[i] = 0ull;
        if (tid == 0) fill_s = kp;
        __syncthreads();
        const int first = pre_s[g1], total = grand_total - first;
        for (int idx = tid; idx < total; idx += NT) {
            int lo = g1, hi = grid_ctas;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pre_s[mid] <= first + idx) lo = mid; else hi = mid;
            }
            const uint64_t key = __ldcg(lists + ((size_t)lo * b_pad + q) * cap + (first + idx - pre_s[lo]));
            if (key > pivot) {
                const int slot = atomicAdd(&fill_s, 1);
                if (slot < kBuf) buf[slot] = key;
            }
        }
        __syncthreads();
        if (fill_s <= kBuf) {
            block_bitonic_sort_desc(buf, max(2 * kp, next_pow2(fill_s)));
            floor_key = buf[kp - 1];
            g0 = grid_ctas;  // done
        } else {  // overflow: keep the pivot batch's winners as the running best and redo the rest in safe batches
            __syncthreads();
            for (int i = kp + tid; i < kBuf; i += NT) buf[i] = 0ull;
            if (tid == 0) fill_s = kp;
            floor_key = pivot;
            g0 = g1;
        }
        __syncthreads();
    }
    while (g0 < grid_ctas) {
        // batch [g0, g1): as many lists as fit in the buffer even if every entry survives the filter
        const int room = kBuf - kp;
        int g1 = g0;
        while (g1 < grid_ctas && pre_s[g1 + 1] - pre_s[g0] <= room) ++g1;
        const int first = pre_s[g0], total = pre_s[g1] - first;
        for (int idx = tid; idx < total; idx += NT) {  // all entries of the batch, independent loads
            int lo = g0, hi = g1;                       // list g with pre_s[g] <= first + idx < pre_s[g + 1]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pre_s[mid] <= first + idx) lo = mid; else hi = mid;
            }
            const uint64_t key = __ldcg(lists + ((size_t)lo * b_pad + q) * cap + (first + idx - pre_s[lo]));
            if (key > floor_key) buf[atomicAdd(&fill_s, 1)] = key;
        }
        __syncthreads();
        const int n_sort = max(2 * kp, next_pow2(fill_s));  // untouched slots beyond fill_s are zero
        block_bitonic_sort_desc(buf, n_sort);
        floor_key = buf[kp - 1];
        __syncthreads();
        for (int i = kp + tid; i < n_sort; i += NT) buf[i] = 0ull;
        if (tid == 0) fill_s = kp;
        __syncthreads();
        g0 = g1;
    }
}

// One CTA per query, run after every corpus segment: (running shortlist of earlier segments) U (this segment's per-CTA
// lists) -> new running shortlist of the kp (= k') best approximate keys, and tau[q] raised to the k'-th best score of
// everything seen so far -- the threshold the next segment's epilogue filters with.
// Nobody needs the shortlist SORTED (the next call needs the set and its minimum, K4 rescoring the set), so this is a
// selection: all candidates that beat the previous k'-th key go to shared memory and an MSB-first radix select (8-bit
// digits, per-warp histograms, bytes every key shares skipped) finds the k'-th largest key in a few passes; the keys
// above it are compacted into slots [0, k'-1) and the k'-th key itself is stored LAST, where the next call looks for the
// floor.  At 4096 queries x k' = 256 this replaced bitonic sorts of 2048 keys: 650-820 us -> see DESIGN.md per call.
// More candidates than fit in shared memory (adversarial orderings) take the general streaming selection instead.
constexpr int kSelMaxKeys = 8192;  // 64 KB of keys

__global__ void __launch_bounds__(kSelThreads)
select_lists_kernel(const uint64_t* __restrict__ lists, const uint32_t* __restrict__ cnt, uint32_t* __restrict__ tau,
                    int grid_ctas, int b_pad, int cap, int kp, int first_segment, uint64_t* __restrict__ short_keys,
                    float* __restrict__ tau_tilde) {
    extern __shared__ __align__(16) uint64_t keys[];  // kSelMaxKeys
    __shared__ int pre_s[kSelThreads + 1];
    __shared__ int warp_tot[kSelThreads / 32];
    __shared__ int fill_s;
    __shared__ unsigned hist[kSelThreads / 32][256];
    __shared__ unsigned long long red_s[kSelThreads / 32];
    __shared__ unsigned sel_digit, sel_need, sel_count;
    __shared__ unsigned long long kstar_s;
    const int q = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // list lengths -> exclusive prefix sum (one list per thread)
    const int mine = tid < grid_ctas ? min((int)cnt[(size_t)tid * b_pad + q], cap) : 0;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    if (tid == 0) fill_s = 0;
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    pre_s[tid + 1] = base + incl;
    if (tid == 0) pre_s[0] = 0;
    __syncthreads();
    const int total = pre_s[grid_ctas];
    const int n_prev = first_segment ? 0 : kp;
    uint64_t kstar = 0ull;

    if (n_prev + total > kSelMaxKeys) {
        __syncthreads();
        select_lists_body<kSelThreads>(lists, cnt, grid_ctas, b_pad, cap, kp, first_segment ? nullptr : short_keys, q, keys,
                                       pre_s, warp_tot, &fill_s);
        for (int i = tid; i < kp; i += kSelThreads) short_keys[(size_t)q * kp + i] = keys[i];  // sorted: the k'-th is last
        kstar = keys[kp - 1];
    } else {
        // gather: previous shortlist + every candidate that beats its k'-th key (0 while the shortlist is not full)
        const uint64_t floor_key = first_segment ? 0ull : short_keys[(size_t)q * kp + kp - 1];
        for (int i = tid; i < n_prev; i += kSelThreads) {
            const uint64_t key = short_keys[(size_t)q * kp + i];
            if (key != 0ull) keys[atomicAdd(&fill_s, 1)] = key;
        }
        for (int idx = tid; idx < total; idx += kSelThreads) {
            int lo = 0, hi = grid_ctas;  // list g with pre_s[g] <= idx < pre_s[g + 1]
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (pre_s[mid] <= idx) lo = mid; else hi = mid;
            }
            const uint64_t key = __ldcg(lists + ((size_t)lo * b_pad + q) * cap + (idx - pre_s[lo]));
            if (key > floor_key) keys[atomicAdd(&fill_s, 1)] = key;
        }
        __syncthreads();
        const int n = fill_s;
        if (n < kp) {  // shortlist not full yet: keep everything, no floor
            for (int i = tid; i < kp; i += kSelThreads) short_keys[(size_t)q * kp + i] = i < n ? keys[i] : 0ull;
        } else {
            // bytes shared by every key carry no information: start below the highest differing bit
            uint64_t diff = 0ull;
            const uint64_t k0 = keys[0];
            for (int i = tid; i < n; i += kSelThreads) diff |= keys[i] ^ k0;
#pragma unroll
            for (int m = 16; m >= 1; m >>= 1) diff |= __shfl_xor_sync(0xffffffffu, diff, m);
            if (lane == 0) red_s[warp] = diff;
            __syncthreads();
            diff = 0ull;
#pragma unroll
            for (int w = 0; w < kSelThreads / 32; ++w) diff |= red_s[w];
            int shift = diff ? ((63 - __clzll((long long)diff)) & ~7) : 0;
            uint64_t prefix = (shift + 8 < 64) ? (k0 >> (shift + 8)) << (shift + 8) : 0ull;  // bits above the first digit
            unsigned need = (unsigned)kp;
            bool found = n == 1;
            if (n == 1) kstar = k0;
            while (!found) {
                for (int i = tid; i < (kSelThreads / 32) * 256; i += kSelThreads) (&hist[0][0])[i] = 0u;
                __syncthreads();
                for (int i = tid; i < n; i += kSelThreads) {
                    const uint64_t key = keys[i];
                    const bool in = shift + 8 >= 64 || ((key ^ prefix) >> (shift + 8)) == 0ull;
                    if (in) atomicAdd(&hist[warp][(unsigned)(key >> shift) & 255u], 1u);
                }
                __syncthreads();
                if (warp == 0) {  // digits in descending order: lane l owns 255 - 8l .. 248 - 8l
                    unsigned c[8], s = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int d = 255 - (8 * lane + j);
                        unsigned v = 0;
#pragma unroll
                        for (int w = 0; w < kSelThreads / 32; ++w) v += hist[w][d];
                        c[j] = v;
                        s += v;
                    }
                    unsigned inc = s;
#pragma unroll
                    for (int off = 1; off < 32; off <<= 1) {
                        const unsigned t = __shfl_up_sync(0xffffffffu, inc, off);
                        if (lane >= off) inc += t;
                    }
                    unsigned run = inc - s;
                    if (run < need && need <= inc) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (run < need && need <= run + c[j]) {
                                sel_digit = (unsigned)(255 - (8 * lane + j));
                                sel_need = need - run;
                                sel_count = c[j];
                            }
                            run += c[j];
                        }
                    }
                }
                __syncthreads();
                prefix |= (uint64_t)sel_digit << shift;
                need = sel_need;
                const unsigned count = sel_count;
                if (shift == 0) {
                    kstar = prefix;  // all 64 bits fixed (keys are unique, so count == 1 here)
                    found = true;
                } else if (count == 1) {  // one key left in the bin: it is the k'-th
                    for (int i = tid; i < n; i += kSelThreads)
                        if (((keys[i] ^ prefix) >> shift) == 0ull) kstar_s = keys[i];
                    __syncthreads();
                    kstar = kstar_s;
                    found = true;
                } else {
                    shift -= 8;
                }
                __syncthreads();
            }
            // keys above the k'-th -> slots [0, k'-1) in any order; the k'-th itself last
            if (tid == 0) fill_s = 0;
            __syncthreads();
            for (int i = tid; i < n; i += kSelThreads) {
                const uint64_t key = keys[i];
                if (key > kstar) short_keys[(size_t)q * kp + atomicAdd(&fill_s, 1)] = key;
            }
            if (tid == 0) short_keys[(size_t)q * kp + kp - 1] = kstar;
        }
    }
    if (tid == 0) {
        // every row discarded so far (epilogue filter, list compaction, this selection) scored <= the k'-th best
        uint32_t t = tau[q];
        if (kstar) t = max(t, (uint32_t)(kstar >> 32));
        tau[q] = t;
        tau_tilde[q] = ord_to_float(t);
    }
}

// After the cold pass (tc_score_kernel with ScoreArgs::cmax set): tau[q] = k'-th largest chunk maximum of query q.
__global__ void __launch_bounds__(128)
tau_from_chunk_maxima_kernel(const uint32_t* __restrict__ cmax, int ncols, int kp, uint32_t* __restrict__ tau) {
    __shared__ int s_cnt[8];
    const uint32_t* vals = cmax + (size_t)blockIdx.x * ncols;
    const int et = threadIdx.x;
    uint32_t v[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = i * 128 + et < ncols ? vals[i * 128 + et] : 0u;
    uint32_t lo = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = lo | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < 10; ++i) c += v[i] >= cand ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        int* sc = s_cnt + ((bit & 1) << 2);
        if ((et & 31) == 0) sc[et >> 5] = c;
        __syncthreads();
        if (sc[0] + sc[1] + sc[2] + sc[3] >= kp) lo = cand;
    }
    if (et == 0) tau[blockIdx.x] = lo > kOrdNegInf ? lo : kOrdNegInf;
}

// K4: one warp per (query, shortlisted row): canonical fp32 dot -> exact key.
template <typename T>
__global__ void __launch_bounds__(256)
rescore_kernel(const T* __restrict__ rows, const float* __restrict__ qn, const uint64_t* __restrict__ short_keys,
               int B, int kp, uint64_t* __restrict__ exact_keys) {
    const int lane = threadIdx.x & 31;
    const size_t item = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (item >= (size_t)B * kp) return;
    const int q = (int)(item / kp);
    const uint64_t key = short_keys[item];
    if (key == 0ull) {
        if (lane == 0) exact_keys[item] = 0ull;
        return;
    }
    const uint32_t id = key_id(key);
    RowFrag<T> x;
    x.load(rows + (size_t)id * kDim, lane);
    float4 qv[kChunks], xv[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        qv[c] = *reinterpret_cast<const float4*>(qn + (size_t)q * kDim + c * 128 + lane * 4);
        xv[c] = x.get(c);
    }
    const float s = warp_butterfly_sum(lane_dot(qv, xv));
    if (lane == 0) exact_keys[item] = pack_key(s, id);
}

// Per query: sort the kp exact keys, emit top-k, verify the exactness margin.
__global__ void __launch_bounds__(kSelThreads)
finalize_kernel(const uint64_t* __restrict__ exact_keys, const float* __restrict__ tau_tilde, const float* __restrict__ eps,
                int kp, int k, int64_t id_base, float* __restrict__ out_scores, int64_t* __restrict__ out_ids,
                uint32_t* __restrict__ flags, uint32_t* __restrict__ flag_count) {
    __shared__ uint64_t buf[256];
    const int q = blockIdx.x;
    for (int i = threadIdx.x; i < kp; i += kSelThreads) buf[i] = exact_keys[(size_t)q * kp + i];
    block_bitonic_sort_desc(buf, kp);
    for (int j = threadIdx.x; j < k; j += kSelThreads) {
        const uint64_t key = j < kp ? buf[j] : 0ull;
        out_scores[(size_t)q * k + j] = key ? key_score(key) : -INFINITY;
        out_ids[(size_t)q * k + j] = key ? (int64_t)key_id(key) + id_base : -1;
    }
    if (threadIdx.x == 0) {
        const float tt = tau_tilde[q];
        bool ok = true;
        if (tt != -INFINITY) {  // something was discarded: the k-th exact score must clear every discarded row's bound
            const uint64_t kth = (k <= kp) ? buf[k - 1] : 0ull;
            ok = kth != 0ull && key_score(kth) > tt + eps[q];
        }
        flags[q] = ok ? 0u : 1u;
        if (!ok) atomicAdd(flag_count, 1u);
    }
}

// ------------------------------------------------------------------------------------------ single-launch pipeline
struct FusedArgs;
template <typename T, int CAP, int KPRIME>
__device__ __forceinline__ void fused_tail(const FusedArgs& a, int G, uint8_t* smem_al);

// Batches of up to 128 queries (one query tile) are HBM-bound: the corpus is streamed once and the tensor pipe idles
// most of the time, so what the segmented pipeline above pays for -- two short cold segments and three selection
// launches, ~200 us next to a 425 us scan at 1M x 768 fp32 -- is pure latency.  tc_fused_kernel does the whole search
// in ONE cooperative launch of one persistent CTA per SM:
//   1. every CTA scores its first tile; the epilogue reads the accumulator once and publishes, per query, the maximum
//      of each 32-row chunk (8 per CTA, masked and out-of-range rows excluded);
//   2. grid barrier; CTA c takes query c and finds the k'-th largest of the G*8 chunk maxima.  The chunks are disjoint,
//      so at least k' allowed rows score >= that value: a valid threshold tau, within a few ranks of the exact k'-th
//      best of the first G*256 rows.  Second grid barrier publishes tau[0..B);
//   3. the first tile -- still in TMEM -- is read again against tau, then the CTA streams its remaining tiles exactly as
//      tc_score_kernel does (the TMA and MMA warps never stop: the second accumulator covers the barriers);
//   4. third grid barrier; CTA c merges the G per-CTA lists of query c, rescoring its k' best rows in the canonical fp32
//      order, sorts, writes the top-k and checks the exactness margin (select_lists + rescore + finalize above).
struct FusedArgs {
    uint32_t n_rows, n_queries, n_tiles, ncols;  // ncols = gridDim.x * 8 chunk maxima per query
    const uint32_t* mask;
    uint32_t* tau;         // [128]: -inf on entry (prep)
    uint32_t* cmax;        // [128][ncols] order-preserving chunk maxima, 0 = no row
    uint32_t* cnt;         // [grid][128]
    uint64_t* lists;       // [grid][128][CAP]
    unsigned int* gbar;    // [3] grid barrier counters, zero on entry (prep)
    const void* rows;      // corpus, for the exact rescoring
    const float* qn;       // [B][768] normalised fp32 queries
    const float* eps;      // [B]
    int k;
    long long id_base;
    float* out_scores;
    int64_t* out_ids;
    uint32_t* flags;
    uint32_t* flag_count;
    unsigned long long* dbg;  // MQVS_TRACE only: [grid][16] globaltimer stamps of the phases, else nullptr
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define MQVS_STAMP(i_) do { if (a.dbg) a.dbg[blockIdx.x * 16 + (i_)] = globaltimer_ns(); } while (0)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// One thread per CTA: publish this CTA's writes, arrive, wait for all `expected` CTAs (cooperative launch: co-resident).
__device__ __forceinline__ void grid_arrive_and_wait(unsigned int* ctr, unsigned int expected) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < expected) {
        __nanosleep(40);
        if (clock64() - t0 > 6000000000ll) __trap();  // seconds: a lost CTA must fail loudly, never hang the device
    }
    __threadfence();
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }  // the 4 epilogue warps

// k-th largest of n <= 1280 values (global, written by other CTAs before a grid barrier), by the 128 epilogue threads:
// binary search over the 32 key bits, one count + named barrier per bit.  Every thread returns the same value.
__device__ __forceinline__ uint32_t kth_largest_u32(const uint32_t* __restrict__ vals, int n, int kth, int et, int* s_cnt) {
    uint32_t v[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) v[i] = i * 128 + et < n ? __ldcg(vals + i * 128 + et) : 0u;
    uint32_t lo = 0;
#pragma unroll 1
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = lo | (1u << bit);
        int c = 0;
#pragma unroll
        for (int i = 0; i < 10; ++i) c += v[i] >= cand ? 1 : 0;
        c = __reduce_add_sync(0xffffffffu, c);
        int* sc = s_cnt + ((bit & 1) << 2);
        if ((et & 31) == 0) sc[et >> 5] = c;
        epi_bar_sync();
        if (sc[0] + sc[1] + sc[2] + sc[3] >= kth) lo = cand;
    }
    return lo;
}

template <typename T, int CAP, int KPRIME>
__global__ void __launch_bounds__(kThreads, 1)
tc_fused_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_x, const FusedArgs a) {
    using K = Kind<T>;
    constexpr int kChunks768 = kDim / K::kElemsPerChunk;
    constexpr uint32_t kIdesc = (1u << 4) | ((uint32_t)K::kFmt << 7) | ((uint32_t)K::kFmt << 10) |
                                ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + (size_t)kStages * kStageBytes);
    const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * kStages;
    const uint32_t bar_tfull = bar_empty + 8 * kStages, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
    int* s_cnt = reinterpret_cast<int*>(smem_al + (size_t)kStages * kStageBytes + 128);  // [2][4]
    uint32_t* hit_stage = reinterpret_cast<uint32_t*>(smem_al + (size_t)kStages * kStageBytes + 256);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t G = gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_q) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        for (int s = 0; s < kStages; ++s) {
            mbar_init(bar_full + 8 * s, 1);
            mbar_init(bar_empty + 8 * s, 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar_tfull + 8 * i, 1);
            mbar_init(bar_tempty + 8 * i, 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    const uint32_t first_tile = blockIdx.x;
    const uint32_t n_my = first_tile < a.n_tiles ? (a.n_tiles - first_tile + G - 1) / G : 0;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t nt = first_tile; nt < a.n_tiles; nt += G)
                for (int kc = 0; kc < kChunks768; ++kc) {
                    mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                    const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                    mbar_expect_tx(bar_full + 8 * stage, kStageBytes);
                    tma_load_2d(sa, &map_q, bar_full + 8 * stage, kc * K::kElemsPerChunk, 0);
                    tma_load_2d(sb, &map_x, bar_full + 8 * stage, kc * K::kElemsPerChunk, (int)(nt * BN));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t w = 0; w < n_my; ++w) {
                const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
                mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + acc * BN;
                for (int kc = 0; kc < kChunks768; ++kc) {
                    mbar_wait(bar_full + 8 * stage, phase);
                    tcgen05_fence_after();
                    const uint32_t sa = smem_base + stage * kStageBytes, sb = sa + kABytes;
                    const uint64_t adesc = smem_desc(sa), bdesc = smem_desc(sb);
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk)
                        tcgen05_mma<K::kTf32>(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, kIdesc, (kc | kk) != 0);
                    tcgen05_commit(bar_empty + 8 * stage);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                tcgen05_commit(bar_tfull + 8 * acc);
            }
        }
    } else {
        // ===================== epilogue =====================
        const int quad = warp & 3;
        const int et = (int)threadIdx.x - 64;  // 0..127
        const uint32_t q = quad * 32 + lane;
        const bool q_ok = q < a.n_queries;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(quad * 32) << 16);
        uint64_t* lst = a.lists + ((size_t)blockIdx.x * BM + q) * CAP;
        uint32_t cnt = 0;

        // ---- 1. chunk maxima of the first tile
        uint32_t* my_cmax = a.cmax + (size_t)(q_ok ? q : 0) * a.ncols + blockIdx.x * 8;
        ChunkTop trk;
        trk.out = my_cmax;
        trk.on = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) trk.v[i] = 0u;
        if (et == 0) MQVS_STAMP(0);
        if (n_my > 0) {
            mbar_wait(bar_tfull, 0);
            tcgen05_fence_after();
            if (et == 0) MQVS_STAMP(1);
            tile_chunk_maxima(lane_taddr, first_tile * BN, a.n_rows, a.mask, trk.v);
            if (q_ok) {
#pragma unroll
                for (int ch = 0; ch < BN / 32; ++ch) my_cmax[ch] = trk.v[ch];
            }
        } else if (q_ok) {
#pragma unroll
            for (int ch = 0; ch < BN / 32; ++ch) my_cmax[ch] = 0u;
        }
        __threadfence();
        epi_bar_sync();
        if (et == 0) { MQVS_STAMP(2); grid_arrive_and_wait(a.gbar + 0, G); MQVS_STAMP(3); }
        epi_bar_sync();

        // ---- 2. CTA c: tau[c] = k'-th largest chunk maximum of query c
        if (blockIdx.x < a.n_queries) {
            const uint32_t t = kth_largest_u32(a.cmax + (size_t)blockIdx.x * a.ncols, (int)a.ncols, KPRIME, et, s_cnt);
            if (et == 0) a.tau[blockIdx.x] = t > kOrdNegInf ? t : kOrdNegInf;
        }
        __threadfence();
        epi_bar_sync();
        if (et == 0) { MQVS_STAMP(4); grid_arrive_and_wait(a.gbar + 1, G); MQVS_STAMP(5); }
        epi_bar_sync();
        float tau = q_ok ? ord_to_float(__ldcg(a.tau + q)) : INFINITY;
        trk.mn = trk.v[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) trk.mn = min(trk.mn, trk.v[i]);

        // ---- 3. first tile again, now against tau, then the rest of this CTA's tiles
        uint32_t w = 0;
        for (uint32_t nt = first_tile; nt < a.n_tiles; nt += G, ++w) {
            const uint32_t acc = w & 1, acc_phase = (w >> 1) & 1;
            if (w > 0) {
                if (q_ok) tau = fmaxf(tau, ord_to_float(__ldcg(a.tau + q)));  // raised by list compactions anywhere
                mbar_wait(bar_tfull + 8 * acc, acc_phase);
                tcgen05_fence_after();
            }
            trk.on = w > 0 && q_ok;  // the first tile's chunks are already in trk
            epilogue_tile<CAP, KPRIME, true>(lane_taddr + acc * BN, nt * BN, a.n_rows, a.mask, hit_stage + (warp - 2) * 1024,
                                             a.tau + q, lane, tau, cnt, lst, &trk);
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * acc);
            // Threshold refresh after 2, 4, 8, ... tiles and after the last one: every (CTA, query) mirrors its 8 largest chunk maxima so far in
            // cmax, so the k'-th largest of a query's G*8 slots is a valid threshold that tightens as the scan proceeds
            // (a query then collects ~k' candidates per doubling of the rows scanned instead of k' N / (256 G)).  Slots are
            // replaced one word at a time and a chunk is only ever written to one slot, so whatever mix of old and new
            // values the reader sees, distinct slots are distinct chunks, i.e. distinct rows scoring at least that much.
            if (blockIdx.x < a.n_queries && w >= 1 && (((w + 1) & w) == 0 || w + 1 == n_my)) {
                const uint32_t t = kth_largest_u32(a.cmax + (size_t)blockIdx.x * a.ncols, (int)a.ncols, KPRIME, et, s_cnt);
                if (et == 0 && t > kOrdNegInf) atomicMax(a.tau + blockIdx.x, t);
            }
        }
        if (q_ok) a.cnt[(size_t)blockIdx.x * BM + q] = cnt;
    }

    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }

    // ---- 4. all lists are complete once every CTA is here; CTA c finishes query c
    __threadfence();
    if (threadIdx.x == 0) { MQVS_STAMP(6); grid_arrive_and_wait(a.gbar + 2, G); MQVS_STAMP(7); }
    __syncthreads();
    if (blockIdx.x >= a.n_queries) return;
    fused_tail<T, CAP, KPRIME>(a, (int)G, smem_al);
}

// ---- tail of tc_fused_kernel: CTA c turns the G per-CTA lists of query c into the final top-k.
// No sort of approximate keys: after the last threshold refresh tau[q] is within a few ranks of the k'-th best score, so
// the candidates that still clear it (~k', counted) ARE the shortlist.  All of them are rescored exactly, and the output
// order comes from ranking each exact key against the others (keys are unique: they embed the row id).  A sorting
// network on one warp costs 10+ us here (single-warp issue latency); this tail is ~6 us at k = 10.
template <typename T, int CAP, int KPRIME>
__device__ __forceinline__ void fused_tail(const FusedArgs& a, int G, uint8_t* smem_al) {
    constexpr int NW = kThreads / 32;
    constexpr int kCcap = 1024;  // shortlist capacity of the direct path (>= 2 k' with room for ties)
    // the operand ring is idle now (every TMA load was consumed by an MMA that has retired): reuse it
    uint64_t* C = reinterpret_cast<uint64_t*>(smem_al);   // 2048 keys (direct path uses kCcap; the fallback all of it)
    uint64_t* ex = C + 2048;                              // kCcap exact keys
    int* pre_s = reinterpret_cast<int*>(ex + kCcap);      // kThreads + 1
    int* warp_tot = pre_s + kThreads + 1;                 // NW
    int* fill_s = warp_tot + NW;
    int* valid_s = fill_s + 1;
    const int qf = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // list lengths -> exclusive prefix sum (one list per thread)
    const int mine = tid < G ? min((int)__ldcg(a.cnt + (size_t)tid * BM + qf), CAP) : 0;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int o = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += o;
    }
    if (lane == 31) warp_tot[warp] = incl;
    if (tid == 0) { *fill_s = 0; *valid_s = 0; }
    __syncthreads();
    int base = 0;
    for (int w = 0; w < warp; ++w) base += warp_tot[w];
    pre_s[tid + 1] = base + incl;
    if (tid == 0) pre_s[0] = 0;
    __syncthreads();
    const int total = pre_s[G];

    // every row that is not a candidate scored < tau[q] (filtered) or <= tau[q] (dropped by a list compaction, which
    // raised tau[q] to its k'-th key), so tau[q] bounds the approximate score of everything outside the shortlist
    uint32_t tq = __ldcg(a.tau + qf);
    for (int idx = tid; idx < total; idx += kThreads) {
        int lo = 0, hi = G;  // list g with pre_s[g] <= idx < pre_s[g + 1]
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (pre_s[mid] <= idx) lo = mid; else hi = mid;
        }
        const uint64_t key = __ldcg(a.lists + ((size_t)lo * BM + qf) * CAP + (idx - pre_s[lo]));
        if ((uint32_t)(key >> 32) >= tq) {
            const int slot = atomicAdd(fill_s, 1);
            if (slot < kCcap) C[slot] = key;
        }
    }
    __syncthreads();
    int nC = *fill_s;
    if (nC > kCcap) {
        // far more survivors than k' (hundreds of exact ties, adversarial orderings): the general streaming selection
        // picks the k' best approximate keys; its k'-th key then bounds everything it left out
        __syncthreads();
        select_lists_body<kThreads>(a.lists, a.cnt, G, BM, CAP, KPRIME, nullptr, qf, C, pre_s, warp_tot, fill_s);
        if (C[KPRIME - 1]) tq = max(tq, (uint32_t)(C[KPRIME - 1] >> 32));
        nC = KPRIME;
    }
    const float tt = ord_to_float(tq);
    if (tid == 0) MQVS_STAMP(8);

    // K4: canonical fp32 rescoring, one warp per row, four rows in flight per warp
    const T* rows = reinterpret_cast<const T*>(a.rows);
    float4 qv[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) qv[c] = *reinterpret_cast<const float4*>(a.qn + (size_t)qf * kDim + c * 128 + lane * 4);
    constexpr int kInFlight = 4;
    int n_valid = 0;
    for (int i0 = warp * kInFlight; i0 < nC; i0 += NW * kInFlight) {
        uint64_t key[kInFlight];
        RowFrag<T> x[kInFlight];
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            key[u] = i0 + u < nC ? C[i0 + u] : 0ull;
            if (key[u] != 0ull) x[u].load(rows + (size_t)key_id(key[u]) * kDim, lane);  // warp-uniform
        }
#pragma unroll
        for (int u = 0; u < kInFlight; ++u) {
            uint64_t out = 0ull;
            if (key[u] != 0ull) {
                float4 xv[kChunks];
#pragma unroll
                for (int c = 0; c < kChunks; ++c) xv[c] = x[u].get(c);
                out = pack_key(warp_butterfly_sum(lane_dot(qv, xv)), key_id(key[u]));
                ++n_valid;
            }
            if (lane == 0 && i0 + u < nC) ex[i0 + u] = out;
        }
    }
    if (lane == 0 && n_valid) atomicAdd(valid_s, n_valid);
    __syncthreads();
    if (tid == 0) MQVS_STAMP(9);

    // rank of each exact key among the others = its output position
    const int n_valid_all = *valid_s;
    for (int i = tid; i < nC; i += kThreads) {
        const uint64_t e = ex[i];
        if (e == 0ull) continue;
        int rank = 0;
        for (int j = 0; j < nC; ++j) rank += ex[j] > e ? 1 : 0;
        if (rank < a.k) {
            a.out_scores[(size_t)qf * a.k + rank] = key_score(e);
            a.out_ids[(size_t)qf * a.k + rank] = (int64_t)key_id(e) + a.id_base;
        }
        if (rank == a.k - 1) {  // the k-th exact score must clear the bound of everything outside the shortlist
            const bool ok = tt == -INFINITY || key_score(e) > tt + a.eps[qf];
            a.flags[qf] = ok ? 0u : 1u;
            if (!ok) atomicAdd(a.flag_count, 1u);
        }
    }
    for (int j = n_valid_all + tid; j < a.k; j += kThreads) {  // fewer candidates than k: pad
        a.out_scores[(size_t)qf * a.k + j] = -INFINITY;
        a.out_ids[(size_t)qf * a.k + j] = -1;
    }
    if (tid == 0 && n_valid_all < a.k) {
        const bool ok = tt == -INFINITY;  // nothing was ever discarded: the candidates are all the allowed rows
        a.flags[qf] = ok ? 0u : 1u;
        if (!ok) atomicAdd(a.flag_count, 1u);
    }
    if (tid == 0) MQVS_STAMP(10);
}

// Raw queries -> everything tc_fused_kernel needs, in one launch: K1's normalisation (bit-identical: same lane shares,
// same order), the fp32 copy the exact rescoring reads, the tensor-core operand matrix (bf16-rounded or fp32) padded to
// 128 rows, eps[q], tau = -inf, and the zeroed grid-barrier counters / flag count.
template <typename T>
__global__ void __launch_bounds__(128)
prep_fused_kernel(const float* __restrict__ q_raw, int dim, int vec4, int B, float* __restrict__ qn, T* __restrict__ q_op,
                  float* __restrict__ eps, uint32_t* __restrict__ tau, unsigned int* __restrict__ gbar,
                  uint32_t* __restrict__ flag_count, unsigned int* __restrict__ bad_epoch, unsigned int epoch) {
    const int lane = threadIdx.x & 31, q = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        gbar[0] = gbar[1] = gbar[2] = 0u;
        *flag_count = 0u;
    }
    float4 v[kChunks];
    if (q < B) {
        load_lane_share(q_raw + (size_t)q * dim, dim, vec4 != 0, lane, v);
        if (!normalize_lane_share(v) && lane == 0) atomicMax(bad_epoch, epoch);
    } else {
#pragma unroll
        for (int c = 0; c < kChunks; ++c) v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float err2 = 0.f;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
        const size_t o = (size_t)q * kDim + c * 128 + lane * 4;
        if (q < B) *reinterpret_cast<float4*>(qn + o) = v[c];
        if constexpr (sizeof(T) == 2) {
            const __nv_bfloat16 h0 = __float2bfloat16_rn(v[c].x), h1 = __float2bfloat16_rn(v[c].y);
            const __nv_bfloat16 h2 = __float2bfloat16_rn(v[c].z), h3 = __float2bfloat16_rn(v[c].w);
            const float d0 = v[c].x - __bfloat162float(h0), d1 = v[c].y - __bfloat162float(h1);
            const float d2 = v[c].z - __bfloat162float(h2), d3 = v[c].w - __bfloat162float(h3);
            err2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
            uint2 pk;
            pk.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            pk.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
            *reinterpret_cast<uint2*>(q_op + o) = pk;
        } else {
            *reinterpret_cast<float4*>(q_op + o) = v[c];
        }
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) err2 += __shfl_xor_sync(0xffffffffu, err2, m);
    if (lane == 0) {
        // same bounds as prep_queries_kernel
        const float e = sizeof(T) == 2 ? sqrtf(err2) * 1.004f * 1.0001f + 1e-4f : 2.0f * 0.0009765625f * 1.001f + 1e-4f;
        if (q < B) eps[q] = e;
        tau[q] = kOrdNegInf;
    }
}

// ------------------------------------------------------------------------------------------ host side
namespace {

PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

// [rows, 768] row-major matrix, box = one 128-byte K-chunk x box_rows rows, 128B swizzle.
bool make_map(CUtensorMap* map, const void* base, uint64_t rows, int dtype, int box_rows) {
    auto enc = get_encode();
    if (!enc || rows == 0) return false;
    const bool f32 = dtype == MQVS_F32;
    cuuint64_t dims[2] = {(cuuint64_t)kDim, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kDim * (f32 ? 4 : 2)};
    cuuint32_t box[2] = {(cuuint32_t)(f32 ? 32 : 64), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// MQVS_TRACE=1: CUDA events after every launch of a tensor-core search, printed by trace_dump() (developer aid)
struct Trace {
    bool on = getenv("MQVS_TRACE") != nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    unsigned long long* dbg = nullptr;  // [grid][8] phase stamps of the last tc_fused_kernel launch
    int dbg_grid = 0;
    unsigned long long* fused_dbg(int grid) {
        if (!dbg && cudaMalloc(&dbg, 256 * 16 * sizeof(unsigned long long)) != cudaSuccess) dbg = nullptr;
        dbg_grid = grid;
        return dbg;
    }
    void mark(const char* what, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, st);
        marks.emplace_back(what, e);
    }
} g_trace;

template <typename T, int CAP, int KPRIME>
cudaError_t launch_score(const CUtensorMap& mq, const CUtensorMap& mx, const ScoreArgs& a, int grid, cudaStream_t st) {
    auto kern = tc_score_kernel<T, CAP, KPRIME>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    kern<<<grid, kThreads, kSmemBytes, st>>>(mq, mx, a);
    return cudaGetLastError();
}

}  // namespace

Plan plan(int k, int B, int sm_count) {
    Plan p{};
    p.kprime = k <= 10 ? 32 : (k <= 32 ? 64 : (k <= 64 ? 128 : 256));
    p.cap = 512;  // the first segment appends a whole 256-row tile per (CTA, query) without compaction
    const int chunk = B < kMaxQueriesPerLaunch ? B : kMaxQueriesPerLaunch;
    p.b_pad = (chunk + BM - 1) / BM * BM;
    p.grid = sm_count;
    return p;
}

size_t op_bytes(const Plan& p, int dtype) { return (size_t)p.b_pad * kDim * (dtype == MQVS_F32 ? 4 : 2); }
size_t lists_bytes(const Plan& p) { return (size_t)p.grid * p.b_pad * p.cap * sizeof(uint64_t); }
size_t cnt_bytes(const Plan& p) { return (size_t)p.grid * p.b_pad * sizeof(uint32_t); }

int launches_per_search(const Plan& p, uint32_t n_rows, int B) {
    uint32_t tiles = (n_rows + BN - 1) / BN, segs = 0;
    const uint32_t m_tiles = (uint32_t)((std::min(B, kMaxQueriesPerLaunch) + BM - 1) / BM);
    for (uint64_t end = first_segment_end(m_tiles, p.grid); ; end = next_segment_end(end, m_tiles)) { ++segs; if (clamp_segment_end(end, tiles) >= tiles) break; }
    const int chunks = (B + kMaxQueriesPerLaunch - 1) / kMaxQueriesPerLaunch;
    return chunks * (int)(1 + 2 /* cold pass + thresholds */ + 2 * segs + 2);
}

const char* search(const Buffers& b, const Plan& p, const void* rows, int dtype, uint32_t n_rows, const float* qn_all, int B_all,
                   int k, const uint32_t* mask, int64_t id_base, float* out_scores_all, int64_t* out_ids_all, cudaStream_t st) {
    CUtensorMap mq, mx;
    if (!make_map(&mq, b.q_op, (uint64_t)p.b_pad, dtype, BM)) return "cuTensorMapEncodeTiled failed for the query matrix";
    if (!make_map(&mx, rows, (uint64_t)n_rows, dtype, BN)) return "cuTensorMapEncodeTiled failed for the corpus";
    const uint32_t n_tiles = (n_rows + BN - 1) / BN;
    if (cudaMemsetAsync(b.flag_count, 0, sizeof(uint32_t), st) != cudaSuccess) return "cudaMemsetAsync failed";

    g_trace.mark("begin", st);
    // queries are processed kMaxQueriesPerLaunch at a time (bounds the candidate-list memory)
    for (int q0 = 0; q0 < B_all; q0 += kMaxQueriesPerLaunch) {
        const int B = std::min(kMaxQueriesPerLaunch, B_all - q0);
        const int b_pad = (B + BM - 1) / BM * BM;
        const float* qn = qn_all + (size_t)q0 * kDim;
        if (dtype == MQVS_F32)
            prep_queries_kernel<float><<<p.b_pad, 32, 0, st>>>(qn, B, p.b_pad, (float*)b.q_op, b.eps + q0, b.tau);
        else
            prep_queries_kernel<__nv_bfloat16><<<p.b_pad, 32, 0, st>>>(qn, B, p.b_pad, (__nv_bfloat16*)b.q_op, b.eps + q0, b.tau);

        g_trace.mark("prep", st);
        ScoreArgs a{};
        a.n_rows = n_rows;
        a.n_queries = (uint32_t)B;
        a.m_tiles = (uint32_t)(b_pad / BM);
        a.mask = mask;
        a.tau = b.tau;
        a.cnt = b.cnt;
        a.lists = b.lists;
        a.b_pad = (uint32_t)p.b_pad;
        a.cmax = nullptr;
        a.ncols = 0;
        cudaError_t e = cudaSuccess;
#define MQVS_TC(KP_)                                                                          \
    if (p.kprime == KP_)                                                                      \
        e = dtype == MQVS_F32 ? launch_score<float, 512, KP_>(mq, mx, a, grid, st)            \
                              : launch_score<__nv_bfloat16, 512, KP_>(mq, mx, a, grid, st);
        // Cold pass: one tile per CTA, no list appends -- the epilogue only records the maximum of every 32-row chunk,
        // and the k'-th largest of a query's G*8 chunk maxima becomes its first threshold (>= k' distinct rows score that
        // much).  Costs one tile per CTA per query tile at tensor speed; the appending cold segment it replaces (4096 rows
        // on 16 CTAs, every score written out, then a selection over 4096 keys per query) took 2.0 ms at config 3.
        // The tiles of the cold pass are scored again by the first segment, now filtered.
        {
            const int grid = (int)std::min<uint32_t>((uint32_t)p.grid, n_tiles);
            a.tile_begin = 0;
            a.tile_end = (uint32_t)grid;
            a.cmax = b.cmax;
            a.ncols = (uint32_t)grid * 8;
            if (cudaMemsetAsync(b.cmax, 0, (size_t)p.b_pad * a.ncols * sizeof(uint32_t), st) != cudaSuccess) return "cudaMemsetAsync failed";
            MQVS_TC(32)
            MQVS_TC(64)
            MQVS_TC(128)
            MQVS_TC(256)
            if (e != cudaSuccess) return cudaGetErrorString(e);
            g_trace.mark("cold score", st);
            tau_from_chunk_maxima_kernel<<<B, 128, 0, st>>>(b.cmax, (int)a.ncols, p.kprime, b.tau);
            g_trace.mark("thresholds", st);
            a.cmax = nullptr;
            a.ncols = 0;
        }
        // Corpus segments of geometrically growing size.  After each one the per-query threshold becomes the k'-th
        // best score over EVERYTHING scored so far (select_lists_kernel), so a segment g times larger than all before
        // it appends only ~g k' candidates per query instead of re-learning its threshold in every CTA.
        static bool sel_attr = false;
        if (!sel_attr) {
            cudaFuncSetAttribute(select_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSelMaxKeys * (int)sizeof(uint64_t));
            sel_attr = true;
        }
        uint32_t begin = 0;
        for (uint64_t end = first_segment_end(a.m_tiles, p.grid); begin < n_tiles; end = next_segment_end(end, a.m_tiles)) {
            a.tile_begin = begin;
            a.tile_end = (uint32_t)clamp_segment_end(end, n_tiles);
            const int grid = (int)std::min<uint32_t>((uint32_t)p.grid, a.tile_end - a.tile_begin);
            if (cudaMemsetAsync(b.cnt, 0, (size_t)grid * p.b_pad * sizeof(uint32_t), st) != cudaSuccess) return "cudaMemsetAsync failed";
            MQVS_TC(32)
            MQVS_TC(64)
            MQVS_TC(128)
            MQVS_TC(256)
            if (e != cudaSuccess) return cudaGetErrorString(e);
            g_trace.mark("score", st);
            select_lists_kernel<<<B, kSelThreads, kSelMaxKeys * sizeof(uint64_t), st>>>(b.lists, b.cnt, b.tau, grid, p.b_pad, p.cap,
                                                                                         p.kprime, begin == 0, b.short_keys, b.tau_tilde);
            g_trace.mark("select", st);
            begin = a.tile_end;
        }
#undef MQVS_TC

        const size_t items = (size_t)B * p.kprime;
        const unsigned rs_grid = (unsigned)((items + 7) / 8);
        if (dtype == MQVS_F32)
            rescore_kernel<float><<<rs_grid, 256, 0, st>>>((const float*)rows, qn, b.short_keys, B, p.kprime, b.exact_keys);
        else
            rescore_kernel<__nv_bfloat16><<<rs_grid, 256, 0, st>>>((const __nv_bfloat16*)rows, qn, b.short_keys, B, p.kprime, b.exact_keys);
        finalize_kernel<<<B, kSelThreads, 0, st>>>(b.exact_keys, b.tau_tilde, b.eps + q0, p.kprime, k, id_base,
                                                   out_scores_all + (size_t)q0 * k, out_ids_all + (size_t)q0 * k,
                                                   b.flags + q0, b.flag_count);
        g_trace.mark("rescore+finalize", st);
    }
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

bool fused_ok(int B, int k, int sm_count) {
    // one query tile; one list / 8 chunk maxima per CTA must fit select_lists_body<192> and kth_largest_u32 (1280 values);
    // CTA c finishes query c, so there must be at least B CTAs
    return B >= 1 && B <= BM && k <= 128 && sm_count >= BM && sm_count <= 160;
}
size_t cmax_bytes(const Plan& p) { return (size_t)p.b_pad * p.grid * 8 * sizeof(uint32_t); }

namespace {
template <typename T, int KP>
cudaError_t launch_fused(const CUtensorMap& mq, const CUtensorMap& mx, const FusedArgs& a, int grid, cudaStream_t st) {
    auto kern = tc_fused_kernel<T, 512, KP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    if (e != cudaSuccess) return e;
    void* args[] = {(void*)&mq, (void*)&mx, (void*)&a};
    // cooperative: the grid barriers need every CTA resident (one per SM), whatever else shares the device
    return cudaLaunchCooperativeKernel((const void*)kern, dim3(grid), dim3(kThreads), args, kSmemBytes, st);
}
}  // namespace

const char* search_fused(const Buffers& b, const Plan& p, const void* rows, int dtype, uint32_t n_rows, const float* q_raw,
                         int dim, float* qn, int B, int k, const uint32_t* mask, int64_t id_base, float* out_scores,
                         int64_t* out_ids, unsigned int* bad_epoch, unsigned int epoch, cudaStream_t st, bool* coop_refused) {
    CUtensorMap mq, mx;
    *coop_refused = false;
    if (!make_map(&mq, b.q_op, (uint64_t)BM, dtype, BM)) return "cuTensorMapEncodeTiled failed for the query matrix";
    if (!make_map(&mx, rows, (uint64_t)n_rows, dtype, BN)) return "cuTensorMapEncodeTiled failed for the corpus";
    g_trace.mark("begin", st);
    const int vec4 = (dim % 4 == 0 && (reinterpret_cast<uintptr_t>(q_raw) & 15) == 0) ? 1 : 0;
    if (dtype == MQVS_F32)
        prep_fused_kernel<float><<<BM / 4, 128, 0, st>>>(q_raw, dim, vec4, B, qn, (float*)b.q_op, b.eps, b.tau, b.gbar,
                                                         b.flag_count, bad_epoch, epoch);
    else
        prep_fused_kernel<__nv_bfloat16><<<BM / 4, 128, 0, st>>>(q_raw, dim, vec4, B, qn, (__nv_bfloat16*)b.q_op, b.eps, b.tau,
                                                                 b.gbar, b.flag_count, bad_epoch, epoch);
    g_trace.mark("prep", st);
    FusedArgs a{};
    a.n_rows = n_rows;
    a.n_queries = (uint32_t)B;
    a.n_tiles = (n_rows + BN - 1) / BN;
    a.ncols = (uint32_t)p.grid * 8;
    a.mask = mask;
    a.tau = b.tau;
    a.cmax = b.cmax;
    a.cnt = b.cnt;
    a.lists = b.lists;
    a.gbar = b.gbar;
    a.rows = rows;
    a.qn = qn;
    a.eps = b.eps;
    a.k = k;
    a.id_base = id_base;
    a.out_scores = out_scores;
    a.out_ids = out_ids;
    a.flags = b.flags;
    a.flag_count = b.flag_count;
    a.dbg = g_trace.on ? g_trace.fused_dbg(p.grid) : nullptr;
    if (a.dbg) cudaMemsetAsync(a.dbg, 0, 256 * 16 * sizeof(unsigned long long), st);
    cudaError_t e = cudaSuccess;
#define MQVS_TCF(KP_)                                                                          \
    if (p.kprime == KP_)                                                                       \
        e = dtype == MQVS_F32 ? launch_fused<float, KP_>(mq, mx, a, p.grid, st)                \
                              : launch_fused<__nv_bfloat16, KP_>(mq, mx, a, p.grid, st);
    MQVS_TCF(32)
    MQVS_TCF(64)
    MQVS_TCF(128)
    MQVS_TCF(256)
#undef MQVS_TCF
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        cudaGetLastError();
        *coop_refused = true;
    }
    if (e != cudaSuccess) return cudaGetErrorString(e);
    g_trace.mark("fused", st);
    e = cudaGetLastError();
    return e == cudaSuccess ? nullptr : cudaGetErrorString(e);
}

void trace_dump() {
    if (!g_trace.on || g_trace.marks.empty()) return;
    cudaEventSynchronize(g_trace.marks.back().second);
    for (size_t i = 1; i < g_trace.marks.size(); ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, g_trace.marks[i - 1].second, g_trace.marks[i].second);
        fprintf(stderr, "[mqvs trace] %-18s %8.1f us\n", g_trace.marks[i].first, ms * 1e3f);
    }
    for (auto& m : g_trace.marks) cudaEventDestroy(m.second);
    g_trace.marks.clear();
    if (g_trace.dbg && g_trace.dbg_grid > 0) {
        // phase stamps of the last single-launch search, relative to the earliest CTA start: min / median / max over CTAs
        const int G = g_trace.dbg_grid;
        std::vector<unsigned long long> h((size_t)G * 16);
        if (cudaMemcpy(h.data(), g_trace.dbg, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost) == cudaSuccess) {
            unsigned long long t0 = ~0ull;
            for (int c = 0; c < G; ++c) t0 = std::min(t0, h[(size_t)c * 16]);
            static const char* names[11] = {"epilogue start", "first tile scored", "chunk maxima out", "barrier 1 passed",
                                            "tau selected", "barrier 2 passed", "scan done", "barrier 3 passed",
                                            "shortlist selected", "rescored", "results written"};
            for (int i = 0; i < 11; ++i) {
                std::vector<double> v;
                for (int c = 0; c < G; ++c)
                    if (h[(size_t)c * 16 + i] >= t0) v.push_back((double)(h[(size_t)c * 16 + i] - t0) * 1e-3);
                if (v.empty()) continue;
                std::sort(v.begin(), v.end());
                fprintf(stderr, "[mqvs trace]   %-18s min %8.1f  med %8.1f  max %8.1f us\n", names[i], v.front(), v[v.size() / 2], v.back());
            }
        }
        g_trace.dbg_grid = 0;
    }
}

}  // namespace tc
}  // namespace mqvs
