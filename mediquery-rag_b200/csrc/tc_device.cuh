// tc_device.cuh -- device code shared by the two tensor-core pipelines (kernels_tc.cu: segmented, large batches;
// kernels_tc_fused.cu: single cooperative launch, B <= 256): tile geometry, tcgen05 / TMA / mbarrier PTX wrappers,
// the per-tile shortlist epilogue, chunk maxima, the warp bitonic sort and the general streaming selection; plus the
// small host helpers both launchers use (tensor maps, the MQVS_TRACE event trace).
#pragma once

#include <cuda.h>
#include <cudaTypedefs.h>

#include <utility>
#include <vector>

#include "common.cuh"
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
// The same wait with a ceiling: a CTA pair whose partner never arrives must fail loudly (a trap the host sees as an error)
// instead of hanging the device -- seconds, where a healthy wait takes microseconds.
__device__ __forceinline__ void mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    long long t0 = 0;
    for (uint32_t spins = 0; !done; ++spins) {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (!done && (spins & 0xfffu) == 0xfffu) {
            if (t0 == 0) t0 = clock64();
            else if (clock64() - t0 > 8000000000ll) __trap();
        }
    }
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
// ---- CTA pairs (cta_group::2): two CTAs of one cluster, on the two SMs of a TPC, issue ONE 256-row MMA together.  Each CTA
// stages its own half of both operands (its 128 queries, its 128 of the tile's 256 corpus rows), the leader (cluster rank 0)
// issues the MMAs, each CTA's TMEM receives its own 128 accumulator rows (all 256 columns).
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load of this CTA's half of a pair operand; the bytes are counted on the LEADER's barrier (`bar_cluster`: its
// shared::cluster address), where the MMA issuer waits
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once the pair's MMAs issued so far have retired
__device__ __forceinline__ void tcgen05_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
template <bool kTf32>
__device__ __forceinline__ void tcgen05_mma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
    if constexpr (kTf32) {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
            : "memory");
    } else {
        asm volatile(
            "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
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
// l2 / ip spaces over raw vectors (common.cuh:rank_score): what an epilogue thread needs to turn the approximate cosines
// of a chunk into approximate ranking scores t = a s - a^2 h (ip: h = 0).  The exact scores come from the rescoring,
// which uses rank_score itself; here only |t~ - t| <= a_max eps matters.
struct TileSpace {
    const float* norms = nullptr;   // [n_rows] ||x|| of the raw rows; nullptr = cosine (scores used as they are)
    const float2* bounds = nullptr; // [ceil(n_rows / 32)] (largest, smallest) norm of every 32-row chunk (norm_bounds_kernel)
    float h = 0.0f;                 // this thread's query: 0.5 / max(||q||, eps) for l2, 0 for ip
};
// Upper bound of every ranking score t = a s - a^2 h of a chunk whose largest cosine is `smax` (this thread's query) and
// whose norms lie in [b.y, b.x]: a s <= b.x smax for smax >= 0, <= b.y smax below; a^2 h >= b.y^2 h.  Rounded UP with slack
// for the fp32 roundings of both sides, so a chunk rejected by it holds no score the exact transform would have kept.
__device__ __forceinline__ float chunk_score_bound(float smax, float2 b, float h) {
    if (!(smax > -INFINITY)) return -INFINITY;  // nothing allowed in the chunk
    const float lin = smax >= 0.0f ? __fmul_ru(b.x, smax) : __fmul_ru(b.y, smax);
    const float quad = __fmul_rd(__fmul_rd(b.y, b.y), h);
    return __fadd_ru(__fsub_ru(lin, quad), 1e-6f * (fabsf(lin) + quad));
}
__device__ __forceinline__ void apply_space(uint32_t (&v)[32], const TileSpace& sp, uint32_t row0, uint32_t n_rows, int lane) {
    const float a_l = __ldg(sp.norms + min(row0 + (uint32_t)lane, n_rows - 1u));  // one coalesced load per chunk and warp
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float a = __shfl_sync(0xffffffffu, a_l, j);
        const float t = __fmaf_rn(a, __uint_as_float(v[j]), -__fmul_rn(__fmul_rn(a, a), sp.h));
        v[j] = v[j] == 0xff800000u ? v[j] : __float_as_uint(t);  // masked / out-of-range rows stay -inf (a may be 0)
    }
}

// The same transform with the tile's 256 norms already in registers (TileNorms below): lane l holds the norms of rows
// l * 8 .. l * 8 + 7 of the tile, so row j of chunk ch lives in register j & 7 of lane 4 ch + (j >> 3).
struct TileNorms {
    float a[8];   // norms of tile rows lane * 8 + i
    float2 b;     // lanes 0 .. 7: (largest, smallest) norm of chunk `lane`
};
// Issued once per tile, before its first chunk is read from TMEM: the per-chunk loads they replace (one dependent global load
// for the bounds, one for the norms, per chunk and warp, nothing to overlap them with) made the l2 / ip epilogue
// latency-bound once the CTA-pair kernel had halved the MMA time per tile (config 3, ip: main segment 40.4 vs 36.7 ms).
__device__ __forceinline__ void load_tile_norms(TileNorms& tn, const TileSpace& sp, uint32_t tile_row0, uint32_t n_rows, int lane) {
    const uint32_t r0 = tile_row0 + (uint32_t)lane * 8u;
    if (tile_row0 + (uint32_t)BN <= n_rows) {
        const float4 lo = __ldg(reinterpret_cast<const float4*>(sp.norms + r0));
        const float4 hi = __ldg(reinterpret_cast<const float4*>(sp.norms + r0) + 1);
        tn.a[0] = lo.x; tn.a[1] = lo.y; tn.a[2] = lo.z; tn.a[3] = lo.w;
        tn.a[4] = hi.x; tn.a[5] = hi.y; tn.a[6] = hi.z; tn.a[7] = hi.w;
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) tn.a[i] = __ldg(sp.norms + min(r0 + (uint32_t)i, n_rows - 1u));
    }
    const uint32_t crow = tile_row0 + (uint32_t)(lane & 7) * 32u;
    tn.b = crow < n_rows ? __ldg(sp.bounds + (crow >> 5)) : make_float2(0.0f, 0.0f);
}
__device__ __forceinline__ void apply_space_tile(uint32_t (&v)[32], const TileNorms& tn, float h, int ch) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const float a = __shfl_sync(0xffffffffu, tn.a[j & 7], 4 * ch + (j >> 3));
        const float t = __fmaf_rn(a, __uint_as_float(v[j]), -__fmul_rn(__fmul_rn(a, a), h));
        v[j] = v[j] == 0xff800000u ? v[j] : __float_as_uint(t);  // masked / out-of-range rows stay -inf (a may be 0)
    }
}

struct ChunkTop {
    uint32_t v[8];
    uint32_t mn;     // min of v
    uint32_t* out;   // this thread's 8 slots of cmax
    bool on;         // record chunks of the tile being processed
};

template <int CAP, int KPRIME, bool TRACK = false>
__device__ __forceinline__ void epilogue_tile(uint32_t taddr, uint32_t tile_row0, uint32_t n_rows,
                                              const uint32_t* __restrict__ mask, uint32_t* st, uint32_t* tau_glob_q, int lane,
                                              float& tau, uint32_t& cnt, uint64_t* lst, ChunkTop* trk = nullptr,
                                              const TileSpace sp = TileSpace()) {
    constexpr int KPL = CAP / 32;
    TileNorms tn;
    if (sp.norms != nullptr) load_tile_norms(tn, sp, tile_row0, n_rows, lane);  // uniform: l2 / ip spaces
    // the tile's 8 allow-mask words in ONE load per warp (lane ch holds the word of chunk ch), before the first chunk is read:
    // a load per chunk sits in front of a branch and of the masking of all 32 scores, with nothing to overlap it with
    uint32_t mw_l = 0xffffffffu;
    if (mask != nullptr) {
        const uint32_t crow = tile_row0 + (uint32_t)(lane & 7) * 32u;
        mw_l = crow < n_rows ? __ldg(mask + (crow >> 5)) : 0u;
    }
#pragma unroll 1
    for (int ch = 0; ch < BN / 32; ++ch) {
        const uint32_t row0 = tile_row0 + ch * 32;
        if (row0 >= n_rows) break;  // uniform
        // allow-mask word of this chunk (uniform over the warp).  Masked and out-of-range rows become -inf BEFORE the
        // fast-reject test: thresholds are learnt on allowed rows only, so at 10 % selectivity unmasked scores would send
        // ten times as many chunks through the hit path only to be dropped there (measured: +17 % on config 5); chunks
        // without any allowed row are not even read from TMEM.
        uint32_t mw = __shfl_sync(0xffffffffu, mw_l, ch);
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
        float mx = max32(v, grp);
        if (sp.norms != nullptr) {  // uniform: l2 / ip spaces
            // most chunks hold nothing near the threshold: reject them with a bound on their ranking scores (two floats
            // per chunk) before paying for the transform of all 32 x 32 scores
            const float2 cb = make_float2(__shfl_sync(0xffffffffu, tn.b.x, ch), __shfl_sync(0xffffffffu, tn.b.y, ch));
            if (__ballot_sync(0xffffffffu, chunk_score_bound(mx, cb, sp.h) >= tau) == 0u) continue;
            apply_space_tile(v, tn, sp.h, ch);
            mx = max32(v, grp);
        }
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
        if constexpr (CAP > 512) {
            // Wide shortlists (k > 128: k' = 512 / 1152, lists of 2048): no in-kernel compaction -- a warp cannot sort 2048 keys in
            // registers.  Thresholds learnt from the chunk maxima keep a list at a few dozen entries; one that fills up anyway
            // (adversarial order, thousands of ties) stops collecting and raises the query's threshold to +inf: the tail
            // then finds no candidates, flags the query, and the repair chain (band pass, exact scan) answers it.
            if (cnt > CAP - 32 && tau != INFINITY) {
                atomicMax(tau_glob_q, 0xFF800000u);  // ord(+inf)
                tau = INFINITY;
            }
            continue;
        }
        // lists that could overflow on the next chunk are compacted now (warp-cooperative)
        unsigned todo = __ballot_sync(0xffffffffu, CAP <= 512 && cnt > CAP - 32);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            __syncwarp();
            const uint64_t* L = reinterpret_cast<const uint64_t*>(__shfl_sync(0xffffffffu, (unsigned long long)lst, src));
            const uint32_t c = min(__shfl_sync(0xffffffffu, cnt, src), (uint32_t)CAP);
            constexpr int KS = CAP <= 512 ? KPL : 1;  // (the wide-list instantiations never get here)
            uint64_t key[KS];
#pragma unroll
            for (int r = 0; r < KS; ++r) key[r] = (uint32_t)(r * 32 + lane) < c ? __ldcg(L + r * 32 + lane) : 0ull;
            warp_sort_desc<KS>(key, lane);
#pragma unroll
            for (int r = 0; r < KS; ++r)
                if (r * 32 + lane < KPRIME) const_cast<uint64_t*>(L)[r * 32 + lane] = key[r];
            const uint64_t kth = __shfl_sync(0xffffffffu, key[((KPRIME - 1) >> 5) % KS], (KPRIME - 1) & 31);
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
                                                  const uint32_t* __restrict__ mask, uint32_t (&out)[BN / 32],
                                                  const TileSpace sp = TileSpace(), int lane = 0) {
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
            if (sp.norms != nullptr) {  // uniform
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = (mw >> j) & 1u ? v[j] : 0xff800000u;
                apply_space(v, sp, row0, n_rows, lane);
            }
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, (mw >> j) & 1u ? __uint_as_float(v[j]) : -INFINITY);
            ord = mw ? ord_u32(m) : 0u;
        }
        out[ch] = ord;
    }
}

// ------------------------------------------------------------------------------------------ streaming selection
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

// ------------------------------------------------------------------------------------------ host helpers
// [rows, 768] row-major matrix, box = one 128-byte K-chunk x box_rows rows, 128B swizzle (kernels_tc.cu).
bool make_map(CUtensorMap* map, const void* base, uint64_t rows, int dtype, int box_rows);

// MQVS_TRACE=1: CUDA events after every launch of a tensor-core search, printed by trace_dump() (developer aid)
struct Trace {
    bool on;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    unsigned long long* dbg = nullptr;  // [grid][16] phase stamps of the last tc_fused_kernel launch
    int dbg_grid = 0;
    Trace();
    unsigned long long* fused_dbg(int grid);
    void mark(const char* what, cudaStream_t st);
};
Trace& trace();

}  // namespace tc
}  // namespace mqvs
