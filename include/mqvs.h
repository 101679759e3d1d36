/*
 * mqvs.h -- C ABI of libmqvs.so, the B200-native exact vector-search backend behind MediQuery-RAG's
 * `vectorstore` object.
 *
 * The reference has NO native interface for this path: its three call sites go through
 * langchain_chroma.Chroma (pure Python) into chromadb / chroma-hnswlib, which are un-vendored.  Each entry
 * point below therefore cites the reference call (file:line under /root/reference) whose work it does; the
 * Python class in mediquery-rag_b200/vectorstore.py binds them with ctypes (INTEGRATION.md shows the stub).
 *
 * Conventions: plain pointers and sizes only; the caller owns every in/out buffer; the library owns the
 * corpus in HBM, scratch, streams.  Every function returns MQVS_OK (0) or a negative code and never throws;
 * mqvs_last_error() gives the thread-local message.  A handle is not re-entrant (one call at a time per
 * handle, enforced by an internal mutex) but may be used from any host thread.
 * "is_device" flags say whether a pointer is device memory on the store's device (1) or host (0).
 */
#ifndef MQVS_H_
#define MQVS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MQVS_VERSION 200

#if defined(__GNUC__)
#define MQVS_API __attribute__((visibility("default")))
#else
#define MQVS_API
#endif

enum { MQVS_OK = 0, MQVS_E_INVALID = -1, MQVS_E_CUDA = -2, MQVS_E_NCCL = -3, MQVS_E_OOM = -4, MQVS_E_IO = -5 };
enum { MQVS_F32 = 0, MQVS_BF16 = 1 };

/* Chroma's `hnsw:space` (collection_metadata; the reference passes none, so upstream it is "l2" -- src/medical_engine.py:52) */
enum { MQVS_METRIC_COSINE = 0, MQVS_METRIC_L2 = 1, MQVS_METRIC_IP = 2 };

/* widest embedding the kernels of THIS BUILD take; rows are zero-padded to MQVS_DIM_PAD elements in HBM
 * (768 = width of shaw/dmeta-embedding-zh, src/medical_engine.py:43).  A build-time constant of the kernels (any multiple
 * of 128): `make -C mediquery-rag_b200/csrc WIDTH=1024` builds libmqvs_w1024.so from the same sources with the same ABI, and
 * mqvs_dim_pad() tells a caller what a loaded library was built for (the Python binding picks the narrowest build that
 * holds the collection's dimension). */
#ifndef MQVS_DIM_PAD
#define MQVS_DIM_PAD 768
#endif
#define MQVS_MAX_K 1024
#define MQVS_MAX_DEVICES 16      /* row shards one handle can drive (one B200 box has 8) */
#define MQVS_SHARD_BLOCK 4096    /* rows per block of the block-cyclic placement of a multi-device handle */

typedef struct mqvs_store mqvs_t;

typedef struct mqvs_stats {
    uint64_t rows;               /* rows stored (all devices) */
    uint64_t searches;           /* mqvs_search calls completed */
    uint64_t queries;            /* sum of B over those calls */
    uint64_t kernel_launches;    /* kernels launched by this handle since creation */
    uint64_t scan_launches;      /* of those, corpus-scan (K2) launches */
    uint64_t scan_bytes;         /* algorithmic corpus bytes streamed by the scans: rows * 768 * sizeof(elem) per pass */
    double   scan_ms;            /* device time of the scan launches (CUDA events on the launch stream).  One exception: a blocking
                                    mqvs_search with host outputs on the single-launch latency path (B <= 4, k <= 16, small stores)
                                    records no events -- they cost 6 us of a 33 us call -- and adds the host-observed time from the
                                    launch call to the kernel's completion word instead; MQVS_TIMING=1 or 2 restores the events */
    double   last_search_ms;     /* device time of the last search: its scan launches (= last_scan_ms), or first kernel to last when the
                                    process runs with MQVS_TIMING=1 (two more CUDA events per search) */
    double   last_scan_ms;       /* device time of the last search's scan launches */
    uint32_t last_scan_passes;   /* corpus passes the last search needed */
    uint32_t last_path;          /* 0 = SIMT fp32 scan, 1 = tensor-core shortlist + fp32 rescoring */
    uint64_t exact_fallbacks;    /* queries whose tensor-core shortlist could not be proven exact and were re-run on the SIMT scan */
    uint64_t unrepaired;         /* flagged queries the device-side repair rounds of a search left flagged.  mqvs_search_async sizes its
                                    rounds to the batch, so this stays 0 there; the blocking mqvs_search repairs any rest from the host */
    uint32_t last_pipeline;      /* how the last search ran: 0 = SIMT scan, normalise + scan(s) + merge launches; 1 = SIMT scan, one
                                    launch (B <= 4, k <= 16); 2 = tensor cores, corpus segments + selection launches;
                                    3 = tensor cores, one cooperative launch (B <= 256) */
    uint32_t reserved_;
    uint64_t band_repairs;       /* flagged queries settled by the band pass (a second tensor-core pass that collects and rescoring every
                                    row within eps of the k-th exact score) instead of an exact scan */
} mqvs_stats;

MQVS_API int mqvs_version(void);
MQVS_API int mqvs_dim_pad(void);   /* MQVS_DIM_PAD of this build */
MQVS_API const char* mqvs_last_error(void);
MQVS_API int mqvs_device_count(int* out);

/* Chroma(persist_directory=..., embedding_function=...) -- src/medical_engine.py:52 -- and the collection
 * Chroma.from_documents creates -- src/ingest_medical.py:106-110.  dim <= MQVS_DIM_PAD.
 * n_dev = 1: one row shard on device dev_ids[0] (NULL = device 0).
 * n_dev > 1: ONE handle, in ONE process, over n_dev row shards on the devices dev_ids[0 .. n_dev) -- what the reference
 * needs, since its `vectorstore` is one module-global object in one process (src/medical_engine.py:52, main.py:42-48) while
 * 100M x 768 rows need the HBM of a whole box.  Rows are placed block-cyclically (global block r / MQVS_SHARD_BLOCK lives
 * on shard block % n_dev, see mqvs_shard_of_row), so shards stay balanced under appends and every row keeps its GLOBAL id
 * through every entry point below: searches scan all shards concurrently (one stream per device), each shard writes its
 * exact top-k straight into the first device's memory over NVLink peer access (a peer copy where that is unavailable), and
 * one merge kernel there finishes the batch (same ordering rule).  Results are bit-identical to a single-shard handle
 * holding the same rows.  A device may be named more than once (several shards on one GPU: how the single-GPU tests
 * exercise this path).  dev_ids[0] is the "home" device: device pointers passed to such a handle (x / q / outputs with
 * *_is_device = 1, mqvs_search_async) must be memory of that device.
 * The alternative layout -- one process per GPU, one single-shard handle each, candidates exchanged with mqvs_xchg_merge or
 * NCCL + mqvs_merge_candidates -- is mediquery-rag_b200/sharded.py. */
MQVS_API int mqvs_create(int n_dev, const int* dev_ids, int dim, int store_dtype, uint64_t capacity_rows, mqvs_t** out);
MQVS_API void mqvs_destroy(mqvs_t* s);
MQVS_API int mqvs_count(mqvs_t* s, uint64_t* out);
MQVS_API int mqvs_reserve(mqvs_t* s, uint64_t capacity_rows);

/* collection.upsert(embeddings=...) under Chroma.from_documents / add_texts -- src/ingest_medical.py:106.
 * x: [n, dim] row-major, x_dtype MQVS_F32 (bf16 input is not accepted).  Runs K1 (L2-normalise + pack).
 * Rows with a non-finite norm are rejected: nothing is appended and MQVS_E_INVALID is returned. */
MQVS_API int mqvs_add(mqvs_t* s, const void* x, int x_dtype, int x_is_device, uint64_t n, uint64_t* first_row_out);

/* Bench/test utility (BASELINE.json configs 2-5 use synthetic 768-d unit vectors; 100M rows cannot be staged
 * through the host): appends n rows v[r, c] = uniform[-1,1) from splitmix64(seed, (first_index + r) * dim + c),
 * generated on the device and passed through the same K1. */
MQVS_API int mqvs_add_synthetic(mqvs_t* s, uint64_t seed, uint64_t first_index, uint64_t n, uint64_t* first_row_out);

/* vectorstore.similarity_search(query, k) -- src/agents/nodes.py:93 (k=5), src/medical_engine.py:70 (k=3) --
 * after the embedder: q [B, dim] fp32 (normalised here), mask = optional allow-bitmask over rows (bit r&31 of
 * word r>>5; ceil(rows/32) words; NULL = all rows), the Chroma `where` filter compiled by the host.
 * Outputs [B, k] row-major: cosine similarity descending, ties -> lower row id; slots beyond the number of
 * allowed rows hold (-inf, -1).  Blocks until the outputs are ready. */
MQVS_API int mqvs_search(mqvs_t* s, const float* q, int q_is_device, int B, int k, const uint32_t* mask, int mask_is_device,
                float* out_scores, int64_t* out_ids, int out_is_device);

/* Pure host arithmetic of the block-cyclic placement (no device needed): global row -> (shard, local row) for a handle of
 * n_dev shards; the inverse is global = ((local / BLOCK) * n_dev + shard) * BLOCK + local % BLOCK. */
MQVS_API int mqvs_shard_of_row(int n_dev, uint64_t row, int* shard_out, uint64_t* local_row_out);

/* Same, enqueued on `cuda_stream` (a cudaStream_t, e.g. torch.cuda.current_stream().cuda_stream) with all
 * buffers on the device; returns after enqueueing.  id_base is added to every returned id (row-sharded
 * stores: the rank's first global row).  Scratch belongs to the handle: enqueue the next call on the same stream
 * (or after this one has finished).  Exact on its own: the device-side repair (band pass, then exact-scan rounds sized to
 * B) settles every query the tensor pipeline could not prove.  Non-finite queries are not reported here (their rows come
 * back as scores of a zero query; use mqvs_search to validate).  On a multi-device handle the buffers and the stream
 * belong to the home device and mask_dev must be the pointer mqvs_filter_build returned (or NULL). */
MQVS_API int mqvs_search_async(mqvs_t* s, const float* q_dev, int B, int k, const uint32_t* mask_dev, int64_t id_base,
                      float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream);

/* Cross-shard merge after the all-gather of per-rank candidates: scores/ids
 * [W, B, k] on `device` -> best k per query [B, k], same ordering rule; empty slots are id < 0. */
MQVS_API int mqvs_merge_candidates(int device, const float* scores_dev, const int64_t* ids_dev, int W, int B, int k,
                          float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream);

/* The same step fused with its transfer, for the ranks of one box (one process per GPU): instead of two NCCL
 * all-gathers + mqvs_merge_candidates, ONE kernel stores this rank's [B, k] candidates into every peer's inbox over
 * NVLink peer memory (CUDA IPC), signals, waits for the peers' signals and merges W*k -> k (same ordering rule).
 * Setup: every rank creates its exchange, publishes its 64-byte handle to the others by any host channel
 * (sharded.py uses torch.distributed.all_gather_object), then connects with the W handles in rank order.
 * Every rank must call mqvs_xchg_merge the same number of times with the same (B, k); B * k <= max_entries. */
typedef struct mqvs_xchg mqvs_xchg_t;
MQVS_API int mqvs_xchg_create(int device, int rank, int world, uint64_t max_entries, mqvs_xchg_t** out);
MQVS_API int mqvs_xchg_get_handle(mqvs_xchg_t* x, void* handle_out, int handle_bytes);
MQVS_API int mqvs_xchg_connect(mqvs_xchg_t* x, const void* handles);
/* Ranks that live in ONE process (one host thread per GPU, or several logical shards on one GPU in the tests) skip
 * IPC: each publishes mqvs_xchg_base() and connects with the W base pointers (peer access is enabled as needed). */
MQVS_API int mqvs_xchg_base(mqvs_xchg_t* x, void** base_out);
MQVS_API int mqvs_xchg_connect_ptrs(mqvs_xchg_t* x, void* const* bases);
MQVS_API int mqvs_xchg_max_entries(mqvs_xchg_t* x, uint64_t* out);
MQVS_API int mqvs_xchg_merge(mqvs_xchg_t* x, const float* scores_dev, const int64_t* ids_dev, int B, int k,
                             float* out_scores_dev, int64_t* out_ids_dev, void* cuda_stream);
/* The same call for ALL ranks of an exchange that live on ONE device (logical shards sharing a GPU): one cooperative launch
 * does every rank's work.  Separate launches must not be used there -- a kernel that spins on a flag another launch has
 * yet to write may never see that launch scheduled.  ranks / buffers: arrays of W entries in rank order.  Synchronous. */
MQVS_API int mqvs_xchg_merge_colocated(mqvs_xchg_t* const* ranks, int W, const float* const* scores_dev,
                                       const int64_t* const* ids_dev, int B, int k, float* const* out_scores_dev,
                                       int64_t* const* out_ids_dev, void* cuda_stream);
MQVS_API void mqvs_xchg_destroy(mqvs_xchg_t* x);

/* Chroma `where` filters evaluated ON THE DEVICE (collection.query(where=...)): metadata
 * columns are uploaded once as typed arrays (strings as dictionary codes the host assigns in string order), a filter is a
 * postfix program of leaf comparisons and AND / OR, and mqvs_filter_build writes the packed allow-bitmask mqvs_search
 * takes (mask_is_device = 1) without the mask ever visiting the host.  Semantics as Chroma's: a row whose field is
 * missing matches only $ne / $nin.  live_words (host, ceil(rows/32) words, nullable) is AND-ed in (tombstones). */
enum { MQVS_OP_EQ = 0, MQVS_OP_NE = 1, MQVS_OP_GT = 2, MQVS_OP_GTE = 3, MQVS_OP_LT = 4, MQVS_OP_LTE = 5, MQVS_OP_IN = 6,
       MQVS_OP_NIN = 7, MQVS_OP_CONTAINS = 8, MQVS_OP_NOT_CONTAINS = 9, MQVS_OP_TEXT_CONTAINS = 10,
       MQVS_OP_TEXT_NOT_CONTAINS = 11, MQVS_OP_AND = 16, MQVS_OP_OR = 17 };
enum { MQVS_COL_I64 = 0, MQVS_COL_F64 = 1, MQVS_COL_TOKENS = 2, MQVS_COL_TEXT = 3 };
typedef struct mqvs_filter_op {
    int32_t op;        /* MQVS_OP_* */
    int32_t col;       /* leaves: column id given to mqvs_column_set */
    int64_t ival;      /* comparison constant for MQVS_COL_I64 columns */
    double fval;       /* comparison constant for MQVS_COL_F64 columns */
    uint32_t set_off;  /* $in / $nin: first element of the set in the constant pool; text leaves: first 8-byte word of the needle */
    uint32_t set_len;  /* $in / $nin: elements of the set; text leaves: BYTES of the needle (1 .. MQVS_MAX_NEEDLE) */
} mqvs_filter_op;
#define MQVS_MAX_NEEDLE 1024
/* values_host: n 8-byte values (int64 or float64); present_host: n bytes (0 = field missing) or NULL = all present.
 * Replaces column col_id (0..63).  n may be smaller than the row count: later rows count as "field missing". */
MQVS_API int mqvs_column_set(mqvs_t* s, int col_id, int col_type, const void* values_host, const uint8_t* present_host,
                             uint64_t n);
/* Token-set columns: the reference stores `tags` as ONE comma-joined string per chunk (src/ingest_medical.py:66-67,75-79:
 * "慢性病，代谢性疾病，发病率"), so a per-tag filter is set membership, not string equality.  The host tokenises (on "," and
 * "，"), assigns token ids from a dictionary and uploads the rows' token lists in CSR form: tokens of row r are
 * tokens_host[offsets_host[r] .. offsets_host[r + 1]).  Leaves MQVS_OP_CONTAINS / MQVS_OP_NOT_CONTAINS test ival (a token id;
 * -1 = a token no row carries) against a row's list; a row without tokens matches only NOT_CONTAINS. */
MQVS_API int mqvs_column_set_tokens(mqvs_t* s, int col_id, const uint64_t* offsets_host, const int64_t* tokens_host, uint64_t n);
/* Text columns -- Chroma's `where_document` ({"$contains": s} / {"$not_contains": s} over the stored documents; the
 * reference's page_content is "问题：{title}\n答案：{content}", src/ingest_medical.py:71): the documents' UTF-8 bytes, all rows
 * concatenated, in CSR form (row r = bytes_host[offsets_host[r] .. offsets_host[r + 1])); present_host: n bytes (0 = the
 * row has no document) or NULL.  Leaves MQVS_OP_TEXT_CONTAINS / MQVS_OP_TEXT_NOT_CONTAINS search every row for the
 * needle stored in the constant pool (set_len bytes from 8-byte word set_off; byte-wise equality on UTF-8 = substring
 * equality on code points); a row without a document matches only NOT_CONTAINS. */
MQVS_API int mqvs_column_set_text(mqvs_t* s, int col_id, const uint64_t* offsets_host, const uint8_t* bytes_host,
                                  const uint8_t* present_host, uint64_t n);
/* prog: n_ops <= 64 operations in postfix order; pool: pool_len 8-byte constants (int64 or float64 bit patterns) the
 * $in / $nin leaves point into.  *mask_dev_out is owned by the handle and valid until the next mqvs_filter_build.
 * Multi-device handles evaluate every shard's slice on its own device; *mask_dev_out is then an opaque token for
 * mqvs_search / mqvs_search_async of the same handle (mask_is_device = 1), not a pointer to read through. */
MQVS_API int mqvs_filter_build(mqvs_t* s, const mqvs_filter_op* prog, int n_ops, const void* pool, uint64_t pool_len,
                               const uint32_t* live_words_host, const uint32_t** mask_dev_out, uint64_t* allowed_out);

/* collection.get(ids, include=["embeddings"]): stored (normalised) rows as fp32, [n, dim] to host. */
MQVS_API int mqvs_get_rows(mqvs_t* s, const int64_t* row_ids, uint64_t n, float* out_host);

/* persist_directory -- src/ingest_medical.py:109, src/medical_engine.py:32-37,52: raw row file + header. */
MQVS_API int mqvs_save(mqvs_t* s, const char* path);
/* Append rows [first_row, count) to a file written earlier by mqvs_save / mqvs_save_append that holds exactly first_row
 * rows (re-running ingest appends -- src/ingest_medical.py:106 on an existing ./medical_db); first_row = 0 rewrites. */
MQVS_API int mqvs_save_append(mqvs_t* s, const char* path, uint64_t first_row);
MQVS_API int mqvs_load(mqvs_t* s, const char* path);

/* Tombstone compaction (collection.delete / upsert leave dead rows that every scan still streams): keep only the rows whose
 * bit is set in keep_words_host (bit r & 31 of word r >> 5, ceil(rows / 32) words), gathered bit-for-bit on the device in
 * ascending order, so surviving row r becomes row popcount(keep[0 .. r)).  Stored norms follow; device-resident metadata
 * columns are dropped.  On a multi-device handle the survivors keep their GLOBAL order and therefore change device where the
 * block-cyclic placement says so: they travel through a pinned host buffer, chunk by chunk (same bits, same result as
 * compacting a single-device handle that holds the same rows). */
MQVS_API int mqvs_compact(mqvs_t* s, const uint32_t* keep_words_host, uint64_t* new_count_out);

/* collection_metadata={"hnsw:space": ...} -- the space rows are RANKED in when the raw vectors given to mqvs_add are not
 * unit vectors (on unit vectors the three spaces order rows identically and MQVS_METRIC_COSINE, the default, serves all
 * of them).  Rows are always stored normalised; with MQVS_METRIC_L2 / MQVS_METRIC_IP the store also keeps ||x|| of every
 * raw row (4 bytes per row, written by K1, saved to `<path>.norms` by mqvs_save) and the scan maps the cosine s of the
 * normalised pair to the space's ranking score before top-k (larger = better in every space):
 *     cosine: s            ip: ||x|| s = <q, x> / ||q||            l2: ||x|| s - ||x||^2 / (2 ||q||) = (||q||^2 - ||q - x||^2) / (2 ||q||)
 * out_scores of mqvs_search hold that ranking score; the caller converts to Chroma's distances with its ||q||
 * (ip: 1 - ||q|| t; l2: ||q||^2 - 2 ||q|| t).  Norms are kept from the first call with a non-cosine metric on, so that
 * call must come while the store is empty; afterwards the metric may be switched freely (e.g. back to
 * MQVS_METRIC_COSINE while every row is a unit vector, see mqvs_norm_deviation).  Non-cosine metrics run on the exact
 * SIMT scan (1-2 queries, small stores, k > 128) or the tensor pipelines, chosen by batch size as for cosine. */
MQVS_API int mqvs_set_metric(mqvs_t* s, int metric);
/* largest | ||x|| - 1 | over the raw rows added (or loaded with their norms) so far: 0 for a collection of exact unit vectors */
MQVS_API int mqvs_norm_deviation(mqvs_t* s, float* out);
/* ||x|| of raw rows [first_row, first_row + n) as stored by K1 (needs a non-cosine metric set before the rows were added) */
MQVS_API int mqvs_get_norms(mqvs_t* s, uint64_t first_row, uint64_t n, float* out_host);

MQVS_API int mqvs_get_stats(mqvs_t* s, mqvs_stats* out);
/* 0 = choose by batch size, 1 = force the SIMT fp32 scan, 2 = force the tensor-core path,
 * 3 = SIMT scan without the single-launch latency kernel, 4 = tensor-core path without the single-launch pipeline
 * for B <= 256 and without the band pass (tests compare the variants) */
MQVS_API int mqvs_set_path(mqvs_t* s, int path);

/* bf16 scoring shadow of an fp32 store (MQVS_F32 handles; on by default, MQVS_NO_SHADOW=1 in the environment or enable = 0
 * turns it off; no effect on MQVS_BF16 handles).  The tensor-core pipelines are bound by the bytes they stream, so an fp32
 * store keeps a second, bf16-rounded copy of its rows (+50 % HBM; built on the device before the first tensor search after
 * an add, and dropped for good if HBM cannot hold it) and streams THAT: half the bytes per row and kind::f16 MMAs at twice
 * the kind::tf32 rate.  Tensor-core scores only ever shortlist; every score that is returned or compared -- K4 rescoring
 * of the shortlist, the band pass, the exact-scan repair, the SIMT paths -- is computed from the stored fp32 rows, and the
 * exactness margin accounts for the rounding of both operands (||q - bf16(q)|| ||x~|| + ||q|| max ||x - bf16(x)||), so
 * results are bit-identical with the shadow on or off.  *bytes_out (nullable) receives the shadow's current size. */
MQVS_API int mqvs_set_shadow(mqvs_t* s, int enable);
MQVS_API int mqvs_shadow_bytes(mqvs_t* s, uint64_t* bytes_out);

#ifdef __cplusplus
}
#endif
#endif /* MQVS_H_ */
