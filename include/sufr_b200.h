/*
 * sufr_b200.h -- C ABI of the B200-native suffix-array / LCP-array constructor.
 *
 * Drop-in boundary for ONE path of TravisWheelerLab/sufr (v0.7.12): libsufr's `create`
 * (`SufrBuilder::<T>::new`, libsufr/src/sufr_builder.rs:143-220, called from
 * `SuffixArray::write`, libsufr/src/suffix_array.rs:460-470, and `sufr::create`,
 * sufr/src/lib.rs:321-371).  Everything here is plain C: pointers, sizes, no CUDA or torch types.
 * A Rust host binds these symbols from an `extern "C"` block (INTEGRATION.md shows the shim).
 *
 * Results are bit-exact with the reference for the SA and LCP arrays, u32 and u64 index widths
 * (see DESIGN.md for the two cases where the reference itself is not a function of its input).
 * There is no CPU fallback: every entry point that computes fails with an error when no CUDA
 * device is available.
 */
#ifndef SUFR_B200_H
#define SUFR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SUFR_B200_ABI_VERSION 2

/* Return codes. */
#define SUFR_B200_OK 0
#define SUFR_B200_ERR_ARGUMENT 1      /* the reference's `bail!` argument errors, same message text */
#define SUFR_B200_ERR_OUT_OF_MEMORY 2
#define SUFR_B200_ERR_INTERNAL 3
#define SUFR_B200_ERR_IO 4            /* "{filename}: {io error}" (sufr_builder.rs:820) */
#define SUFR_B200_ERR_UNSUPPORTED 5
#define SUFR_B200_ERR_CUDA 100        /* 100 + cudaError_t */

/* Where a buffer lives. */
#define SUFR_B200_MEM_HOST 0
#define SUFR_B200_MEM_DEVICE 1

/* Per-device context: one CUDA stream and one device memory pool.  One context per GPU per process. */
typedef struct SufrB200Ctx SufrB200Ctx;

/*
 * Mirrors `SufrBuilderArgs` (libsufr/src/types.rs:527-582) field by field.  Options become a
 * presence flag + value (max_query_len) or a nullable pointer (path, seed_mask).
 */
typedef struct SufrB200Args {
    const uint8_t* text;          /* types.rs:530  raw text bytes; borrowed for the call */
    uint64_t text_len;
    const char* path;             /* types.rs:533  NULL = None ("out.sufr" when a file is written) */
    uint8_t low_memory;           /* types.rs:536  accepted and ignored, like the reference builder */
    uint8_t has_max_query_len;    /* types.rs:540  Option tag */
    uint8_t is_dna;               /* types.rs:546 */
    uint8_t allow_ambiguity;      /* types.rs:550 */
    uint8_t ignore_softmask;      /* types.rs:554 */
    uint8_t reserved[3];
    uint64_t max_query_len;
    const uint64_t* sequence_starts;   /* types.rs:558 */
    const char* const* sequence_names; /* types.rs:562  num_sequences NUL-terminated UTF-8 strings */
    uint64_t num_sequences;
    uint64_t num_partitions;      /* types.rs:573  accepted; does not change the result */
    const char* seed_mask;        /* types.rs:577  NULL = None, else "1101..." */
    uint64_t random_seed;         /* types.rs:581  accepted; does not change the result */
    /* Sharding, one process per GPU: this call builds key-range shard `rank` of `world_size`.
     * Single GPU: rank 0, world_size 1. */
    int32_t rank;
    int32_t world_size;
} SufrB200Args;

/* Phase timings measured with CUDA events on the context's stream (milliseconds). */
typedef struct SufrB200Timings {
    double h2d_ms;        /* host -> device copy of the text (0 when the text was already resident) */
    double encode_ms;     /* text transform, alphabet, packing, N-run scan   (sufr_builder.rs:144-195) */
    double keys_ms;       /* first key word of every suffix + shard selection (partition assignment, :404-487) */
    double sort_ms;       /* radix sort of (key, position)                    (sort, :495-598) */
    double refine_ms;     /* equal-key groups: next-word / prefix-doubling refinement (merge compares, :634-767) */
    double lcp_ms;        /* LCP values not already produced by the sort      (find_lcp, :268-334) */
    double finish_ms;     /* N-run tie rule, suffix filter, widening          (:446-449, :305-307) */
    double d2h_ms;        /* device -> host copy of text / SA / LCP (0 for device results) */
    double total_ms;      /* first kernel to last kernel, excluding h2d/d2h */
    /* The dominant kernel (radix-sort downsweep of the main sort), timed per launch with CUDA events: */
    double dominant_kernel_ms;        /* sum over its launches in this build */
    uint64_t dominant_kernel_launches;
    uint64_t dominant_kernel_bytes;   /* algorithmic bytes of ONE launch: elements x 2 x (8 B key + 4 B position) */
} SufrB200Timings;

typedef struct SufrB200Result {
    uint32_t index_bits;      /* 32 or 64: width of the SA / LCP elements */
    uint32_t memory;          /* SUFR_B200_MEM_HOST (pinned) or SUFR_B200_MEM_DEVICE */
    uint64_t text_len;
    uint64_t num_suffixes;    /* suffixes in THIS shard (== total_suffixes when world_size == 1) */
    uint64_t total_suffixes;  /* SufrBuilder.num_suffixes over all shards.  Sharded full sorts balance the shards on a */
    uint64_t shard_offset;    /*   sampled histogram: then these two are (num_suffixes, 0) and the CALLER overwrites them with */
                              /*   the sum / prefix sum of the ranks' num_suffixes (the same exchange the seam repair needs) */
    uint64_t first_suffix;    /* SA[0] / SA[num_suffixes-1] of this shard: inputs of the seam repair */
    uint64_t last_suffix;     /*   (sufr_builder.rs:893-902); undefined when num_suffixes == 0 */
    uint8_t* text;            /* transformed text (SufrBuilder.text), text_len bytes; NULL in HOST results of ranks > 0 */
    void* sa;                 /* num_suffixes elements of index_bits */
    void* lcp;                /* num_suffixes elements; lcp[0] of shard > 0 needs sufr_b200_patch_seam */
    uint64_t* n_ranges;       /* host: num_n_ranges pairs [start, end) (SufrBuilder.n_ranges) */
    uint64_t num_n_ranges;
    SufrB200Timings timings;
    uint64_t kernel_launches; /* kernels launched by this build */
    uint64_t peak_device_bytes;
    uint32_t alphabet_size;   /* distinct bytes in the transformed text */
    uint32_t bits_per_symbol;
    uint32_t refine_rounds;   /* next-word refinement rounds */
    uint32_t doubling_rounds; /* prefix-doubling rounds (0 unless the text has deep repeats) */
    uint64_t h2d_bytes;       /* bytes this build copied host -> device (0 for a device-resident text) */
    uint64_t d2h_bytes;       /* bytes it copied device -> host: large host results travel compactly (LCP as */
                              /*   bytes + exceptions, 64-bit SA as u32) and are widened by host threads */
    uint32_t position_bits;   /* width of text positions INSIDE the build: 32, or 64 for texts of u32::MAX bytes and */
    uint32_t reserved2;       /*   more (suffix_array.rs:460-470; SUFR_B200_DEBUG_POS64 forces 64 on short texts) */
    void* owner;              /* internal */
} SufrB200Result;

/* -- context ------------------------------------------------------------------------------- */
int sufr_b200_ctx_create(int device, SufrB200Ctx** out);
void sufr_b200_ctx_destroy(SufrB200Ctx* ctx);
/* Pre-size the device memory pool for texts of up to `text_len` bytes (optional). */
int sufr_b200_ctx_reserve(SufrB200Ctx* ctx, uint64_t text_len, uint32_t index_bits);
/* Give pooled device memory that holds no live result back to the driver. */
void sufr_b200_ctx_trim(SufrB200Ctx* ctx);

/* -- build: replaces SufrBuilder::<u32>::new / SufrBuilder::<u64>::new up to (not including) write()
 *    (sufr_builder.rs:143-217).  `text_memory` says where args->text lives, `result_memory` where the
 *    outputs should be left.  index_bits: 32, 64, or 0 = the reference's dispatch
 *    (u32 iff text_len < u32::MAX, suffix_array.rs:460-470). */
int sufr_b200_build(SufrB200Ctx* ctx, const SufrB200Args* args, uint32_t index_bits, int text_memory,
                    int result_memory, SufrB200Result* out);
void sufr_b200_result_free(SufrB200Ctx* ctx, SufrB200Result* result);

/* Seam repair between shards (sufr_builder.rs:893-902): sets lcp[0] of this shard to the LCP of
 * (`prev_last_suffix`, first_suffix).  No-op for an empty shard. */
int sufr_b200_patch_seam(SufrB200Ctx* ctx, const SufrB200Args* args, SufrB200Result* result,
                         uint64_t prev_last_suffix);

/* -- verify: full check of a finished DEVICE result against the transformed text, independent of the data
 *    structures of the build (sufr_b200/csrc/verify.cuh): every SA entry is an indexed position
 *    (sufr_builder.rs:446-449) and none occurs twice; EVERY adjacent pair is in the order the mode defines and
 *    every LCP value is the mode's value (full sort; N-run rule :305-307, :701-712; --max-query-len; seed mask).
 *    Sharded builds: every rank verifies its shard, `has_prev` / `prev_last_suffix` add the seam pair.
 *    The result is correct iff all error counters are 0 and the ranks' num_suffixes sum to expected_suffixes. */
typedef struct SufrB200VerifyReport {
    uint64_t pairs_checked;     /* adjacent pairs compared (num_suffixes - 1, + 1 with has_prev) */
    uint64_t order_errors;
    uint64_t lcp_errors;
    uint64_t out_of_range;      /* SA entries >= text_len */
    uint64_t not_indexed;       /* SA entries the reference would not index */
    uint64_t duplicates;        /* positions that occur more than once in this array */
    uint64_t first_bad_rank;    /* smallest rank with an order / LCP error, UINT64_MAX if none */
    uint64_t max_lcp;
    uint64_t lcp_sum;
    uint64_t expected_suffixes; /* positions of the text the reference indexes */
    uint64_t deferred_pairs;    /* pairs sharing >= 2048 bytes: settled by a second pass, see method */
    uint32_t method;            /* 0 direct comparison only; 1 deferred pairs compared directly without a bound;
                                   2 deferred pairs by successor ranks (inverse suffix array) + Kasai LCP in text order */
    uint32_t reserved;
    double ms;                  /* device time of the check */
} SufrB200VerifyReport;
int sufr_b200_verify(SufrB200Ctx* ctx, const SufrB200Args* args, const SufrB200Result* result, int has_prev,
                     uint64_t prev_last_suffix, SufrB200VerifyReport* out);

/* -- the consumer of the LCP array: LCP-subsampled suffix array + batched search over a
 *    device-resident index.  An index borrows text / SA / LCP of a DEVICE result, which must outlive it.
 *      subsample  SufrFile::subsample_suffix_array (sufr_file.rs:429-456): the entries with lcp < max_query_len and
 *                 their ranks ("compressed" in-memory suffix array)
 *      search     SufrSearch::search (sufr_search.rs:104-350) for a batch of queries, one GPU thread per query:
 *                 queries = concatenated bytes, offsets[num_queries + 1]; has/max_query_len = the run-time `-m`;
 *                 use_subsample != 0 searches the subsampled array and maps the hits back through the ranks.
 *                 rank_begin / rank_end (host, num_queries each) = half-open rank range in the full suffix array
 *                 (SearchResultLocations::ranks), both UINT64_MAX when the query does not occur; count = end - begin.
 *      suffixes   SA[rank_begin .. rank_begin + count) as u64 (what `locate` reports), host output. */
typedef struct SufrB200Index SufrB200Index;
int sufr_b200_index_create(SufrB200Ctx* ctx, const SufrB200Args* args, const SufrB200Result* device_result,
                           SufrB200Index** out);
int sufr_b200_index_subsample(SufrB200Index* index, uint64_t max_query_len, uint64_t* kept);
int sufr_b200_index_search(SufrB200Index* index, const uint8_t* queries, const uint64_t* offsets, uint64_t num_queries,
                           int has_max_query_len, uint64_t max_query_len, int use_subsample, uint64_t* rank_begin,
                           uint64_t* rank_end);
int sufr_b200_index_suffixes(SufrB200Index* index, uint64_t rank_begin, uint64_t count, uint64_t* out);
void sufr_b200_index_free(SufrB200Index* index);

/* -- write: replaces SufrBuilder::write (sufr_builder.rs:817-918), version-6 `.sufr` layout.
 *    Single shard: writes the whole file.  Sharded: every rank calls it with the same path; rank 0
 *    writes header, text and the names tail, every rank pwrites its SA / LCP slice at its offset.
 *    `result` must be a HOST result. */
int sufr_b200_write(const SufrB200Args* args, const SufrB200Result* result);

/* -- create: SufrBuilder::new as a single call (build on `device`, then write args->path or
 *    "out.sufr").  The suffix and LCP arrays stream from device memory into the file through a ring
 *    of pinned buffers (copy and pwrite overlap; no host copy of them is allocated), so the result is
 *    what the reference's builder struct holds after `new` (sufr_builder.rs:38-89): counts, the
 *    transformed text, n_ranges -- `sa` and `lcp` are NULL, they are in the file.  timings.d2h_ms is
 *    the wall time of that streaming write.  Free with sufr_b200_result_free(NULL, out).
 *    `out` may be NULL.  Single process only (world_size <= 1): sharded builds go through
 *    sufr_b200_build / sufr_b200_patch_seam / sufr_b200_write. */
int sufr_b200_create(const SufrB200Args* args, int device, SufrB200Result* out);

/* -- create on several GPUs of one box in ONE call (sufr::create is one call too, sufr/src/lib.rs:321-371): a host
 *    thread per device builds key-range shard r of num_devices; the text is uploaded once and replicated over NVLink
 *    (peer copies), seams are repaired from the shards' (count, first, last), and every device streams its slice of
 *    the suffix / LCP arrays into args->path at its offset.  index_bits: 32, 64 or 0 = the reference's dispatch
 *    (suffix_array.rs:460-470).  args->rank / world_size must be 0 / <= 1.  The result is that of sufr_b200_create
 *    (counts over all shards; timings: the slowest shard).  num_devices == 1 is a plain single-GPU create. */
int sufr_b200_create_multi(const SufrB200Args* args, const int* devices, int num_devices, uint32_t index_bits,
                           SufrB200Result* out);

/* -- helpers shared with the host-side mirror ---------------------------------------------- */
/* SeedMask::new (types.rs:80-97): returns the weight, or -1 if the mask is invalid.
 * bytes/positions/differences may be NULL; otherwise they need strlen(mask) entries. */
int64_t sufr_b200_seed_mask(const char* mask, uint8_t* bytes, uint64_t* positions, uint64_t* differences);
/* find_lcp_full_offset (util.rs:19-37); mask == NULL means "not a mask sort". */
uint64_t sufr_b200_find_lcp_full_offset(uint64_t lcp, const char* mask);

/* read_sequence_file (util.rs:51-89): FASTA / FASTQ -> text with delimiters and trailing '$'. */
typedef struct SufrB200Sequences {
    uint8_t* seq;
    uint64_t seq_len;
    uint64_t* start_positions;
    char** sequence_names;
    uint64_t num_sequences;
} SufrB200Sequences;
int sufr_b200_read_sequence_file(const char* path, uint8_t sequence_delimiter, SufrB200Sequences* out);
void sufr_b200_sequences_free(SufrB200Sequences* seqs);

/* Synthetic workload generators used by bench.py (device buffers, deterministic in `seed`). */
int sufr_b200_synth_dna(SufrB200Ctx* ctx, uint8_t* device_text, uint64_t text_len, uint64_t seed,
                        const uint64_t* record_starts, uint64_t num_records, uint8_t delimiter);

/* Message of the last error on the calling thread. */
const char* sufr_b200_last_error(void);
int sufr_b200_abi_version(void);
/* Number of CUDA devices visible (0 when there is none or the driver is missing). */
int sufr_b200_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* SUFR_B200_H */
