/* khoice_b200.h -- C ABI of libkhoice_b200.so: the B200 (sm_100a) engine behind the k-mer
 * discriminatory-power path of khoice (experiment type 1).
 *
 * The reference has no FFI: its "plugin interface" for this path is a process boundary -- Snakemake rules
 * exec the KMC 3.2.1 command line tools and exchange files
 * (/root/reference/workflow/rules/exp_type_1.smk:156-259).  Each entry point below cites the rule / command
 * line it replaces.  Every function returns KHB_OK (0) or a negative error code; nothing throws across
 * the boundary; khb_last_error() returns the message.  There is NO CPU fallback: without a CUDA device
 * khb_init() fails with KHB_ERR_NODEV.
 *
 * Conventions
 *   - `d_` pointers are device memory owned by the caller (khb_alloc, or any CUDA allocation such as a
 *     torch tensor's data_ptr); they must be 16-byte aligned.  `h_` pointers are host memory.
 *   - all work is enqueued on the context's stream; functions that return host-visible results
 *     synchronise that stream before returning, the others may return early (call khb_sync()).
 *   - k-mer words: k <= 32 -> one uint64 per k-mer; 33 <= k <= 64 -> 16 bytes per k-mer, low word first.
 *     The value of a k-mer is its base-4 number (A=0,C=1,G=2,T=3), first base most significant; the
 *     canonical form is min(k-mer, reverse complement) (/root/reference/src/merge_lists.py:60-73).
 *     The all-ones word is the "no k-mer here" sentinel (never canonical).
 *   - one context per (process, device); calls on one context must be serialised by the caller.
 */
#ifndef KHOICE_B200_H
#define KHOICE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KHB_ABI_VERSION 1
#define KHB_API __attribute__((visibility("default")))

#define KHB_OK 0
#define KHB_ERR_ARG (-1)      /* invalid argument                                   */
#define KHB_ERR_CUDA (-2)     /* CUDA runtime error (sticky for the context)        */
#define KHB_ERR_NOMEM (-3)    /* host or device allocation failed                   */
#define KHB_ERR_NODEV (-4)    /* no usable CUDA device (there is no CPU fallback)   */
#define KHB_ERR_CAPACITY (-5) /* a caller-provided buffer is too small              */
#define KHB_ERR_STATE (-6)    /* call sequence error (e.g. across stage before any group) */

/* FASTA staging granule: every file staged for khb_pack_fasta starts at a multiple of this. */
#define KHB_FASTA_TILE 16384
/* Counter saturation of the union (`-cs5000`, exp_type_1.smk:61,84) and default histogram length. */
#define KHB_COUNTER_MAX 5000

/* Sizes (in words) of the packed arrays for a stream of at most `cap_symbols` symbols. */
static inline size_t khb_codes_words(size_t cap_symbols) { return cap_symbols / 32 + 4; } /* uint64 words */
static inline size_t khb_valid_words(size_t cap_symbols) { return cap_symbols / 32 + 4; } /* uint32 words */
static inline int khb_key_bytes(int k) { return k <= 32 ? 8 : 16; }

typedef struct khb_ctx khb_ctx;

/* ---- context ------------------------------------------------------------------------------------- */
KHB_API int khb_abi_version(void);
KHB_API int khb_init(int device, khb_ctx **ctx);
KHB_API int khb_destroy(khb_ctx *ctx);
KHB_API const char *khb_last_error(const khb_ctx *ctx); /* ctx may be NULL: message of the last failed khb_init */
KHB_API int khb_device_info(khb_ctx *ctx, int *num_sms, size_t *free_bytes, size_t *total_bytes);
KHB_API uint64_t khb_launch_count(const khb_ctx *ctx); /* kernels launched on this context so far */
KHB_API void *khb_stream(khb_ctx *ctx);                /* the cudaStream_t all work is enqueued on */

/* Per-kernel timing for the roofline report: when enabled, every launch of the kernels below is bracketed by
 * CUDA events on the context's stream.  khb_profile_enable() also clears what was recorded so far.
 * khb_profile_read() sums, for one kernel id, the launches, their device time and their ALGORITHMIC bytes
 * (compulsory traffic of the kernel's contract). */
#define KHB_KERNEL_PACK 0       /* fasta_summary + fasta_scan + fasta_pack (one record per khb_pack_fasta) */
#define KHB_KERNEL_EXTRACT 1    /* extract64 / extract128 */
#define KHB_KERNEL_RADIX_HIST 2 /* radix_hist + radix_scan */
#define KHB_KERNEL_ONESWEEP 3   /* onesweep_kernel, one record per digit pass */
#define KHB_KERNEL_UNIQUE 4     /* unique_kernel */
#define KHB_KERNEL_RLE 5        /* rle_hist_kernel */
#define KHB_KERNEL_PARTITION 6  /* partition_kernel (both modes) */
#define KHB_KERNEL_HASH_INSERT 7 /* hash_insert_kernel: K2 fused with the group hash table (hashset.cu) */
#define KHB_KERNEL_HASH_COUNT 8  /* hash_count_kernel: table scan -> histogram + group set */
#define KHB_KERNEL_BIN_PARTITION 9 /* mb_partition_kernel: symbol stream -> super-k-mer records in minimizer bins (bins.cu) */
#define KHB_KERNEL_BIN_COUNT 10    /* mb_count_kernel: per-bin shared-memory counting -> histogram + group set */
#define KHB_KERNEL_BIN_ACROSS 11   /* across-group stage bin by bin: segment events ordered by bin + mb_across_kernel (one record per call) */
KHB_API int khb_profile_enable(khb_ctx *ctx, int on);
KHB_API int khb_profile_read(khb_ctx *ctx, int kernel_id, uint64_t *launches, double *ms, uint64_t *alg_bytes);

/* ---- memory ------------------------------------------------------------------------------------------ */
KHB_API int khb_alloc(khb_ctx *ctx, size_t bytes, void **d_ptr);
KHB_API int khb_free(khb_ctx *ctx, void *d_ptr);
KHB_API int khb_alloc_host(khb_ctx *ctx, size_t bytes, void **h_ptr); /* pinned */
KHB_API int khb_free_host(khb_ctx *ctx, void *h_ptr);
KHB_API int khb_memcpy_h2d(khb_ctx *ctx, void *d_dst, const void *h_src, size_t bytes); /* async on the ctx stream */
KHB_API int khb_memcpy_d2h(khb_ctx *ctx, void *h_dst, const void *d_src, size_t bytes); /* async on the ctx stream */
KHB_API int khb_memcpy_d2d(khb_ctx *ctx, void *d_dst, const void *d_src, size_t bytes); /* async on the ctx stream */
KHB_API int khb_memset(khb_ctx *ctx, void *d_dst, int value, size_t bytes);
KHB_API int khb_sync(khb_ctx *ctx);

/* ---- kernels (device buffers in, device buffers out) ----------------------------------------------- */

/* Lay `n_files` FASTA texts (host memory) out in d_fasta, each at a KHB_FASTA_TILE-aligned offset,
 * separated by "\n>\n..." filler so that no window spans two files.  h_begin[n_files+1] receives the
 * byte offset of every file (h_begin[n_files] = staged size, a multiple of KHB_FASTA_TILE).
 * khb_staged_size() returns that size without copying.  Replaces: kmc reading {genome}.fna.gz
 * (exp_type_1.smk:158,163); inflating .gz is the caller's job. */
KHB_API size_t khb_staged_size(int n_files, const size_t *h_sizes);
KHB_API int khb_stage_fasta(khb_ctx *ctx, int n_files, const uint8_t *const *h_files, const size_t *h_sizes,
                    uint8_t *d_fasta, size_t d_capacity, uint64_t *h_begin);

/* K1: FASTA text -> 2-bit codes + validity bits (`kmc -fm` input stage, exp_type_1.smk:163).
 * nbytes must be a multiple of KHB_FASTA_TILE.  d_codes / d_valid hold khb_codes_words(cap_symbols) /
 * khb_valid_words(cap_symbols) words, cap_symbols >= nbytes.  d_tile_base[nbytes/KHB_FASTA_TILE + 1]
 * receives the symbol offset of every text tile (so the symbol range of a staged file starts at
 * d_tile_base[h_begin[f] / KHB_FASTA_TILE]); d_counts[0] = symbols in the stream, d_counts[1] = break
 * symbols (one per header line; bases = d_counts[0] - d_counts[1]). */
KHB_API int khb_pack_fasta(khb_ctx *ctx, const uint8_t *d_fasta, size_t nbytes, uint64_t *d_codes, uint32_t *d_valid,
                   size_t cap_symbols, uint64_t *d_tile_base, uint64_t *d_counts);

/* K2: canonical k-mer of the window starting at every symbol (`kmc -k{k}`, both strands, exp_type_1.smk:163).
 * d_keys[i] (khb_key_bytes(k) bytes each, n_symbols entries) = canonical k-mer or the sentinel. */
KHB_API int khb_extract_kmers(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, size_t n_symbols, int k,
                      void *d_keys);

/* K2, fused-path flavour: d_keys[i] = h(canonical k-mer), h = the bijective mixer of khb_common.cuh (uniform top
 * bits, sentinel preserved).  khb_remix_keys applies h (inverse = 0) or its inverse (inverse = 1) in place. */
KHB_API int khb_extract_kmers_hashed(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, size_t n_symbols, int k,
                             void *d_keys);
KHB_API int khb_remix_keys(khb_ctx *ctx, void *d_keys, size_t n, int k, int inverse);

/* Prefix plan of the fused path: instead of all ceil(2k/8) digits, hashed keys are sorted by the *npass digits
 * starting at bit *first_bit only (enough prefix bits for segments of at most n_max keys, plus the spare bit above
 * the key so that sentinels never share a prefix with a k-mer); equality inside a prefix run is then resolved by
 * comparison (khb_resolve_*).  first_bit = 0 means a full sort (small k, or k = 32 / 64 which have no spare bit). */
KHB_API int khb_prefix_plan(int k, uint64_t n_max, int *first_bit, int *npass);
KHB_API int khb_sort_key_bits(khb_ctx *ctx, void *d_keys, void *d_tmp, const uint64_t *h_seg_off, int n_segments, int key_bytes,
                      int first_bit, int npass, int *result_in_tmp);
/* K4 / K5 / K6 on prefix-sorted keys (prefix = bits >= prefix_shift; 0 = fully sorted input): distinct keys in
 * order of first occurrence, resp. histogram of multiplicities (+ optional distinct keys).  Every segment of the
 * input must end with at least one sentinel or be the only segment. */
KHB_API int khb_resolve_unique(khb_ctx *ctx, const void *d_sorted, size_t n, int k, int prefix_shift, void *d_out, uint64_t *h_count);
KHB_API int khb_resolve_count(khb_ctx *ctx, const void *d_sorted, size_t n, int k, int prefix_shift, uint32_t cs, uint32_t nbins,
                      uint64_t *h_hist, void *d_out_keys, uint64_t *h_runs);

/* K3: sort every segment [h_seg_off[s], h_seg_off[s+1]) of d_keys independently (LSD radix, ceil(2k/8)
 * passes).  d_tmp is a same-sized ping-pong buffer; *result_in_tmp tells where the result is. */
KHB_API int khb_sort_keys(khb_ctx *ctx, void *d_keys, void *d_tmp, const uint64_t *h_seg_off, int n_segments, int k,
                  int *result_in_tmp);

/* K4: distinct non-sentinel keys of a sorted array (`kmc -ci1` + `kmc_tools transform ... set_counts 1`,
 * exp_type_1.smk:163,173).  d_out may not alias d_sorted.  *h_count = number of keys written. */
KHB_API int khb_unique(khb_ctx *ctx, const void *d_sorted, size_t n, int k, void *d_out, uint64_t *h_count);

/* K5/K6: run lengths of a sorted array = union with counter sum, saturated at `cs`
 * (`kmc_tools complex (set1 + ... + setN)` with -cs5000, exp_type_1.smk:52-61,182 and :75-84,250) and their
 * count-of-counts histogram (`kmc_tools transform ... histogram`, exp_type_1.smk:191,259).
 * h_hist[nbins+1]: h_hist[c] = number of distinct keys whose counter is c (index 0 unused).
 * d_out_keys / d_out_counts (optional, may be NULL) receive the distinct keys / their counters.
 * *h_runs = number of distinct keys. */
KHB_API int khb_count_runs(khb_ctx *ctx, const void *d_sorted, size_t n, int k, uint32_t cs, uint32_t nbins,
                   uint64_t *h_hist, void *d_out_keys, uint32_t *d_out_counts, uint64_t *h_runs);

/* ---- fused stages (what the drop-in rules call) ------------------------------------------------- */

/* Statistics of one fused stage (all counts are pipeline outputs, used for the roofline accounting). */
typedef struct khb_stats {
    uint64_t fasta_bytes;     /* staged text bytes (incl. filler)            */
    uint64_t bases;           /* sequence symbols (valid + invalid)          */
    uint64_t windows;         /* keys produced by K2 (= stream symbols)      */
    uint64_t genome_distinct; /* sum over genomes of distinct k-mers (S_G)   */
    uint64_t distinct;        /* distinct k-mers of the group / across groups */
    float ms_total;           /* device time of the stage, CUDA events        */
    float ms_h2d, ms_pack, ms_extract, ms_sort1, ms_unique, ms_sort2, ms_count;
    int passes_genome;        /* radix passes of the per-genome sort          */
    int passes_group;         /* radix passes of the group / across sort      */
} khb_stats;

/* One species group, from FASTA text in HOST memory to its step_4 histogram: rules
 * build_kmc_database_on_genome, transform_genome_to_set, within_group_union,
 * within_group_union_histogram (exp_type_1.smk:156-191) for one (k, group).  If keep_set != 0 the group's
 * distinct k-mer set (rule build_group_kmer_set, exp_type_1.smk:233-241) is retained on the device for
 * khb_across_groups().  h_hist has nbins+1 entries. */
KHB_API int khb_group_from_fasta(khb_ctx *ctx, int k, int n_genomes, const uint8_t *const *h_files,
                         const size_t *h_sizes, uint32_t nbins, uint64_t *h_hist, int keep_set,
                         khb_stats *stats);

/* Double-buffered ingestion: start copying the text of the NEXT group to the device on a second stream while the
 * current group is being processed.  A following khb_group_from_fasta call with the same file list uses the
 * prefetched copy (it waits for it on the device) instead of copying again. */
KHB_API int khb_group_prefetch_fasta(khb_ctx *ctx, int n_genomes, const uint8_t *const *h_files, const size_t *h_sizes);

/* Same stage with the text already staged in device memory by khb_stage_fasta (h_begin as returned by it). */
KHB_API int khb_group_from_staged(khb_ctx *ctx, int k, int n_genomes, const uint8_t *d_fasta, const uint64_t *h_begin,
                          uint32_t nbins, uint64_t *h_hist, int keep_set, khb_stats *stats);

/* Pack once, sweep k: the reference runs 30 values of k over the same genomes (workflow/Snakefile:36) and lets KMC
 * re-read and re-parse every .fna.gz for each of them (exp_type_1.smk:156-163).  khb_pack_group stages and packs a
 * group ONCE (K1) and keeps the 2-bit stream resident (3/8 byte per base); khb_group_from_packed runs K2..K5 for one k
 * on it.  Results are identical to khb_group_from_fasta. */
typedef struct khb_packed khb_packed;
KHB_API int khb_pack_group(khb_ctx *ctx, int n_genomes, const uint8_t *const *h_files, const size_t *h_sizes, khb_packed **out);
KHB_API int khb_packed_info(const khb_packed *pk, uint64_t *n_symbols, uint64_t *bases, uint64_t *device_bytes);
KHB_API int khb_packed_free(khb_ctx *ctx, khb_packed *pk);
KHB_API int khb_group_from_packed(khb_ctx *ctx, int k, const khb_packed *pk, uint32_t nbins, uint64_t *h_hist, int keep_set,
                          khb_stats *stats);

/* How the group stage (khb_group_from_*) finds the k-mers genomes share.  AUTO: the direct-address table for small k
 * (presence.cu); minimizer bins + per-bin shared-memory counting for 17 <= k <= 31 (bins.cu: no sort, ~2.3 bytes per window
 * through HBM; a group whose bins do not fit falls back to the sort); else one prefix sort of all windows with the genome id
 * as payload.  SINGLE_SORT forces that sort wherever AUTO would take the bins, BINS is AUTO spelled out,
 * TWO_SORT the KMC-shaped chain sort, unique, sort, count.  HASH (k <= 31, <= 448 genomes; other groups take the AUTO
 * route): K2 fused with an open-addressing table of (k-mer, genome bit set) records instead of a sort (hashset.cu) --
 * measured 20 % slower than the single sort on config 2 (profiles/r1s3_hash_group_stage.md), kept as a checked
 * alternative.  All modes give identical results.  The initial mode comes from the environment variable
 * KHB_GROUP_MODE (single-sort | two-sort | hash | bins). */
#define KHB_GROUP_AUTO 0
#define KHB_GROUP_SINGLE_SORT 1
#define KHB_GROUP_TWO_SORT 2
#define KHB_GROUP_HASH 3
#define KHB_GROUP_BINS 4
KHB_API int khb_set_group_mode(khb_ctx *ctx, int mode);
/* Groups the hash path handed back to the sort path because a probe sequence hit its limit (performance counter). */
KHB_API uint64_t khb_hash_overflows(const khb_ctx *ctx);
/* Performance counters of the minimizer-bin path: groups it handed to the sort path (a bin could not be counted in shared
 * memory), bins it redid in hash classes after their table filled up, groups it partitioned a second time with exact
 * region sizes because a bin region overflowed. */
KHB_API void khb_bins_counters(const khb_ctx *ctx, uint64_t *fallbacks, uint64_t *big_bins, uint64_t *repartitions);
/* How khb_across_groups ran so far: bin by bin over the segments the groups' end-of-bin passes left in the store (every retained
 * group came through the minimizer bins with the same number of bins; only with KHB_ACROSS_MODE=bins in the environment of the first
 * call -- measured slower than the sort), or by the prefix sort of the store (default). */
KHB_API void khb_across_counters(const khb_ctx *ctx, uint64_t *by_bins, uint64_t *by_sort);

/* Across-group union-sum + histogram over the retained group sets: rules across_group_union and
 * across_group_union_histogram (exp_type_1.smk:243-259) for one k. */
KHB_API int khb_across_groups(khb_ctx *ctx, uint32_t nbins, uint64_t *h_hist, khb_stats *stats);

/* Retained group sets: count / size, copy out (host), append (from host, e.g. received from a peer or
 * read back from a step_6 file), reset. */
KHB_API int khb_group_sets_info(khb_ctx *ctx, int *k, int *n_groups, uint64_t *n_keys);
KHB_API int khb_group_sets_device(khb_ctx *ctx, void **d_keys, uint64_t *n_keys); /* device view of the retained keys */
KHB_API int khb_group_sets_append_device(khb_ctx *ctx, int k, const void *d_keys, uint64_t n_keys, int n_groups, int hashed);
KHB_API int khb_group_sets_hashed(khb_ctx *ctx);           /* 1 if the retained keys are h(canonical k-mer) */
KHB_API int khb_group_sets_export(khb_ctx *ctx, void *h_out); /* retained keys as canonical k-mer values, to host */
KHB_API int khb_group_sets_append_host(khb_ctx *ctx, int k, const void *h_keys, uint64_t n_keys, int n_groups);
KHB_API int khb_group_sets_reset(khb_ctx *ctx);

/* ---- experiment type 2: pivot analysis (next row N1 of SURVEY.md section 8f) -------------------------------------- */

/* One group of experiment type 2 for one k: `pk` holds the group's rest-of-set genomes followed by the PIVOT genome as
 * its LAST member (khb_pack_group order).  Replaces, for one (k, dataset): build_kmc_database_on_{genome,pivot}_exp_type_2,
 * transform_{genome,pivot}_to_set_exp_type2, within_group_union_exp_type2, pivot_intersect_within_group_exp_type2,
 * pivot_subtract_within_group_exp_type2 and within_group_histogram_exp_type2 (exp_type_2.smk:297-393):
 *   h_hist[c], c = 1..nbins = number of pivot k-mers x with 1 + #{rest genomes containing x} = c
 *   -> h_hist[1] is the size of `kmc_tools simple pivot rest kmers_subtract` (its histogram has that one row),
 *      h_hist[c >= 2] is the histogram of `kmc_tools simple pivot rest intersect -ocsum`.
 * If keep_sets != 0 the rest-of-set union (transform_rest_of_set_to_single_counts, exp_type_2.smk:428-438) and the pivot's
 * k-mer set stay on the device for khb_pivot_across().  Do not mix with khb_group_from_* in one store
 * (khb_group_sets_reset() clears both). */
KHB_API int khb_pivot_group_from_packed(khb_ctx *ctx, int k, const khb_packed *pk, uint32_t nbins, uint64_t *h_hist,
                                int keep_sets, khb_stats *stats);
KHB_API int khb_pivot_sets_info(khb_ctx *ctx, int *n_pivots, uint64_t *n_pivot_keys, uint64_t *n_union_keys);

/* Across groups, for every retained pivot j = 1..G at once: rules across_group_union_for_pivot_exp_type2,
 * pivot_{intersect,subtract}_across_group_exp_type2 and across_group_histogram_exp_type2 (exp_type_2.smk:440-508).
 *   h_hists[(j-1) * (nbins+1) + c] = number of k-mers x of pivot j with 1 + #{groups i != j whose rest-of-set union
 *   contains x} = c   (c = 1: kmers_subtract, c >= 2: intersect -ocsum). */
KHB_API int khb_pivot_across(khb_ctx *ctx, uint32_t nbins, uint64_t *h_hists, khb_stats *stats);

/* Rule-compatible `kmc_tools simple A B intersect|kmers_subtract` (exp_type_2.smk:354-380): d_index[i] = position of
 * d_a[i] in the sorted, duplicate-free d_b, or UINT64_MAX.  Keys are canonical k-mer values (khb_key_bytes(k) each). */
KHB_API int khb_sorted_lookup(khb_ctx *ctx, const void *d_a, uint64_t n_a, const void *d_b, uint64_t n_b, int k,
                      uint64_t *d_index);

/* Experiment type 4 (confusion matrix, exp_type_4.smk:217-270 + src/merge_lists.py:14-33): for every k-mer of
 * n_query_sets query sets (the pivots' k-mers; canonical values, ASCENDING and duplicate-free inside each set, concatenated
 * in d_queries with host offsets h_query_off[n_query_sets+1]) the set of retained group sets that contain it -- what the
 * reference obtains from G x G `kmc_tools simple ... intersect` runs, G x G text dumps and a Python dictionary join.
 * The retained group sets are the groups given to khb_group_from_* with keep_set since the last reset, group g occupying
 * keys [h_group_off[g], h_group_off[g+1]) of the store (khb_group_sets_info after each group).
 * d_mask[(i) * mask_words + w] bit b is set iff group 64 w + b contains query key i.  mask_words >= ceil(n_groups / 64), <= 4. */
KHB_API int khb_group_membership(khb_ctx *ctx, int n_groups, const uint64_t *h_group_off, const void *d_queries, int n_query_sets,
                         const uint64_t *h_query_off, uint64_t *d_mask, int mask_words);

/* Experiment type 6, read level (src/merge_lists.py:149-181 as called by rule run_merge_list_exp6, exp_type_6.smk:327-346):
 * votes of every read for every dataset.  d_index[i] = position of window i's canonical k-mer in the pivot's ascending
 * distinct k-mer list (khb_sorted_lookup of the K2 output against that list; UINT64_MAX = no k-mer), d_mask = that list's
 * membership masks (khb_group_membership), read r owns the windows [d_read_first[r], d_read_first[r] + d_read_nwin[r]).
 * d_votes[r * n_groups + d] = sum over the read's windows, IN WINDOW ORDER, of 1 / len(matches) for the windows whose k-mer
 * is in dataset d (IEEE doubles, the same additions in the same order as the reference's Python floats);
 * d_unmatched[r] = windows whose k-mer is in no dataset.  The argmax with random tie-breaking stays on the host. */
KHB_API int khb_read_votes(khb_ctx *ctx, const uint64_t *d_index, const uint64_t *d_mask, int mask_words, int n_groups,
                   const uint64_t *d_read_first, const uint32_t *d_read_nwin, uint64_t n_reads, double *d_votes,
                   uint32_t *d_unmatched);

/* The first pass of the minimizer-bin group stage (bins.cu) on its own, for tests: partition the k-mer windows of a packed stream
 * (the windows khb_extract_kmers emits) into n_bins bins by minimizer -- the minimum over a window's canonical 13-mers of a 32-bit
 * hash, re-mixed, scaled to [0, n_bins) -- as super-k-mer records: maximal runs of consecutive windows with one minimum, cut at every
 * multiple of KHB_BINS_TILE window starts and into pieces of at most 32 windows.  Region r = bin * ceil(n_genomes / 64) + genome / 64.
 * h_records[r] / h_windows[r]: records and windows of every region (the test suite states the same rules in numpy).
 * 17 <= k <= 63, k != 32.  d_seg_off: symbol offset of every genome, n_genomes + 1 entries, device memory. */
#define KHB_BINS_TILE 4096
KHB_API int khb_bins_partition(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, uint64_t n_symbols,
                       const uint64_t *d_seg_off, int n_genomes, int k, uint32_t n_bins, uint32_t *h_records, uint32_t *h_windows);

/* K7: split n keys into n_parts buckets by splitmix64(key) % n_parts (multi-GPU hash-range partition of the
 * k-mer space).  d_out receives the keys grouped by bucket, h_part_off[n_parts+1] the bucket offsets. */
KHB_API int khb_partition_by_hash(khb_ctx *ctx, const void *d_keys, uint64_t n, int k, int n_parts, void *d_out,
                          uint64_t *h_part_off);

/* ---- multi-GPU exchange over peer memory (csrc/peer.cu; no reference counterpart: the reference is single-host CPU) ----
 * The across-group stage needs every copy of a k-mer on one GPU.  Instead of partition + NCCL all-to-all + copy, every rank
 * owns a receive buffer of `world` regions (region s is written by rank s only), maps its peers' buffers (CUDA IPC over
 * NVLink) and khb_peer_push() -- one kernel behind every group's K5 -- stores each new key of the local group-set store
 * straight into its owner's region.  Protocol per round, on every rank:
 *     khb_group_sets_reset, khb_peer_begin;  { khb_group_from_*(keep_set = 1), khb_peer_push } per group;
 *     khb_peer_counts (waits for the own pushes);  exchange the count table between the ranks (any collective: it is the
 *     barrier);  khb_peer_import(counts sent to me);  khb_across_groups;  all-reduce of the histogram (orders the next round).
 * If *overflow is set on any rank, a region was too small: redo the round over khb_partition_by_hash + NCCL (the local store
 * is untouched) and set the exchange up again with larger regions. */
KHB_API int khb_peer_alloc(khb_ctx *ctx, int world, int rank, int key_bytes, uint64_t region_keys, unsigned char *handle_out /* 64 bytes */);
KHB_API int khb_peer_open(khb_ctx *ctx, const unsigned char *handles /* world x 64 bytes, rank order */);
KHB_API int khb_peer_begin(khb_ctx *ctx);
KHB_API int khb_peer_push(khb_ctx *ctx);
KHB_API int khb_peer_counts(khb_ctx *ctx, uint64_t *h_counts /* [world] */, int *overflow);
KHB_API int khb_peer_import(khb_ctx *ctx, const uint64_t *h_recv_counts /* [world] */, int k, int n_groups, int hashed);
/* khb_peer_import + khb_across_groups without the copy in between: the pushed regions are sorted where they lie (the first radix pass
 * gathers them) and counted; h_hist = the step_8 rows of this rank's hash range.  The retained group sets are forgotten. */
KHB_API int khb_peer_across(khb_ctx *ctx, const uint64_t *h_recv_counts /* [world] */, int k, int n_groups, int hashed, uint32_t nbins,
                    uint64_t *h_hist, khb_stats *stats);
KHB_API int khb_peer_unmap(khb_ctx *ctx); /* drop the peers' mappings; barrier between the ranks; then khb_peer_close frees the own buffer */
KHB_API int khb_peer_close(khb_ctx *ctx);
KHB_API uint64_t khb_peer_region_keys(const khb_ctx *ctx);

/* ---- ONE group on several GPUs (csrc/team.cu, csrc/bins.cu; no reference counterpart) ----------------------------------------
 * Steps 1-4 of one group (/root/reference/workflow/rules/exp_type_1.smk:156-191) sharded over the `team_size` <= 8 members of a
 * team, for job shapes with fewer (or unevenly many) groups than GPUs.  The group's genomes are split into contiguous slices, one per
 * member, every slice padded to whole chunks of 64 genome ids (slice t starts at id 64 * chunk_base[t]); the minimizer bins are split
 * into `team_size` ranges of owners.  Each member packs its slice, partitions it into super-k-mer records (pass P of bins.cu) and
 * copies every region, as one run, into its own area of the record buffer of the bin's owner -- peer memory over NVLink (CUDA IPC),
 * no collective, no remote atomic: a (bin, chunk) region has exactly one writer -- followed by the region's start and size.  After ONE barrier between the members
 * (any collective of the caller; it also carries the overflow flags) every member counts the bins it owns (passes C and B): its
 * h_hist holds the step_4 rows of ITS bins' k-mers, the sum over the members is the group's histogram, and its distinct keys go to
 * the group-set store / the across-group exchange like those of a whole group.
 * Every member must pass the same khb_team_group except chunk_base.  Two receive buffers alternate (parity), so a member may
 * partition the next group while another still counts this one.  Protocol per group, on every member:
 *     khb_team_partition_*(..., h_info);  barrier + exchange h_info[0..1] in the team;  if any member reports h_info[0] != 0 (a region
 *     or an area overflowed) repeat with a larger region_cap / area_pct (every member);  khb_team_count. */
typedef struct khb_team_group {
    int32_t n_genomes_total; /* genomes of the whole group                                                        */
    int32_t n_chunks_total;  /* sum over the members of ceil(slice genomes / 64)                                   */
    int32_t chunk_base;      /* first chunk of THIS member's slice                                                 */
    int32_t parity;          /* 0 / 1: which receive buffer this group uses (alternate from group to group)        */
    uint64_t n_sym_total;    /* symbols of the whole group (an estimate will do: it sizes bins and tables)         */
    double rho;              /* distinct k-mers per window measured on an earlier group of this shape, 0 = unknown */
    uint32_t region_cap;     /* records per (bin, chunk) region of the members' local partitions, 0 = planner default */
    uint32_t area_pct;       /* room of one (sender, owner) area of the receive buffers in percent of the mean share, 0 = 250 */
} khb_team_group;
KHB_API int khb_team_alloc(khb_ctx *ctx, int team_size, int member, uint64_t half_bytes, unsigned char *handle_out /* 64 bytes */);
KHB_API int khb_team_open(khb_ctx *ctx, const unsigned char *handles /* team_size x 64 bytes, member order */);
KHB_API int khb_team_unmap(khb_ctx *ctx); /* drop the members' mappings; barrier in the team; then khb_team_close frees the own buffers */
KHB_API int khb_team_close(khb_ctx *ctx);
/* Geometry the planner derives from (k, tg): bins of the group, records per region, bytes ONE receive buffer must hold. */
KHB_API int khb_team_plan(khb_ctx *ctx, int k, const khb_team_group *tg, uint32_t *n_bins, uint32_t *region_cap, uint64_t *half_bytes);
/* K1 + pass P of this member's slice, then its regions packed into the owners' buffers.  h_info[4]: [0] bit 0: one of this member's
 * regions overflowed (repeat with a larger region_cap), bit 1: one of its areas in an owner's buffer did (repeat with a larger area_pct),
 * [1] records asked for in its fullest region, [2] symbols, [3] bases of the slice.  Returns after the member's stores have completed. */
KHB_API int khb_team_partition_fasta(khb_ctx *ctx, int k, int n_genomes, const uint8_t *const *h_files, const size_t *h_sizes,
                             const khb_team_group *tg, uint64_t *h_info);
KHB_API int khb_team_partition_staged(khb_ctx *ctx, int k, int n_genomes, const uint8_t *d_fasta, const uint64_t *h_begin,
                              const khb_team_group *tg, uint64_t *h_info);
KHB_API int khb_team_partition_packed(khb_ctx *ctx, int k, const khb_packed *pk, const khb_team_group *tg, uint64_t *h_info);
/* Passes C and B over the bins this member owns (after the team's barrier). */
KHB_API int khb_team_count(khb_ctx *ctx, int k, const khb_team_group *tg, uint32_t nbins, uint64_t *h_hist, int keep_set, khb_stats *stats);

#ifdef __cplusplus
}
#endif
#endif /* KHOICE_B200_H */
