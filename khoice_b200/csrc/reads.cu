// reads.cu -- experiment type 6, read level: per-read votes for the confusion matrix.
//
// The reference classifies every simulated read of an out-pivot genome (src/merge_lists.py:149-181, called with -r by rule
// run_merge_list_exp6, /root/reference/workflow/rules/exp_type_6.smk:327-346): for every k-mer of the read, in read order,
//     matches = datasets whose rest-of-set union holds the canonical k-mer;   votes[d] += 1 / len(matches)  for d in matches
// and the read goes to argmax(votes) (ties: random.choice, host side).  The additions are Python floats (IEEE doubles) and
// their ORDER is part of the result, so one thread owns one (read, dataset) pair and adds in window order; 1.0 / len is the
// same correctly rounded double division as Python's.
//   d_index[i]   position of window i's canonical k-mer in the pivot's ascending distinct k-mer list (khb_sorted_lookup of
//                the K2 output against Engine.kmer_counts' keys), UINT64_MAX for windows without a k-mer
//   d_mask       group-membership bit masks of that list (khb_group_membership), mask_words u64 per k-mer
//   d_read_first first window of every read in the symbol stream, d_read_nwin its number of windows (len - k + 1, or 0)
// Output: d_votes[r * n_groups + d] (double), d_unmatched[r] = windows of read r whose k-mer is in no dataset.
#include "khb_common.cuh"

__global__ void __launch_bounds__(128)
read_votes_kernel(const u64 *__restrict__ index, const u64 *__restrict__ mask, int mask_words, int n_groups, const u64 *__restrict__ read_first,
                  const u32 *__restrict__ read_nwin, u64 n_reads, double *__restrict__ votes, u32 *__restrict__ unmatched)
{
    const u64 t = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    const u64 r = t / (u64)n_groups;
    const u32 d = (u32)(t % (u64)n_groups);
    if (r >= n_reads) return;
    const u64 first = read_first[r];
    const u32 nwin = read_nwin[r];
    double v = 0.0;
    u32 none = 0;
    for (u32 j = 0; j < nwin; j++) {
        const u64 ix = index[first + j];
        if (ix == ~0ull) continue;
        const u64 *m = mask + ix * (u64)mask_words;
        u32 len = 0;
        for (int w = 0; w < mask_words; w++) len += (u32)__popcll(m[w]);
        if (len == 0) {
            none++;
            continue;
        }
        if ((m[d >> 6] >> (d & 63u)) & 1ull) v += 1.0 / (double)len;
    }
    votes[r * (u64)n_groups + d] = v;
    if (d == 0) unmatched[r] = none;
}

extern "C" int khb_read_votes(khb_ctx *ctx, const uint64_t *d_index, const uint64_t *d_mask, int mask_words, int n_groups, const uint64_t *d_read_first,
                              const uint32_t *d_read_nwin, uint64_t n_reads, double *d_votes, uint32_t *d_unmatched)
{
    KHB_CHECK_CTX(ctx);
    if (n_groups < 1 || mask_words < 1 || mask_words > 4 || n_groups > 64 * mask_words || !d_votes || !d_unmatched)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_read_votes: bad arguments (n_groups=%d, mask_words=%d)", n_groups, mask_words);
    if (!n_reads) return KHB_OK;
    const u64 threads = n_reads * (u64)n_groups;
    const u64 blocks = div_up(threads, 128);
    if (blocks > 0x7fffffffull) return khb_fail(ctx, KHB_ERR_ARG, "khb_read_votes: %llu reads x %d datasets exceed the grid", (u64)n_reads, n_groups);
    read_votes_kernel<<<(unsigned)blocks, 128, 0, ctx->stream>>>((const u64 *)d_index, (const u64 *)d_mask, mask_words, n_groups, (const u64 *)d_read_first,
                                                               d_read_nwin, n_reads, d_votes, d_unmatched);
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}
