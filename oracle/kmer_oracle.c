/* kmer_oracle.c -- CPU restatement of the khoice exp-type-1 k-mer path.
 *
 * TEST INFRASTRUCTURE, NOT PRODUCT.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this library.  The product (khoice_b200/) never
 * imports, links or executes anything under oracle/.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in the third-party binary KMC 3.2.1
 * (refresh-bio/KMC, bioconda build h9ee0642_0, pinned at
 * /root/reference/workflow/envs/khoice_exps.yaml:97) whose source is NOT under /root/reference and
 * which is not installed in this image; the reference holds no golden vectors for it.  What is
 * restated here is KMC's published behaviour for the flags the reference's call sites use
 * (/root/reference/workflow/rules/exp_type_1.smk:163,173,182,191,241,250,259), anchored on the
 * reference's only in-tree definition of a canonical k-mer (/root/reference/src/merge_lists.py:60-73)
 * and on the inline invariants listed in SURVEY.md section 4.  Each KMC behaviour is a named rule
 * (R1..R9) so that it can be flipped in one place if a real `kmc` disagrees.
 *
 * Build:  make -C oracle      ->  oracle/libkmer_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* R3: valid symbols are ACGTacgt (A=0,C=1,G=2,T=3); any other byte breaks the window. */
static inline int ko_code(uint8_t c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return -1;
    }
}

#define KEY_T uint64_t
#define SUF 64
#include "ko_body.inc"
#undef KEY_T
#undef SUF

#define KEY_T unsigned __int128
#define SUF 128
#include "ko_body.inc"
#undef KEY_T
#undef SUF

#define KO_API __attribute__((visibility("default")))

/* Width in bytes of a k-mer word: 8 for k<=32, 16 for k<=64 (little-endian: lo word first). */
KO_API int ko_key_bytes(int k) { return k <= 32 ? 8 : 16; }

/* Canonical k-mers of every valid window of one FASTA buffer, input order. */
KO_API int64_t ko_kmers(const uint8_t *buf, size_t n, int k, void *out, size_t cap, uint64_t *n_symbols)
{
    if (k < 1 || k > 64) return -2;
    return k <= 32 ? ko_kmers_64(buf, n, k, (uint64_t *)out, cap, n_symbols)
                   : ko_kmers_128(buf, n, k, (unsigned __int128 *)out, cap, n_symbols);
}

KO_API int64_t ko_sort_unique(void *keys, size_t n, int k)
{
    return k <= 32 ? ko_sort_unique_64((uint64_t *)keys, n, k)
                   : ko_sort_unique_128((unsigned __int128 *)keys, n, k);
}

KO_API int64_t ko_union_sum(void *keys, size_t n, int k, uint32_t *counts, uint32_t cs)
{
    return k <= 32 ? ko_union_sum_64((uint64_t *)keys, n, k, counts, cs)
                   : ko_union_sum_128((unsigned __int128 *)keys, n, k, counts, cs);
}

/* R8: `kmc_tools transform X histogram` (exp_type_1.smk:191,259): hist[c] = #k-mers with counter c,
 * c in [1, nbins]; hist[0] unused. */
KO_API void ko_histogram(const uint32_t *counts, size_t n, uint64_t *hist, size_t nbins)
{
    memset(hist, 0, (nbins + 1) * sizeof(uint64_t));
    for (size_t i = 0; i < n; i++)
        if (counts[i] <= nbins) hist[counts[i]]++;
}

/* Whole exp-type-1 arithmetic for one k, used as the timed CPU baseline and by the end-to-end tests.
 *   bufs[i], lens[i] : FASTA text of genome i;  group_of[i] in [0, n_groups)
 *   within_hist      : n_groups x (nbins+1) uint64, row g = step_4 histogram of group g
 *   across_hist      : (nbins+1) uint64 = step_8 histogram
 *   stats[0..4]      : total symbols, total valid k-mers, sum of per-genome distinct (S_G summed),
 *                      sum of per-group distinct (S), overall distinct (D)
 * Parallelism: OpenMP over genomes for the per-genome sets, then every union on all cores by key range (the reference gets
 * its parallelism from snakemake --cores running independent rule instances, exp_type_1.smk:156-191, and from KMC's own
 * threads).  Returns 0 or <0. */
KO_API int ko_exp1(const uint8_t *const *bufs, const size_t *lens, const int32_t *group_of, int n_genomes,
                   int n_groups, int k, uint32_t cs, size_t nbins, uint64_t *within_hist,
                   uint64_t *across_hist, uint64_t *stats)
{
    if (k < 1 || k > 64 || n_genomes < 0 || n_groups < 1) return -2;
    const size_t W = (size_t)ko_key_bytes(k);
    void **gset = (void **)calloc((size_t)n_genomes, sizeof(void *));
    int64_t *gcnt = (int64_t *)calloc((size_t)n_genomes, sizeof(int64_t));
    uint64_t *gsym = (uint64_t *)calloc((size_t)n_genomes, sizeof(uint64_t));
    int64_t *gval = (int64_t *)calloc((size_t)n_genomes, sizeof(int64_t));
    void **grpset = (void **)calloc((size_t)n_groups, sizeof(void *));
    int64_t *grpcnt = (int64_t *)calloc((size_t)n_groups, sizeof(int64_t));
    int err = 0;

    /* step_1 + step_2: per-genome distinct canonical k-mer sets */
#pragma omp parallel for schedule(dynamic, 1)
    for (int i = 0; i < n_genomes; i++) {
        size_t cap = lens[i] + 1;
        void *keys = malloc(cap * W);
        if (!keys) { err = -1; continue; }
        int64_t nk = ko_kmers(bufs[i], lens[i], k, keys, cap, &gsym[i]);
        gval[i] = nk;
        int64_t nu = ko_sort_unique(keys, (size_t)nk, k);
        if (nu < 0) { err = -1; nu = 0; }
        gset[i] = keys;
        gcnt[i] = nu;
    }
    if (err) goto done;

    /* step_3 + step_4 (+ step_6): per-group union-sum, histogram, group set.  One group after the other, every group on all
     * cores (key-range parallel merge): the sampled workloads have few groups, and a loop over groups would leave cores idle */
    for (int g = 0; g < n_groups; g++) {
        size_t tot = 0;
        int m = 0;
        for (int i = 0; i < n_genomes; i++) if (group_of[i] == g) { tot += (size_t)gcnt[i]; m++; }
        uint8_t *cat = (uint8_t *)malloc((tot + 1) * W);
        uint32_t *cnts = (uint32_t *)malloc((tot + 1) * sizeof(uint32_t));
        void **members = (void **)malloc((size_t)(m + 1) * sizeof(void *));
        int64_t *msize = (int64_t *)malloc((size_t)(m + 1) * sizeof(int64_t));
        if (!cat || !cnts || !members || !msize) { err = -1; free(cat); free(cnts); free(members); free(msize); break; }
        m = 0;
        for (int i = 0; i < n_genomes; i++)
            if (group_of[i] == g) { members[m] = gset[i]; msize[m] = gcnt[i]; m++; }
        int64_t nd = k <= 32 ? ko_union_sum_par_64((uint64_t *const *)members, msize, m, k, cs, (uint64_t *)cat, cnts)
                             : ko_union_sum_par_128((unsigned __int128 *const *)members, msize, m, k, cs, (unsigned __int128 *)cat, cnts);
        free(members);
        free(msize);
        if (nd < 0) { err = -1; nd = 0; }
        ko_histogram(cnts, (size_t)nd, within_hist + (size_t)g * (nbins + 1), nbins);
        free(cnts);
        grpset[g] = cat;
        grpcnt[g] = nd;
    }
    if (err) goto done;

    /* step_7 + step_8: across-group union-sum, histogram */
    {
        size_t tot = 0;
        for (int g = 0; g < n_groups; g++) tot += (size_t)grpcnt[g];
        uint8_t *cat = (uint8_t *)malloc((tot + 1) * W);
        uint32_t *cnts = (uint32_t *)malloc((tot + 1) * sizeof(uint32_t));
        if (!cat || !cnts) { err = -1; free(cat); free(cnts); goto done; }
        int64_t nd = k <= 32 ? ko_union_sum_par_64((uint64_t *const *)grpset, grpcnt, n_groups, k, cs, (uint64_t *)cat, cnts)
                             : ko_union_sum_par_128((unsigned __int128 *const *)grpset, grpcnt, n_groups, k, cs, (unsigned __int128 *)cat, cnts);
        if (nd < 0) { err = -1; nd = 0; }
        ko_histogram(cnts, (size_t)nd, across_hist, nbins);
        if (stats) {
            stats[0] = stats[1] = stats[2] = 0;
            for (int i = 0; i < n_genomes; i++) { stats[0] += gsym[i]; stats[1] += (uint64_t)gval[i]; stats[2] += (uint64_t)gcnt[i]; }
            stats[3] = tot;
            stats[4] = (uint64_t)nd;
        }
        free(cat);
        free(cnts);
    }
done:
    for (int i = 0; i < n_genomes; i++) free(gset[i]);
    for (int g = 0; g < n_groups; g++) free(grpset[g]);
    free(gset); free(gcnt); free(gsym); free(gval); free(grpset); free(grpcnt);
    return err;
}

KO_API int ko_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* The host cores the checker may use.  torchrun exports OMP_NUM_THREADS=1 to every rank; bench.py's reference arm runs on
 * rank 0 alone and asks for all cores explicitly. */
KO_API void ko_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
