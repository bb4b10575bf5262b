// superkmer.cu -- EXPERIMENTAL first kernel of the design step DESIGN.md section 7 names next (not on any product path yet):
// partition the k-mer windows of a packed symbol stream by minimizer into bins, as super-k-mers.  This file holds the COUNT
// pass: per bin the number of windows and the number of super-k-mers, which sizes the per-bin buffers of the scatter pass.
// Reference statement: the test suite holds it (tests/test_gpu_superkmer.py; same mix / bin functions; windows are the ones K2 emits: rules R1-R5 of SURVEY.md
// 8c, reference call site /root/reference/workflow/rules/exp_type_1.smk:156-163).
//
// A CTA takes a tile of SK_TILE consecutive window starts.  Phase 1: every thread hashes the canonical m-mers of its symbol
// positions (two funnel shifts + BREV like extract64_kernel, then the 64-bit mixer) into shared memory, SK_TILE + k - m of
// them.  Phase 2: every thread walks SK_PER consecutive windows with a running minimum over the k - m + 1 hashes of a window
// (a rescan only when the minimum leaves the window: about one extra shared-memory read per window on random sequence),
// turns the minimum into a bin and stores it (0xFFFF'FFFF for windows without a k-mer).  Phase 3: every window that starts a
// super-k-mer -- first of the tile, or bin / validity differs from its predecessor -- walks to the end of its run and adds
// (1, length) to its bin's counters: one pair of global reductions per super-k-mer (~11 windows), not per window.
#include <stdlib.h>

#include <vector>

#include "khb_common.cuh"

#define SK_BLOCK 256
#define SK_PER 17   // odd: thread chunks start 17 entries apart, so the lanes of a warp spread over the shared-memory banks
#define SK_TILE (SK_BLOCK * SK_PER)
#define SK_MIX_C 0x9E3779B97F4A7C15ull
#define SK_BIN_C 0xD6E8FEB86659FD93ull

__device__ __forceinline__ u64 sk_mix(u64 x)
{
    x = (x ^ (x >> 15)) * SK_MIX_C;
    return x ^ (x >> 29);
}

// canonical value of the m-mer starting at symbol p of the MSB-first 2-bit stream (garbage if a symbol in it is invalid)
__device__ __forceinline__ u64 sk_canonical(const u64 *__restrict__ codes, u64 p, int m)
{
    const u64 w = p >> 5;
    const u32 o = (u32)(p & 31);
    const u64 c0 = __ldg(codes + w), c1 = __ldg(codes + w + 1);
    const u64 x = o ? ((c0 << (2 * o)) | (c1 >> (64 - 2 * o))) : c0;
    const int rs = 64 - 2 * m;
    const u64 fwd = x >> rs;
    u64 r = __brevll(~x) << rs >> rs;
    r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
    return fwd < r ? fwd : r;
}

// Phases 1 and 2 of a tile, shared by the kernels below: hashes of the tile's canonical m-mers into hm[], then the bin of every
// window into bins[] (~0u = no k-mer).  Phase 2 is one window per thread and round (consecutive lanes on consecutive windows,
// k - m + 1 conflict-free shared-memory reads each); the first version walked SK_PER consecutive windows per thread with a
// running minimum -- fewer reads, but one long dependent chain per thread.
__device__ __forceinline__ void sk_tile_bins(const u64 *__restrict__ codes, const u32 *__restrict__ valid, u64 n_sym, int k, int m, int log2_bins,
                                             u64 tile0, u64 *hm, u32 *bins)
{
    const int w = k - m + 1;
    const u32 tid = threadIdx.x;
    const u32 n_hash = SK_TILE + w - 1;
    for (u32 j = tid; j < n_hash; j += SK_BLOCK) {
        const u64 p = tile0 + j;
        hm[j] = p + m <= n_sym ? sk_mix(sk_canonical(codes, p, m)) : ~0ull;
    }
    __syncthreads();
    const u64 ones_k = k == 64 ? ~0ull : ((1ull << k) - 1ull);
    for (u32 j = tid; j < SK_TILE; j += SK_BLOCK) {
        const u64 i = tile0 + j;
        bool ok = i + k <= n_sym;
        if (ok) {
            const u64 q = i >> 5;
            const u32 o = (u32)(i & 31);
            const u64 vv = ((u64)__ldg(valid + q) << 32) | (u64)__ldg(valid + q + 1);
            ok = (((vv << o) >> (64 - k)) == ones_k);   // k validity bits from bit o of the MSB-first stream (k <= 32 here)
        }
        u64 cur = hm[j];
        for (int d = 1; d < w; d++) {
            const u64 h = hm[j + d];
            cur = h < cur ? h : cur;
        }
        bins[j] = ok ? (u32)((cur * SK_BIN_C) >> (64 - log2_bins)) : ~0u;
    }
    __syncthreads();
}

extern "C" int khb_superkmer_count(khb_ctx *, const uint64_t *, const uint32_t *, uint64_t, int, int, int, uint32_t *, uint32_t *);

__global__ void __launch_bounds__(SK_BLOCK)
superkmer_count_kernel(const u64 *__restrict__ codes, const u32 *__restrict__ valid, u64 n_sym, int k, int m, int log2_bins,
                       u32 *__restrict__ bin_windows, u32 *__restrict__ bin_superkmers)
{
    extern __shared__ __align__(8) unsigned char sk_smem[];
    const int w = k - m + 1;                       // m-mers per window
    u64 *hm = (u64 *)sk_smem;                      // [SK_TILE + w - 1] hashes of the m-mers at symbols tile0 ..
    u32 *bins = (u32 *)(hm + SK_TILE + 32);        // [SK_TILE] bin of every window of the tile, ~0u = no k-mer
    const u32 tid = threadIdx.x;
    const u64 tile0 = (u64)blockIdx.x * SK_TILE;
    sk_tile_bins(codes, valid, n_sym, k, m, log2_bins, tile0, hm, bins);
    for (u32 j = tid; j < SK_TILE; j += SK_BLOCK) {   // consecutive lanes on consecutive windows
        const u32 b = bins[j];
        if (b == ~0u) continue;
        if (j > 0 && bins[j - 1] == b) continue;   // not a start (super-k-mers are cut at tile boundaries)
        u32 len = 1;
        while (j + len < SK_TILE && bins[j + len] == b) len++;
        const u32 cap = 65u - (u32)k;              // a record holds k - 1 + len <= 64 symbols: longer runs are several records
        atomicAdd(&bin_superkmers[b], (len + cap - 1) / cap);
        atomicAdd(&bin_windows[b], len);
    }
}


// ---- EXPERIMENT, passes B and C: scatter (canonical k-mer, genome) records into their bins, count every bin in shared memory ----
// Pass B repeats pass A's tile work; the thread that owns the start of a super-k-mer reserves its run's slots in the bin with
// ONE atomicAdd (the bin's region starts at bin_off[b], from an exclusive scan of pass A's window counts) and notes every
// window's destination in shared memory; then ALL threads compute the canonical k-mers, consecutive lanes on consecutive
// windows, and store them (the first version let the start's owner write its ~11 records serially: 12.5 ms per group).
__device__ __forceinline__ u32 sk_segment_of(const u64 *__restrict__ seg_off, int nseg, u64 i)
{
    int lo = 0, hi = nseg;  // invariant: seg_off[lo] <= i < seg_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= i) lo = mid; else hi = mid;
    }
    return (u32)lo;
}

__global__ void __launch_bounds__(SK_BLOCK)
superkmer_scatter_kernel(const u64 *__restrict__ codes, const u32 *__restrict__ valid, u64 n_sym, int k, int m, int log2_bins,
                         const u64 *__restrict__ seg_off, int nseg, const u64 *__restrict__ bin_off, u32 *__restrict__ bin_cursor,
                         u64 *__restrict__ out_keys, unsigned short *__restrict__ out_gid)
{
    extern __shared__ __align__(8) unsigned char sk_smem[];
    const int w = k - m + 1;
    u64 *hm = (u64 *)sk_smem;
    u64 *dest = hm + SK_TILE + 32;                 // [SK_TILE] slot of every window in the output arrays
    u32 *bins = (u32 *)(dest + SK_TILE);
    const u32 tid = threadIdx.x;
    const u64 tile0 = (u64)blockIdx.x * SK_TILE;
    sk_tile_bins(codes, valid, n_sym, k, m, log2_bins, tile0, hm, bins);
    // phase 3: the owner of a super-k-mer's start reserves the run's slots and notes every window's destination in shared memory
    for (u32 j = tid; j < SK_TILE; j += SK_BLOCK) {   // consecutive lanes on consecutive windows
        const u32 b = bins[j];
        if (b == ~0u) continue;
        if (j > 0 && bins[j - 1] == b) continue;
        u32 len = 1;
        while (j + len < SK_TILE && bins[j + len] == b) len++;
        const u64 at = __ldg(bin_off + b) + atomicAdd(&bin_cursor[b], len);
        for (u32 e = 0; e < len; e++) dest[j + e] = at + e;
    }
    __syncthreads();
    // phase 4: all threads, consecutive lanes on consecutive windows: canonical k-mer + genome -> the window's slot
    const u32 g_first = sk_segment_of(seg_off, nseg, tile0 < n_sym ? tile0 : n_sym - 1);
    const u64 g_first_end = __ldg(seg_off + g_first + 1);
    for (u32 j = tid; j < SK_TILE; j += SK_BLOCK) {
        if (bins[j] == ~0u) continue;
        const u64 i = tile0 + j;
        const u64 at = dest[j];
        out_keys[at] = sk_canonical(codes, i, k);
        out_gid[at] = (unsigned short)(i < g_first_end ? g_first : sk_segment_of(seg_off, nseg, i));
    }
}

// Pass C: one CTA per bin.  An open-addressing table of (canonical k-mer, 64 genome bits) in shared memory; every record is
// one probe + one bit update; afterwards the popcount of every slot is c(x) = number of genomes containing x.
#define SKC_BLOCK 256
#define SKC_SLOTS 4096          // keys + bits: 64 KiB
__global__ void __launch_bounds__(SKC_BLOCK)
superkmer_bin_count_kernel(const u64 *__restrict__ keys, const unsigned short *__restrict__ gid, const u64 *__restrict__ bin_off,
                           const u32 *__restrict__ bin_windows, u32 nbins, u64 *__restrict__ hist, u64 *__restrict__ totals /* [0] distinct, [1] pairs, [2] overflowed bins */)
{
    extern __shared__ __align__(8) unsigned char skc_smem[];
    u64 *tkey = (u64 *)skc_smem;                 // [SKC_SLOTS]
    u32 *tbits = (u32 *)(tkey + SKC_SLOTS);      // [SKC_SLOTS][2]
    __shared__ u32 s_over;
    __shared__ u32 s_hist[65];                   // c <= 64 genomes; one flush per CTA instead of one global atomic per k-mer
    const u32 tid = threadIdx.x, b = blockIdx.x;
    const u32 n = bin_windows[b];
    if (n == 0) return;
    for (u32 i = tid; i < SKC_SLOTS; i += SKC_BLOCK) { tkey[i] = ~0ull; tbits[2 * i] = 0u; tbits[2 * i + 1] = 0u; }
    if (tid < 65) s_hist[tid] = 0;
    if (tid == 0) s_over = 0;
    __syncthreads();
    const u64 off = bin_off[b];
    for (u32 i = tid; i < n; i += SKC_BLOCK) {
        const u64 key = keys[off + i];
        const u32 g = gid[off + i];
        u32 slot = (u32)((key * SK_MIX_C) >> 40) & (SKC_SLOTS - 1);
        u32 probes = 0;
        for (;;) {
            const u64 c = *(volatile u64 *)&tkey[slot];
            if (c == key) break;
            if (c == ~0ull) {
                const u64 old = atomicCAS((unsigned long long *)&tkey[slot], ~0ull, key);
                if (old == ~0ull || old == key) break;
            }
            slot = (slot + 1) & (SKC_SLOTS - 1);
            if (++probes >= SKC_SLOTS) { s_over = 1; break; }
        }
        if (probes < SKC_SLOTS) atomicOr(&tbits[2 * slot + (g >> 5)], 1u << (g & 31u));
    }
    __syncthreads();
    u32 my_distinct = 0, my_pairs = 0;
    for (u32 i = tid; i < SKC_SLOTS; i += SKC_BLOCK) {
        if (tkey[i] == ~0ull) continue;
        const u32 c = __popc(tbits[2 * i]) + __popc(tbits[2 * i + 1]);
        my_distinct++;
        my_pairs += c;
        atomicAdd(&s_hist[c], 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_distinct += __shfl_xor_sync(0xffffffffu, my_distinct, o);
        my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
    }
    if ((tid & 31u) == 0) {
        if (my_distinct) atomicAdd((unsigned long long *)&totals[0], (unsigned long long)my_distinct);
        if (my_pairs) atomicAdd((unsigned long long *)&totals[1], (unsigned long long)my_pairs);
    }
    if (tid == 0 && s_over) atomicAdd((unsigned long long *)&totals[2], 1ull);
    __syncthreads();
    if (tid < 65 && s_hist[tid] && tid <= nbins) atomicAdd((unsigned long long *)&hist[tid], (unsigned long long)s_hist[tid]);
}

// EXPERIMENT: the group stage through minimizer bins (passes A, B, C; <= 64 genomes, k <= 32, canonical unmixed keys).
// h_hist[nbins + 1], h_totals[3] = distinct k-mers, sum of per-genome distinct, bins whose table overflowed (then the counts are
// incomplete), h_ms[3] = device time of the three passes.  d_seg_off: device copy of the genomes' symbol offsets [n_genomes + 1].
extern "C" int khb_superkmer_group(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, uint64_t n_symbols, const uint64_t *d_seg_off, int n_genomes,
                                   int k, int m, int log2_bins, uint32_t nbins, uint64_t *h_hist, uint64_t *h_totals, float *h_ms)
{
    KHB_CHECK_CTX(ctx);
    if (n_genomes < 1 || n_genomes > 64 || nbins < 1 || nbins > 8192 || !h_hist || !h_totals || !d_seg_off)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_superkmer_group: 1..64 genomes, 1..8192 histogram rows");
    if (k < 1 || k > 32 || m < 1 || m > k || log2_bins < 1 || log2_bins > 20) return khb_fail(ctx, KHB_ERR_ARG, "khb_superkmer_group: bad k / m / log2_bins");
    const size_t nb = (size_t)1 << log2_bins;
    void *p;
    int rc = khb_scratch_get(ctx, SCR_AUX, nb * (4 + 4 + 4 + 8) + ((size_t)nbins + 1 + 3) * 8 + 256, &p);
    if (rc) return rc;
    u32 *d_win = (u32 *)p, *d_sk = d_win + nb, *d_cur = d_sk + nb;
    u64 *d_off = (u64 *)(d_cur + nb);
    u64 *d_hist = d_off + nb, *d_tot = d_hist + nbins + 1;
    void *pk, *pg;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, (n_symbols + 4) * 8, &pk))) return rc;
    if ((rc = khb_scratch_get(ctx, SCR_PAY_A, (n_symbols + 8) * 2, &pg))) return rc;
    cudaEvent_t ev[4];
    for (int i = 0; i < 4; i++) KHB_CUDA(ctx, cudaEventCreate(&ev[i]));
    KHB_CUDA(ctx, cudaEventRecord(ev[0], ctx->stream));
    if ((rc = khb_superkmer_count(ctx, d_codes, d_valid, n_symbols, k, m, log2_bins, d_win, d_sk))) return rc;
    KHB_CUDA(ctx, cudaEventRecord(ev[1], ctx->stream));
    std::vector<u32> h_win(nb);
    std::vector<u64> h_off(nb);
    KHB_CUDA(ctx, cudaMemcpyAsync(h_win.data(), d_win, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    u64 run = 0;
    for (size_t i = 0; i < nb; i++) { h_off[i] = run; run += h_win[i]; }
    KHB_CUDA(ctx, cudaMemcpyAsync(d_off, h_off.data(), nb * 8, cudaMemcpyHostToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_cur, 0, nb * 4, ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins + 1 + 3) * 8, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // h_off is pageable
    KHB_CUDA(ctx, cudaEventRecord(ev[1], ctx->stream));
    if (n_symbols) {
        const u64 tiles = div_up(n_symbols, SK_TILE);
        const size_t shm = (size_t)(SK_TILE + 32) * sizeof(u64) + (size_t)SK_TILE * (sizeof(u64) + sizeof(u32));
        cudaFuncSetAttribute(superkmer_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        superkmer_scatter_kernel<<<(unsigned)tiles, SK_BLOCK, shm, ctx->stream>>>((const u64 *)d_codes, d_valid, n_symbols, k, m, log2_bins, (const u64 *)d_seg_off, n_genomes,
                                                                                 d_off, d_cur, (u64 *)pk, (unsigned short *)pg);
        KHB_LAUNCH_CHECK(ctx);
    }
    KHB_CUDA(ctx, cudaEventRecord(ev[2], ctx->stream));
    {
        const size_t shm = (size_t)SKC_SLOTS * 16;
        cudaFuncSetAttribute(superkmer_bin_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        superkmer_bin_count_kernel<<<(unsigned)nb, SKC_BLOCK, shm, ctx->stream>>>((const u64 *)pk, (const unsigned short *)pg, d_off, d_win, nbins, d_hist, d_tot);
        KHB_LAUNCH_CHECK(ctx);
    }
    KHB_CUDA(ctx, cudaEventRecord(ev[3], ctx->stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, d_hist, ((size_t)nbins + 1 + 3) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, ctx->h_mail, ((size_t)nbins + 1) * 8);
    memcpy(h_totals, ctx->h_mail + nbins + 1, 3 * 8);
    if (h_ms) {
        float a = 0.f;
        // pass A is timed by the caller-visible khb_superkmer_count as well; here: A incl. the host scan, B, C
        cudaEventElapsedTime(&a, ev[0], ev[1]); h_ms[0] = a;
        cudaEventElapsedTime(&a, ev[1], ev[2]); h_ms[1] = a;
        cudaEventElapsedTime(&a, ev[2], ev[3]); h_ms[2] = a;
    }
    for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
    return KHB_OK;
}

// ---- EXPERIMENT, compact variant: the bins hold SUPER-K-MER RECORDS (24 bytes: genome, length, 64 symbols), ~2.2 bytes per
// window instead of 10; the per-bin counting kernel expands them in registers ----------------------------------------------
__global__ void __launch_bounds__(SK_BLOCK)
superkmer_scatter_compact_kernel(const u64 *__restrict__ codes, const u32 *__restrict__ valid, u64 n_sym, int k, int m, int log2_bins,
                                 const u64 *__restrict__ seg_off, int nseg, const u64 *__restrict__ bin_roff, u32 *__restrict__ bin_cursor,
                                 u64 *__restrict__ rec /* 3 words per record */, u32 bin_cap /* > 0: ONE-pass mode, bin b owns records [b * cap, (b + 1) * cap) */,
                                 u64 *__restrict__ overflow)
{
    extern __shared__ __align__(8) unsigned char sk_smem[];
    const int w = k - m + 1;
    u64 *hm = (u64 *)sk_smem;
    u32 *bins = (u32 *)(hm + SK_TILE + 32);
    const u32 tid = threadIdx.x;
    const u64 tile0 = (u64)blockIdx.x * SK_TILE;
    sk_tile_bins(codes, valid, n_sym, k, m, log2_bins, tile0, hm, bins);
    const u32 cap = 65u - (u32)k;
    for (u32 j = tid; j < SK_TILE; j += SK_BLOCK) {   // consecutive lanes on consecutive windows
        const u32 b = bins[j];
        if (b == ~0u) continue;
        if (j > 0 && bins[j - 1] == b) continue;
        u32 len = 1;
        while (j + len < SK_TILE && bins[j + len] == b) len++;
        const u32 pieces = (len + cap - 1) / cap;
        const u32 at = atomicAdd(&bin_cursor[b], pieces);
        if (bin_cap && at + pieces > bin_cap) {            // the bin's fixed region is full: the caller redoes the group with exact sizes
            *overflow = 1ull;
            continue;
        }
        u64 r = (bin_cap ? (u64)b * bin_cap : __ldg(bin_roff + b)) + at;
        const u64 g = sk_segment_of(seg_off, nseg, tile0 + j);
        for (u32 s0 = 0; s0 < len; s0 += cap, r++) {
            const u32 pl = len - s0 < cap ? len - s0 : cap;
            const u64 i = tile0 + j + s0;                  // first symbol of the piece; it spans k - 1 + pl <= 64 symbols
            const u64 q = i >> 5;
            const u32 o = (u32)(i & 31);
            const u64 c0 = __ldg(codes + q), c1 = __ldg(codes + q + 1), c2 = __ldg(codes + q + 2);
            rec[3 * r] = (g << 48) | ((u64)pl << 40);
            rec[3 * r + 1] = o ? ((c0 << (2 * o)) | (c1 >> (64 - 2 * o))) : c0;
            rec[3 * r + 2] = o ? ((c1 << (2 * o)) | (c2 >> (64 - 2 * o))) : c1;
        }
    }
}

__global__ void __launch_bounds__(SKC_BLOCK)
superkmer_bin_expand_count_kernel(const u64 *__restrict__ rec, const u64 *__restrict__ bin_roff, const u32 *__restrict__ bin_records, int k, u32 nbins,
                                  u64 *__restrict__ hist, u64 *__restrict__ totals, u32 bin_cap)
{
    extern __shared__ __align__(8) unsigned char skc_smem[];
    u64 *tkey = (u64 *)skc_smem;
    u32 *tbits = (u32 *)(tkey + SKC_SLOTS);
    __shared__ u32 s_over;
    __shared__ u32 s_hist[65];
    const u32 tid = threadIdx.x, b = blockIdx.x;
    const u32 n = bin_cap && bin_records[b] > bin_cap ? bin_cap : bin_records[b];
    if (n == 0) return;
    for (u32 i = tid; i < SKC_SLOTS; i += SKC_BLOCK) { tkey[i] = ~0ull; tbits[2 * i] = 0u; tbits[2 * i + 1] = 0u; }
    if (tid < 65) s_hist[tid] = 0;
    if (tid == 0) s_over = 0;
    __syncthreads();
    const u64 off = bin_cap ? (u64)b * bin_cap : bin_roff[b];
    const int rs = 64 - 2 * k;
    // 16 lanes share a record and take its windows e = lane, lane + 16, ...: a record holds ~11 windows on average, so one
    // thread per record (the first version: 5.3 ms per group) leaves most lanes waiting for the longest record of the warp
    const u32 sub = tid & 15u, grp = tid >> 4;
    for (u32 i = grp; i < n; i += SKC_BLOCK / 16) {
        const u64 w0 = rec[3 * (off + i)], w1 = rec[3 * (off + i) + 1], w2 = rec[3 * (off + i) + 2];
        const u32 g = (u32)(w0 >> 48), len = (u32)(w0 >> 40) & 0xffu;
        for (u32 e = sub; e < len; e += 16) {
            const u64 x = e == 0 ? w1 : e < 32 ? ((w1 << (2 * e)) | (w2 >> (64 - 2 * e))) : (w2 << (2 * (e - 32)));
            const u64 fwd = x >> rs;
            u64 r = __brevll(~x) << rs >> rs;
            r = ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
            const u64 key = fwd < r ? fwd : r;
            u32 slot = (u32)((key * SK_MIX_C) >> 40) & (SKC_SLOTS - 1);
            u32 probes = 0;
            for (;;) {
                const u64 c = *(volatile u64 *)&tkey[slot];
                if (c == key) break;
                if (c == ~0ull) {
                    const u64 old = atomicCAS((unsigned long long *)&tkey[slot], ~0ull, key);
                    if (old == ~0ull || old == key) break;
                }
                slot = (slot + 1) & (SKC_SLOTS - 1);
                if (++probes >= SKC_SLOTS) { s_over = 1; break; }
            }
            if (probes < SKC_SLOTS) atomicOr(&tbits[2 * slot + (g >> 5)], 1u << (g & 31u));
        }
    }
    __syncthreads();
    u32 my_distinct = 0, my_pairs = 0;
    for (u32 i = tid; i < SKC_SLOTS; i += SKC_BLOCK) {
        if (tkey[i] == ~0ull) continue;
        const u32 c = __popc(tbits[2 * i]) + __popc(tbits[2 * i + 1]);
        my_distinct++;
        my_pairs += c;
        atomicAdd(&s_hist[c], 1u);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_distinct += __shfl_xor_sync(0xffffffffu, my_distinct, o);
        my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
    }
    if ((tid & 31u) == 0) {
        if (my_distinct) atomicAdd((unsigned long long *)&totals[0], (unsigned long long)my_distinct);
        if (my_pairs) atomicAdd((unsigned long long *)&totals[1], (unsigned long long)my_pairs);
    }
    if (tid == 0 && s_over) atomicAdd((unsigned long long *)&totals[2], 1ull);
    __syncthreads();
    if (tid < 65 && s_hist[tid] && tid <= nbins) atomicAdd((unsigned long long *)&hist[tid], (unsigned long long)s_hist[tid]);
}

// EXPERIMENT: as khb_superkmer_group, with super-k-mer records in the bins instead of expanded k-mers.
extern "C" int khb_superkmer_group_compact(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, uint64_t n_symbols, const uint64_t *d_seg_off,
                                           int n_genomes, int k, int m, int log2_bins, uint32_t nbins, uint64_t *h_hist, uint64_t *h_totals, float *h_ms)
{
    KHB_CHECK_CTX(ctx);
    if (n_genomes < 1 || n_genomes > 64 || nbins < 1 || nbins > 8192 || !h_hist || !h_totals || !d_seg_off)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_superkmer_group_compact: 1..64 genomes, 1..8192 histogram rows");
    if (k < 2 || k > 32 || m < 1 || m > k || log2_bins < 1 || log2_bins > 20) return khb_fail(ctx, KHB_ERR_ARG, "khb_superkmer_group_compact: bad k / m / log2_bins");
    const size_t nb = (size_t)1 << log2_bins;
    void *p;
    int rc = khb_scratch_get(ctx, SCR_AUX, nb * (4 + 4 + 4 + 8) + ((size_t)nbins + 1 + 3) * 8 + 256, &p);
    if (rc) return rc;
    u32 *d_win = (u32 *)p, *d_sk = d_win + nb, *d_cur = d_sk + nb;
    u64 *d_off = (u64 *)(d_cur + nb);
    u64 *d_hist = d_off + nb, *d_tot = d_hist + nbins + 1;
    cudaEvent_t ev[4];
    for (int i = 0; i < 4; i++) KHB_CUDA(ctx, cudaEventCreate(&ev[i]));
    KHB_CUDA(ctx, cudaEventRecord(ev[0], ctx->stream));
    // KHB_SUPERKMER_ONEPASS=1: no count pass -- every bin owns a fixed region of 8 x the mean number of records (+ 64); a bin
    // that would overflow raises totals[2] (the counts are then incomplete and the caller has to fall back to exact sizes)
    const char *one = getenv("KHB_SUPERKMER_ONEPASS");
    const u32 bin_cap = (one && atoi(one)) ? (u32)(8 * (n_symbols / 10 / nb + 1) + 64) : 0u;
    void *pr;
    u32 *d_records = d_sk;
    if (bin_cap) {
        if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, ((u64)nb * bin_cap + 4) * 24, &pr))) return rc;
        d_records = d_cur;                                  // the cursors ARE the record counts
    } else {
        if ((rc = khb_superkmer_count(ctx, d_codes, d_valid, n_symbols, k, m, log2_bins, d_win, d_sk))) return rc;
        std::vector<u32> h_sk(nb);
        std::vector<u64> h_off(nb);
        KHB_CUDA(ctx, cudaMemcpyAsync(h_sk.data(), d_sk, nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        u64 run = 0;
        for (size_t i = 0; i < nb; i++) { h_off[i] = run; run += h_sk[i]; }
        if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, (run + 4) * 24, &pr))) return rc;
        KHB_CUDA(ctx, cudaMemcpyAsync(d_off, h_off.data(), nb * 8, cudaMemcpyHostToDevice, ctx->stream));
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // h_off is pageable
    }
    KHB_CUDA(ctx, cudaMemsetAsync(d_cur, 0, nb * 4, ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins + 1 + 3) * 8, ctx->stream));
    KHB_CUDA(ctx, cudaEventRecord(ev[1], ctx->stream));
    if (n_symbols) {
        const u64 tiles = div_up(n_symbols, SK_TILE);
        const size_t shm = (size_t)(SK_TILE + 32) * sizeof(u64) + (size_t)SK_TILE * sizeof(u32);
        cudaFuncSetAttribute(superkmer_scatter_compact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        superkmer_scatter_compact_kernel<<<(unsigned)tiles, SK_BLOCK, shm, ctx->stream>>>((const u64 *)d_codes, d_valid, n_symbols, k, m, log2_bins, (const u64 *)d_seg_off,
                                                                                         n_genomes, d_off, d_cur, (u64 *)pr, bin_cap, d_tot + 2);
        KHB_LAUNCH_CHECK(ctx);
    }
    KHB_CUDA(ctx, cudaEventRecord(ev[2], ctx->stream));
    {
        const size_t shm = (size_t)SKC_SLOTS * 16;
        cudaFuncSetAttribute(superkmer_bin_expand_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        superkmer_bin_expand_count_kernel<<<(unsigned)nb, SKC_BLOCK, shm, ctx->stream>>>((const u64 *)pr, d_off, d_records, k, nbins, d_hist, d_tot, bin_cap);
        KHB_LAUNCH_CHECK(ctx);
    }
    KHB_CUDA(ctx, cudaEventRecord(ev[3], ctx->stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, d_hist, ((size_t)nbins + 1 + 3) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, ctx->h_mail, ((size_t)nbins + 1) * 8);
    memcpy(h_totals, ctx->h_mail + nbins + 1, 3 * 8);
    if (h_ms) {
        float a = 0.f;
        cudaEventElapsedTime(&a, ev[0], ev[1]); h_ms[0] = a;
        cudaEventElapsedTime(&a, ev[1], ev[2]); h_ms[1] = a;
        cudaEventElapsedTime(&a, ev[2], ev[3]); h_ms[2] = a;
    }
    for (int i = 0; i < 4; i++) cudaEventDestroy(ev[i]);
    return KHB_OK;
}

// Count pass over a packed stream (khb_pack_fasta layout).  d_bin_windows / d_bin_superkmers: u32 [1 << log2_bins], zeroed here.
extern "C" int khb_superkmer_count(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, uint64_t n_symbols, int k, int m, int log2_bins,
                                   uint32_t *d_bin_windows, uint32_t *d_bin_superkmers)
{
    KHB_CHECK_CTX(ctx);
    if (k < 1 || k > 32 || m < 1 || m > k || log2_bins < 1 || log2_bins > 24 || !d_bin_windows || !d_bin_superkmers)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_superkmer_count: need 1 <= m <= k <= 32 and 1 <= log2_bins <= 24");
    const size_t nb = (size_t)1 << log2_bins;
    KHB_CUDA(ctx, cudaMemsetAsync(d_bin_windows, 0, nb * sizeof(u32), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_bin_superkmers, 0, nb * sizeof(u32), ctx->stream));
    if (!n_symbols) return KHB_OK;
    const u64 tiles = div_up(n_symbols, SK_TILE);
    const size_t shm = (size_t)(SK_TILE + 32) * sizeof(u64) + (size_t)SK_TILE * sizeof(u32);
    cudaFuncSetAttribute(superkmer_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    superkmer_count_kernel<<<(unsigned)tiles, SK_BLOCK, shm, ctx->stream>>>((const u64 *)d_codes, d_valid, n_symbols, k, m, log2_bins, d_bin_windows, d_bin_superkmers);
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}
