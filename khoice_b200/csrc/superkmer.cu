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
    const u32 n_hash = SK_TILE + w - 1;
    for (u32 j = tid; j < n_hash; j += SK_BLOCK) {
        const u64 p = tile0 + j;
        hm[j] = p + m <= n_sym ? sk_mix(sk_canonical(codes, p, m)) : ~0ull;
    }
    __syncthreads();
    const u64 ones_k = k == 64 ? ~0ull : ((1ull << k) - 1ull);
    const u32 base = tid * SK_PER;
    u64 cur = ~0ull;
    u32 cur_at = 0;                                // position (in hm) of the current minimum
    bool have = false;
#pragma unroll 1
    for (u32 t = 0; t < SK_PER; t++) {
        const u32 j = base + t;                    // window j of the tile covers hm[j .. j + w)
        const u64 i = tile0 + j;
        bool ok = i + k <= n_sym;
        if (ok) {
            const u64 q = i >> 5;
            const u32 o = (u32)(i & 31);
            const u64 vv = ((u64)__ldg(valid + q) << 32) | (u64)__ldg(valid + q + 1);
            ok = (((vv << o) >> (64 - k)) == ones_k);   // k validity bits from bit o of the MSB-first stream (k <= 32 here)
        }
        if (!have || cur_at < j) {                 // (re)scan the whole window
            cur = hm[j];
            cur_at = j;
            for (int d = 1; d < w; d++) {
                const u64 h = hm[j + d];
                if (h < cur) { cur = h; cur_at = j + d; }
            }
            have = true;
        } else {
            const u64 h = hm[j + w - 1];           // the one m-mer that entered
            if (h < cur) { cur = h; cur_at = j + w - 1; }
        }
        bins[j] = ok ? (u32)((cur * SK_BIN_C) >> (64 - log2_bins)) : ~0u;
    }
    __syncthreads();
#pragma unroll 1
    for (u32 t = 0; t < SK_PER; t++) {
        const u32 j = base + t;
        const u32 b = bins[j];
        if (b == ~0u) continue;
        if (j > 0 && bins[j - 1] == b) continue;   // not a start (super-k-mers are cut at tile boundaries)
        u32 len = 1;
        while (j + len < SK_TILE && bins[j + len] == b) len++;
        atomicAdd(&bin_superkmers[b], 1u);
        atomicAdd(&bin_windows[b], len);
    }
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
    static bool attr = false;
    if (!attr) {
        cudaFuncSetAttribute(superkmer_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
        attr = true;
    }
    superkmer_count_kernel<<<(unsigned)tiles, SK_BLOCK, shm, ctx->stream>>>((const u64 *)d_codes, d_valid, n_symbols, k, m, log2_bins, d_bin_windows, d_bin_superkmers);
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}
