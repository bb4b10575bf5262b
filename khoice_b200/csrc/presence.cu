// presence.cu -- direct-address path for small k.
//
// For small k the canonical k-mer space is smaller than the data: at k = 7 a group of 50 x 5 Mbp genomes has 2.5e8
// windows but at most 4^7 / 2 = 8192 distinct k-mers, so sorting windows would move every k-mer thirty thousand times.
// Instead the canonical value addresses a presence table directly:
//   table[c][g / 32] bit (g % 32)   <=>   genome g of the group contains k-mer c         (4^k x ceil(N / 32) words)
//   presence_kernel        K2 fused with the table update: canonical k-mer of every window (same bit arithmetic as
//                          extract64_kernel), a plain load tests the bit and only a missing bit costs an atomicOr
//   presence_count_kernel  one thread per table row: c(x) = popcount of the row = the counter of
//                          `kmc_tools complex (set1 + ... + setN)` (exp_type_1.smk:175-182) -> histogram, distinct
//                          k-mers (mixed like the sort path's, so the across-group stage is unchanged), sum of c.
//                          PIVOT (exp_type_2.smk:354-380): rows whose last genome bit is set are the pivot's k-mers.
// Same results as the sort path (tests sweep k across the switch-over); chosen by count_stage() when the table fits.
// Algorithmic bytes: B/4 + B/8 (packed stream) + 2 x 4^k x ceil(N/32) x 4 (table zero-fill + read).
#include "khb_common.cuh"

__device__ __forceinline__ u64 pr_swap_pairs(u64 r)
{
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}

__device__ __forceinline__ u32 pr_segment_of(const u64 *__restrict__ seg_off, int nseg, u64 i)
{
    int lo = 0, hi = nseg;  // invariant: seg_off[lo] <= i < seg_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= i) lo = mid; else hi = mid;
    }
    return (u32)lo;
}

__global__ void __launch_bounds__(256)
presence_kernel(const u64 *__restrict__ codes, const u32 *__restrict__ valid, size_t n_sym, int k, const u64 *__restrict__ seg_off,
                int nseg, u32 *__restrict__ table, int words)
{
    const u64 ones_k = (1ull << k) - 1ull;
    const int rs = 64 - 2 * k;
    // every CTA walks one contiguous chunk (see extract64_kernel): the cached genome interval almost always answers
    const size_t per = ((n_sym + gridDim.x - 1) / gridDim.x + blockDim.x - 1) / blockDim.x * blockDim.x;
    const size_t i_end = (size_t)(blockIdx.x + 1) * per < n_sym ? (size_t)(blockIdx.x + 1) * per : n_sym;
    u64 seg_lo = 1, seg_hi = 0;
    u32 seg_g = 0;
    for (size_t i = (size_t)blockIdx.x * per + threadIdx.x; i < i_end; i += blockDim.x) {
        const size_t m = i >> 5;
        const u32 o = (u32)(i & 31);
        const u64 c0 = __ldg(codes + m), c1 = __ldg(codes + m + 1);
        const u64 vv = ((u64)__ldg(valid + m) << 32) | (u64)__ldg(valid + m + 1);
        if (((vv << o) >> (64 - k)) != ones_k) continue;
        const u64 x = o ? ((c0 << (2 * o)) | (c1 >> (64 - 2 * o))) : c0;
        const u64 fwd = x >> rs;
        const u64 rc = pr_swap_pairs(__brevll(~x) << rs >> rs);
        const u64 can = fwd < rc ? fwd : rc;
        if (i < seg_lo || i >= seg_hi) {
            seg_g = pr_segment_of(seg_off, nseg, i);
            seg_lo = __ldg(seg_off + seg_g);
            seg_hi = __ldg(seg_off + seg_g + 1);
        }
        u32 *w = table + can * (u64)words + (seg_g >> 5);
        const u32 bit = 1u << (seg_g & 31u);
        if (!(*(volatile u32 *)w & bit)) atomicOr(w, bit);
    }
}

template <bool PIVOT>
__global__ void __launch_bounds__(256)
presence_count_kernel(const u32 *__restrict__ table, u64 n_rows, int words, int k, int hashed, u32 cs, u32 nbins, u32 pivot_gid,
                      u64 *__restrict__ hist, u64 *__restrict__ out_keys, u64 *__restrict__ d_cursor, u64 *__restrict__ d_pairs,
                      u64 *__restrict__ out_pivot, u64 *__restrict__ d_pcursor)
{
    extern __shared__ u32 sh_hist[];  // [nbins+1]
    const u32 tid = threadIdx.x, lane = lane_id();
    for (u32 i = tid; i <= nbins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    u32 my_pairs = 0;
    const u64 rows_padded = (n_rows + 31) & ~31ull;  // whole warps stay together for the ballots
    for (u64 r = (u64)blockIdx.x * blockDim.x + tid; r < rows_padded; r += (u64)gridDim.x * blockDim.x) {
        u32 c = 0;
        bool pv = false;
        if (r < n_rows) {
            const u32 *row = table + r * (u64)words;
            for (int w = 0; w < words; w++) c += __popc(row[w]);
            if (PIVOT) pv = (row[pivot_gid >> 5] >> (pivot_gid & 31u)) & 1u;
        }
        const bool emit = c > 0 && !(PIVOT && pv && c == 1);
        if (c > 0) {
            my_pairs += c - (pv ? 1u : 0u);
            if (!PIVOT || pv) {
                const u32 cc = c > cs ? cs : c;
                if (cc <= nbins) atomicAdd(&sh_hist[cc], 1u);
            }
        }
        const u64 key = hashed ? kmer_mix64(r, k) : r;
        const u32 em = __ballot_sync(0xffffffffu, emit);
        if (em) {
            u64 base = 0;
            if (lane == 0) base = atomicAdd(d_cursor, (u64)__popc(em));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (emit && out_keys != nullptr) out_keys[base + __popc(em & lanemask_lt())] = key;
        }
        if (PIVOT) {
            const u32 pm = __ballot_sync(0xffffffffu, pv);
            if (pm) {
                u64 base = 0;
                if (lane == 0) base = atomicAdd(d_pcursor, (u64)__popc(pm));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (pv && out_pivot != nullptr) out_pivot[base + __popc(pm & lanemask_lt())] = key;
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
    if (lane == 0 && my_pairs) atomicAdd(d_pairs, (u64)my_pairs);
    __syncthreads();
    for (u32 i = tid; i <= nbins; i += blockDim.x) {
        const u32 v = sh_hist[i];
        if (v) atomicAdd(&hist[i], (u64)v);
    }
}

// Bytes of the presence table for (k, n_genomes), or 0 if the direct-address path does not apply.
size_t khb_presence_table_bytes(int k, int n_genomes)
{
    if (k < 1 || k > 15 || n_genomes < 1) return 0;
    return ((size_t)1 << (2 * k)) * (size_t)((n_genomes + 31) / 32) * sizeof(u32);
}

// K2..K5 of one group through the presence table.  Outputs as khb_pairs_count_impl.
int khb_presence_count_impl(khb_ctx *ctx, const u64 *d_codes, const u32 *d_valid, size_t n_sym, int k, int hashed, const u64 *d_seg_off,
                            int n_genomes, u32 *d_table, u32 cs, u32 nbins, u64 *d_hist, void *d_out_keys, u64 *d_runs, u64 *d_pairs,
                            int pivot, void *d_out_pivot, u64 *d_pruns)
{
    const int words = (n_genomes + 31) / 32;
    const u64 n_rows = 1ull << (2 * k);
    KHB_CUDA(ctx, cudaMemsetAsync(d_table, 0, n_rows * (size_t)words * sizeof(u32), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_runs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_pairs, 0, sizeof(u64), ctx->stream));
    if (pivot) KHB_CUDA(ctx, cudaMemsetAsync(d_pruns, 0, sizeof(u64), ctx->stream));
    if (n_sym) {
        size_t blocks = div_up(n_sym, 256);
        const size_t cap = (size_t)ctx->num_sms * 32;
        if (blocks > cap) blocks = cap;
        khb_prof_begin(ctx, KHB_K_EXTRACT);
        presence_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(d_codes, d_valid, n_sym, k, d_seg_off, n_genomes, d_table, words);
        KHB_LAUNCH_CHECK(ctx);
        khb_prof_end(ctx, KHB_K_EXTRACT, (u64)n_sym / 4 + n_sym / 8 + n_rows * words * 4);
    }
    u64 grid = div_up(n_rows, 256);
    if (grid > (u64)ctx->num_sms * 8) grid = (u64)ctx->num_sms * 8;
    const size_t shm = ((size_t)nbins + 1) * sizeof(u32);
    khb_prof_begin(ctx, KHB_K_RLE);
    if (pivot)
        presence_count_kernel<true><<<(unsigned)grid, 256, shm, ctx->stream>>>(d_table, n_rows, words, k, hashed, cs, nbins, (u32)(n_genomes - 1), d_hist,
                                                                               (u64 *)d_out_keys, d_runs, d_pairs, (u64 *)d_out_pivot, d_pruns);
    else
        presence_count_kernel<false><<<(unsigned)grid, 256, shm, ctx->stream>>>(d_table, n_rows, words, k, hashed, cs, nbins, 0u, d_hist,
                                                                                (u64 *)d_out_keys, d_runs, d_pairs, nullptr, nullptr);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_RLE, n_rows * words * 4);
    return KHB_OK;
}
