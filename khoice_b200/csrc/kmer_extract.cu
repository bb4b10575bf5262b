// kmer_extract.cu -- K2: canonical k-mer of every window of the packed symbol stream.
//
// Replaces KMC's splitter / k-mer enumeration (`kmc -k{k}` without -b, reference call site
// /root/reference/workflow/rules/exp_type_1.smk:163; rule R5 of SURVEY.md section 8c): the value of a
// k-mer is its base-4 number with the first base most significant, the canonical form is
// min(k-mer, reverse complement) (the reference's own statement: /root/reference/src/merge_lists.py:60-73).
//
// Because K1 packs MSB-first, the forward k-mer of the window starting at symbol i is the bit range
// [2i, 2i+2k) of the code stream: two funnel shifts, no rolling dependency.  The reverse complement is
// a complement + bit reversal (BREV) + swap inside each 2-bit pair.  A window is valid iff its k
// validity bits are all ones.  Output index = window start, so the kernel is a pure streaming map:
// invalid windows get the all-ones sentinel, which is never a canonical k-mer (T^k's reverse complement
// A^k = 0 is smaller) and therefore sorts behind every real key and is dropped by the unique pass.
//
// Algorithmic bytes: B/4 + B/8 read, W bytes per window written (W = 8 for k <= 32, 16 for k <= 64).
#include "khb_common.cuh"

__device__ __forceinline__ u64 swap_pairs(u64 r)
{
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}

// Index of the segment (genome) that holds window i: largest g with seg_off[g] <= i.  The table is tiny and L1 resident.
__device__ __forceinline__ u32 segment_of(const u64 *__restrict__ seg_off, int nseg, u64 i)
{
    int lo = 0, hi = nseg;  // invariant: seg_off[lo] <= i < seg_off[hi] (seg_off[nseg] = n_sym)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= i) lo = mid; else hi = mid;
    }
    return (u32)lo;
}

// k <= 32.  One thread produces the two windows starting at i and i+1 (i even) -> one 16-byte store.
__global__ void __launch_bounds__(256)
extract64_kernel(const u64 *__restrict__ codes, const u32 *__restrict__ valid, size_t n_sym, int k, int hashed,
                 ulonglong2 *__restrict__ out, unsigned short *__restrict__ gids, const u64 *__restrict__ seg_off, int nseg,
                 u32 *__restrict__ hist /* [hist_npass][256] digit counts of the emitted keys, or null */, int hist_npass, int hist_first_bit)
{
    // hist: the radix sort that follows needs the digit histogram of every pass; counting while the keys are still in
    // registers saves the sort's own histogram sweep (one more read of all keys)
    extern __shared__ u32 sh_hist[];  // [hist_npass][256]
    if (hist != nullptr) {
        for (int i = threadIdx.x; i < hist_npass * 256; i += blockDim.x) sh_hist[i] = 0;
        __syncthreads();
    }
    const size_t npairs = (n_sym + 1) >> 1;
    const u64 ones_k = k == 64 ? ~0ull : ((1ull << k) - 1ull);
    const int rs = 64 - 2 * k;  // right shift that brings the 2k window bits to the low end
    // every CTA walks ONE contiguous chunk, so consecutive iterations stay inside one genome and the cached
    // segment interval [seg_lo, seg_hi) almost always answers the genome-id question without a search
    const size_t per = ((npairs + gridDim.x - 1) / gridDim.x + blockDim.x - 1) / blockDim.x * blockDim.x;
    const size_t pr_end = (size_t)(blockIdx.x + 1) * per < npairs ? (size_t)(blockIdx.x + 1) * per : npairs;
    u64 seg_lo = 1, seg_hi = 0;
    u32 seg_g = 0;
    for (size_t pr = (size_t)blockIdx.x * per + threadIdx.x; pr < pr_end; pr += blockDim.x) {
        const size_t i = pr << 1;
        const size_t m = i >> 5;
        const u32 o = (u32)(i & 31);  // even, <= 30
        const u64 c0 = __ldg(codes + m), c1 = __ldg(codes + m + 1);
        const u64 vv = ((u64)__ldg(valid + m) << 32) | (u64)__ldg(valid + m + 1);
        u64 res[2];
#pragma unroll
        for (int t = 0; t < 2; t++) {
            const u32 ot = o + t;  // <= 31
            const u64 x = ot ? ((c0 << (2 * ot)) | (c1 >> (64 - 2 * ot))) : c0;
            const u64 fwd = x >> rs;
            u64 rc = swap_pairs(__brevll(~x) << rs >> rs);
            // note: ~x has the complemented window in its top 2k bits; brev moves them to the low 2k bits
            // (reversed), the shift pair clears the bits that came from below the window.
            const bool ok = (((vv << ot) >> (64 - k)) == ones_k) && (i + t < n_sym);
            u64 can = fwd < rc ? fwd : rc;
            if (hashed) can = kmer_mix64(can, k);
            res[t] = ok ? can : ~0ull;
        }
        if (i + 1 < n_sym) {
            out[pr] = make_ulonglong2(res[0], res[1]);
        } else {
            ((u64 *)out)[i] = res[0];
        }
        if (hist != nullptr) {
            for (int p = 0; p < hist_npass; p++) {
                const int sh = hist_first_bit + 8 * p;
                atomicAdd(&sh_hist[p * 256 + ((u32)(res[0] >> sh) & 0xffu)], 1u);
                if (i + 1 < n_sym) atomicAdd(&sh_hist[p * 256 + ((u32)(res[1] >> sh) & 0xffu)], 1u);
            }
        }
        if (gids != nullptr) {
            if (i < seg_lo || i >= seg_hi) {
                seg_g = segment_of(seg_off, nseg, i);
                seg_lo = __ldg(seg_off + seg_g);
                seg_hi = __ldg(seg_off + seg_g + 1);
            }
            const u32 g0 = seg_g;
            const u32 g1 = (i + 1 >= seg_hi && i + 1 < n_sym) ? segment_of(seg_off, nseg, i + 1) : g0;
            if (i + 1 < n_sym) ((ushort2 *)gids)[pr] = make_ushort2((unsigned short)g0, (unsigned short)g1);
            else gids[i] = (unsigned short)g0;
        }
    }
    if (hist != nullptr) {
        __syncthreads();
        for (int i = threadIdx.x; i < hist_npass * 256; i += blockDim.x) {
            const u32 c = sh_hist[i];
            if (c) atomicAdd(&hist[i], c);
        }
    }
}

// 33 <= k <= 64.  One thread per window -> one 16-byte store (lo, hi).
__global__ void __launch_bounds__(256)
extract128_kernel(const u64 *__restrict__ codes, const u32 *__restrict__ valid, size_t n_sym, int k, int hashed,
                  ulonglong2 *__restrict__ out, unsigned short *__restrict__ gids, const u64 *__restrict__ seg_off, int nseg,
                  u32 *__restrict__ hist, int hist_npass, int hist_first_bit)
{
    extern __shared__ u32 sh_hist[];  // [hist_npass][256], see extract64_kernel
    if (hist != nullptr) {
        for (int i = threadIdx.x; i < hist_npass * 256; i += blockDim.x) sh_hist[i] = 0;
        __syncthreads();
    }
    const u64 ones_k = k == 64 ? ~0ull : ((1ull << k) - 1ull);
    const int rs = 128 - 2 * k;  // 0..62
    const size_t per = ((n_sym + gridDim.x - 1) / gridDim.x + blockDim.x - 1) / blockDim.x * blockDim.x;
    const size_t i_end = (size_t)(blockIdx.x + 1) * per < n_sym ? (size_t)(blockIdx.x + 1) * per : n_sym;
    u64 seg_lo = 1, seg_hi = 0;
    u32 seg_g = 0;
    for (size_t i = (size_t)blockIdx.x * per + threadIdx.x; i < i_end; i += blockDim.x) {
        const size_t m = i >> 5;
        const u32 o = (u32)(i & 31);
        const u64 c0 = __ldg(codes + m), c1 = __ldg(codes + m + 1), c2 = __ldg(codes + m + 2);
        const u64 va = ((u64)__ldg(valid + m) << 32) | (u64)__ldg(valid + m + 1);
        const u64 vb = ((u64)__ldg(valid + m + 2) << 32) | (u64)__ldg(valid + m + 3);
        // 128-bit window starting at bit 2o of (c0:c1:c2)
        const u64 xh = o ? ((c0 << (2 * o)) | (c1 >> (64 - 2 * o))) : c0;
        const u64 xl = o ? ((c1 << (2 * o)) | (c2 >> (64 - 2 * o))) : c1;
        // forward value = (xh:xl) >> rs
        const u64 fh = xh >> rs;
        const u64 fl = rs ? ((xl >> rs) | (xh << (64 - rs))) : xl;
        // reverse complement: bit-reverse the complemented 128-bit window, keep its low 2k bits, fix pairs
        const u64 nh = ~xh, nl = ~xl;
        const u64 bh = __brevll(nl), bl = __brevll(nh);   // reversed (bh:bl); window bits now lowest
        u64 rh, rl;
        if (rs) {
            // clear the 128-2k top bits that came from below the window
            rh = (bh << rs) >> rs;
            rl = bl;
        } else {
            rh = bh;
            rl = bl;
        }
        rh = swap_pairs(rh);
        rl = swap_pairs(rl);
        const u64 vwin = o ? ((va << o) | (vb >> (64 - o))) : va;
        const bool ok = ((vwin >> (64 - k)) == ones_k);
        const bool f_lt = fh < rh || (fh == rh && fl < rl);
        u64 cl = f_lt ? fl : rl, ch = f_lt ? fh : rh;
        if (hashed) kmer_mix128(ch, cl, k);
        ulonglong2 r;
        r.x = ok ? cl : ~0ull;   // lo
        r.y = ok ? ch : ~0ull;   // hi
        out[i] = r;
        if (hist != nullptr) {
            const Key128 kk{r.x, r.y};
            for (int p = 0; p < hist_npass; p++) atomicAdd(&sh_hist[p * 256 + key_digit(kk, hist_first_bit + 8 * p)], 1u);
        }
        if (gids != nullptr) {
            if (i < seg_lo || i >= seg_hi) {
                seg_g = segment_of(seg_off, nseg, i);
                seg_lo = __ldg(seg_off + seg_g);
                seg_hi = __ldg(seg_off + seg_g + 1);
            }
            gids[i] = (unsigned short)seg_g;
        }
    }
    if (hist != nullptr) {
        __syncthreads();
        for (int i = threadIdx.x; i < hist_npass * 256; i += blockDim.x) {
            const u32 c = sh_hist[i];
            if (c) atomicAdd(&hist[i], c);
        }
    }
}

// keys[i] <- h(keys[i]) or h^-1(keys[i]) in place (sentinels stay sentinels)
__global__ void __launch_bounds__(256) remix64_kernel(u64 *__restrict__ keys, size_t n, int k, int inverse)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const u64 v = keys[i];
        if (v != ~0ull) keys[i] = inverse ? kmer_unmix64(v, k) : kmer_mix64(v, k);
    }
}
__global__ void __launch_bounds__(256) remix128_kernel(ulonglong2 *__restrict__ keys, size_t n, int k, int inverse)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        ulonglong2 v = keys[i];
        if ((v.x & v.y) != ~0ull) {
            u64 hi = v.y, lo = v.x;
            if (inverse) kmer_unmix128(hi, lo, k); else kmer_mix128(hi, lo, k);
            keys[i] = make_ulonglong2(lo, hi);
        }
    }
}

int khb_remix_impl(khb_ctx *ctx, void *d_keys, size_t n, int k, int inverse)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_remix: k=%d outside 1..64", k);
    if (n == 0) return KHB_OK;
    size_t blocks = div_up(n, 256);
    const size_t cap = (size_t)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
    if (k <= 32) remix64_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((u64 *)d_keys, n, k, inverse);
    else remix128_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((ulonglong2 *)d_keys, n, k, inverse);
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}

// d_gids / d_seg_off (optional): also write, per window, the index of the segment (genome) it belongs to.
int khb_extract_kmers_impl(khb_ctx *ctx, const u64 *d_codes, const u32 *d_valid, size_t n_sym, int k, int hashed, void *d_keys,
                           unsigned short *d_gids, const u64 *d_seg_off, int nseg, u32 *d_hist, int hist_npass, int hist_first_bit)
{
    // d_hist (optional, zeroed by the caller): [hist_npass][256] digit counts of the keys written, for the sort that follows
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_extract_kmers: k=%d outside 1..64", k);
    if (n_sym == 0) return KHB_OK;
    const size_t work = k <= 32 ? (n_sym + 1) / 2 : n_sym;
    size_t blocks = div_up(work, 256);
    const size_t cap = (size_t)ctx->num_sms * 32;
    if (blocks > cap) blocks = cap;
    const size_t shm = d_hist ? (size_t)hist_npass * 256 * sizeof(u32) : 0;  // at most 16 passes: 16 KiB
    khb_prof_begin(ctx, KHB_K_EXTRACT);
    if (k <= 32)
        extract64_kernel<<<(unsigned)blocks, 256, shm, ctx->stream>>>(d_codes, d_valid, n_sym, k, hashed, (ulonglong2 *)d_keys, d_gids, d_seg_off, nseg,
                                                                      d_hist, hist_npass, hist_first_bit);
    else
        extract128_kernel<<<(unsigned)blocks, 256, shm, ctx->stream>>>(d_codes, d_valid, n_sym, k, hashed, (ulonglong2 *)d_keys, d_gids, d_seg_off, nseg,
                                                                       d_hist, hist_npass, hist_first_bit);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_EXTRACT, (u64)n_sym / 4 + n_sym / 8 + (u64)n_sym * ((k <= 32 ? 8 : 16) + (d_gids ? 2 : 0)));
    return KHB_OK;
}
