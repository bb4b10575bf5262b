// bins.cu -- the minimizer-bin group stage (default group path for 17 <= k <= 63, k != 32): what `kmc` + `kmc_tools complex` compute for
// one group (/root/reference/workflow/rules/exp_type_1.smk:156-191: per-genome canonical k-mer sets, their counter-summing union, its
// histogram) WITHOUT a sort.  KMC's own idea, rebuilt for one B200: windows that share a minimizer go to the same bin as compact
// super-k-mer records; a bin is small enough to be counted by ONE CTA in a shared-memory table of (k-mer, genome bits).
//
//   pass P  mb_partition_kernel   packed symbol stream -> regions (one per bin and chunk of 64 genomes) of 24- / 32-byte super-k-mer records
//                                 (tile of 4096 window starts; 32-bit hashes of the canonical 13-mers; sliding minimum over the
//                                 k - 12 hashes of a window by the van Herk / Gil-Werman prefix/suffix trick in shared memory;
//                                 a maximal run of windows with one minimum = one record; ONE global atomicAdd per record)
//   pass C  mb_count_kernel       persistent CTAs stream their bins' records through a double-buffered shared-memory ring with
//                                 cp.async.bulk + mbarrier (the next bin is in flight while this one is counted).  Two levels: every
//                                 record is looked up in an index of the bin's DISTINCT records (content -> genome mask); only the
//                                 distinct records are expanded into canonical k-mers and inserted into an open-addressing table in
//                                 shared memory with their whole genome mask.  End of a bin: popcount per claimed slot -> step_4
//                                 histogram, distinct keys (mixed like K2's) appended to the group-set store, table left clean
//   pass B  mb_bigbin_kernel      (bin, hash class) pairs whose table filled up, redone window by window with the class split further
//   mb_region_scan_kernel         exact region offsets for the second partition attempt after a region overflowed
//   mb_region_scan4_kernel,       ONE group on several GPUs (a team, team.cu; khb_bins_team_*): a member partitions its slice of the genomes
//   mb_region_push_kernel         locally, then one warp per region packs the region into the sender's area of the record buffer of the bin's
//                                 owner (peer memory over NVLink); the owner counts its bins with passes C and B unchanged
//   mb_across_kernel (+ ma_*)     OPT-IN across-group stage bin by bin over the key segments the groups left in the store (exact, but
//                                 slower than the prefix sort: KHB_ACROSS_MODE=bins)
//
// HBM traffic: ~2.4 bytes per window written and read once (records) + 8 / 16 bytes per DISTINCT k-mer -- against ~100 bytes per window
// of the prefix sort.  The bound is shared-memory work and instruction issue, not HBM.
// Every result is exact; a region that overflows makes the caller partition once more with the counted sizes, and only a class of a
// bin that cannot be split small enough for the table sends the group to the sort path.
#include <stdlib.h>

#include "khb_common.cuh"

#define MB_TILE 4096          // window starts per partition CTA
#define MB_BLOCK 256
#define MB_HALO 64            // k - m <= 63 extra hashes behind a tile
#define MB_NH (MB_TILE + MB_HALO)
#define MB_MAXW 32            // windows per record at most: k - 1 + 32 symbols fit KW = 2 words for k <= 32

#define MC_BLOCK 256
#define MC_R 256              // records per stage of the ring (one per thread)
#define MC_PROBES 48
#define MC_WCAP 4096          // windows of distinct records laid out at a time
#define MC_PENDING 0xffffffffu // an index slot whose owner is still writing its record

__device__ __forceinline__ u32 mb_mix32(u32 x)
{
    x *= 0x9E3779B1u;
    x ^= x >> 15;
    x *= 0x85EBCA77u;
    x ^= x >> 13;
    x *= 0xC2B2AE3Du;
    x ^= x >> 16;
    return x | 1u;            // 0 is reserved for "no m-mer here"
}
__device__ __forceinline__ u32 mb_bin_of(u32 minhash, u32 nbins)
{
    u32 b = minhash * 0xD6E8FEB9u;
    b ^= b >> 16;
    b *= 0x7FEB352Du;
    return __umulhi(b, nbins);
}

#define MB_SEGS 512           // genomes whose symbol offsets a partition CTA keeps in shared memory
__device__ __forceinline__ u32 mb_segment_of_shared(const u64 *seg_off, int nseg, u64 i)
{
    int lo = 0, hi = nseg;  // invariant: seg_off[lo] <= i < seg_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (seg_off[mid] <= i) lo = mid; else hi = mid;
    }
    return (u32)lo;
}
__device__ __forceinline__ u32 mb_segment_of(const u64 *__restrict__ seg_off, int nseg, u64 i)
{
    int lo = 0, hi = nseg;  // invariant: seg_off[lo] <= i < seg_off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(seg_off + mid) <= i) lo = mid; else hi = mid;
    }
    return (u32)lo;
}

// Region r of the record buffer: the fixed layout gives every region `cap` records at r * cap; after a region overflowed, the group is
// partitioned again into regions of exactly the sizes the first attempt counted (roff[r] .. roff[r + 1], from mb_region_scan_kernel).
__device__ __forceinline__ u64 mb_region_base(const u64 *__restrict__ roff, u32 r, u32 cap) { return roff ? __ldg(roff + r) : (u64)r * cap; }
// (cap = MB_CAP_DENSE with roff: a team's receive buffer -- the regions lie wherever their senders packed them, roff[r] is the start of
// region r only, and the region sizes the senders delivered are exact)
#define MB_CAP_DENSE 0xFFFFFFFFu
__device__ __forceinline__ u32 mb_region_cap(const u64 *__restrict__ roff, u32 r, u32 cap)
{
    return roff && cap != MB_CAP_DENSE ? (u32)(__ldg(roff + r + 1) - __ldg(roff + r)) : cap;
}

// exclusive prefix sums of the region sizes a first partition attempt counted (rounded up to even: 16-byte aligned regions), one CTA
__global__ void __launch_bounds__(1024)
mb_region_scan_kernel(const u32 *__restrict__ cursor, u64 n_regions, u64 *__restrict__ roff)
{
    __shared__ u64 ws[33];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u64 base = 0; base < n_regions; base += 1024) {
        const u64 i = base + threadIdx.x;
        const u64 v = i < n_regions ? (((u64)cursor[i] + 1ull) & ~1ull) : 0ull;
        u64 total;
        const u64 ex = block_excl_sum<u64>(v, ws, &total);
        if (i < n_regions) roff[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) roff[n_regions] = carry;
}

// the same, four regions per thread and round (a team member packs ~10^5 regions per group: 32 rounds instead of 128)
__global__ void __launch_bounds__(1024)
mb_region_scan4_kernel(const u32 *__restrict__ cursor, u64 n_regions, u64 *__restrict__ roff)
{
    __shared__ u64 ws[33];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u64 base = 0; base < n_regions; base += 4096) {
        const u64 i0 = base + 4ull * threadIdx.x;
        u64 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++) v[j] = i0 + j < n_regions ? (((u64)cursor[i0 + j] + 1ull) & ~1ull) : 0ull;
        u64 total;
        u64 ex = carry + block_excl_sum<u64>(v[0] + v[1] + v[2] + v[3], ws, &total);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i0 + j < n_regions) roff[i0 + j] = ex;
            ex += v[j];
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) roff[n_regions] = carry;
}

// the fullest region of the group (records asked for, whether they fitted or not): sizes the regions of the next group of this shape
__global__ void __launch_bounds__(256)
mb_region_max_kernel(const u32 *__restrict__ cursor, u64 n_regions, u64 *__restrict__ out)
{
    u32 m = 0;
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n_regions; i += (u64)gridDim.x * blockDim.x) m = cursor[i] > m ? cursor[i] : m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const u32 n = __shfl_xor_sync(0xffffffffu, m, o);
        m = n > m ? n : m;
    }
    if ((threadIdx.x & 31u) == 0 && m) atomicMax((unsigned long long *)out, (unsigned long long)m);
}

// ---- pass P -----------------------------------------------------------------------------------------------------------------
// Record layout (KW + 1 words of 64 bits, KW = 2 for k <= 32): word 0 = genome << 48 | windows << 40; words 1..KW = the record's
// k - 1 + windows <= 32 * KW symbols from its first window start, MSB first, zero behind them (window e's k-mer = bits [2e, 2e + 2k)).
// gid_base: added to the genome ids written into the records (a team member's slice of a sharded group starts at id 64 * chunk_base;
// team.cu) -- the regions are still indexed by the slice's own chunks.
template <int KW>
__global__ void __launch_bounds__(MB_BLOCK)
mb_partition_kernel(const u64 *__restrict__ codes, const u32 *__restrict__ valid, u64 n_sym, u64 last_cw, u64 last_vw, int k, int m, u32 nbins,
                    const u64 *__restrict__ seg_off, int nseg, u32 nchunks, u32 *__restrict__ cursor, u64 *__restrict__ rec, u32 cap, u32 capw,
                    const u64 *__restrict__ roff, u64 *__restrict__ flags, u32 gid_base)
{
    constexpr int CW = MB_TILE / 32 + KW + 2;     // 64-bit code words staged per tile
    constexpr int VW = MB_TILE / 32 + 4;
    __shared__ u32 A[MB_NH];                      // hashes -> suffix minima -> per-window minimum
    __shared__ __align__(16) u32 P[MB_NH];        // prefix minima
    __shared__ u32 sc[2 * CW + 2];                // the tile's symbols as 16-symbol chunks in stream order
    __shared__ u32 sv[VW];
    __shared__ u32 mv[VW];                        // bit (31 - i) of mv[q]: the m symbols from position 32 q + i on are all valid
    constexpr u32 WSEG = MB_TILE / (MB_BLOCK / 32);                  // windows per warp in phase 3
    unsigned short (*blist)[WSEG + 2] = (unsigned short (*)[WSEG + 2])P;   // per warp: the boundaries of its windows, in window order (P is free by then)
    __shared__ u32 wcnt[MB_BLOCK / 32], wnext[MB_BLOCK / 32];        // ... how many, and the first boundary behind its windows
    __shared__ u64 sseg[MB_SEGS + 1];             // the genomes' symbol offsets (groups of up to MB_SEGS genomes)
    const u32 tid = threadIdx.x;
    const u64 tile0 = (u64)blockIdx.x * MB_TILE;  // a multiple of 32
    const int w = k - m + 1;                      // m-mers per window
    const u32 B = (w & 1) ? (u32)w : (u32)(w > 1 ? w - 1 : 1);   // block length of the prefix/suffix minima: w or w - 1, whichever is odd
    const u32 NH = MB_TILE + (u32)w - 1u;
    {
        const u64 q0 = tile0 >> 5;
        for (u32 i = tid; i < CW; i += MB_BLOCK) {
            const u64 q = q0 + i < last_cw ? q0 + i : last_cw;
            const u64 c = __ldg(codes + q);
            sc[2 * i] = (u32)(c >> 32);
            sc[2 * i + 1] = (u32)c;
        }
        for (u32 i = tid; i < VW; i += MB_BLOCK) sv[i] = q0 + i <= last_vw ? __ldg(valid + q0 + i) : 0u;
        if (nseg <= MB_SEGS)
            for (u32 i = tid; i <= (u32)nseg; i += MB_BLOCK) sseg[i] = __ldg(seg_off + i);
    }
    __syncthreads();
    // validity of every m-mer start, 32 positions per word: AND of the validity stream with itself shifted by 1, 2, 4, 8 (MSB first)
    for (u32 i = tid; i < VW; i += MB_BLOCK) {
        const u64 x = ((u64)sv[i] << 32) | (i + 1 < VW ? sv[i + 1] : 0u);
        const u64 y1 = x & (x << 1), y2 = y1 & (y1 << 2), y4 = y2 & (y2 << 4);
        u64 res = ~0ull;
        int offb = 0, rem = m;
        if (rem >= 8) { res &= y4 << offb; offb += 8; rem -= 8; }
        if (rem >= 4) { res &= y2 << offb; offb += 4; rem -= 4; }
        if (rem >= 2) { res &= y1 << offb; offb += 2; rem -= 2; }
        if (rem >= 1) { res &= x << offb; }
        mv[i] = (u32)(res >> 32);
    }
    __syncthreads();
    // phase 1: hash of the canonical m-mer at every symbol position of the tile (0 = not an m-mer)
    {
        const u32 sh = 32u - 2u * (u32)m, mask2m = (1u << (2 * m)) - 1u;
        const u64 left = n_sym - tile0;            // symbols from the tile's first to the end of the stream
        const u32 j_end = left >= (u64)NH + (u64)m ? NH : (left >= (u64)m ? (u32)(left - (u64)m) + 1u : 0u);   // m-mers start below j_end
        for (u32 j = tid; j < NH; j += MB_BLOCK) {
            const u32 t = j >> 4, s = (j & 15u) * 2u;
            const u32 x = __funnelshift_l(sc[t + 1], sc[t], s);
            const u32 fwd = x >> sh;
            u32 r = __brev(~x);
            r = ((r >> 1) & 0x55555555u) | ((r & 0x55555555u) << 1);
            r &= mask2m;
            const u32 c = fwd < r ? fwd : r;
            const bool ok = ((mv[j >> 5] << (j & 31u)) >> 31) != 0u && j < j_end;
            A[j] = ok ? mb_mix32(c) : 0u;
        }
    }
    __syncthreads();
    // phase 2a: per block of B hashes the running minimum from the left (P) and from the right (A, in place); lanes are B words
    // apart and B is odd, so the accesses of a warp fall into 32 different banks
    {
        const u32 nblk = (NH + B - 1) / B;
        for (u32 b = tid; b < nblk; b += MB_BLOCK) {
            const u32 lo = b * B, hi = lo + B < NH ? lo + B : NH;
            u32 run = 0xFFFFFFFFu;
            for (u32 p = lo; p < hi; p++) {
                const u32 h = A[p];
                run = h < run ? h : run;
                P[p] = run;
            }
            run = 0xFFFFFFFFu;
            for (u32 p = hi; p-- > lo;) {
                const u32 h = A[p];
                run = h < run ? h : run;
                A[p] = run;
            }
        }
    }
    __syncthreads();
    // phase 2b: minimum over the w hashes of window j = min(suffix minimum at j, prefix minimum at j + w - 1); 0 = no k-mer here
    for (u32 j = tid; j < MB_TILE; j += MB_BLOCK) {
        const u32 a = A[j], p = P[j + w - 1];
        A[j] = a < p ? a : p;
    }
    __syncthreads();
    // phase 3a: every warp lists the boundaries of its own 512 windows -- windows whose minimum differs from their predecessor's -- in
    // window order (one ballot per 32 windows, a running count in a register); segment r of a warp = [blist[r], blist[r + 1]) is a run of
    // equal minima or of non-windows, and the last one ends at the first boundary behind the warp's windows
    const u32 lane = tid & 31u, wrp = tid >> 5;
    {
        u32 n = 0;
#pragma unroll 4
        for (u32 it = 0; it < WSEG / 32; it++) {
            const u32 j = wrp * WSEG + it * 32 + lane;
            const u32 cur = A[j];
            const bool bnd = j == 0 || A[j - 1] != cur;
            const u32 mk = __ballot_sync(0xffffffffu, bnd);
            if (bnd) blist[wrp][n + __popc(mk & lanemask_lt())] = (unsigned short)j;
            n += __popc(mk);
        }
        if (lane == 0) wcnt[wrp] = n;
    }
    __syncthreads();
    if (tid == 0) {
        u32 nx = MB_TILE;
        for (int v = MB_BLOCK / 32 - 1; v >= 0; v--) {
            wnext[v] = nx;
            if (wcnt[v]) nx = blist[v][0];
        }
    }
    __syncthreads();
    // phase 3b: one thread per run: reserve the record slots in the bin with ONE atomicAdd, write the record(s)
    const bool segs_shared = nseg <= MB_SEGS;
    const u32 g_first = segs_shared ? mb_segment_of_shared(sseg, nseg, tile0) : mb_segment_of(seg_off, nseg, tile0);
    const u64 g_first_end = segs_shared ? sseg[g_first + 1] : __ldg(seg_off + g_first + 1);
    const u32 nbnd = wcnt[wrp];
    for (u32 r = lane; r < nbnd; r += 32) {
        const u32 j = blist[wrp][r];
        const u32 mh = A[j];
        if (mh == 0u) continue;
        const u32 len = (r + 1 < nbnd ? (u32)blist[wrp][r + 1] : wnext[wrp]) - j;
        const u64 i0 = tile0 + j;
        const u64 g = i0 < g_first_end ? g_first : segs_shared ? mb_segment_of_shared(sseg, nseg, i0) : mb_segment_of(seg_off, nseg, i0);
        const u32 region = mb_bin_of(mh, nbins) * nchunks + (u32)(g >> 6);   // one region per bin and chunk of 64 genomes
        const u32 pieces = len <= capw ? 1u : (len + capw - 1) / capw;
        const u32 at = atomicAdd(&cursor[region], pieces);
        if (at + pieces > mb_region_cap(roff, region, cap)) {   // the region is full: the caller partitions again with the sizes counted here
            *flags = 1ull;
            continue;
        }
        u64 *dst = rec + (mb_region_base(roff, region, cap) + at) * (KW + 1);
        const u64 gid = g + gid_base;
        for (u32 s0 = 0; s0 < len; s0 += capw) {
            const u32 pl = len - s0 < capw ? len - s0 : capw;
            const u32 rel = j + s0, t = rel >> 4, s = (rel & 15u) * 2u;
            const int nbits = 2 * (k - 1 + (int)pl);   // the record's symbols; what follows is zeroed so that equal runs give equal records
            u64 W[KW + 1];
            W[0] = (gid << 48) | ((u64)pl << 40);
#pragma unroll
            for (int e = 0; e < KW; e++) {
                const u32 hi = __funnelshift_l(sc[t + 2 * e + 1], sc[t + 2 * e], s);
                const u32 lo = __funnelshift_l(sc[t + 2 * e + 2], sc[t + 2 * e + 1], s);
                const int keep = nbits - 64 * e;
                const u64 mask = keep >= 64 ? ~0ull : keep <= 0 ? 0ull : (~0ull << (64 - keep));
                W[e + 1] = (((u64)hi << 32) | lo) & mask;
            }
#pragma unroll
            for (int e = 0; e <= KW; e++) *dst++ = W[e];
        }
    }
}

// ---- pass C -----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64 *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(u64 *bar, u32 bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity)
{
    u32 done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// 1-D bulk copy global -> shared (the TMA unit moves the bytes; completion is counted on the mbarrier).  16-byte aligned, multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}

// One end-of-bin emission into the group-set store: `count` distinct keys of bin `bin` starting at key index `base`.  While every group of a
// store came through the bins with the same number of bins, these segments let the across-group stage count bin by bin (mb_across_kernel).
struct mb_event {
    u32 bin, count, base_lo, base_hi;
};
struct mb_evlog {
    mb_event *buf;      // device
    u64 *count;         // device counter
    u64 cap;            // events that fit
    u64 store_base;     // key index of the group's first key inside the store
};
__device__ __forceinline__ void mb_log_event(const mb_evlog &ev, u32 bin, u32 count, u64 base, u64 *flags)
{
    if (!ev.buf) return;
    const u64 e = atomicAdd((unsigned long long *)ev.count, 1ull);
    if (e >= ev.cap) {
        atomicOr((unsigned long long *)flags, 4ull);   // not an error: the store merely loses its bin-by-bin description
        return;
    }
    const u64 b = ev.store_base + base;
    mb_event o;
    o.bin = bin;
    o.count = count;
    o.base_lo = (u32)b;
    o.base_hi = (u32)(b >> 32);
    ev.buf[e] = o;
}

struct mc_desc {
    u32 bin, count, flags, chunk, cls, ncls;
};
enum { MCF_LAST = 1 /* last stage of a pass over the bin's records */, MCF_SCAN = 2 /* ... of the last genome chunk: emit the table */, MCF_DONE = 4 };

// what thread 0 knows about the stream of stages it feeds into the ring.  A bin is streamed ncls x nchunks times: once per hash
// class of its k-mers (a bin with more distinct k-mers than the table holds is counted class by class) and per chunk of 64 genomes.
struct mc_iter {
    u32 bin, n, n_next, off, chunk, cls, ncls;
};
// ---- k-mers of either width: KW = 2 record words -> one 64-bit word (k <= 32), KW = 3 -> 128 bits (k <= 63) ------------------------
template <int KW> struct mb_kmer;
template <> struct mb_kmer<2> { u64 v; };
template <> struct mb_kmer<3> { u64 hi, lo; };

__device__ __forceinline__ u64 mb_pairswap(u64 r) { return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1); }

// canonical k-mer of window e (< 32) of a record (R[1..KW] = its symbols, MSB first)
__device__ __forceinline__ mb_kmer<2> mb_expand(const u64 *R, u32 e, int k, mb_kmer<2> *)
{
    const u64 w0 = R[1], w1 = R[2];
    const u64 x = e ? ((w0 << (2 * e)) | (w1 >> (64 - 2 * e))) : w0;
    const int rs = 64 - 2 * k;
    const u64 fwd = x >> rs;
    const u64 r = mb_pairswap(__brevll(~x) << rs >> rs);
    mb_kmer<2> o;
    o.v = fwd < r ? fwd : r;
    return o;
}
__device__ __forceinline__ mb_kmer<3> mb_expand(const u64 *R, u32 e, int k, mb_kmer<3> *)
{
    const u64 w0 = R[1], w1 = R[2], w2 = R[3];
    const u64 xh = e ? ((w0 << (2 * e)) | (w1 >> (64 - 2 * e))) : w0;
    const u64 xl = e ? ((w1 << (2 * e)) | (w2 >> (64 - 2 * e))) : w1;
    const int rs = 128 - 2 * k;                    // 2 .. 62 for 33 <= k <= 63
    const u64 fh = xh >> rs, fl = (xl >> rs) | (xh << (64 - rs));
    // reverse complement: all 128 bits reversed and complemented, pairs swapped back; the k-mer's bits end up in the low 2k bits
    const u64 rl = mb_pairswap(__brevll(~xh));
    const u64 rh = mb_pairswap(__brevll(~xl)) & ((1ull << (2 * k - 64)) - 1ull);
    mb_kmer<3> o;
    const bool fwd_less = fh < rh || (fh == rh && fl < rl);
    o.hi = fwd_less ? fh : rh;
    o.lo = fwd_less ? fl : rl;
    return o;
}
__device__ __forceinline__ u64 mb_hash(const mb_kmer<2> &a) { return a.v * 0x9E3779B97F4A7C15ull; }
__device__ __forceinline__ u64 mb_hash(const mb_kmer<3> &a) { return (a.lo ^ (a.hi * 0xBF58476D1CE4E5B9ull)) * 0x9E3779B97F4A7C15ull; }
template <int KW> __device__ __forceinline__ u32 mb_class_hash(const mb_kmer<KW> &a)
{
    const u64 h = mb_hash(a);
    return (u32)((h ^ (h >> 29)) * 0xD6E8FEB86659FD93ull >> 32);
}

#define MB_EMPTY (~0ull)
#define MB_TAG_PENDING 1u

// Find or claim the table slot of `key` (a claimed slot is appended to slots[]); -1 when the probe sequence is too long (table too full).
// 64-bit keys: one compare-and-swap on the key word.  128-bit keys: a 32-bit tag per slot is claimed first (PENDING), the key written,
// the tag published; a thread that meets a pending slot waits for its owner (independent thread scheduling: the owner may sit in the
// same warp) -- the owner publishes without any warp-level synchronisation in between.
// claimed != nullptr: a freshly claimed slot is only reported (*claimed = true) and the caller appends it to slots[] itself, one shared-memory
// atomic per warp instead of one per key (mb_append_claimed) -- every new key of a CTA otherwise hits the same counter.
__device__ __forceinline__ int mb_find_slot(u64 *tkey, u64 *, u32 *, unsigned short *slots, u32 s_log2, const mb_kmer<2> &key, u32 *s_distinct,
                                            bool *claimed = nullptr)
{
    const u32 smask = (1u << s_log2) - 1u;
    u32 slot = (u32)(mb_hash(key) >> 40) & smask;
    for (u32 probes = 0;; probes++) {
        const u64 c = *(volatile u64 *)&tkey[slot];
        if (c == key.v) return (int)slot;
        if (c == MB_EMPTY) {
            const u64 old = atomicCAS((unsigned long long *)&tkey[slot], MB_EMPTY, key.v);
            if (old == MB_EMPTY) {
                if (claimed) *claimed = true;
                else slots[atomicAdd(s_distinct, 1u)] = (unsigned short)slot;   // the occupied slots, densely: the end-of-bin pass never looks at an empty one
                return (int)slot;
            }
            if (old == key.v) return (int)slot;
        }
        if (probes >= MC_PROBES) return -1;
        slot = (slot + 1) & smask;
    }
}
__device__ __forceinline__ int mb_find_slot(u64 *thi, u64 *tlo, u32 *ttag, unsigned short *slots, u32 s_log2, const mb_kmer<3> &key, u32 *s_distinct,
                                            bool *claimed = nullptr)
{
    const u32 smask = (1u << s_log2) - 1u;
    const u64 h = mb_hash(key);
    const u32 mytag = (u32)h | 2u;                 // never 0 (empty) or 1 (pending)
    u32 slot = (u32)(h >> 40) & smask;
    for (u32 probes = 0;; probes++) {
        u32 t = *(volatile u32 *)&ttag[slot];
        if (t == 0u) {
            t = atomicCAS(&ttag[slot], 0u, MB_TAG_PENDING);
            if (t == 0u) {
                thi[slot] = key.hi;
                tlo[slot] = key.lo;
                __threadfence_block();
                *(volatile u32 *)&ttag[slot] = mytag;
                if (claimed) *claimed = true;
                else slots[atomicAdd(s_distinct, 1u)] = (unsigned short)slot;
                return (int)slot;
            }
        }
        while (t == MB_TAG_PENDING) t = *(volatile u32 *)&ttag[slot];
        if (t == mytag && *(volatile u64 *)&thi[slot] == key.hi && *(volatile u64 *)&tlo[slot] == key.lo) return (int)slot;
        if (probes >= MC_PROBES) return -1;
        slot = (slot + 1) & smask;
    }
}
// All 32 lanes of a warp, converged: the lanes that claimed a slot append it to the dense list with ONE atomic on the shared counter.
__device__ __forceinline__ void mb_append_claimed(unsigned short *slots, u32 *s_distinct, bool claimed, int slot)
{
    const u32 m = __ballot_sync(0xffffffffu, claimed);
    if (!m) return;
    const u32 lane = threadIdx.x & 31u, leader = (u32)__ffs(m) - 1u;
    u32 base = 0;
    if (lane == leader) base = atomicAdd(s_distinct, (u32)__popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (claimed) slots[base + __popc(m & ((1u << lane) - 1u))] = (unsigned short)slot;
}
// OR a 64-bit genome mask into a slot's bits (two 32-bit shared-memory atomics, skipped when nothing new)
__device__ __forceinline__ void mb_or_mask(u64 *word, u64 mask)
{
    u32 *bw = (u32 *)word;
    const u32 lo = (u32)mask, hi = (u32)(mask >> 32);
    if (lo && (lo & ~*(volatile u32 *)bw)) atomicOr(bw, lo);
    if (hi && (hi & ~*(volatile u32 *)(bw + 1))) atomicOr(bw + 1, hi);
}

// Shared-memory carve-up of the counting kernels (dynamic): k-mer table (keys, 64 genome bits, counts when there are more genomes), the ring
// of record stages, the distinct records of the bin (content + genome mask) with their hash index,
// per-record window offsets, window -> record map, histogram.
struct mc_geom {
    u32 s_log2;      // k-mer table slots
    u32 dcap;        // distinct records held at a time
    u32 rt_log2;     // slots of the record index (>= 2 * dcap)
    u32 nchunks;     // chunks of 64 genomes
    u32 hrows;       // histogram rows kept in shared memory (+ row 0)
};
struct mc_smem {
    u64 *tkey, *tlo, *tbits, *ring, *dstore, *dmask;   // tkey: the key (KW = 2) or its high word; tlo: low word (KW = 3 only)
    u32 *ttag, *rtab, *hist;                            // ttag: slot tags (KW = 3 only)
    unsigned short *tcnt, *slots, *dwoff, *map;
};
__host__ __device__ inline size_t mc_smem_bytes(int KW, const mc_geom &g, int block)
{
    const size_t S = (size_t)1 << g.s_log2;
    return S * 16 + (KW == 3 ? S * 12 : 0) + 2 * (size_t)MC_R * (KW + 1) * 8 + (size_t)g.dcap * (KW + 1) * 8 + (size_t)g.dcap * 8 + ((size_t)4 << g.rt_log2) +
           (((size_t)g.hrows + 2) & ~(size_t)1) * 4 + (g.nchunks > 1 ? S * 2 : 0) + S * 2 + (size_t)g.dcap * 2 + (size_t)MC_WCAP * 2 + 64;
}
__device__ __forceinline__ mc_smem mc_carve(unsigned char *base, int KW, const mc_geom &g)
{
    const size_t S = (size_t)1 << g.s_log2;
    mc_smem s;
    s.tkey = (u64 *)base;
    s.tlo = s.tkey + S;
    s.tbits = s.tlo + (KW == 3 ? S : 0);
    s.ring = s.tbits + S;
    s.dstore = s.ring + 2 * (size_t)MC_R * (KW + 1);
    s.dmask = s.dstore + (size_t)g.dcap * (KW + 1);
    s.ttag = (u32 *)(s.dmask + g.dcap);
    s.rtab = s.ttag + (KW == 3 ? S : 0);
    s.hist = s.rtab + ((size_t)1 << g.rt_log2);
    s.tcnt = (unsigned short *)(s.hist + (((size_t)g.hrows + 2) & ~(size_t)1));   // [S] when there are several chunks of genomes
    s.slots = s.tcnt + (g.nchunks > 1 ? S : 0);                                   // [S]
    s.dwoff = s.slots + S;                                                         // [dcap]
    s.map = s.dwoff + g.dcap;                                                      // [MC_WCAP]
    return s;
}

template <int KW> __device__ __forceinline__ int mb_slot_of(const mc_smem &sm, u32 s_log2, const mb_kmer<KW> &key, u32 *s_distinct, bool *claimed = nullptr)
{
    return mb_find_slot(sm.tkey, sm.tlo, sm.ttag, sm.slots, s_log2, key, s_distinct, claimed);
}
// Insert `key` with genome bit gb (of the 64 of the current chunk of genomes).  False when the table is too full.
template <int KW> __device__ __forceinline__ bool mb_insert(const mc_smem &sm, u32 s_log2, const mb_kmer<KW> &key, u32 gb, u32 *s_distinct)
{
    const int slot = mb_slot_of<KW>(sm, s_log2, key, s_distinct);
    if (slot < 0) return false;
    u32 *bw = (u32 *)&sm.tbits[slot] + (gb >> 5);
    const u32 bm = 1u << (gb & 31u);
    if (!(*(volatile u32 *)bw & bm)) atomicOr(bw, bm);
    return true;
}
template <int KW> __device__ __forceinline__ void mb_slot_reset(const mc_smem &sm, u32 i)
{
    if (KW == 3) sm.ttag[i] = 0u;
    else sm.tkey[i] = MB_EMPTY;
}
// Multi-GPU, first half: the mixed key of slot i goes back into the slot, its owner and its rank among the bin's keys for that owner into
// the slot's bit word (the bits were counted already); second half (mb_route_store) after the owners' ranges were reserved.
template <int KW> __device__ __forceinline__ u32 mb_route_owner(const mc_smem &sm, u32 i, int k, u32 world)
{
    u32 o;
    if (KW == 3) {
        u64 hi = sm.tkey[i], lo = sm.tlo[i];
        kmer_mix128(hi, lo, k);
        sm.tkey[i] = hi;
        sm.tlo[i] = lo;
        Key128 kk;
        kk.lo = lo;
        kk.hi = hi;
        o = part_of(kk, world);
    } else {
        Key64 kk;
        kk.v = kmer_mix64(sm.tkey[i], k);
        sm.tkey[i] = kk.v;
        o = part_of(kk, world);
    }
    return o;
}
template <int KW> __device__ __forceinline__ void mb_route_note(const mc_smem &sm, u32 i, int k, u32 world, u32 *s_ocnt)
{
    const u32 o = mb_route_owner<KW>(sm, i, k, world);
    const u32 r = atomicAdd(&s_ocnt[o], 1u);
    sm.tbits[i] = ((u64)o << 32) | r;
}
template <int KW> __device__ __forceinline__ void mb_route_store(const mc_smem &sm, u32 i, const khb_peer_route &route, const u64 *s_obase, void *out, u64 at)
{
    const u64 w = sm.tbits[i];
    const u32 o = (u32)(w >> 32), r = (u32)w;
    const u64 pos = s_obase[o] + r;
    if (KW == 3) {
        Key128 kk;
        kk.lo = sm.tlo[i];
        kk.hi = sm.tkey[i];
        if (pos < route.cap) ((Key128 *)route.dst[o])[pos] = kk;
        if (out) ((Key128 *)out)[at] = kk;
    } else {
        const u64 v = sm.tkey[i];
        if (pos < route.cap) ((u64 *)route.dst[o])[pos] = v;
        if (out) ((u64 *)out)[at] = v;
    }
    sm.tbits[i] = 0ull;
}
// between the halves: one reservation per owner on the sender-side cursor (no remote atomics), and the owners' first positions in the
// bin's owner-sorted order (warp 0: exclusive scan of the <= 64 counts) -- the second half writes every owner's keys as ONE contiguous
// run of consecutive threads, which is what NVLink wants (scattered 8-byte stores reached 334 GB/s on 8 B200s, runs three times that)
__device__ __forceinline__ void mb_route_reserve(const khb_peer_route &route, u32 *s_ocnt, u64 *s_obase, u32 *s_ostart)
{
    const u32 o = threadIdx.x;
    if (o < 32) {
        const u32 c0 = o < route.world ? s_ocnt[o] : 0u, c1 = o + 32 < route.world ? s_ocnt[o + 32] : 0u;
        const u32 i0 = warp_incl_sum(c0);
        const u32 t0 = __shfl_sync(0xffffffffu, i0, 31);
        const u32 i1 = warp_incl_sum(c1);
        s_ostart[o] = i0 - c0;
        s_ostart[o + 32] = t0 + i1 - c1;
    }
    if (o < route.world) {
        const u32 c = s_ocnt[o];
        s_obase[o] = c ? atomicAdd((unsigned long long *)&route.cursor[o], (unsigned long long)c) : 0ull;
        if (c && s_obase[o] + c > route.cap) route.cursor[64] = 1ull;   // region full: the caller redoes the round over NCCL
    }
}

// the k-mer of slot i, mixed like K2's keys (kmer_mix64 / kmer_mix128), to position `at` of the group-set store
template <int KW> __device__ __forceinline__ void mb_emit(const mc_smem &sm, u32 i, void *out, u64 at, int k)
{
    if (KW == 3) {
        u64 hi = sm.tkey[i], lo = sm.tlo[i];
        kmer_mix128(hi, lo, k);
        Key128 o;
        o.lo = lo;
        o.hi = hi;
        ((Key128 *)out)[at] = o;
    } else {
        ((u64 *)out)[at] = kmer_mix64(sm.tkey[i], k);
    }
}

__device__ __forceinline__ u64 mb_record_hash(const u64 *w, int KW, u32 len)
{
    u64 h = w[0] * 0x9E3779B97F4A7C15ull;
    h ^= h >> 32;
    for (int i = 1; i < KW; i++) {
        h += w[i];
        h *= 0xBF58476D1CE4E5B9ull;
        h ^= h >> 29;
    }
    h += len;
    h *= 0x94D049BB133111EBull;
    return h ^ (h >> 32);
}

// Genomes of one group are near copies of each other, so most super-k-mer records of a bin are byte-identical across genomes.  The
// kernel therefore counts in two levels: every record is looked up in an index of the bin's DISTINCT records (content + genome mask:
// one hash, one compare, one bit per record), and only the distinct records are expanded into k-mers, each k-mer taking the whole
// genome mask of its record.  Unrelated genomes lose nothing but the lookup.
template <int KW, bool MULTI, int BLOCK>
__global__ void __launch_bounds__(BLOCK)
mb_count_kernel(const u64 *__restrict__ rec, const u32 *__restrict__ cursor, u32 nbins, u32 cap, int k, mc_geom geo, u32 n_genomes, u32 cs,
                u32 thr1 /* records one pass over a bin may hold */, u32 over_cap, u64 *__restrict__ hist, void *__restrict__ out_keys,
                u64 *__restrict__ d_cursor, u64 *__restrict__ d_pairs, u32 *__restrict__ over_list, u32 *__restrict__ over_count, u64 *__restrict__ d_stat,
                const u64 *__restrict__ roff, mb_evlog evlog, khb_peer_route route)
{
    extern __shared__ __align__(16) unsigned char mc_raw[];
    __shared__ __align__(8) u64 bars[2];
    __shared__ mc_desc desc[2];
    __shared__ u32 s_over, s_distinct, s_dcount, s_wtotal, s_dfull;
    __shared__ u32 s_cn[MULTI ? 128 : 1];   // thread 0: records per chunk of the bin it is feeding
    __shared__ u32 s_ocnt[64], s_ostart[64]; // multi-GPU: keys of the bin per owner, their first position in owner-sorted order ...
    __shared__ u64 s_obase[64];             // ... and where they go in the owner's region
    __shared__ u64 s_base;
    const u32 s_log2 = geo.s_log2, S = 1u << s_log2, nchunks = geo.nchunks, hrows = geo.hrows, dcap = geo.dcap, RT = 1u << geo.rt_log2;
    const mc_smem sm = mc_carve(mc_raw, KW, geo);
    const u32 tid = threadIdx.x;
    if (d_stat[0] & 1ull) return;   // a bin region overflowed: its slots hold stale bytes, and the caller redoes the group anyway
    for (u32 i = tid; i < S; i += BLOCK) {
        mb_slot_reset<KW>(sm, i);
        sm.tbits[i] = 0ull;
        if (MULTI) sm.tcnt[i] = 0u;
    }
    for (u32 i = tid; i < RT; i += BLOCK) sm.rtab[i] = 0u;
    for (u32 i = tid; i <= hrows; i += BLOCK) sm.hist[i] = 0u;
    if (tid == 0) {
        mbar_init(&bars[0], 1);
        mbar_init(&bars[1], 1);
        s_over = 0;
        s_distinct = 0;
        s_dcount = 0;
        s_wtotal = 0;
        s_dfull = 0;
        for (int o = 0; o < 64; o++) s_ocnt[o] = 0u;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // The stream of stages.  Region u = bin * nchunks + chunk holds the records of the bin that come from chunk `chunk` of 64 genomes; a bin
    // is streamed class by class (hash classes of its k-mers, when the table cannot hold them all) and, inside a class, chunk by chunk.
    mc_iter it;
    it.bin = blockIdx.x;
    it.off = 0;
    it.cls = 0;
    it.ncls = 1;
    it.chunk = 0;
    it.n = 0;        // records of the current region
    it.n_next = 0;   // one chunk: records of this CTA's next bin, loaded one bin ahead
    u32 it_last = 0; // several chunks: the last non-empty chunk of the current bin
    u64 my_records = 0;
    auto region_records = [&](u32 u, u32 lim) -> u32 {
        if (u >= lim) return 0u;
        const u32 c = __ldg(cursor + u);
        const u32 rcap = mb_region_cap(roff, u, cap);
        return c < rcap ? c : rcap;
    };
    auto open_bin = [&]() {   // counts of the bin it.bin (several chunks: into s_cn); false when the bin is empty
        if (!MULTI) return it.n != 0;
        u32 total = 0;
        for (u32 c = 0; c < nchunks; c++) {
            const u32 n = region_records(it.bin * nchunks + c, nbins * nchunks);
            s_cn[c] = n;
            total += n;
            if (n) it_last = c;
        }
        if (!total) return false;
        it.chunk = 0;
        while (s_cn[it.chunk] == 0) it.chunk++;
        it.n = s_cn[it.chunk];
        it.ncls = total <= thr1 ? 1u : (total + thr1 - 1) / thr1;
        my_records += total;
        return true;
    };
    auto next_bin = [&]() {
        for (;;) {
            it.bin += gridDim.x;
            if (!MULTI) {
                it.n = it.n_next;
                it.n_next = region_records(it.bin + gridDim.x, nbins);
            }
            if (it.bin >= nbins) return;
            if (open_bin()) break;
        }
        if (!MULTI) {
            it.ncls = it.n <= thr1 ? 1u : (it.n + thr1 - 1) / thr1;
            my_records += it.n;
        }
    };
    auto produce = [&](u32 buf) {   // thread 0 only: the next stage of this CTA's stream goes into ring buffer `buf`
        if (it.bin >= nbins) {
            desc[buf].flags = MCF_DONE;
            mbar_arrive(&bars[buf]);
            return;
        }
        const u32 count = it.n - it.off < MC_R ? it.n - it.off : MC_R;
        const bool last = it.off + count == it.n;
        const bool last_chunk = !MULTI || it.chunk == it_last;
        desc[buf].bin = it.bin;
        desc[buf].count = count;
        desc[buf].chunk = it.chunk;
        desc[buf].cls = it.cls;
        desc[buf].ncls = it.ncls;
        desc[buf].flags = (last ? MCF_LAST : 0u) | (last && last_chunk ? MCF_SCAN : 0u);
        const u32 bytes = (count * (u32)((KW + 1) * 8) + 15u) & ~15u;   // an odd count of 24-byte records: 8 bytes of the next record ride along
        mbar_arrive_expect_tx(&bars[buf], bytes);
        const u64 region = MULTI ? (u64)it.bin * nchunks + it.chunk : (u64)it.bin;
        bulk_g2s(sm.ring + (size_t)buf * MC_R * (KW + 1), rec + (mb_region_base(roff, (u32)region, cap) + it.off) * (KW + 1), bytes, &bars[buf]);
        it.off += count;
        if (last) {
            it.off = 0;
            if (!last_chunk) {
                do it.chunk++; while (s_cn[it.chunk] == 0);
                it.n = s_cn[it.chunk];
            } else if (++it.cls == it.ncls) {
                it.cls = 0;
                next_bin();
            } else if (MULTI) {
                it.chunk = 0;
                while (s_cn[it.chunk] == 0) it.chunk++;
                it.n = s_cn[it.chunk];
            }
        }
    };
    if (tid == 0) {
        if (!MULTI) {
            it.n = region_records(it.bin, nbins);
            it.n_next = region_records(it.bin + gridDim.x, nbins);
            if (it.bin < nbins && it.n) {
                it.ncls = it.n <= thr1 ? 1u : (it.n + thr1 - 1) / thr1;
                my_records += it.n;
            } else if (it.bin < nbins) {
                next_bin();
            }
        } else if (it.bin < nbins && !open_bin()) {
            next_bin();
        }
        produce(0);
        produce(1);
    }
    const u32 c_all = n_genomes < cs ? n_genomes : cs;   // the count of a k-mer every genome holds
    u32 n_one = 0, n_all = 0;
    u64 pairs = 0;

    // Expand the distinct records collected so far into the k-mer table (every k-mer takes its record's genome mask), then forget them.
    // The windows of the distinct records were laid out densely while the records were collected (map[t] = record of window t,
    // dwoff[record] = its first t), so the threads simply share them: no prefix sum, no barrier before the loop.
    u64 my_drecords = 0;
    auto flush_records = [&](u32 cls, u32 ncls) {
        const u32 nw = s_wtotal < MC_WCAP ? s_wtotal : MC_WCAP;
        const u32 nd = s_dcount < dcap ? s_dcount : dcap;
        my_drecords += nd;
        for (u32 id = tid; id < nd; id += BLOCK) {                                 // window t of the layout belongs to record map[t]
            const u32 len = (u32)sm.dstore[(size_t)id * (KW + 1)], woff = sm.dwoff[id];
            for (u32 e = 0; e < len; e++) sm.map[woff + e] = (unsigned short)id;
        }
        __syncthreads();
        bool ok = true;
        for (u32 t0 = 0; t0 < nw; t0 += BLOCK) {                                // every lane of a warp stays in the loop: the append below is warp-wide
            const u32 t = t0 + tid;
            const u32 id = t < nw ? sm.map[t] : 0xffffu;                           // 0xffff: reserved by a record that found no room
            int slot = -1;
            bool claimed = false;
            if (id != 0xffffu) {
                const u64 *R = sm.dstore + (size_t)id * (KW + 1);
                const u32 e = t - sm.dwoff[id];
                const mb_kmer<KW> key = mb_expand(R, e, k, (mb_kmer<KW> *)nullptr);
                if (ncls == 1 || __umulhi(mb_class_hash<KW>(key), ncls) == cls) {
                    slot = mb_slot_of<KW>(sm, s_log2, key, &s_distinct, &claimed);
                    if (slot < 0) ok = false;
                    else mb_or_mask(sm.tbits + slot, sm.dmask[id]);
                }
            }
            mb_append_claimed(sm.slots, &s_distinct, claimed, slot);
        }
        if (!ok) s_over = 1;
        __syncthreads();
        for (u32 i = tid; i < RT; i += BLOCK) sm.rtab[i] = 0u;
        if (tid == 0) {
            s_dcount = 0;
            s_wtotal = 0;
            s_dfull = 0;
        }
        __syncthreads();
    };

    for (u32 s = 0;; s++) {
        const u32 buf = s & 1u;
        mbar_wait(&bars[buf], (s >> 1) & 1u);
        const mc_desc d = desc[buf];
        if (d.flags & MCF_DONE) break;
        if (!s_over) {
            // level 1: every record of the stage -> the index of distinct records
            const u64 *R = sm.ring + ((size_t)buf * MC_R + tid) * (KW + 1);
            bool todo = tid < d.count;
            u32 g = 0, len = 0;
            u64 w[KW], h = 0;
            if (todo) {
                const u64 hdr = R[0];
                g = (u32)(hdr >> 48);
                len = (u32)(hdr >> 40) & 0xffu;
                len = len < MB_MAXW ? len : MB_MAXW;
#pragma unroll
                for (int i = 0; i < KW; i++) w[i] = R[1 + i];
                h = mb_record_hash(w, KW, len);
            }
            // Rounds: a record finds its content in the index (-> one bit), or claims an empty index slot and becomes a new distinct
            // record, or meets a slot whose owner is still writing and looks again in the next round (copies of one record from many
            // genomes sit in the same stage: only ONE of them may allocate).
            for (;;) {
                if (todo) {
                    u32 slot = (u32)h & (RT - 1);
                    for (;;) {
                        u32 v = *(volatile u32 *)&sm.rtab[slot];
                        if (v == 0u) {
                            v = atomicCAS(&sm.rtab[slot], 0u, MC_PENDING);
                            if (v == 0u) {   // this thread owns the slot
                                const u32 woff = atomicAdd(&s_wtotal, len);
                                const u32 wend = woff + len < MC_WCAP ? woff + len : MC_WCAP;
                                u32 id = ~0u;
                                if (woff + len <= MC_WCAP) id = atomicAdd(&s_dcount, 1u);
                                if (id >= dcap) {         // no room (records or windows): expand what is there, then this record again
                                    for (u32 t = woff; t < wend; t++) sm.map[t] = 0xffffu;
                                    s_dfull = 1;
                                    *(volatile u32 *)&sm.rtab[slot] = 0u;
                                    break;
                                }
                                sm.dwoff[id] = (unsigned short)woff;     // its windows are entered into the map at flush time, by all threads
                                u64 *D = sm.dstore + (size_t)id * (KW + 1);
                                D[0] = (u64)len;
#pragma unroll
                                for (int i = 0; i < KW; i++) D[1 + i] = w[i];
                                sm.dmask[id] = 1ull << (g & 63u);
                                __threadfence_block();
                                *(volatile u32 *)&sm.rtab[slot] = id + 1u;
                                todo = false;
                                break;
                            }
                        }
                        if (v == MC_PENDING) break;      // its owner publishes before the next round
                        const u64 *D = sm.dstore + (size_t)(v - 1u) * (KW + 1);
                        bool same = (u32)*(volatile u64 *)&D[0] == len;
#pragma unroll
                        for (int i = 0; i < KW; i++) same = same && *(volatile u64 *)&D[1 + i] == w[i];
                        if (same) {
                            u32 *mw = (u32 *)&sm.dmask[v - 1u] + ((g >> 5) & 1u);
                            const u32 bm = 1u << (g & 31u);
                            if (!(*(volatile u32 *)mw & bm)) atomicOr(mw, bm);
                            todo = false;
                            break;
                        }
                        slot = (slot + 1) & (RT - 1);
                    }
                }
                if (!__syncthreads_or(todo)) break;
                if (s_dfull) {
                    __syncthreads();
                    flush_records(d.cls, d.ncls);
                }
            }
        }
        __syncthreads();
        if (tid == 0) produce(buf);
        if (!(d.flags & MCF_LAST)) continue;
        // the last stage of this class of the bin: expand, then every occupied slot is one distinct k-mer of the group
        if (!s_over) flush_records(d.cls, d.ncls);
        if (s_distinct > S - S / 4) s_over = 1;     // same value in every thread (s_distinct is stable between the barriers)
        if (!(d.flags & MCF_SCAN)) {                // end of a chunk of 64 genomes, more to come: fold the bits into the counts
            if (!s_over) {
                const u32 nk = s_distinct;
                for (u32 j = tid; j < nk; j += BLOCK) {
                    const u32 i = sm.slots[j];
                    const u64 b = sm.tbits[i];
                    if (b) {
                        sm.tcnt[i] += (unsigned short)__popcll(b);
                        sm.tbits[i] = 0ull;
                    }
                }
            }
            __syncthreads();
            continue;
        }
        if (s_over) {
            __syncthreads();
            for (u32 i = tid; i < S; i += BLOCK) {
                mb_slot_reset<KW>(sm, i);
                sm.tbits[i] = 0ull;
                if (MULTI) sm.tcnt[i] = 0u;
            }
            for (u32 i = tid; i < RT; i += BLOCK) sm.rtab[i] = 0u;
            if (tid == 0) {
                const u32 at = atomicAdd(over_count, 1u);
                if (at < over_cap) {
                    over_list[3 * at] = d.bin;
                    over_list[3 * at + 1] = d.cls;
                    over_list[3 * at + 2] = d.ncls;
                } else {
                    atomicOr((unsigned long long *)d_stat, 2ull);
                }
                s_over = 0;
                s_distinct = 0;
                s_dcount = 0;
                s_wtotal = 0;
                s_dfull = 0;
            }
            __syncthreads();
            continue;
        }
        const u32 nd_keys = s_distinct;
        const bool routed = route.world && out_keys;
        if (!routed) {      // (routed: the bin's place in the local store is reserved together with its places at the owners, further down)
            if (tid == 0 && nd_keys) {
                s_base = atomicAdd((unsigned long long *)d_cursor, (unsigned long long)nd_keys);
                if (out_keys) mb_log_event(evlog, d.bin, nd_keys, s_base, d_stat);
            }
            __syncthreads();
        }
        const u64 base = routed ? 0ull : s_base;
        for (u32 j = tid; j < nd_keys; j += BLOCK) {
            const u32 i = sm.slots[j];
            u32 c = (u32)__popcll(sm.tbits[i]);
            sm.tbits[i] = 0ull;
            if (MULTI) {
                c += sm.tcnt[i];
                sm.tcnt[i] = 0u;
            }
            pairs += c;
            const u32 cc = c > cs ? cs : c;
            if (cc == 1u) n_one++;
            else if (cc == c_all) n_all++;
            else if (cc <= hrows) atomicAdd(&sm.hist[cc], 1u);
            if (routed) {
                mb_route_note<KW>(sm, i, k, route.world, s_ocnt);     // multi-GPU: the key leaves for its owner below
            } else {
                if (out_keys) mb_emit<KW>(sm, i, out_keys, base + j, k);
                mb_slot_reset<KW>(sm, i);
            }
        }
        __syncthreads();
        if (routed) {
            // ONE round trip to L2 per bin for all reservations: the owners' ranges (threads 0 .. world - 1) and the local store's (thread 64).
            // With the local reservation up front (as in the unrouted case) every bin waited for two dependent atomics: +0.2 ms per config-2 group
            if (tid == 64 && nd_keys) {
                s_base = atomicAdd((unsigned long long *)d_cursor, (unsigned long long)nd_keys);
                mb_log_event(evlog, d.bin, nd_keys, s_base, d_stat);
            }
            mb_route_reserve(route, s_ocnt, s_obase, s_ostart);
            __syncthreads();
            const u64 lbase = s_base;
            if (route.flags & 1u) {
                // KHB_PEER_SORTED=1: the bin's keys in owner-sorted order (the window map is free at this point), so that every owner's keys
                // leave as one run of consecutive threads.  Off by default: the NVLink side is far from its limit here (~140 GB/s asked of
                // it) and the extra pass costs instructions (2 B200s: 2.32 against 2.28 ms per group, profiles/r2_bench_history.md)
                for (u32 j = tid; j < nd_keys; j += BLOCK) {
                    const u32 i = sm.slots[j];
                    const u64 w = sm.tbits[i];
                    sm.map[s_ostart[(u32)(w >> 32)] + (u32)w] = (unsigned short)i;
                }
                __syncthreads();
            }
            if (tid < 64) s_ocnt[tid] = 0u;
            for (u32 j = tid; j < nd_keys; j += BLOCK) {
                const u32 i = (route.flags & 1u) ? sm.map[j] : sm.slots[j];
                mb_route_store<KW>(sm, i, route, s_obase, out_keys, lbase + j);
                mb_slot_reset<KW>(sm, i);
            }
            __syncthreads();
        }
        if (tid == 0) s_distinct = 0;
        __syncthreads();
    }
    // flush this CTA's counts
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_one += __shfl_xor_sync(0xffffffffu, n_one, o);
        n_all += __shfl_xor_sync(0xffffffffu, n_all, o);
        pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    }
    if ((tid & 31u) == 0) {
        if (n_one && 1u <= hrows) atomicAdd(&sm.hist[1], n_one);
        if (n_all && c_all <= hrows) atomicAdd(&sm.hist[c_all], n_all);
        if (pairs) atomicAdd((unsigned long long *)d_pairs, (unsigned long long)pairs);
    }
    if (tid == 0 && my_records) {
        atomicAdd((unsigned long long *)&d_stat[1], (unsigned long long)my_records);
        atomicAdd((unsigned long long *)&d_stat[4], (unsigned long long)my_drecords);
    }
    __syncthreads();
    for (u32 i = tid; i <= hrows; i += BLOCK) {
        const u32 v = sm.hist[i];
        if (v) atomicAdd((unsigned long long *)&hist[i], (unsigned long long)v);
    }
}

// ---- pass B: the bins whose table filled up, redone with the keys split into P hash classes, one class per pass ------------------
template <int KW, bool MULTI>
__global__ void __launch_bounds__(MC_BLOCK)
mb_bigbin_kernel(const u64 *__restrict__ rec, const u32 *__restrict__ cursor, u32 cap, int k, mc_geom geo, u32 n_genomes, u32 cs,
                 u64 *__restrict__ hist, void *__restrict__ out_keys, u64 *__restrict__ d_cursor, u64 *__restrict__ d_pairs,
                 const u32 *__restrict__ over_list, const u32 *__restrict__ over_count, u32 over_cap, u64 *__restrict__ flags, const u64 *__restrict__ roff,
                 mb_evlog evlog, khb_peer_route route)
{
    extern __shared__ __align__(16) unsigned char mc_raw[];
    __shared__ u32 s_over, s_distinct, ws[33];
    __shared__ u32 s_ocnt[64], s_ostart[64];
    __shared__ u64 s_obase[64];
    __shared__ u64 s_base;
    const u32 s_log2 = geo.s_log2, S = 1u << s_log2, nchunks = geo.nchunks, hrows = geo.hrows;
    const mc_smem sm = mc_carve(mc_raw, KW, geo);
    const u32 tid = threadIdx.x;
    const u32 n_over = *over_count < over_cap ? *over_count : over_cap;
    if (flags[0] & 1ull) return;
    if (blockIdx.x == 0 && tid == 0) flags[2] = n_over;
    if (blockIdx.x >= n_over) return;
    for (u32 i = tid; i < S; i += MC_BLOCK) {
        mb_slot_reset<KW>(sm, i);
        sm.tbits[i] = 0ull;
        if (MULTI) sm.tcnt[i] = 0u;
    }
    for (u32 i = tid; i <= hrows; i += MC_BLOCK) sm.hist[i] = 0u;
    if (tid < 64) s_ocnt[tid] = 0u;
    if (tid == 0) {
        s_over = 0;
        s_distinct = 0;
    }
    __syncthreads();
    u64 pairs = 0;
    for (u32 li = blockIdx.x; li < n_over; li += gridDim.x) {
        const u32 bin = over_list[3 * li], cls0 = over_list[3 * li + 1], ncls0 = over_list[3 * li + 2];
        const u32 nch = MULTI ? nchunks : 1u;
        // work list of hash classes (modulus M, residue r): keys with h % M == r.  A class whose distinct keys do not fit the table is
        // split into (2M, r) and (2M, r + M), which together are exactly that class -- nothing of it was emitted yet.
        u32 stM[40], stR[40];
        int sp = 0;
        stM[sp] = 2; stR[sp++] = 1;
        stM[sp] = 2; stR[sp++] = 0;
        while (sp > 0) {
            const u32 M = stM[--sp], r0 = stR[sp];
            for (u32 chunk = 0; chunk < nch; chunk++) {
                const u32 region = bin * nch + chunk;
                const u32 lim = mb_region_cap(roff, region, cap);
                const u32 n = cursor[region] < lim ? cursor[region] : lim;
                const u64 *rb = rec + mb_region_base(roff, region, cap) * (KW + 1);
                const u32 sub = tid & 15u, grp = tid >> 4;   // 16 lanes share a record
                bool ok = true;
                for (u32 r = grp; r < n; r += MC_BLOCK / 16) {
                    const u64 *R = rb + (size_t)r * (KW + 1);
                    const u64 h = R[0];
                    const u32 g = (u32)(h >> 48);
                    u32 len = (u32)(h >> 40) & 0xffu;
                    len = len < MB_MAXW ? len : MB_MAXW;
                    for (u32 e = sub; e < len; e += 16) {
                        const mb_kmer<KW> key = mb_expand(R, e, k, (mb_kmer<KW> *)nullptr);
                        if (ncls0 > 1 && __umulhi(mb_class_hash<KW>(key), ncls0) != cls0) continue;
                        if ((u32)(mb_hash(key) >> 20) % M != r0) continue;
                        ok = mb_insert<KW>(sm, s_log2, key, g & 63u, &s_distinct) && ok;
                    }
                }
                if (!ok) s_over = 1;
                __syncthreads();
                if (s_distinct > S - S / 4) s_over = 1;
                if (s_over) break;
                if (MULTI && chunk + 1 < nchunks) {
                    const u32 nk = s_distinct;
                    for (u32 j = tid; j < nk; j += MC_BLOCK) {
                        const u32 i = sm.slots[j];
                        const u64 b = sm.tbits[i];
                        if (b) {
                            sm.tcnt[i] += (unsigned short)__popcll(b);
                            sm.tbits[i] = 0ull;
                        }
                    }
                    __syncthreads();
                }
            }
            const bool over = s_over != 0;
            __syncthreads();
            if (over) {
                for (u32 i = tid; i < S; i += MC_BLOCK) {
                    mb_slot_reset<KW>(sm, i);
                    sm.tbits[i] = 0ull;
                    if (MULTI) sm.tcnt[i] = 0u;
                }
                if (tid == 0) {
                    s_over = 0;
                    s_distinct = 0;
                }
                __syncthreads();
                if (M >= (1u << 16) || sp + 2 > 40) {      // cannot happen for a bin that fits its region; the caller redoes the group by sorting
                    if (tid == 0) atomicOr((unsigned long long *)flags, 2ull);
                    sp = 0;
                    break;
                }
                stM[sp] = 2 * M; stR[sp++] = r0 + M;
                stM[sp] = 2 * M; stR[sp++] = r0;
                continue;
            }
            const u32 nd_keys = s_distinct;
            if (tid == 0 && nd_keys) {
                s_base = atomicAdd((unsigned long long *)d_cursor, (unsigned long long)nd_keys);
                if (out_keys) mb_log_event(evlog, bin, nd_keys, s_base, flags);
            }
            __syncthreads();
            const u64 base = s_base;
            for (u32 j = tid; j < nd_keys; j += MC_BLOCK) {
                const u32 i = sm.slots[j];
                u32 c = (u32)__popcll(sm.tbits[i]);
                sm.tbits[i] = 0ull;
                if (MULTI) {
                    c += sm.tcnt[i];
                    sm.tcnt[i] = 0u;
                }
                pairs += c;
                const u32 cc = c > cs ? cs : c;
                if (cc <= hrows) atomicAdd(&sm.hist[cc], 1u);
                if (route.world && out_keys) {
                    mb_route_note<KW>(sm, i, k, route.world, s_ocnt);
                } else {
                    if (out_keys) mb_emit<KW>(sm, i, out_keys, base + j, k);
                    mb_slot_reset<KW>(sm, i);
                }
            }
            __syncthreads();
            if (route.world && out_keys) {
                mb_route_reserve(route, s_ocnt, s_obase, s_ostart);
                __syncthreads();
                if (tid < 64) s_ocnt[tid] = 0u;
                for (u32 j = tid; j < nd_keys; j += MC_BLOCK) {
                    const u32 i = sm.slots[j];
                    mb_route_store<KW>(sm, i, route, s_obase, out_keys, base + j);
                    mb_slot_reset<KW>(sm, i);
                }
                __syncthreads();
            }
            if (tid == 0) s_distinct = 0;
            __syncthreads();
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pairs += __shfl_xor_sync(0xffffffffu, pairs, o);
    if ((tid & 31u) == 0 && pairs) atomicAdd((unsigned long long *)d_pairs, (unsigned long long)pairs);
    __syncthreads();
    for (u32 i = tid; i <= hrows; i += MC_BLOCK) {
        const u32 v = sm.hist[i];
        if (v) atomicAdd((unsigned long long *)&hist[i], (unsigned long long)v);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
void khb_prof_patch(khb_ctx *ctx, int id, u64 alg_bytes);

static long long mb_env(const char *name, long long dflt)
{
    const char *e = getenv(name);
    return e && *e ? atoll(e) : dflt;
}

// Does the minimizer-bin path apply?  (64-bit keys with a spare value, enough m-mers per window for super-k-mers to pay.)
int khb_bins_eligible(int k, int n_genomes, u64 n_sym)
{
    return k >= 17 && k <= 63 && k != 32 && n_genomes >= 1 && n_genomes <= 8191 && n_sym >= 1 && n_sym < (1ull << 40);
}

// ---- the planner: bins, regions, tables -- a pure function of its inputs, so that the members of a team (team.cu) arrive at the same geometry
struct mb_plan_in {
    int k, n_genomes;       // genomes of the WHOLE group
    u32 nchunks;            // chunks of <= 64 genomes (0: ceil(n_genomes / 64))
    u64 n_sym;              // symbols of the whole group
    double rho_prev;        // distinct k-mers per window measured on the previous group (0: a guess from the group size)
    u32 nb_fixed;           // != 0: this many bins
    u32 cap_fixed;          // != 0: this many records per region
    u32 nbins_hist;
    int hint_k, hint_genomes;          // shape of the group before this one ...
    u64 hint_regions, hint_max;        // ... its regions and the records asked for in its fullest one
};
struct mb_plan {
    int KW, m;
    u32 capw, nb, cap, thr1, over_cap;
    mc_geom geo;
    u64 n_regions;
    size_t rec_bytes;
    double rho_w, est_records;
};
static int mb_make_plan(khb_ctx *ctx, const mb_plan_in &in, mb_plan *out)
{
    const int k = in.k, n_genomes = in.n_genomes;
    const u64 n_sym = in.n_sym;
    const int KW = k <= 32 ? 2 : 3;               // record words of symbols: k - 1 + 32 windows fit 64 (k <= 32) or 96 symbols
    const int m = 13, w = k - m + 1;
    const u32 capw = (u32)(32 * KW + 1 - k) < MB_MAXW ? (u32)(32 * KW + 1 - k) : MB_MAXW;
    double avg_len = (w + 1) * 0.5;               // windows per super-k-mer on random sequence, before the cuts at capw and at tile ends
    if (avg_len > capw) avg_len = capw;
    // distinct k-mers per window: measured on the previous group (+ 25 %), else a guess from the group size
    double rho_w = in.rho_prev > 0.0 ? in.rho_prev * 1.25 : (n_genomes >= 32 ? 0.2 : n_genomes >= 16 ? 0.25 : n_genomes >= 4 ? 0.5 : 1.0);
    const long long rho_pct = mb_env("KHB_BINS_RHO_PCT", 0);   // test hook: distinct k-mers per 100 windows
    if (rho_pct > 0) rho_w = rho_pct / 100.0;
    if (rho_w > 1.0) rho_w = 1.0;
    if (rho_w < 0.005) rho_w = 0.005;
    mc_geom geo;
    geo.nchunks = in.nchunks ? in.nchunks : (u32)div_up((size_t)n_genomes, 64);
    // Bin geometry.  All copies of a k-mer -- one per genome that holds it, avg_len windows around each -- land in one bin together, so a
    // region's load (one bin, one chunk of <= 64 genomes) comes in lumps of ~avg_len x genomes windows; ~6 lumps per region keep the largest
    // region within a few times the mean.  But a bin's distinct k-mers should fill less than half of the largest table of which two CTAs fit
    // an SM (2^12 slots of 64-bit, 2^11 of 128-bit k-mers); where the two pull apart the table wins, the regions get more room above their
    // mean, and a region that still overflows makes the group be partitioned a second time with exact sizes.
    const u32 s_max = KW == 3 ? 11u : 12u;
    const double lump = (double)(n_genomes < 64 ? n_genomes : 64) * avg_len;
    double wpb_d = (double)geo.nchunks * 6.0 * lump;
    const double wpb_tab = 0.45 * (double)(1u << s_max) / rho_w;
    if (wpb_d > wpb_tab) wpb_d = wpb_tab;
    if (wpb_d < 1024.0) wpb_d = 1024.0;
    const u64 wpb = (u64)mb_env("KHB_BINS_WPB", (long long)wpb_d);
    u64 nb64 = div_up(n_sym, wpb ? wpb : 1024);
    if (nb64 < 16) nb64 = 16;
    if (in.nb_fixed) nb64 = in.nb_fixed;          // the store's earlier groups were binned with this many bins: keep the bins comparable
    if (nb64 > (1ull << 28)) return khb_fail(ctx, KHB_ERR_ARG, "minimizer-bin group stage: group too large");
    const u32 nb = (u32)nb64;
    {   // table slots: the mean bin's distinct k-mers fill ~45 % (larger bins are counted in hash classes)
        const double want = (double)(n_sym / nb + 1) * rho_w / 0.45;
        u32 l2 = 10;
        while (l2 < s_max && (double)(1u << l2) < want) l2++;
        geo.s_log2 = (u32)mb_env("KHB_BINS_SLOTS_LOG2", l2);
        if (geo.s_log2 < 8 || geo.s_log2 > 13) return khb_fail(ctx, KHB_ERR_ARG, "KHB_BINS_SLOTS_LOG2 outside 8..13");
    }
    geo.hrows = in.nbins_hist < (u32)n_genomes ? in.nbins_hist : (u32)n_genomes;
    geo.dcap = (u32)mb_env("KHB_BINS_DCAP", geo.nchunks > 1 ? 256 : 384);
    if (geo.dcap < 64 || geo.dcap > 4096) return khb_fail(ctx, KHB_ERR_ARG, "KHB_BINS_DCAP outside 64..4096");
    geo.rt_log2 = 7;
    while ((1u << geo.rt_log2) < 2 * geo.dcap) geo.rt_log2++;
    const u32 s_log2 = geo.s_log2;
    const double est_records = (double)n_sym / avg_len * 1.15 + (double)div_up(n_sym, MB_TILE);
    // one region per bin and chunk of 64 genomes; the fewer lumps a region holds on average, the more room above the mean it gets
    const u64 n_regions = (u64)nb * geo.nchunks;
    const double slack = (double)mb_env("KHB_BINS_SLACK_PCT", (double)(n_sym / nb) >= 5.0 * lump * geo.nchunks ? 400 : 600) / 100.0;
    u32 cap = ((u32)(est_records / (double)n_regions * slack) + 64u) & ~1u;   // even: every region starts 16-byte aligned
    // ... unless the group before this one had the same shape: then its fullest region (+ 40 %) is the better guide
    if (in.hint_k == k && in.hint_genomes == n_genomes && in.hint_regions > 0 && !mb_env("KHB_BINS_SLACK_PCT", 0) &&
        (double)n_regions > 0.8 * (double)in.hint_regions && (double)n_regions < 1.25 * (double)in.hint_regions)
        cap = ((u32)((double)in.hint_max * 1.4 * (double)in.hint_regions / (double)n_regions) + 64u) & ~1u;
    if (in.cap_fixed) cap = (in.cap_fixed + 1u) & ~1u;
    // Records one pass over a bin may hold so that its distinct k-mers fill at most ~55 % of the table; larger bins are counted in
    // several hash classes.
    double thr = 0.55 * (double)(1u << s_log2) / (rho_w * avg_len);
    out->KW = KW;
    out->m = m;
    out->capw = capw;
    out->nb = nb;
    out->cap = cap;
    out->thr1 = thr < 8.0 ? 8u : thr > 1e9 ? 1000000000u : (u32)thr;
    out->over_cap = 4 * nb;
    out->geo = geo;
    out->n_regions = n_regions;
    out->rec_bytes = (size_t)n_regions * cap * (KW + 1) * 8;
    out->rho_w = rho_w;
    out->est_records = est_records;
    return KHB_OK;
}

// Passes C and B over `nb_count` bins whose regions lie in `rec` (region u = bin * nchunks + chunk at u * cap, or at roff[u]) with `cur[u]` records each.
static int mb_launch_count(khb_ctx *ctx, const mb_plan &pl, int k, const u64 *rec, const u32 *cur, u32 nb_count, u32 n_genomes, u32 cs, u64 *d_hist, void *d_out_keys,
                           u64 *d_runs, u64 *d_pairs, u32 *d_over_list, u32 *d_over_count, u64 *d_stat, const u64 *roff, mb_evlog evlog, khb_peer_route route, u64 n_sym)
{
    const int KW = pl.KW;
    mc_geom geo = pl.geo;
    const bool multi = geo.nchunks > 1;
    size_t shm = mc_smem_bytes(KW, geo, 256);
    if (shm > 227 * 1024) return khb_fail(ctx, KHB_ERR_ARG, "minimizer-bin group stage: %zu bytes of shared memory per CTA", shm);
    // two CTAs of 256 threads per SM at least; where the tables leave room for one CTA only, that one has 512 threads
    const bool wide = 2 * (shm + 1024) > 227 * 1024;
    const void *fn, *bfn;
    if (KW == 2) {
        fn = multi ? (wide ? (const void *)mb_count_kernel<2, true, 512> : (const void *)mb_count_kernel<2, true, 256>)
                   : (wide ? (const void *)mb_count_kernel<2, false, 512> : (const void *)mb_count_kernel<2, false, 256>);
        bfn = multi ? (const void *)mb_bigbin_kernel<2, true> : (const void *)mb_bigbin_kernel<2, false>;
    } else {
        fn = multi ? (wide ? (const void *)mb_count_kernel<3, true, 512> : (const void *)mb_count_kernel<3, true, 256>)
                   : (wide ? (const void *)mb_count_kernel<3, false, 512> : (const void *)mb_count_kernel<3, false, 256>);
        bfn = multi ? (const void *)mb_bigbin_kernel<3, true> : (const void *)mb_bigbin_kernel<3, false>;
    }
    const int block = wide ? 512 : 256;
    shm = mc_smem_bytes(KW, geo, block);
    int per_sm = 0;
    KHB_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    KHB_CUDA(ctx, cudaFuncSetAttribute(bfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    KHB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, block, shm));
    if (per_sm < 1) return khb_fail(ctx, KHB_ERR_ARG, "minimizer-bin group stage: the counting kernel does not fit an SM (%zu bytes of shared memory)", shm);
    const long long want = mb_env("KHB_BINS_CTAS_PER_SM", 0);
    if (want > 0 && want < per_sm) per_sm = (int)want;
    u32 grid = (u32)ctx->num_sms * (u32)per_sm;
    if (grid > nb_count) grid = nb_count;
    if (grid < 1) grid = 1;
    const u64 *c_rec = rec;
    const u32 *c_cur = cur;
    u32 n_gen = n_genomes, cs_ = cs, nb_ = nb_count, cap_ = pl.cap, thr_ = pl.thr1, ocap_ = pl.over_cap;
    int k_ = k;
    void *keys_ = d_out_keys;
    const u32 *c_list = d_over_list, *c_cnt = d_over_count;
    void *cargs[] = {&c_rec, &c_cur, &nb_, &cap_, &k_, &geo, &n_gen, &cs_, &thr_, &ocap_, &d_hist, &keys_, &d_runs, &d_pairs, &d_over_list, &d_over_count, &d_stat, &roff, &evlog, &route};
    if (mb_env("KHB_BINS_VERBOSE", 0))
        fprintf(stderr, "[bins] k=%d genomes=%u windows=%llu bins=%u (of %u) chunks=%u cap=%u slots=2^%u dcap=%u thr1=%u rho_w=%.3f shm=%zu ctas/sm=%d block=%d\n", k, n_genomes,
                (unsigned long long)n_sym, nb_count, pl.nb, geo.nchunks, pl.cap, geo.s_log2, geo.dcap, pl.thr1, pl.rho_w, shm, per_sm, block);
    khb_prof_begin(ctx, KHB_K_BIN_COUNT);
    KHB_CUDA(ctx, cudaLaunchKernel(fn, dim3(grid), dim3(block), cargs, shm, ctx->stream));
    khb_prof_end(ctx, KHB_K_BIN_COUNT, 0);
    KHB_LAUNCH_CHECK(ctx);
    void *bargs[] = {&c_rec, &c_cur, &cap_, &k_, &geo, &n_gen, &cs_, &d_hist, &keys_, &d_runs, &d_pairs, &c_list, &c_cnt, &ocap_, &d_stat, &roff, &evlog, &route};
    KHB_CUDA(ctx, cudaLaunchKernel(bfn, dim3((u32)ctx->num_sms * 2u), dim3(MC_BLOCK), bargs, shm, ctx->stream));
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}

// The group stage through minimizer bins.  d_stat: u64[4] in device memory, written here: [0] flags (1: a bin region overflowed,
// 2: a bin could not be counted -- either way the outputs are incomplete and the caller redoes the group another way), [1] records,
// [2] bins redone by mb_bigbin_kernel, [3] records asked for in the fullest region, [4] distinct records.  d_hist[nbins_hist + 1], d_runs (distinct k-mers = keys appended to d_out_keys), d_pairs
// (sum over genomes of their distinct k-mers) are zeroed here.  exact != 0: the call before this one, on the same group, ended with flag 1;
// partition again into regions of exactly the sizes that attempt counted (they are still in the context's scratch).
int khb_bins_count_impl(khb_ctx *ctx, const u64 *d_codes, const u32 *d_valid, u64 n_sym, int k, const u64 *d_seg_off, int n_genomes, u32 cs, u32 nbins_hist,
                        u64 *d_hist, void *d_out_keys, u64 *d_runs, u64 *d_pairs, u64 *d_stat, int exact, u32 nb_fixed, u32 *nb_used, void *ev_buf,
                        u64 *ev_count, u64 ev_cap, u64 store_base, khb_peer_route route)
{
    if (!khb_bins_eligible(k, n_genomes, n_sym)) return khb_fail(ctx, KHB_ERR_ARG, "minimizer-bin group stage: k=%d / %d genomes not supported", k, n_genomes);
    mb_plan_in pin;
    pin.k = k;
    pin.n_genomes = n_genomes;
    pin.nchunks = 0;
    pin.n_sym = n_sym;
    pin.rho_prev = ctx->bins_rho;
    pin.nb_fixed = nb_fixed;
    pin.cap_fixed = 0;
    pin.nbins_hist = nbins_hist;
    pin.hint_k = ctx->bins_hint_k;
    pin.hint_genomes = ctx->bins_hint_genomes;
    pin.hint_regions = ctx->bins_hint_regions;
    pin.hint_max = ctx->bins_hint_max;
    mb_plan pl;
    int rc = mb_make_plan(ctx, pin, &pl);
    if (rc) return rc;
    const int KW = pl.KW, m = pl.m;
    const u32 capw = pl.capw, nb = pl.nb, cap = pl.cap, thr1 = pl.thr1, over_cap = pl.over_cap;
    const mc_geom geo = pl.geo;
    const u64 n_regions = pl.n_regions;
    const size_t rec_bytes = pl.rec_bytes;
    const double rho_w = pl.rho_w;
    if (nb_used) *nb_used = nb;
    const bool multi = geo.nchunks > 1;
    void *p;
    if ((rc = khb_scratch_get(ctx, SCR_AUX, (size_t)n_regions * 12 + (size_t)nb * 48 + 512, &p))) return rc;
    u32 *d_over_count = (u32 *)p;                 // [0] bins in the list
    u32 *d_cur = d_over_count + 16, *d_over_list = d_cur + n_regions;
    u64 *d_roff = (u64 *)(((uintptr_t)(d_over_list + 3 * (size_t)over_cap) + 15) & ~(uintptr_t)15);   // [n_regions + 1], exact layout only
    const u64 *roff = nullptr;
    void *pr;
    if (exact) {
        mb_region_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_cur, n_regions, d_roff);
        KHB_LAUNCH_CHECK(ctx);
        KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, d_roff + n_regions, 8, cudaMemcpyDeviceToHost, ctx->stream));
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        const u64 total = ctx->h_mail[0];
        if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, (size_t)total * (KW + 1) * 8 + 64, &pr))) return rc;
        roff = d_roff;
    } else if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, rec_bytes + 64, &pr))) {
        return rc;
    }
    KHB_CUDA(ctx, cudaMemsetAsync(d_over_count, 0, 64 + (size_t)n_regions * 4, ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins_hist + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_runs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_pairs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_stat, 0, 5 * sizeof(u64), ctx->stream));
    const u64 last_w = n_sym / 32 + 3;            // khb_codes_words / khb_valid_words: n / 32 + 4 words each
    {
        const u64 tiles = div_up(n_sym, MB_TILE);
        khb_prof_begin(ctx, KHB_K_BIN_PARTITION);
        if (KW == 2)
            mb_partition_kernel<2><<<(unsigned)tiles, MB_BLOCK, 0, ctx->stream>>>(d_codes, d_valid, n_sym, last_w, last_w, k, m, nb, d_seg_off, n_genomes, geo.nchunks,
                                                                                  d_cur, (u64 *)pr, cap, capw, roff, d_stat, 0u);
        else
            mb_partition_kernel<3><<<(unsigned)tiles, MB_BLOCK, 0, ctx->stream>>>(d_codes, d_valid, n_sym, last_w, last_w, k, m, nb, d_seg_off, n_genomes, geo.nchunks,
                                                                                  d_cur, (u64 *)pr, cap, capw, roff, d_stat, 0u);
        khb_prof_end(ctx, KHB_K_BIN_PARTITION, n_sym * 3 / 8);
        KHB_LAUNCH_CHECK(ctx);
    }
    {
        mb_evlog evlog;
        evlog.buf = (mb_event *)ev_buf;
        evlog.count = ev_count;
        evlog.cap = ev_cap;
        evlog.store_base = store_base;
        if ((rc = mb_launch_count(ctx, pl, k, (const u64 *)pr, d_cur, nb, (u32)n_genomes, cs, d_hist, d_out_keys, d_runs, d_pairs, d_over_list, d_over_count, d_stat, roff,
                                  evlog, route, n_sym))) return rc;
        mb_region_max_kernel<<<(unsigned)ctx->num_sms * 4u, 256, 0, ctx->stream>>>(d_cur, n_regions, d_stat + 3);
        KHB_LAUNCH_CHECK(ctx);
        ctx->bins_last_regions = n_regions;
    }
    return KHB_OK;
}

// ---- one group on several GPUs: the same two passes, the records crossing NVLink in between ------------------------------------------
// (include/khoice_b200.h: khb_team_*).  The members of a team plan from the same numbers (mb_make_plan is a pure function), so every member
// knows where the owner of a bin expects the records of (bin, chunk): buffer `parity` of the owner holds bpo x nchunks_total regions of `cap`
// records, then the table of region sizes.
struct mb_team_layout {
    u32 bpo;              // bins per owner
    u64 n_regions;        // regions of one owner: bpo x chunks of the whole group
    u64 area;             // records ONE sender may pack into one owner's buffer (even)
    size_t roff_off;      // offset of the region starts (u64 per region, in records) inside a receive buffer
    size_t cur_off;       // ... of the region sizes (u32 per region)
    size_t need;          // bytes one receive buffer must hold
};
// A receive buffer: `team` sender areas of `area` records each, packed densely by their senders (a region = one contiguous run, rounded up to an
// even number of records), then the table of region starts and the table of region sizes.  Dense on purpose: the first version kept the
// slack layout of the local partition (region u at u * cap, ~20 % filled) and reached ~200 GB/s over NVLink; the senders' stores want
// consecutive lines of consecutive pages.
static mb_team_layout mb_team_layout_of(const mb_plan &pl, int team, u32 area_pct)
{
    mb_team_layout L;
    L.bpo = (u32)div_up((size_t)pl.nb, (size_t)team);
    L.n_regions = (u64)L.bpo * pl.geo.nchunks;
    const double pct = area_pct ? (double)area_pct : 250.0;     // above the mean share of a (sender, owner) pair: uneven slices, uneven bins
    L.area = ((u64)(pl.est_records / (double)team / (double)team * pct / 100.0) + 2 * L.n_regions + 1024) & ~1ull;
    const size_t rec_bytes = (size_t)L.area * team * (pl.KW + 1) * 8;
    L.roff_off = (rec_bytes + 255) & ~(size_t)255;
    L.cur_off = L.roff_off + ((L.n_regions * 8 + 255) & ~(size_t)255);
    L.need = L.cur_off + ((L.n_regions * 4 + 255) & ~(size_t)255);
    return L;
}
static int mb_team_plan(khb_ctx *ctx, int k, const khb_team_group *tg, u32 nbins_hist, mb_plan *pl)
{
    if (!tg || tg->n_genomes_total < 1 || tg->n_chunks_total < 1 || tg->n_chunks_total > 128 || tg->chunk_base < 0 || tg->chunk_base >= tg->n_chunks_total ||
        (tg->parity != 0 && tg->parity != 1) || !tg->n_sym_total)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_team_*: bad group description");
    if (!khb_bins_eligible(k, tg->n_genomes_total, tg->n_sym_total))
        return khb_fail(ctx, KHB_ERR_ARG, "khb_team_*: a group is sharded through the minimizer bins only (17 <= k <= 63, k != 32); k=%d", k);
    mb_plan_in pin;
    pin.k = k;
    pin.n_genomes = tg->n_genomes_total;
    pin.nchunks = (u32)tg->n_chunks_total;
    pin.n_sym = tg->n_sym_total;
    pin.rho_prev = tg->rho;
    pin.nb_fixed = 0;
    pin.cap_fixed = tg->region_cap;
    pin.nbins_hist = nbins_hist;
    pin.hint_k = 0;
    pin.hint_genomes = 0;
    pin.hint_regions = 0;
    pin.hint_max = 0;
    return mb_make_plan(ctx, pin, pl);
}
int khb_bins_team_plan_impl(khb_ctx *ctx, int k, const khb_team_group *tg, u32 nbins_hist, u32 *nb, u32 *cap, u64 *half_bytes)
{
    khb_team *tm = ctx->team;
    if (!tm) return khb_fail(ctx, KHB_ERR_STATE, "khb_team_plan: khb_team_alloc first");
    mb_plan pl;
    int rc = mb_team_plan(ctx, k, tg, nbins_hist, &pl);
    if (rc) return rc;
    const mb_team_layout L = mb_team_layout_of(pl, tm->size, tg->area_pct);
    if (nb) *nb = pl.nb;
    if (cap) *cap = pl.cap;
    if (half_bytes) *half_bytes = L.need;
    return KHB_OK;
}

// Every region of this member's partition -> the record buffer of the bin's owner, packed densely into this sender's area there, followed by
// its start and its size (the owner's counting pass reads them like the exact layout of a local partition).  loff: exclusive prefix sums of
// this member's region sizes rounded up to even (mb_region_scan_kernel); the regions of one owner are a contiguous index range, so a region's
// place inside the area is its prefix minus the prefix of the owner's first region.  One warp per region; a region is one contiguous run of
// a few KB moved as full 16-byte vectors, 512 bytes per warp instruction (records stored one by one from the partition pass itself -- the
// first version: 32 scattered bytes per thread -- crossed NVLink at 4 GB/s).
__global__ void __launch_bounds__(256)
mb_region_push_kernel(const u64 *__restrict__ rec, const u32 *__restrict__ lcur, const u64 *__restrict__ loff, u32 nb, u32 nchunks_mine, u32 cap, u32 rec_words,
                      mb_shard sh, u64 *__restrict__ info)
{
    const u32 lane = threadIdx.x & 31u;
    const u64 n = (u64)nb * nchunks_mine;
    const u64 warps = ((u64)gridDim.x * blockDim.x) >> 5;
    u64 moved = 0;
    for (u64 i = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n; i += warps) {
        const u32 bin = (u32)(i / nchunks_mine), lc = (u32)(i - (u64)bin * nchunks_mine);
        const u32 owner = bin / sh.bpo, lb = bin - owner * sh.bpo;
        const u64 dst_region = (u64)lb * sh.nchunks_total + sh.chunk_base + lc;
        const u32 asked = __ldg(lcur + i);
        u32 c = asked < cap ? asked : cap;
        const u64 at = __ldg(loff + i) - __ldg(loff + (u64)owner * sh.bpo * nchunks_mine);     // inside this sender's area of the owner's buffer
        if (at + ((c + 1u) & ~1u) > sh.area) {     // the area is full: the team repeats the group with larger buffers
            if (lane == 0) atomicOr((unsigned long long *)info, 2ull);
            c = 0;
        }
        const u64 dst_at = (u64)sh.member * sh.area + (c ? at : 0ull);
        if (lane == 0) {
            sh.roff[owner][dst_region] = dst_at;
            sh.cur[owner][dst_region] = c;
        }
        moved += c;
        const uint4 *src = (const uint4 *)(rec + i * cap * rec_words);               // cap is even: every region starts 16-byte aligned
        uint4 *dst = (uint4 *)(sh.rec[owner] + dst_at * rec_words);                  // so does every packed region (even sizes)
        const u32 nv = (c * rec_words * 8u + 15u) >> 4;                              // an odd count of 24-byte records: 8 bytes of slack ride along
        u32 v = lane;
        for (; v + 96 < nv; v += 128) {
            const uint4 a = src[v], b = src[v + 32], d = src[v + 64], e = src[v + 96];
            dst[v] = a;
            dst[v + 32] = b;
            dst[v + 64] = d;
            dst[v + 96] = e;
        }
        for (; v < nv; v += 32) dst[v] = src[v];
    }
    if (lane == 0 && moved) atomicAdd((unsigned long long *)(info + 1), (unsigned long long)moved);   // records of this member's slice
}

// Pass P of this member's slice (n_genomes genomes, n_sym symbols) of a sharded group.  d_info: u64[4] device, [0] flags (1: one of this
// member's regions overflowed), [3] records asked for in its fullest region.
int khb_bins_team_partition_impl(khb_ctx *ctx, const u64 *d_codes, const u32 *d_valid, u64 n_sym, int k, const u64 *d_seg_off, int n_genomes,
                                 const khb_team_group *tg, u64 *d_info)
{
    khb_team *tm = ctx->team;
    if (!tm || !tm->opened) return khb_fail(ctx, KHB_ERR_STATE, "khb_team_partition: khb_team_alloc / khb_team_open first");
    mb_plan pl;
    int rc = mb_team_plan(ctx, k, tg, KHB_COUNTER_MAX, &pl);
    if (rc) return rc;
    const u32 nchunks_mine = (u32)div_up((size_t)n_genomes, 64);
    if (n_genomes < 1 || (u32)tg->chunk_base + nchunks_mine > (u32)tg->n_chunks_total)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_team_partition: %d genomes from chunk %d do not fit the group's %d chunks", n_genomes, tg->chunk_base, tg->n_chunks_total);
    const mb_team_layout L = mb_team_layout_of(pl, tm->size, tg->area_pct);
    if (L.need > tm->half_bytes)
        return khb_fail(ctx, KHB_ERR_CAPACITY, "khb_team_partition: the group needs %zu bytes per receive buffer, the team allocated %zu", L.need, tm->half_bytes);
    mb_shard sh;
    memset(&sh, 0, sizeof(sh));
    sh.team = (u32)tm->size;
    sh.bpo = L.bpo;
    sh.nchunks_total = pl.geo.nchunks;
    sh.chunk_base = (u32)tg->chunk_base;
    sh.member = (u32)tm->member;
    sh.area = L.area;
    for (int t = 0; t < tm->size; t++) {
        char *half = (char *)tm->peer_base[t] + (size_t)tg->parity * tm->half_bytes;
        sh.rec[t] = (u64 *)half;
        sh.roff[t] = (u64 *)(half + L.roff_off);
        sh.cur[t] = (u32 *)(half + L.cur_off);
    }
    void *p, *pr;
    const u64 n_local = (u64)pl.nb * nchunks_mine;
    const u32 RW = (u32)pl.KW + 1u;
    if ((rc = khb_scratch_get(ctx, SCR_TEAM, n_local * 12 + 256, &p))) return rc;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, (size_t)n_local * pl.cap * RW * 8 + 64, &pr))) return rc;
    u64 *d_loff = (u64 *)p;                       // [n_local + 1] prefix sums of the region sizes
    u32 *d_lcur = (u32 *)(d_loff + n_local + 2);
    KHB_CUDA(ctx, cudaMemsetAsync(d_lcur, 0, n_local * 4, ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_info, 0, 4 * sizeof(u64), ctx->stream));
    if (n_sym) {
        khb_prof_begin(ctx, KHB_K_BIN_PARTITION);
        // the slice is partitioned into LOCAL regions (bin, own chunk), the genome ids already those of the whole group
        const u64 tiles = div_up(n_sym, MB_TILE), last_w = n_sym / 32 + 3;
        if (pl.KW == 2)
            mb_partition_kernel<2><<<(unsigned)tiles, MB_BLOCK, 0, ctx->stream>>>(d_codes, d_valid, n_sym, last_w, last_w, k, pl.m, pl.nb, d_seg_off, n_genomes, nchunks_mine,
                                                                                  d_lcur, (u64 *)pr, pl.cap, pl.capw, nullptr, d_info, 64u * (u32)tg->chunk_base);
        else
            mb_partition_kernel<3><<<(unsigned)tiles, MB_BLOCK, 0, ctx->stream>>>(d_codes, d_valid, n_sym, last_w, last_w, k, pl.m, pl.nb, d_seg_off, n_genomes, nchunks_mine,
                                                                                  d_lcur, (u64 *)pr, pl.cap, pl.capw, nullptr, d_info, 64u * (u32)tg->chunk_base);
        khb_prof_end(ctx, KHB_K_BIN_PARTITION, n_sym * 3 / 8);
        KHB_LAUNCH_CHECK(ctx);
    }
    // ... and every region leaves for its owner as one run, packed (d_info[0] & 2: an area was too small, d_info[1]: records moved)
    khb_prof_begin(ctx, KHB_K_PARTITION);
    mb_region_scan4_kernel<<<1, 1024, 0, ctx->stream>>>(d_lcur, n_local, d_loff);
    KHB_LAUNCH_CHECK(ctx);
    mb_region_push_kernel<<<(unsigned)ctx->num_sms * 8u, 256, 0, ctx->stream>>>((const u64 *)pr, d_lcur, d_loff, pl.nb, nchunks_mine, pl.cap, RW, sh, d_info);
    khb_prof_end(ctx, KHB_K_PARTITION, 0);
    KHB_LAUNCH_CHECK(ctx);
    mb_region_max_kernel<<<(unsigned)ctx->num_sms * 4u, 256, 0, ctx->stream>>>(d_lcur, n_local, d_info + 3);
    KHB_LAUNCH_CHECK(ctx);
    if (mb_env("KHB_BINS_VERBOSE", 0))
        fprintf(stderr, "[team] member %d/%d: k=%d slice=%d genomes (chunks %d..%u of %d) windows=%llu of %llu bins=%u (%u per owner) cap=%u buffer=%zu of %zu bytes\n", tm->member,
                tm->size, k, n_genomes, tg->chunk_base, tg->chunk_base + nchunks_mine - 1, tg->n_chunks_total, (unsigned long long)n_sym, (unsigned long long)tg->n_sym_total,
                pl.nb, L.bpo, pl.cap, L.need, tm->half_bytes);
    return KHB_OK;
}

// Passes C and B over the bins this member owns; outputs like khb_bins_count_impl's (d_stat[0] & 2: a bin could not be counted).
int khb_bins_team_count_impl(khb_ctx *ctx, int k, const khb_team_group *tg, u32 cs, u32 nbins_hist, u64 *d_hist, void *d_out_keys, u64 *d_runs,
                             u64 *d_pairs, u64 *d_stat, khb_peer_route route)
{
    khb_team *tm = ctx->team;
    if (!tm || !tm->opened) return khb_fail(ctx, KHB_ERR_STATE, "khb_team_count: khb_team_alloc / khb_team_open first");
    mb_plan pl;
    int rc = mb_team_plan(ctx, k, tg, nbins_hist, &pl);
    if (rc) return rc;
    const mb_team_layout L = mb_team_layout_of(pl, tm->size, tg->area_pct);
    if (L.need > tm->half_bytes) return khb_fail(ctx, KHB_ERR_CAPACITY, "khb_team_count: the group needs %zu bytes per receive buffer, the team allocated %zu", L.need, tm->half_bytes);
    const u64 first = (u64)tm->member * L.bpo;
    const u32 nb_mine = first >= pl.nb ? 0u : (pl.nb - first < L.bpo ? (u32)(pl.nb - first) : L.bpo);
    char *half = (char *)tm->recv + (size_t)tg->parity * tm->half_bytes;
    void *p;
    if ((rc = khb_scratch_get(ctx, SCR_AUX, (size_t)pl.nb * 48 + 512, &p))) return rc;
    u32 *d_over_count = (u32 *)p, *d_over_list = d_over_count + 16;
    KHB_CUDA(ctx, cudaMemsetAsync(d_over_count, 0, 64, ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins_hist + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_runs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_pairs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_stat, 0, 5 * sizeof(u64), ctx->stream));
    if (!nb_mine) return KHB_OK;
    mb_evlog evlog;
    evlog.buf = nullptr;
    evlog.count = nullptr;
    evlog.cap = 0;
    evlog.store_base = 0;
    pl.cap = MB_CAP_DENSE;      // the regions lie where their senders packed them: exact starts, exact sizes
    return mb_launch_count(ctx, pl, k, (const u64 *)half, (const u32 *)(half + L.cur_off), nb_mine, (u32)tg->n_genomes_total, cs, d_hist, d_out_keys, d_runs, d_pairs,
                           d_over_list, d_over_count, d_stat, (const u64 *)(half + L.roff_off), evlog, route, tg->n_sym_total);
}

// ---- the across-group stage, bin by bin -------------------------------------------------------------------------------------------
// `kmc_tools complex` over the groups' sets + its histogram (/root/reference/workflow/rules/exp_type_1.smk:243-259) without the sort: a
// k-mer's bin is the same in every group, so the segments the groups' end-of-bin passes left in the store (mb_event) are first
// ordered by bin (a counting sort of the events, not of the keys), then one CTA per bin counts in how many segments every key occurs --
// each group holds a key at most once -- in a shared-memory table of (key, count).  Bins with more keys than the table takes are
// counted in hash classes (several passes over the bin's segments, which then come from L2).
#define MA_BLOCK 256
#define MA_EVMAX 512          // segments of a bin held in shared memory at a time
#define MA_UNROLL 4
#define MA_PROBES 1024        // linear probing in a table that is at most 44 % full: never reached

__global__ void __launch_bounds__(256)
ma_event_hist_kernel(const mb_event *__restrict__ ev, u64 n_ev, u32 nb, u32 *__restrict__ bin_events, u32 *__restrict__ bin_keys, u64 *__restrict__ flags)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n_ev; e += (u64)gridDim.x * blockDim.x) {
        const mb_event v = ev[e];
        if (v.bin >= nb) {
            *flags = 1ull;
            continue;
        }
        atomicAdd(&bin_events[v.bin], 1u);
        atomicAdd(&bin_keys[v.bin], v.count);
    }
}
__global__ void __launch_bounds__(1024)
ma_scan_kernel(const u32 *__restrict__ cnt, u64 n, u64 *__restrict__ off)
{
    __shared__ u64 ws[33];
    __shared__ u64 carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (u64 base = 0; base < n; base += 1024) {
        const u64 i = base + threadIdx.x;
        const u64 v = i < n ? (u64)cnt[i] : 0ull;
        u64 total;
        const u64 ex = block_excl_sum<u64>(v, ws, &total);
        if (i < n) off[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = carry;
}
__global__ void __launch_bounds__(256)
ma_event_scatter_kernel(const mb_event *__restrict__ ev, u64 n_ev, u32 nb, const u64 *__restrict__ off, u32 *__restrict__ fill, mb_event *__restrict__ out)
{
    for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < n_ev; e += (u64)gridDim.x * blockDim.x) {
        const mb_event v = ev[e];
        if (v.bin >= nb) continue;
        out[off[v.bin] + atomicAdd(&fill[v.bin], 1u)] = v;
    }
}

// KW2 = 1: 64-bit keys (u64), 2: 128-bit keys (Key128: lo, hi).  Persistent CTAs, one bin at a time: the bin's segments are copied into
// shared memory by cp.async.bulk (one copy per segment, one mbarrier for all of them; a 64-bit segment that starts at an odd key index is
// copied from the key before it, so that every copy is 16-byte aligned), then every hash class of the bin is counted from shared memory.
#define MA_STAGE_BYTES 40960  // keys of a bin (or of a chunk of its segments) in shared memory
template <int KW2>
__global__ void __launch_bounds__(MA_BLOCK)
mb_across_kernel(const void *__restrict__ store, const mb_event *__restrict__ ev, const u64 *__restrict__ off, const u32 *__restrict__ bin_keys, u32 nb,
                 u32 s_log2, u32 cs, u32 hrows, u32 n_groups, u64 *__restrict__ hist, u64 *__restrict__ d_distinct, u64 *__restrict__ flags)
{
    extern __shared__ __align__(16) unsigned char ma_raw[];
    __shared__ u32 e_start[MA_EVMAX];     // first real key of the segment inside the stage buffer (in keys)
    __shared__ u32 e_pre[MA_EVMAX + 1];   // real keys of the segments before it
    __shared__ __align__(8) u64 bar;
    __shared__ u32 s_over, s_distinct, s_ne;
    __shared__ u64 s_next;                // first segment of the next chunk
    constexpr u32 KB = KW2 == 2 ? 16u : 8u;
    constexpr u32 STAGE_KEYS = MA_STAGE_BYTES / KB;
    const u32 S = 1u << s_log2, smask = S - 1u, tid = threadIdx.x;
    unsigned char *stage = ma_raw;                                // [MA_STAGE_BYTES]
    u64 *tkey = (u64 *)(ma_raw + MA_STAGE_BYTES);                 // key (KW2 = 1) or its low word
    u64 *thi = tkey + S;                                          // high word (KW2 = 2)
    u32 *tcnt = (u32 *)(thi + (KW2 == 2 ? S : 0));
    u32 *ttag = tcnt + S;                                         // KW2 = 2 only
    u32 *shist = ttag + (KW2 == 2 ? S : 0);
    unsigned short *slots = (unsigned short *)(shist + ((hrows + 2) & ~1u));
    for (u32 i = tid; i < S; i += MA_BLOCK) {
        if (KW2 == 2) ttag[i] = 0u; else tkey[i] = MB_EMPTY;
        tcnt[i] = 0u;
    }
    for (u32 i = tid; i <= hrows; i += MA_BLOCK) shist[i] = 0u;
    if (tid == 0) {
        mbar_init(&bar, 1);
        s_over = 0;
        s_distinct = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const u32 thr = S / 4 + S / 8 + S / 16;                       // keys one pass may bring (>= its distinct keys): the table stays below 44 % full
    const u32 c_all = n_groups < cs ? n_groups : cs;
    u32 n_one = 0, n_all = 0, phase = 0;
    u64 distinct = 0;
    // the next bin's numbers are loaded while this one is counted
    u32 b = blockIdx.x;
    u32 total = b < nb ? __ldg(bin_keys + b) : 0u;
    u64 e0 = b < nb ? __ldg(off + b) : 0ull, e1 = b < nb ? __ldg(off + b + 1) : 0ull;
    while (b < nb) {
        const u32 bn = b + gridDim.x;
        const u32 total_n = bn < nb ? __ldg(bin_keys + bn) : 0u;
        const u64 e0_n = bn < nb ? __ldg(off + bn) : 0ull, e1_n = bn < nb ? __ldg(off + bn + 1) : 0ull;
        const u32 ncls = total <= thr ? 1u : (total + thr - 1) / thr;
        bool resident = false;                                    // the whole bin sits in the stage buffer (loaded once, counted class by class)
        bool gave_up = false;
        for (u32 cls = 0; cls < ncls && !gave_up; cls++) {
            bool ok = true;
            for (u64 ec = e0; ec < e1;) {                         // chunks of segments that fit the stage buffer (normally ONE per bin)
                if (!resident) {
                    __syncthreads();                              // the stage buffer and the segment tables are free
                    if (tid < 32) {
                        // one warp lays the segments out, 32 at a time, lane i taking segment ec + i0 + i, until the buffer is full
                        u32 keys_run = 0, real_run = 0, n_take = 0;
                        bool full = false;
                        for (u32 i0 = 0; i0 < MA_EVMAX && ec + i0 < e1 && !full; i0 += 32) {
                            const u64 ei = ec + i0 + tid;
                            mb_event v;
                            v.count = 0;
                            v.base_lo = v.base_hi = 0;
                            const bool have = ei < e1 && i0 + tid < MA_EVMAX;
                            if (have) v = ev[ei];
                            const u64 base = ((u64)v.base_hi << 32) | v.base_lo;
                            const u32 lead = (KW2 == 1 && have) ? (u32)(base & 1ull) : 0u;   // the key in front, for 16-byte alignment
                            const u32 padded = !have ? 0u : KW2 == 1 ? ((v.count + lead + 1u) & ~1u) : v.count;   // keys copied
                            const u32 inc_p = warp_incl_sum(padded), inc_r = warp_incl_sum(have ? v.count : 0u);
                            const u32 my_start = keys_run + inc_p - padded;
                            const bool fits = have && my_start + padded <= STAGE_KEYS;
                            const u32 fit_mask = __ballot_sync(0xffffffffu, fits), have_mask = __ballot_sync(0xffffffffu, have);
                            const u32 first_bad = __ffs(have_mask & ~fit_mask);              // 1-based lane of the first segment that does not fit
                            const bool take = have && (first_bad == 0 || tid + 1 < first_bad);
                            if (take) {
                                const u32 idx = i0 + tid;
                                e_start[idx] = my_start + lead;
                                e_pre[idx] = real_run + inc_r - v.count;
                                bulk_g2s(stage + (size_t)my_start * KB, (const unsigned char *)store + (base - lead) * KB, padded * KB, &bar);
                            }
                            const u32 take_mask = __ballot_sync(0xffffffffu, take);
                            const u32 n_t = __popc(take_mask);
                            const u32 last = n_t ? (u32)(31 - __clz(take_mask)) : 0u;
                            const u32 add_p = __shfl_sync(0xffffffffu, inc_p, last), add_r = __shfl_sync(0xffffffffu, inc_r, last);
                            if (n_t) {
                                keys_run += add_p;
                                real_run += add_r;
                            }
                            n_take += n_t;
                            full = first_bad != 0;
                        }
                        if (tid == 0) {
                            e_pre[n_take] = real_run;
                            s_ne = n_take;
                            s_next = ec + n_take;
                            if (n_take) mbar_arrive_expect_tx(&bar, keys_run * KB);
                        }
                    }
                    __syncthreads();
                    if (!s_ne) {                                  // a single segment larger than the stage buffer: the caller sorts instead
                        if (tid == 0) atomicOr((unsigned long long *)flags, 2ull);
                        gave_up = true;
                        break;
                    }
                    mbar_wait(&bar, phase);
                    phase ^= 1u;
                    if (ec == e0 && s_next >= e1) resident = true;
                }
                const u32 ne = s_ne;
                const u32 nk = e_pre[ne];
                for (u32 t0 = 0; t0 < nk; t0 += MA_BLOCK) {       // every lane of a warp stays in the loop: the append below is warp-wide
                    const u32 t = t0 + tid;
                    int found = -1;
                    bool claimed = false;
                    if (t < nk) {
                        u32 lo = 0, hi = ne;                      // segment of key t: e_pre[lo] <= t < e_pre[lo + 1]
                        while (hi - lo > 1) {
                            const u32 mid = (lo + hi) >> 1;
                            if (e_pre[mid] <= t) lo = mid; else hi = mid;
                        }
                        const u32 at = e_start[lo] + (t - e_pre[lo]);
                        u64 klo, khi = 0;
                        if (KW2 == 2) {
                            const Key128 kk = ((const Key128 *)stage)[at];
                            klo = kk.lo;
                            khi = kk.hi;
                        } else {
                            klo = ((const u64 *)stage)[at];
                        }
                        const u64 h = (klo ^ (khi * 0xBF58476D1CE4E5B9ull)) * 0x9E3779B97F4A7C15ull;   // the keys are mixed already: spread the class and the slot bits
                        if (ncls == 1 || __umulhi((u32)(h >> 8), ncls) == cls) {
                            u32 slot = (u32)(h >> 40) & smask;
                            if (KW2 == 1) {
                                for (u32 probes = 0; probes <= MA_PROBES; probes++) {
                                    const u64 c = *(volatile u64 *)&tkey[slot];
                                    if (c == klo) { found = (int)slot; break; }
                                    if (c == MB_EMPTY) {
                                        const u64 old = atomicCAS((unsigned long long *)&tkey[slot], MB_EMPTY, klo);
                                        if (old == MB_EMPTY) {
                                            claimed = true;
                                            found = (int)slot;
                                            break;
                                        }
                                        if (old == klo) { found = (int)slot; break; }
                                    }
                                    slot = (slot + 1) & smask;
                                }
                            } else {
                                const u32 mytag = (u32)h | 2u;
                                for (u32 probes = 0; probes <= MA_PROBES; probes++) {
                                    u32 tg = *(volatile u32 *)&ttag[slot];
                                    if (tg == 0u) {
                                        tg = atomicCAS(&ttag[slot], 0u, MB_TAG_PENDING);
                                        if (tg == 0u) {
                                            tkey[slot] = klo;
                                            thi[slot] = khi;
                                            __threadfence_block();
                                            *(volatile u32 *)&ttag[slot] = mytag;
                                            claimed = true;
                                            found = (int)slot;
                                            break;
                                        }
                                    }
                                    while (tg == MB_TAG_PENDING) tg = *(volatile u32 *)&ttag[slot];
                                    if (tg == mytag && *(volatile u64 *)&tkey[slot] == klo && *(volatile u64 *)&thi[slot] == khi) { found = (int)slot; break; }
                                    slot = (slot + 1) & smask;
                                }
                            }
                            if (found < 0) ok = false;
                            else atomicAdd(&tcnt[found], 1u);
                        }
                    }
                    mb_append_claimed(slots, &s_distinct, claimed, found);
                }
                ec = resident ? e1 : s_next;
            }
            if (!ok) s_over = 1;
            __syncthreads();
            if (s_distinct > S - S / 4) s_over = 1;
            const u32 nd = s_distinct;
            if (s_over || gave_up) {                              // the caller sorts instead
                if (tid == 0) atomicOr((unsigned long long *)flags, 2ull);
                for (u32 i = tid; i < S; i += MA_BLOCK) {
                    if (KW2 == 2) ttag[i] = 0u; else tkey[i] = MB_EMPTY;
                    tcnt[i] = 0u;
                }
            } else {
                for (u32 j = tid; j < nd; j += MA_BLOCK) {
                    const u32 i = slots[j];
                    const u32 c = tcnt[i];
                    tcnt[i] = 0u;
                    if (KW2 == 2) ttag[i] = 0u; else tkey[i] = MB_EMPTY;
                    const u32 cc = c > cs ? cs : c;
                    if (cc == 1u) n_one++;
                    else if (cc == c_all) n_all++;
                    else if (cc <= hrows) atomicAdd(&shist[cc], 1u);
                }
                if (tid == 0) distinct += nd;
            }
            __syncthreads();
            if (tid == 0) {
                s_over = 0;
                s_distinct = 0;
            }
            __syncthreads();
        }
        b = bn;
        total = total_n;
        e0 = e0_n;
        e1 = e1_n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        n_one += __shfl_xor_sync(0xffffffffu, n_one, o);
        n_all += __shfl_xor_sync(0xffffffffu, n_all, o);
    }
    if ((tid & 31u) == 0) {
        if (n_one && 1u <= hrows) atomicAdd(&shist[1], n_one);
        if (n_all && c_all <= hrows && c_all != 1u) atomicAdd(&shist[c_all], n_all);
    }
    if (tid == 0 && distinct) atomicAdd((unsigned long long *)d_distinct, (unsigned long long)distinct);
    __syncthreads();
    for (u32 i = tid; i <= hrows; i += MA_BLOCK) {
        const u32 v = shist[i];
        if (v) atomicAdd((unsigned long long *)&hist[i], (unsigned long long)v);
    }
}

// Across-group histogram from the store's segment events.  d_hist[nbins_hist + 1] and d_runs (distinct k-mers overall) are zeroed here;
// d_flag: != 0 afterwards means the result is incomplete and the caller has to sort instead.
int khb_bins_across_impl(khb_ctx *ctx, int k, const void *d_store, const void *d_events, u64 n_events, u32 nb, int n_groups, u32 cs, u32 nbins_hist,
                         u64 *d_hist, u64 *d_runs, u64 *d_flag)
{
    const int KW2 = k <= 32 ? 1 : 2;
    const u32 hrows = nbins_hist < (u32)n_groups ? nbins_hist : (u32)n_groups;
    const u32 s_log2 = (u32)mb_env("KHB_ACROSS_SLOTS_LOG2", 12);
    if (s_log2 < 8 || s_log2 > 13) return khb_fail(ctx, KHB_ERR_ARG, "KHB_ACROSS_SLOTS_LOG2 outside 8..13");
    const size_t S = (size_t)1 << s_log2;
    const size_t shm = MA_STAGE_BYTES + S * 8 + (KW2 == 2 ? S * 12 : 0) + S * 4 + (((size_t)hrows + 2) & ~(size_t)1) * 4 + S * 2 + 64;
    int rc;
    void *p;
    const size_t aux = (size_t)nb * 12 + ((size_t)nb + 1) * 8 + n_events * sizeof(mb_event) + 256;
    if ((rc = khb_scratch_get(ctx, SCR_AUX, aux, &p))) return rc;
    u64 *d_off = (u64 *)p;                                        // [nb + 1]
    mb_event *d_sorted = (mb_event *)(d_off + nb + 2);            // [n_events] (16-byte aligned: nb + 2 words of 8 bytes ... see below)
    d_sorted = (mb_event *)(((uintptr_t)d_sorted + 15) & ~(uintptr_t)15);
    u32 *d_bin_events = (u32 *)(d_sorted + n_events), *d_bin_keys = d_bin_events + nb, *d_fill = d_bin_keys + nb;
    KHB_CUDA(ctx, cudaMemsetAsync(d_bin_events, 0, (size_t)nb * 12, ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins_hist + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_runs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_flag, 0, sizeof(u64), ctx->stream));
    if (!n_events) return KHB_OK;
    const unsigned eg = (unsigned)(div_up(n_events, 256) < 4096 ? div_up(n_events, 256) : 4096);
    khb_prof_begin(ctx, KHB_K_BIN_ACROSS);
    ma_event_hist_kernel<<<eg, 256, 0, ctx->stream>>>((const mb_event *)d_events, n_events, nb, d_bin_events, d_bin_keys, d_flag);
    KHB_LAUNCH_CHECK(ctx);
    ma_scan_kernel<<<1, 1024, 0, ctx->stream>>>(d_bin_events, nb, d_off);
    KHB_LAUNCH_CHECK(ctx);
    ma_event_scatter_kernel<<<eg, 256, 0, ctx->stream>>>((const mb_event *)d_events, n_events, nb, d_off, d_fill, d_sorted);
    KHB_LAUNCH_CHECK(ctx);
    const void *fn = KW2 == 1 ? (const void *)mb_across_kernel<1> : (const void *)mb_across_kernel<2>;
    KHB_CUDA(ctx, cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    int per_sm = 0;
    KHB_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, MA_BLOCK, shm));
    if (per_sm < 1) return khb_fail(ctx, KHB_ERR_ARG, "across-group stage by bins: %zu bytes of shared memory do not fit an SM", shm);
    u32 grid = (u32)ctx->num_sms * (u32)per_sm;
    if (grid > nb) grid = nb;
    const mb_event *c_ev = d_sorted;
    const u64 *c_off = d_off;
    const u32 *c_keys = d_bin_keys;
    u32 nb_ = nb, sl = s_log2, cs_ = cs, hr = hrows, ng = (u32)n_groups;
    void *args[] = {&d_store, &c_ev, &c_off, &c_keys, &nb_, &sl, &cs_, &hr, &ng, &d_hist, &d_runs, &d_flag};
    KHB_CUDA(ctx, cudaLaunchKernel(fn, dim3(grid), dim3(MA_BLOCK), args, shm, ctx->stream));
    khb_prof_end(ctx, KHB_K_BIN_ACROSS, 0);
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}

// ---- the partition alone, for the intermediate-level parity test (tests/test_gpu_bins.py: the same rules stated in numpy) -----------
__global__ void __launch_bounds__(256)
mb_region_windows_kernel(const u64 *__restrict__ rec, const u32 *__restrict__ cursor, u64 n_regions, u32 cap, int rec_words, u32 *__restrict__ windows)
{
    const u64 r = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_regions) return;
    const u32 n = cursor[r] < cap ? cursor[r] : cap;
    u32 wsum = 0;
    for (u32 i = 0; i < n; i++) wsum += (u32)(rec[(r * cap + i) * rec_words] >> 40) & 0xffu;
    windows[r] = wsum;
}

extern "C" KHB_API int khb_bins_partition(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, uint64_t n_symbols,
                                                                       const uint64_t *d_seg_off, int n_genomes, int k, uint32_t n_bins,
                                                                       uint32_t *h_records, uint32_t *h_windows)
{
    KHB_CHECK_CTX(ctx);
    if (!khb_bins_eligible(k, n_genomes, n_symbols ? n_symbols : 1) || n_bins < 1 || n_bins > (1u << 22) || !h_records || !h_windows || !d_seg_off)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_bins_partition: 17 <= k <= 63 (k != 32), 1 .. 2^22 bins");
    const int KW = k <= 32 ? 2 : 3, m = 13, w = k - m + 1;
    const u32 capw = (u32)(32 * KW + 1 - k) < MB_MAXW ? (u32)(32 * KW + 1 - k) : MB_MAXW;
    const u32 nchunks = (u32)div_up((size_t)n_genomes, 64);
    const u64 n_regions = (u64)n_bins * nchunks;
    double avg_len = (w + 1) * 0.5;
    if (avg_len > capw) avg_len = capw;
    const u32 cap = ((u32)((double)n_symbols / avg_len * 1.15 / (double)n_regions * 8.0) + 256u) & ~1u;
    int rc;
    void *p, *pr;
    if ((rc = khb_scratch_get(ctx, SCR_AUX, n_regions * 8 + 64, &p))) return rc;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, n_regions * cap * (size_t)(KW + 1) * 8 + 64, &pr))) return rc;
    u32 *d_cur = (u32 *)p, *d_win = d_cur + n_regions;
    u64 *d_flag = (u64 *)(d_win + n_regions);
    KHB_CUDA(ctx, cudaMemsetAsync(p, 0, n_regions * 8 + 64, ctx->stream));
    if (n_symbols) {
        const u64 tiles = div_up(n_symbols, MB_TILE), last_w = n_symbols / 32 + 3;
        if (KW == 2)
            mb_partition_kernel<2><<<(unsigned)tiles, MB_BLOCK, 0, ctx->stream>>>((const u64 *)d_codes, d_valid, n_symbols, last_w, last_w, k, m, n_bins, (const u64 *)d_seg_off,
                                                                                  n_genomes, nchunks, d_cur, (u64 *)pr, cap, capw, nullptr, d_flag, 0u);
        else
            mb_partition_kernel<3><<<(unsigned)tiles, MB_BLOCK, 0, ctx->stream>>>((const u64 *)d_codes, d_valid, n_symbols, last_w, last_w, k, m, n_bins, (const u64 *)d_seg_off,
                                                                                  n_genomes, nchunks, d_cur, (u64 *)pr, cap, capw, nullptr, d_flag, 0u);
        KHB_LAUNCH_CHECK(ctx);
        mb_region_windows_kernel<<<(unsigned)div_up(n_regions, 256), 256, 0, ctx->stream>>>((const u64 *)pr, d_cur, n_regions, cap, KW + 1, d_win);
        KHB_LAUNCH_CHECK(ctx);
    }
    u64 flag = 0;
    KHB_CUDA(ctx, cudaMemcpyAsync(h_records, d_cur, n_regions * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(h_windows, d_win, n_regions * 4, cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(&flag, d_flag, 8, cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (flag) return khb_fail(ctx, KHB_ERR_STATE, "khb_bins_partition: a region overflowed its diagnostic buffer");
    return KHB_OK;
}
