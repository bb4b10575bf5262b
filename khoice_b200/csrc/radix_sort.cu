// radix_sort.cu -- K3: segmented LSD radix sort of k-mer words (8-bit digits, one scatter pass per
// digit, decoupled look-back), the engine behind
//   * per-genome dedup          (kmc + `set_counts 1`,   /root/reference/workflow/rules/exp_type_1.smk:156-173)
//   * per-group union-sum       (`kmc_tools complex`,     exp_type_1.smk:175-182)
//   * across-group union-sum    (`kmc_tools complex`,     exp_type_1.smk:243-250)
// KMC does these with a disk-based bin sort and streaming N-way merges; on a B200 a group's keys fit in
// HBM, so each stage is one batched sort.  Only the low ceil(2k/8) bytes of a key are ever non-zero,
// so only that many passes run.
//
// One launch of onesweep_kernel per digit handles ALL segments (genomes) at once:
//   tile  = 512 threads x ITEMS keys, held in registers in warp-striped order (coalesced 8/16-byte loads)
//   rank  = __match_any_sync on the digit + warp-private shared-memory counters  (stable)
//   scan  = 256 digit columns: across warps, then across digits
//   chain = per-(tile, digit) look-back words (lookback.cuh); the first hop is issued before the
//           shared-memory reorder so its L2 latency is hidden
//   store = keys are reordered through shared memory so that every digit run leaves as one contiguous,
//           coalesced global write
// Digit histograms for every pass come from one upfront sweep (radix_hist_kernel), scanned per
// (segment, pass) by radix_scan_kernel.
//
// Algorithmic bytes per key of width W: W (histogram read) + P x 2W (read + write per pass).
#include "khb_common.cuh"
#include <stdlib.h>
#include <vector>

#include "lookback.cuh"

#define RS_BLOCK 512
#define RS_WARPS (RS_BLOCK / 32)

// Histogram kernel tiling (independent of the scatter-pass variant).
template <typename Key> struct RsCfg;
template <> struct RsCfg<Key64> { static constexpr int ITEMS = 12; };
template <> struct RsCfg<Key128> { static constexpr int ITEMS = 8; };

// Scatter-pass variants, selectable with KHB_SORT_VARIANT for tuning runs (default = the measured best).
struct RsVariant {
    int block, items;
};

template <typename Key> __device__ __forceinline__ Key key_max();
template <> __device__ __forceinline__ Key64 key_max<Key64>() { return Key64{~0ull}; }
template <> __device__ __forceinline__ Key128 key_max<Key128>() { return Key128{~0ull, ~0ull}; }

// Segment table (device): seg_off[s] = first key of segment s, seg_off[nseg] = end;
// seg_tile[s] = number of tiles in segments < s.
__device__ __forceinline__ int find_segment(const u64 *__restrict__ seg_tile, int nseg, u64 tile)
{
    int lo = 0, hi = nseg;  // invariant: seg_tile[lo] <= tile < seg_tile[hi]
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (seg_tile[mid] <= tile) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- upfront digit histograms ------------------------------------------------------------------
template <typename Key>
__global__ void __launch_bounds__(RS_BLOCK)
radix_hist_kernel(const Key *__restrict__ in, const u64 *__restrict__ seg_off, const u64 *__restrict__ seg_tile,
                  int nseg, u64 ntiles, int npass, int first_bit, u32 *__restrict__ hist /* [nseg][npass][256] */, u32 TILE,
                  const u64 *__restrict__ gather /* see onesweep_kernel */)
{
    extern __shared__ u32 sh[];  // [npass][256]
    const u32 tid = threadIdx.x;
    for (int i = tid; i < npass * 256; i += RS_BLOCK) sh[i] = 0;
    __syncthreads();
    // contiguous tile range per CTA so that a flush happens only when the segment changes
    const u64 per = (ntiles + gridDim.x - 1) / gridDim.x;
    const u64 t0 = (u64)blockIdx.x * per;
    const u64 t1 = t0 + per < ntiles ? t0 + per : ntiles;
    int cur = -1;
    for (u64 t = t0; t < t1; t++) {
        const int seg = find_segment(seg_tile, nseg, t);
        const int hseg = gather ? 0 : seg;     // gathered pieces are ONE logical segment: one histogram
        if (hseg != cur) {
            if (cur >= 0) {
                __syncthreads();
                for (int i = tid; i < npass * 256; i += RS_BLOCK) {
                    const u32 c = sh[i];
                    if (c) atomicAdd(&hist[(size_t)cur * npass * 256 + i], c);
                    sh[i] = 0;
                }
                __syncthreads();
            }
            cur = hseg;
        }
        const u64 begin = seg_off[seg] + (t - seg_tile[seg]) * TILE;
        const u64 end = seg_off[seg + 1];
        const u32 n = (u32)(end - begin < (u64)TILE ? end - begin : (u64)TILE);
        const Key *src = gather ? in + gather[seg] + (t - seg_tile[seg]) * TILE : in + begin;
#pragma unroll 4
        for (u32 i = tid; i < n; i += RS_BLOCK) {
            const Key key = src[i];
            for (int p = 0; p < npass; p++) atomicAdd(&sh[p * 256 + key_digit(key, first_bit + 8 * p)], 1u);
        }
    }
    __syncthreads();
    if (cur >= 0)
        for (int i = tid; i < npass * 256; i += RS_BLOCK) {
            const u32 c = sh[i];
            if (c) atomicAdd(&hist[(size_t)cur * npass * 256 + i], c);
        }
}

// Exclusive scan of every 256-bin histogram, in place: one warp per (segment, pass).
__global__ void radix_scan_kernel(u32 *__restrict__ hist, int nhist)
{
    const int h = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (h >= nhist) return;
    u32 *p = hist + (size_t)h * 256;
    const u32 lane = lane_id();
    u32 carry = 0;
    for (int c = 0; c < 8; c++) {
        const u32 v = p[c * 32 + lane];
        const u32 inc = warp_incl_sum(v);
        p[c * 32 + lane] = carry + inc - v;
        carry += __shfl_sync(0xffffffffu, inc, 31);
    }
}

// ---- one scatter pass ------------------------------------------------------------------------------
// peers = lanes of the warp whose digit equals mine.  MATCH 0: MATCH.ANY instruction; MATCH 1: eight ballots
// (one per digit bit; fixed-latency ALU/vote work that pipelines across the ITEMS rounds).
template <int MATCH>
__device__ __forceinline__ u32 match_digit(u32 d)
{
    if (MATCH == 0) return __match_any_sync(0xffffffffu, d);
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; b++) {
        const bool bit = (d >> b) & 1u;
        const u32 bal = __ballot_sync(0xffffffffu, bit);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// Walk back over the predecessors of tile `rel` in one digit column until an inclusive prefix is found, then
// publish this tile's inclusive prefix.  win[] holds the LBW nearest predecessors (already fetched); further hops
// fetch LBW2 entries at a time (independent loads in flight).  Returns the exclusive prefix.
template <int LBW>
__device__ __forceinline__ u32 lookback_finish(u64 *lbcol, u64 rel, const u64 (&win)[LBW], u32 count, u32 epoch)
{
    u32 excl = 0;
    if (rel == 0) return 0;
    constexpr int LBW2 = 8;
    u64 r0 = rel;  // entries [0, r0) remain to be examined, nearest first
    bool done = false;
#pragma unroll
    for (int i = 0; i < LBW; i++) {
        if (done || r0 <= (u64)i) continue;
        u64 e = win[i];
        u32 st = lb_status(e, epoch);
        while (st == 0) {
            e = lb_load(lbcol + (r0 - 1 - i) * 256);
            st = lb_status(e, epoch);
        }
        excl += (u32)e;
        if (st == LB_PREFIX) done = true;
    }
    if (r0 <= LBW) done = true;
    r0 = r0 > LBW ? r0 - LBW : 0;
    while (!done) {
        u64 w8[LBW2];
#pragma unroll
        for (int i = 0; i < LBW2; i++) w8[i] = (r0 > (u64)i) ? lb_load(lbcol + (r0 - 1 - i) * 256) : 0ull;
#pragma unroll
        for (int i = 0; i < LBW2; i++) {
            if (done || r0 <= (u64)i) continue;
            u64 e = w8[i];
            u32 st = lb_status(e, epoch);
            while (st == 0) {
                e = lb_load(lbcol + (r0 - 1 - i) * 256);
                st = lb_status(e, epoch);
            }
            excl += (u32)e;
            if (st == LB_PREFIX) done = true;
        }
        if (r0 <= LBW2) done = true;
        r0 = r0 > LBW2 ? r0 - LBW2 : 0;
    }
    lb_store(lbcol + rel * 256, lb_pack(LB_PREFIX, (u64)excl + count, epoch));
    return excl;
}

#ifdef KHB_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[16];
#define PHASE_MARK(i)                                                      \
    do {                                                                   \
        if (threadIdx.x == 0) {                                            \
            const long long now__ = clock64();                             \
            atomicAdd(&g_phase_cycles[i], (unsigned long long)(now__ - t_prev__)); \
            t_prev__ = now__;                                              \
        }                                                                  \
    } while (0)
#else
#define PHASE_MARK(i)
#endif

// PAY = 1: every key carries a 16-bit payload (the genome id in the single-sort group path) through the pass.
template <typename Key, int BLOCK, int ITEMS, int MINB, int MATCH, int PAY, int EARLY>
__global__ void __launch_bounds__(BLOCK, MINB)
onesweep_kernel(const Key *__restrict__ in, Key *__restrict__ out, const unsigned short *__restrict__ pin,
                unsigned short *__restrict__ pout, const u64 *__restrict__ seg_off,
                const u64 *__restrict__ seg_tile, int nseg, int pass_row, int shift, int npass,
                const u32 *__restrict__ bin_base /* [nseg][npass][256], exclusive */, u64 *__restrict__ lookback,
                u32 *__restrict__ ticket, u32 epoch, int debug_nolb, const u64 *__restrict__ gather)
{
    // gather != nullptr (first pass of the multi-GPU across-group stage, api.cu: khb_peer_across): the "segments" are PIECES of one logical
    // array -- piece s holds the keys [seg_off[s], seg_off[s + 1]) of it and lies at in + gather[s] (a region of the peer receive buffer,
    // written by rank s) -- so the input is read piece by piece, a partial tile at the end of every piece, while the prefix chain, the
    // digit offsets and the output are those of a single segment: the pass sorts the concatenation without anybody concatenating it.
    constexpr int TILE = BLOCK * ITEMS;
    constexpr int NW = BLOCK / 32;
    constexpr int LBW = 2;  // look-back window: predecessors fetched per hop
    static_assert(BLOCK >= 256, "256 digit threads needed");
    static_assert(TILE < 65536, "16-bit tile offsets");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Key *sorted = (Key *)smem_raw;                                             // [TILE]
    unsigned short *wcnt = (unsigned short *)(smem_raw + sizeof(Key) * TILE);  // [NW][256]
    u32 *glob_off = (u32 *)(wcnt + NW * 256);                                  // [256] offset of sorted[j] in the segment, minus j
    u32 *ws = glob_off + 256;                                                  // [36] scan scratch
    u32 *mm = ws + 36;                                                         // [NW][256] peer masks (MATCH >= 2 only)
    u32 *tcnt = mm + (MATCH >= 2 ? NW * 256 : 0);                              // [256] early tile digit counts (EARLY only)
    unsigned short *sortedp = (unsigned short *)(tcnt + (EARLY ? 256 : 0));    // [TILE] payloads in sorted order (PAY only)
    __shared__ u32 s_tile;
    __shared__ int s_seg;

    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
#ifdef KHB_PHASE_TIMING
    long long t_prev__ = clock64();
#endif
    if (tid == 0) {
        const u32 t = atomicAdd(ticket, 1u);
        s_tile = t;
        s_seg = find_segment(seg_tile, nseg, t);
    }
    for (int i = tid; i < NW * 256 / 2; i += BLOCK) ((u32 *)wcnt)[i] = 0;
    if (MATCH >= 2)
        for (int i = tid; i < NW * 256; i += BLOCK) mm[i] = 0;
    if (EARLY)
        for (int i = tid; i < 256; i += BLOCK) tcnt[i] = 0;
    __syncthreads();
    const u64 tile = s_tile;
    const int seg = s_seg;
    const u64 piece_rel = tile - seg_tile[seg];
    const u64 first_tile = gather ? 0ull : seg_tile[seg];
    const u64 seg_begin = gather ? 0ull : seg_off[seg];
    const u64 rel = tile - first_tile;  // position of this tile in its segment's chain
    const u64 begin = seg_off[seg] + piece_rel * TILE;
    const u64 seg_end = seg_off[seg + 1];
    const u32 n = (u32)(seg_end - begin < (u64)TILE ? seg_end - begin : (u64)TILE);
    const int oseg = gather ? 0 : seg;  // segment of the digit offsets
    PHASE_MARK(0);  // ticket + segment lookup

    // load: warp-striped, memory order = (warp, item, lane)
    Key keys[ITEMS];
    u32 rank2[(ITEMS + 1) / 2];  // two 16-bit ranks per register
#pragma unroll
    for (int r = 0; r < (ITEMS + 1) / 2; r++) rank2[r] = 0;
    const u32 wbase = warp * (32 * ITEMS);
    const Key *src = (gather ? in + gather[seg] + piece_rel * TILE : in + begin) + wbase + lane;
    u32 pay2[PAY ? (ITEMS + 1) / 2 : 1];  // two payloads per register
    if (wbase + 32 * ITEMS <= n) {
#pragma unroll
        for (int r = 0; r < ITEMS; r++) keys[r] = src[r * 32];
    } else {
#pragma unroll
        for (int r = 0; r < ITEMS; r++) keys[r] = (wbase + r * 32 + lane < n) ? src[r * 32] : key_max<Key>();
    }
    if (PAY) {
        const unsigned short *psrc = pin + begin + wbase + lane;
#pragma unroll
        for (int r = 0; r < (ITEMS + 1) / 2; r++) pay2[r] = 0;
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const u32 v = (wbase + r * 32 + lane < n) ? (u32)psrc[r * 32] : 0u;
            pay2[r >> 1] |= v << (16 * (r & 1));
        }
    }
#ifdef KHB_PHASE_TIMING
    if (tid == 0 && key_digit(keys[0], 0) == 999u) g_phase_cycles[15] = 1;  // force the loads to complete here
#endif
    PHASE_MARK(1);  // key loads
    // EARLY: count the tile's digits with plain shared-memory atomics and publish the aggregate BEFORE the (much
    // longer) stable ranking, so that successors never wait for this tile's ranking and the prefix chain resolves
    // while everybody is still ranking.
    u32 count = 0, dstart = 0, excl = 0;
    u64 *lbcol = lookback + first_tile * 256 + tid;  // column `tid` of the segment's look-back rows
    u64 win[LBW];
    if (EARLY) {
#pragma unroll
        for (int r = 0; r < ITEMS; r++) atomicAdd(&tcnt[key_digit(keys[r], shift)], 1u);
        __syncthreads();
        if (tid < 256) count = tcnt[tid];
        u32 total0;
        dstart = block_excl_sum<u32>(count, ws, &total0);
        if (tid < 256) {
            if (tid == 255) count -= (u32)(TILE - n);  // padding is not data
            lb_store(lbcol + rel * 256, lb_pack(rel == 0 ? LB_PREFIX : LB_AGG, count, epoch));
#pragma unroll
            for (int i = 0; i < LBW; i++) win[i] = (rel > (u64)i) ? lb_load(lbcol + (rel - 1 - i) * 256) : 0ull;
            // EARLY 2: resolve the chain NOW -- the inclusive prefix of this tile is published a few thousand cycles
            // after its ticket instead of after ranking + reorder, so the frontier stays close behind the newest tile
            if (EARLY == 2) excl = lookback_finish<LBW>(lbcol, rel, win, count, epoch);
        }
    }
    // rank within the warp's chunk (stable)
    unsigned short *mycnt = wcnt + warp * 256;
    if (MATCH >= 2) {
        // peers through shared memory: every lane ORs its lane bit into the word of its digit.  MATCH 3 / 4 send
        // every 2nd / 3rd round through eight ballots instead, to balance the shared-memory pipe against the ALUs.
        u32 *mymm = mm + warp * 256;
        const u32 lbit = 1u << lane;
        u32 bpeers[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const bool by_ballot = (MATCH == 3 && (r & 1)) || (MATCH == 4 && (r % 3) == 0);
            bpeers[r] = by_ballot ? match_digit<1>(key_digit(keys[r], shift)) : 0u;
        }
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const bool by_ballot = (MATCH == 3 && (r & 1)) || (MATCH == 4 && (r % 3) == 0);
            const u32 d = key_digit(keys[r], shift);
            u32 peers;
            if (by_ballot) {
                peers = bpeers[r];
            } else {
                atomicOr(&mymm[d], lbit);
                __syncwarp();
                peers = mymm[d];
            }
            const u32 c = mycnt[d];
            __syncwarp();
            const u32 below = __popc(peers & lanemask_lt());
            if (below == 0) {
                if (!by_ballot) mymm[d] = 0;
                mycnt[d] = (unsigned short)(c + __popc(peers));
            }
            rank2[r >> 1] |= (c + below) << (16 * (r & 1));
            __syncwarp();
        }
    } else {
        u32 peers[ITEMS];
#pragma unroll
        for (int r = 0; r < ITEMS; r++) peers[r] = match_digit<MATCH>(key_digit(keys[r], shift));
#pragma unroll
        for (int r = 0; r < ITEMS; r++) {
            const u32 d = key_digit(keys[r], shift);
            const u32 c = mycnt[d];
            __syncwarp();
            const u32 below = __popc(peers[r] & lanemask_lt());
            if (below == 0) mycnt[d] = (unsigned short)(c + __popc(peers[r]));
            rank2[r >> 1] |= (c + below) << (16 * (r & 1));
            __syncwarp();
        }
    }
    PHASE_MARK(2);  // ranking (warp 0)
    __syncthreads();
    PHASE_MARK(3);  // wait for the other warps
    // per digit: exclusive scan across warps -> tile count
    if (!EARLY) {
        if (tid < 256) {
#pragma unroll
            for (int w = 0; w < NW; w++) count += wcnt[w * 256 + tid];
        }
        // exclusive scan across digits (padding keys of a partial tile carry digit 255 and stay at the end)
        u32 total;
        dstart = block_excl_sum<u32>(count, ws, &total);
    }
    if (tid < 256) {
        // per-warp start of every digit inside the sorted tile
        u32 run = dstart;
#pragma unroll
        for (int w = 0; w < NW; w++) {
            const u32 c = wcnt[w * 256 + tid];
            wcnt[w * 256 + tid] = (unsigned short)run;
            run += c;
        }
        if (!EARLY) {
            if (tid == 255) count -= (u32)(TILE - n);  // padding is not data
            lb_store(lbcol + rel * 256, lb_pack(rel == 0 ? LB_PREFIX : LB_AGG, count, epoch));
            // first window of predecessors: issued now, consumed after the reorder below
#pragma unroll
            for (int i = 0; i < LBW; i++) win[i] = (rel > (u64)i) ? lb_load(lbcol + (rel - 1 - i) * 256) : 0ull;
        }
    }
    PHASE_MARK(4);  // scans + publish
    __syncthreads();
    // reorder through shared memory
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const u32 d = key_digit(keys[r], shift);
        const u32 pos = wcnt[warp * 256 + d] + ((rank2[r >> 1] >> (16 * (r & 1))) & 0xffffu);
        sorted[pos] = keys[r];
        if (PAY) sortedp[pos] = (unsigned short)(pay2[r >> 1] >> (16 * (r & 1)));
    }
    PHASE_MARK(5);  // reorder
    // finish the look-back (256 digit threads) unless it already ran right after the early count (EARLY 2)
    if (tid < 256) {
        if (EARLY != 2) excl = debug_nolb ? 0u : lookback_finish<LBW>(lbcol, rel, win, count, epoch);  // debug_nolb: timing experiment only (wrong output)
        glob_off[tid] = bin_base[((size_t)oseg * npass + pass_row) * 256 + tid] + excl - dstart;
    }
    PHASE_MARK(6);  // look-back (thread 0's column)
    __syncthreads();
    PHASE_MARK(7);  // wait for all look-backs
    Key *dst = out + seg_begin;
    unsigned short *pdst = PAY ? pout + seg_begin : nullptr;
    // coalesced store: position j of the sorted tile goes to segment offset glob_off[digit] + j
#pragma unroll 4
    for (u32 j = tid; j < n; j += BLOCK) {
        const Key key = sorted[j];
        const u32 o = glob_off[key_digit(key, shift)] + j;
        dst[o] = key;
        if (PAY) pdst[o] = sortedp[j];
    }
    PHASE_MARK(8);  // store
}

// ---- host side ---------------------------------------------------------------------------------------
// First pass over gathered pieces (onesweep_kernel: gather): d_in[s] = where piece s lies inside the input buffer; from the second pass on the
// keys are one contiguous segment described by d_off1 / d_tile1 (two entries each) with ntiles1 tiles.
struct RsGather {
    const u64 *d_in, *d_off1, *d_tile1;
    u64 ntiles1;
};
template <typename Key, int BLOCK, int ITEMS, int MINB, int MATCH, int PAY = 0, int EARLY = 0>
static int launch_passes(khb_ctx *ctx, Key *src, Key *dst, unsigned short *psrc, unsigned short *pdst, const u64 *d_off, const u64 *d_tile, int nseg, int npass, int first_bit,
                         u64 ntiles, u64 n_keys, const u32 *d_hist, u64 *d_lb, u32 *d_ticket, const RsGather *gat)
{
    constexpr int TILE = BLOCK * ITEMS;
    constexpr int NW = BLOCK / 32;
    const size_t shm = sizeof(Key) * TILE + NW * 256 * sizeof(unsigned short) + 256 * sizeof(u32) + 36 * sizeof(u32) +
                       (MATCH >= 2 ? NW * 256 * sizeof(u32) : 0) + (EARLY ? 256 * sizeof(u32) : 0) + (PAY ? TILE * sizeof(unsigned short) : 0);
    // the opt-in is a per-device function attribute: set on every call (cheap) so that every context on every GPU has it
    KHB_CUDA(ctx, cudaFuncSetAttribute(onesweep_kernel<Key, BLOCK, ITEMS, MINB, MATCH, PAY, EARLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
#ifdef KHB_EXPERIMENTS
    static int nolb = -1;   // timing experiment only (WRONG output): skip the look-back.  Not in the product build.
    if (nolb < 0) {
        const char *e = getenv("KHB_SORT_DEBUG_NOLB");
        nolb = e ? atoi(e) : 0;
    }
#else
    const int nolb = 0;
#endif
    for (int pass = 0; pass < npass; pass++) {
        khb_prof_begin(ctx, KHB_K_ONESWEEP);
        if (gat && pass > 0)
            onesweep_kernel<Key, BLOCK, ITEMS, MINB, MATCH, PAY, EARLY><<<(unsigned)gat->ntiles1, BLOCK, shm, ctx->stream>>>(
                src, dst, psrc, pdst, gat->d_off1, gat->d_tile1, 1, pass, first_bit + 8 * pass, npass, d_hist, d_lb, d_ticket + pass, (u32)(pass + 1), nolb, nullptr);
        else
            onesweep_kernel<Key, BLOCK, ITEMS, MINB, MATCH, PAY, EARLY><<<(unsigned)ntiles, BLOCK, shm, ctx->stream>>>(
                src, dst, psrc, pdst, d_off, d_tile, nseg, pass, first_bit + 8 * pass, npass, d_hist, d_lb, d_ticket + pass, (u32)(pass + 1), nolb, gat ? gat->d_in : nullptr);
        KHB_LAUNCH_CHECK(ctx);
        khb_prof_end(ctx, KHB_K_ONESWEEP, 2 * n_keys * (sizeof(Key) + (PAY ? 2 : 0)));  // read + write every key (+ payload) once
        Key *t = src; src = dst; dst = t;
        unsigned short *pt = psrc; psrc = pdst; pdst = pt;
    }
#ifdef KHB_PHASE_TIMING
    {
        unsigned long long h[16];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof(h));
        unsigned long long tot = 0;
        for (int i = 0; i < 9; i++) tot += h[i];
        const char *names[9] = {"ticket+seg", "load", "rank", "sync1", "scan+publish", "reorder", "lookback", "sync2", "store"};
        fprintf(stderr, "[phase] tiles*passes=%llu avg cycles/tile:", (unsigned long long)(ntiles * npass));
        for (int i = 0; i < 9; i++) fprintf(stderr, " %s=%.0f", names[i], (double)h[i] / (double)(ntiles * npass));
        fprintf(stderr, " total=%.0f | per tile: wide hops=%.2f wide spins=%.2f wide entries=%.2f first-window spins=%.2f\n", (double)tot / (double)(ntiles * npass),
                (double)h[10] / (ntiles * npass), (double)h[11] / (ntiles * npass), (double)h[12] / (ntiles * npass), (double)h[13] / (ntiles * npass));
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_phase_cycles, h, sizeof(h));
    }
#endif
    return KHB_OK;
}

static int sort_variant()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("KHB_SORT_VARIANT");
        v = e ? atoi(e) : 5;
    }
    return v;
}

// tile size (keys) of the scatter-pass variant `v` for key width W
static u32 variant_tile(int v, size_t W)
{
    if (W == 16) return v == 31 ? 512 * 6 : 512 * 8;
    // 64-bit keys: 14 keys per thread by default (7168-key tiles, 2 CTAs per SM); 12 for the tuning variants listed in
    // dispatch_passes, 16 for variant 41 (112 bytes of spills, measured slower)
    if (v == 41) return 512 * 16;
    return (v == 0 || v == 1 || v == 12 || v == 13 || v == 18 || v == 42) ? 512 * 12 : 512 * 14;
}

template <typename Key>
static int dispatch_passes(khb_ctx *ctx, int v, Key *src, Key *dst, unsigned short *psrc, unsigned short *pdst, const u64 *d_off,
                           const u64 *d_tile, int nseg, int npass, int first_bit, u64 ntiles, u64 n_keys, const u32 *d_hist,
                           u64 *d_lb, u32 *d_ticket, const RsGather *gat);

#define KHB_PASS_ARGS ctx, src, dst, psrc, pdst, d_off, d_tile, nseg, npass, first_bit, ntiles, n_keys, d_hist, d_lb, d_ticket, gat
template <>
int dispatch_passes<Key64>(khb_ctx *ctx, int v, Key64 *src, Key64 *dst, unsigned short *psrc, unsigned short *pdst, const u64 *d_off,
                           const u64 *d_tile, int nseg, int npass, int first_bit, u64 ntiles, u64 n_keys, const u32 *d_hist,
                           u64 *d_lb, u32 *d_ticket, const RsGather *gat)
{
    if (psrc) {  // keys + 16-bit payload
        if (v == 41) return launch_passes<Key64, 512, 16, 2, 2, 1>(KHB_PASS_ARGS);
        if (v == 13) return launch_passes<Key64, 512, 12, 2, 4, 1>(KHB_PASS_ARGS);
        if (v == 12) return launch_passes<Key64, 512, 12, 2, 3, 1>(KHB_PASS_ARGS);
        if (v == 0 || v == 1 || v == 18 || v == 42) return launch_passes<Key64, 512, 12, 2, 2, 1>(KHB_PASS_ARGS);
        return launch_passes<Key64, 512, 14, 2, 2, 1>(KHB_PASS_ARGS);  // default: 82.5 ms per config-2 step against 83.6 with 12 keys per thread
    }
    switch (v) {
    case 41: return launch_passes<Key64, 512, 16, 2, 2>(KHB_PASS_ARGS);
    case 42: return launch_passes<Key64, 512, 12, 2, 2>(KHB_PASS_ARGS);         // the default's ranking with 12 keys per thread
    case 0: return launch_passes<Key64, 512, 12, 2, 0>(KHB_PASS_ARGS);         // MATCH.ANY (1.46 TB/s: ~1 MATCH.ANY per 100 cycles per SM)
    case 1: return launch_passes<Key64, 512, 12, 2, 1>(KHB_PASS_ARGS);         // eight ballots (2.33 TB/s: ALU-bound, 3.9 warp-instr/key)
    case 13: return launch_passes<Key64, 512, 12, 2, 4>(KHB_PASS_ARGS);        // every 3rd round by ballots, the others by atomicOr (2.64 TB/s)
    case 18: return launch_passes<Key64, 512, 12, 2, 2, 0, 2>(KHB_PASS_ARGS);  // early count + early look-back (2.45 TB/s)
    default: return launch_passes<Key64, 512, 14, 2, 2>(KHB_PASS_ARGS);        // shared-memory atomicOr peer masks, 14 keys per thread -- default
    }
}

template <>
int dispatch_passes<Key128>(khb_ctx *ctx, int v, Key128 *src, Key128 *dst, unsigned short *psrc, unsigned short *pdst, const u64 *d_off,
                            const u64 *d_tile, int nseg, int npass, int first_bit, u64 ntiles, u64 n_keys, const u32 *d_hist,
                            u64 *d_lb, u32 *d_ticket, const RsGather *gat)
{
    if (v == 31) {  // 6 keys per thread (3072-key tiles): the earlier default, 3.4 % slower per pass (k = 47: 146.4 vs 141.5 ms per step)
        if (psrc) return launch_passes<Key128, 512, 6, 2, 2, 1>(KHB_PASS_ARGS);
        return launch_passes<Key128, 512, 6, 2, 2>(KHB_PASS_ARGS);
    }
    // 8 keys per thread: 4096-key tiles, 2 CTAs x 97 KB of shared memory per SM, 64 registers without spills
    if (psrc) return launch_passes<Key128, 512, 8, 2, 2, 1>(KHB_PASS_ARGS);
    return launch_passes<Key128, 512, 8, 2, 2>(KHB_PASS_ARGS);
}
#undef KHB_PASS_ARGS

template <typename Key>
static int sort_impl(khb_ctx *ctx, Key *d_keys, Key *d_tmp, unsigned short *d_pay, unsigned short *d_pay_tmp, const u64 *h_seg_off,
                     int nseg, int first_bit, int npass, int *result_in_tmp, int hist_ready, const u64 *h_gather = nullptr)
{
    // h_gather != nullptr: the nseg "segments" are pieces of ONE array to be sorted; piece s lies at d_keys + h_gather[s] (keys).  The
    // first pass reads them in place and writes d_tmp contiguously; from then on d_tmp and d_keys[0 .. n) ping-pong as usual (the caller
    // guarantees that d_keys holds n keys from its start and that the pieces may be overwritten once they are read).
    // hist_ready: the SCR_HIST scratch already holds the raw digit counts [1][npass][256] of the single segment (K2 counted
    // them while it produced the keys, khb_sort_hist_buffer), so the histogram sweep over the keys is skipped
    const int v = sort_variant();
    const u32 TILE = variant_tile(v, sizeof(Key));
    *result_in_tmp = 0;
    if (nseg <= 0) return KHB_OK;
    // tile table
    u64 *h_tab = (u64 *)malloc(sizeof(u64) * 2 * ((size_t)nseg + 1));
    if (!h_tab) return khb_fail(ctx, KHB_ERR_NOMEM, "sort: host table");
    u64 *h_off = h_tab, *h_tile = h_tab + nseg + 1;
    u64 ntiles = 0, n_keys = 0;
    for (int s = 0; s < nseg; s++) {
        if (h_seg_off[s + 1] < h_seg_off[s]) { free(h_tab); return khb_fail(ctx, KHB_ERR_ARG, "sort: segment offsets not monotone"); }
        if (h_seg_off[s + 1] - h_seg_off[s] >= (1ull << 32)) { free(h_tab); return khb_fail(ctx, KHB_ERR_ARG, "sort: segment %d has >= 2^32 keys", s); }
        h_off[s] = h_seg_off[s];
        h_tile[s] = ntiles;
        ntiles += div_up(h_seg_off[s + 1] - h_seg_off[s], TILE);
        n_keys += h_seg_off[s + 1] - h_seg_off[s];
    }
    h_off[nseg] = h_seg_off[nseg];
    h_tile[nseg] = ntiles;
    if (ntiles == 0) { free(h_tab); return KHB_OK; }
    if (ntiles >= (1ull << 32)) { free(h_tab); return khb_fail(ctx, KHB_ERR_ARG, "sort: too many tiles"); }

    void *p;
    int rc = khb_scratch_get(ctx, SCR_MISC, sizeof(u64) * 3 * ((size_t)nseg + 1) + 64 + 256, &p);
    if (rc) { free(h_tab); return rc; }
    u64 *d_off = (u64 *)p, *d_tile = d_off + nseg + 1;
    u64 *d_gat = d_tile + nseg + 1;              // [nseg + 1] piece positions, then the single-segment tables [2] + [2] (gathered input only)
    u32 *d_ticket = (u32 *)(d_gat + nseg + 1 + 8);  // 64 tickets max (one per pass)
    // synchronous small copy: h_tab is pageable and freed right after
    KHB_CUDA(ctx, cudaMemcpyAsync(d_off, h_tab, sizeof(u64) * 2 * ((size_t)nseg + 1), cudaMemcpyHostToDevice, ctx->stream));
    RsGather gat_s, *gat = nullptr;
    u64 h_one[4] = {0, n_keys, 0, div_up(n_keys, TILE)};
    if (h_gather) {
        if (d_pay) { free(h_tab); return khb_fail(ctx, KHB_ERR_ARG, "sort: gathered input carries no payload"); }
        KHB_CUDA(ctx, cudaMemcpyAsync(d_gat, h_gather, sizeof(u64) * (size_t)nseg, cudaMemcpyHostToDevice, ctx->stream));
        KHB_CUDA(ctx, cudaMemcpyAsync(d_gat + nseg + 1, h_one, sizeof(h_one), cudaMemcpyHostToDevice, ctx->stream));
        gat_s.d_in = d_gat;
        gat_s.d_off1 = d_gat + nseg + 1;
        gat_s.d_tile1 = d_gat + nseg + 3;
        gat_s.ntiles1 = h_one[3];
        gat = &gat_s;
    }
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free(h_tab);
    KHB_CUDA(ctx, cudaMemsetAsync(d_ticket, 0, 64 * sizeof(u32), ctx->stream));

    const size_t hist_bytes = (size_t)(gat ? 1 : nseg) * npass * 256 * sizeof(u32);
    rc = khb_scratch_get(ctx, SCR_HIST, hist_bytes, &p);
    if (rc) return rc;
    u32 *d_hist = (u32 *)p;
    if (hist_ready && nseg != 1) return khb_fail(ctx, KHB_ERR_ARG, "sort: precomputed histograms need a single segment");
    if (!hist_ready) KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, hist_bytes, ctx->stream));
    const size_t lb_bytes = (size_t)ntiles * 256 * sizeof(u64);
    rc = khb_scratch_get(ctx, SCR_LOOKBACK, lb_bytes, &p);
    if (rc) return rc;
    u64 *d_lb = (u64 *)p;
    KHB_CUDA(ctx, cudaMemsetAsync(d_lb, 0, lb_bytes, ctx->stream));

    {
        u64 grid = (u64)ctx->num_sms * 4;
        if (grid > ntiles) grid = ntiles;
        const size_t shm = (size_t)npass * 256 * sizeof(u32);
        khb_prof_begin(ctx, KHB_K_RADIX_HIST);
        if (!hist_ready) {
            radix_hist_kernel<Key><<<(unsigned)grid, RS_BLOCK, shm, ctx->stream>>>(d_keys, d_off, d_tile, nseg, ntiles, npass, first_bit, d_hist, TILE,
                                                                                   gat ? gat->d_in : nullptr);
            KHB_LAUNCH_CHECK(ctx);
        }
        const int nhist = (gat ? 1 : nseg) * npass;
        radix_scan_kernel<<<(unsigned)div_up(nhist, 8), 256, 0, ctx->stream>>>(d_hist, nhist);
        KHB_LAUNCH_CHECK(ctx);
        khb_prof_end(ctx, KHB_K_RADIX_HIST, hist_ready ? 0 : n_keys * sizeof(Key));
    }
    rc = dispatch_passes<Key>(ctx, v, d_keys, d_tmp, d_pay, d_pay_tmp, d_off, d_tile, nseg, npass, first_bit, ntiles, n_keys, d_hist, d_lb, d_ticket, gat);
    if (rc) return rc;
    *result_in_tmp = (npass & 1);
    return KHB_OK;
}

// Sort every segment by the `npass` 8-bit digits starting at bit `first_bit` (stable LSD).  d_pay / d_pay_tmp
// (optional, both or none): a 16-bit payload per key that travels with it (ping-pong like the keys).
// Zeroed [npass][256] digit-count buffer that a key producer may fill before khb_sort_bits_impl(..., hist_ready = 1).
int khb_sort_hist_buffer(khb_ctx *ctx, int npass, u32 **d_hist)
{
    void *p;
    const size_t bytes = (size_t)npass * 256 * sizeof(u32);
    int rc = khb_scratch_get(ctx, SCR_HIST, bytes, &p);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemsetAsync(p, 0, bytes, ctx->stream));
    *d_hist = (u32 *)p;
    return KHB_OK;
}

// One array given as `npieces` pieces inside d_buf (piece s: h_len[s] keys at d_buf + h_pos[s]; d_buf holds at least the sum of the lengths from
// its start, and the pieces may be overwritten): sorted by the digits like a single segment, without concatenating it first.
int khb_sort_gathered_impl(khb_ctx *ctx, void *d_buf, void *d_tmp, const u64 *h_pos, const u64 *h_len, int npieces, int key_bytes, int first_bit, int npass,
                           int *result_in_tmp)
{
    if ((key_bytes != 8 && key_bytes != 16) || npieces < 1 || npass < 1 || npass > 16 || first_bit < 0 || first_bit + 8 * npass > 8 * key_bytes + 7)
        return khb_fail(ctx, KHB_ERR_ARG, "sort (gathered): bad arguments");
    std::vector<u64> off((size_t)npieces + 1, 0);
    for (int s = 0; s < npieces; s++) off[s + 1] = off[s] + h_len[s];
    if (off[npieces] >= (1ull << 32)) return khb_fail(ctx, KHB_ERR_ARG, "sort (gathered): >= 2^32 keys");
    *result_in_tmp = 0;
    return key_bytes == 8 ? sort_impl<Key64>(ctx, (Key64 *)d_buf, (Key64 *)d_tmp, nullptr, nullptr, off.data(), npieces, first_bit, npass, result_in_tmp, 0, h_pos)
                          : sort_impl<Key128>(ctx, (Key128 *)d_buf, (Key128 *)d_tmp, nullptr, nullptr, off.data(), npieces, first_bit, npass, result_in_tmp, 0, h_pos);
}

int khb_sort_bits_impl(khb_ctx *ctx, void *d_keys, void *d_tmp, const u64 *h_seg_off, int nseg, int key_bytes, int first_bit, int npass,
                       int *result_in_tmp, unsigned short *d_pay, unsigned short *d_pay_tmp, int hist_ready)
{
    if (key_bytes != 8 && key_bytes != 16) return khb_fail(ctx, KHB_ERR_ARG, "sort: key_bytes=%d", key_bytes);
    if (npass < 0 || npass > 16 || first_bit < 0 || first_bit + 8 * npass > 8 * key_bytes + 7)
        return khb_fail(ctx, KHB_ERR_ARG, "sort: bad digit range first_bit=%d npass=%d", first_bit, npass);
    if ((d_pay == nullptr) != (d_pay_tmp == nullptr)) return khb_fail(ctx, KHB_ERR_ARG, "sort: payload needs both buffers");
    *result_in_tmp = 0;
    if (npass == 0) return KHB_OK;
    return key_bytes == 8 ? sort_impl<Key64>(ctx, (Key64 *)d_keys, (Key64 *)d_tmp, d_pay, d_pay_tmp, h_seg_off, nseg, first_bit, npass, result_in_tmp, hist_ready)
                          : sort_impl<Key128>(ctx, (Key128 *)d_keys, (Key128 *)d_tmp, d_pay, d_pay_tmp, h_seg_off, nseg, first_bit, npass, result_in_tmp, hist_ready);
}

// Full sort of k-mer words: all ceil(2k/8) digits.
int khb_sort_keys_impl(khb_ctx *ctx, void *d_keys, void *d_tmp, const u64 *h_seg_off, int nseg, int k, int *result_in_tmp)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_sort_keys: k=%d outside 1..64", k);
    return khb_sort_bits_impl(ctx, d_keys, d_tmp, h_seg_off, nseg, k <= 32 ? 8 : 16, 0, (2 * k + 7) / 8, result_in_tmp, nullptr, nullptr, 0);
}
