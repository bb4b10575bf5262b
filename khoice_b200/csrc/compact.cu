// compact.cu -- K4 adjacent-unique compaction and K5/K6 run-length count + count-of-counts histogram.
//
//   unique_kernel    : sorted keys -> distinct keys (sentinels dropped).  With per-genome segments laid
//                      end to end this yields the concatenation of the per-genome k-mer SETS, i.e. the
//                      result of `kmc` + `kmc_tools transform ... set_counts 1`
//                      (/root/reference/workflow/rules/exp_type_1.smk:156-173).  Every segment ends with
//                      sentinel keys (the window starting at the genome's trailing break symbol), so the
//                      first key of a segment never equals its predecessor and no segment table is needed.
//   rle_hist_kernel  : sorted keys -> for every run of equal keys its length c (saturated at `cs`, the
//                      `-cs5000` of exp_type_1.smk:61,84), hist[c]++, optionally the distinct keys and
//                      their counters.  On the concatenated per-genome sets of a group c is the number of
//                      genomes containing the k-mer = the counter `kmc_tools complex (set1+...+setN)`
//                      produces (exp_type_1.smk:175-182) and hist is what `kmc_tools transform ... histogram`
//                      prints (exp_type_1.smk:184-191); on the concatenated group sets it is the
//                      across-group table and histogram (exp_type_1.smk:243-259).
//
// Both kernels are single-pass: persistent CTAs take tiles from an atomic ticket, flags are turned into
// positions with warp ballots + popc (keys stay in registers in warp-striped order), and the running
// output offset is chained from tile to tile with decoupled look-back (lookback.cuh).
// Algorithmic bytes: unique W*n + W*distinct;  rle W*n (+ W*runs if keys are emitted, + 4*runs for counters).
#include "khb_common.cuh"
#include "lookback.cuh"
#include <stdlib.h>

#define CP_BLOCK 512
#define CP_WARPS (CP_BLOCK / 32)
#define CP_ITEMS 8
#define CP_TILE (CP_BLOCK * CP_ITEMS)
// count_prefix_kernel: smaller CTAs, four per SM (latency-bound phases overlap across CTAs)
#ifndef CQ_BLOCK
#define CQ_BLOCK 128
#endif
#define CQ_WARPS (CQ_BLOCK / 32)
#define CQ_TILE (CQ_BLOCK * CP_ITEMS)

__device__ __forceinline__ Key64 shfl_key(const Key64 &k, int src)
{
    return Key64{__shfl_sync(0xffffffffu, k.v, src)};
}
__device__ __forceinline__ Key128 shfl_key(const Key128 &k, int src)
{
    return Key128{__shfl_sync(0xffffffffu, k.lo, src), __shfl_sync(0xffffffffu, k.hi, src)};
}
template <typename Key> __device__ __forceinline__ Key sentinel_key();
template <> __device__ __forceinline__ Key64 sentinel_key<Key64>() { return Key64{~0ull}; }
template <> __device__ __forceinline__ Key128 sentinel_key<Key128>() { return Key128{~0ull, ~0ull}; }

// Load a tile warp-striped; keys past n become sentinels.
template <typename Key>
__device__ __forceinline__ void load_tile(const Key *__restrict__ in, u64 begin, u64 n, u32 wbase, u32 lane, Key (&keys)[CP_ITEMS])
{
#pragma unroll
    for (int r = 0; r < CP_ITEMS; r++) {
        const u64 g = begin + wbase + r * 32 + lane;
        keys[r] = g < n ? in[g] : sentinel_key<Key>();
    }
}

// Key at global index g-1 for the element held by (round r, lane); `before` is the key preceding the
// warp's chunk (only read when the chunk does not start at index 0).
template <typename Key>
__device__ __forceinline__ Key prev_key(const Key (&keys)[CP_ITEMS], int r, u32 lane, const Key &before)
{
    Key up = shfl_key(keys[r], (int)((lane + 31) & 31));         // lane-1 (lane 0 gets lane 31: replaced below)
    Key last = r > 0 ? shfl_key(keys[r > 0 ? r - 1 : 0], 31) : before;
    return lane == 0 ? last : up;
}

template <typename Key>
__global__ void __launch_bounds__(CP_BLOCK)
unique_kernel(const Key *__restrict__ in, u64 n, Key *__restrict__ out, u64 *__restrict__ lookback,
              u32 *__restrict__ ticket, u32 epoch, u64 *__restrict__ d_count)
{
    __shared__ u64 ws[33];
    __shared__ u32 s_tile;
    __shared__ u64 s_base;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u64 ntiles = (n + CP_TILE - 1) / CP_TILE;
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const u64 tile = s_tile;
        if (tile >= ntiles) break;
        const u64 begin = tile * CP_TILE;
        const u32 wbase = warp * (32 * CP_ITEMS);
        Key keys[CP_ITEMS];
        load_tile(in, begin, n, wbase, lane, keys);
        const u64 chunk0 = begin + wbase;
        Key before = sentinel_key<Key>();
        if (chunk0 > 0 && chunk0 < n) before = in[chunk0 - 1];
        u32 ball[CP_ITEMS];
        u32 wcount = 0;
#pragma unroll
        for (int r = 0; r < CP_ITEMS; r++) {
            const u64 g = chunk0 + r * 32 + lane;
            const Key pk = prev_key(keys, r, lane, before);
            const bool keep = g < n && !key_is_sentinel(keys[r]) && (g == 0 || !key_eq(keys[r], pk));
            ball[r] = __ballot_sync(0xffffffffu, keep);
            wcount += __popc(ball[r]);
        }
        u64 total;
        const u64 woff = block_excl_sum<u64>(lane == 0 ? (u64)wcount : 0ull, ws, &total);
        // woff is only meaningful on lane 0 of each warp; broadcast
        const u64 warp_off = __shfl_sync(0xffffffffu, woff, 0);
        if (warp == 0) {
            u64 excl = 0;
            if (tile == 0) {
                if (lane == 0) lb_store(lookback, lb_pack(LB_PREFIX, total, epoch));
            } else {
                if (lane == 0) lb_store(lookback + tile, lb_pack(LB_AGG, total, epoch));
                excl = lb_walk_warp(lookback, tile, 0, epoch);
                if (lane == 0) lb_store(lookback + tile, lb_pack(LB_PREFIX, excl + total, epoch));
            }
            if (lane == 0) {
                s_base = excl;
                if (tile == ntiles - 1) *d_count = excl + total;
            }
        }
        __syncthreads();
        u64 pos = s_base + warp_off;
#pragma unroll
        for (int r = 0; r < CP_ITEMS; r++) {
            if ((ball[r] >> lane) & 1u) out[pos + __popc(ball[r] & lanemask_lt())] = keys[r];
            pos += __popc(ball[r]);
        }
        __syncthreads();  // s_tile / s_base are rewritten by the next iteration
    }
}

// Index of the first element of the run of `key` that reaches back from index `at` (exclusive): the
// smallest j <= at such that in[j..at) are all equal to key.  Warp-cooperative.
template <typename Key>
__device__ __forceinline__ u64 run_head_before(const Key *__restrict__ in, u64 at, const Key &key, u32 lane)
{
    u64 j = at;
    while (j > 0) {
        const bool have = j > lane;
        bool same = false;
        if (have) same = key_eq(in[j - 1 - lane], key);
        const u32 diff = __ballot_sync(0xffffffffu, !same);  // lanes beyond index 0 count as "different"
        if (diff) {
            j -= (u32)(__ffs(diff) - 1);
            break;
        }
        j -= 32;
    }
    return j;
}

template <typename Key>
__global__ void __launch_bounds__(CP_BLOCK)
rle_hist_kernel(const Key *__restrict__ in, u64 n, u32 cs, u32 nbins, u64 *__restrict__ hist,
                Key *__restrict__ out_keys, u32 *__restrict__ out_counts, u64 *__restrict__ lookback,
                u32 *__restrict__ ticket, u32 epoch, u64 *__restrict__ d_runs)
{
    extern __shared__ u32 sh_hist[];  // [nbins+1]
    __shared__ u64 ws[33];
    __shared__ u32 s_tile;
    __shared__ u64 s_base;
    __shared__ u64 s_head0;            // global index of the head of the run that is open at the tile start
    __shared__ u32 s_wlast[CP_WARPS];  // per warp: local index+1 of its last head, 0 = none
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u64 ntiles = (n + CP_TILE - 1) / CP_TILE;
    const bool emit = out_keys != nullptr || out_counts != nullptr;
    for (u32 i = tid; i <= nbins; i += CP_BLOCK) sh_hist[i] = 0;
    u64 my_runs = 0;
    for (;;) {
        if (tid == 0) s_tile = atomicAdd(ticket, 1u);
        __syncthreads();
        const u64 tile = s_tile;
        if (tile >= ntiles) break;
        const u64 begin = tile * CP_TILE;
        const u32 wbase = warp * (32 * CP_ITEMS);
        Key keys[CP_ITEMS];
        load_tile(in, begin, n, wbase, lane, keys);
        const u64 chunk0 = begin + wbase;
        Key before = sentinel_key<Key>();
        if (chunk0 > 0 && chunk0 < n) before = in[chunk0 - 1];
        const u64 after_idx = chunk0 + 32 * CP_ITEMS;
        Key after = sentinel_key<Key>();
        if (after_idx < n) after = in[after_idx];
        if (warp == 0) {
            // the run open at the tile start (only matters if the first key continues it)
            Key k0 = shfl_key(keys[0], 0);
            u64 h0 = begin;
            if (begin > 0 && begin < n && !key_is_sentinel(k0)) h0 = run_head_before(in, begin, k0, lane);
            if (lane == 0) s_head0 = h0;
        }
        u32 hball[CP_ITEMS], tball[CP_ITEMS], hrun[CP_ITEMS];
        u32 carry = 0, wheads = 0;
#pragma unroll
        for (int r = 0; r < CP_ITEMS; r++) {
            const u64 g = chunk0 + r * 32 + lane;
            const Key pk = prev_key(keys, r, lane, before);
            Key dn = shfl_key(keys[r], (int)((lane + 1) & 31));
            Key nxt = r + 1 < CP_ITEMS ? shfl_key(keys[r + 1 < CP_ITEMS ? r + 1 : r], 0) : after;
            const Key nk = lane == 31 ? nxt : dn;
            const bool valid = g < n && !key_is_sentinel(keys[r]);
            const bool head = valid && (g == 0 || !key_eq(keys[r], pk));
            const bool tail = valid && (g + 1 >= n || !key_eq(keys[r], nk));
            hball[r] = __ballot_sync(0xffffffffu, head);
            tball[r] = __ballot_sync(0xffffffffu, tail);
            wheads += __popc(hball[r]);
            const u32 loc = wbase + r * 32 + lane + 1;  // local index + 1
            u32 h = warp_incl_max<u32>(head ? loc : 0u);
            h = h > carry ? h : carry;
            hrun[r] = h;
            carry = __shfl_sync(0xffffffffu, h, 31);
        }
        if (lane == 0) s_wlast[warp] = carry;
        u64 total;
        const u64 woff = block_excl_sum<u64>(lane == 0 ? (u64)wheads : 0ull, ws, &total);  // syncs: s_wlast, s_head0 visible
        const u64 warp_off = __shfl_sync(0xffffffffu, woff, 0);
        if (emit && warp == 0) {
            u64 excl = 0;
            if (tile == 0) {
                if (lane == 0) lb_store(lookback, lb_pack(LB_PREFIX, total, epoch));
            } else {
                if (lane == 0) lb_store(lookback + tile, lb_pack(LB_AGG, total, epoch));
                excl = lb_walk_warp(lookback, tile, 0, epoch);
                if (lane == 0) lb_store(lookback + tile, lb_pack(LB_PREFIX, excl + total, epoch));
            }
            if (lane == 0) s_base = excl;
        }
        if (tid == 0) my_runs += total;
        u32 wprefix = 0;  // last head (local index+1) in earlier warps of this tile
        for (u32 w = 0; w < warp; w++) wprefix = s_wlast[w] > wprefix ? s_wlast[w] : wprefix;
        const u64 head0 = s_head0;
        if (emit) __syncthreads();  // s_base
        u64 pos = (emit ? s_base : 0ull) + warp_off;  // heads before this warp's chunk (global ordinal)
#pragma unroll
        for (int r = 0; r < CP_ITEMS; r++) {
            const u64 g = chunk0 + r * 32 + lane;
            const u32 below = __popc(hball[r] & lanemask_lt());
            const bool head = (hball[r] >> lane) & 1u;
            if ((tball[r] >> lane) & 1u) {
                const u32 h = hrun[r] ? hrun[r] : wprefix;
                const u64 hg = h ? begin + (h - 1) : head0;
                u64 len = g - hg + 1;
                const u32 c = len > (u64)cs ? cs : (u32)len;
                if (c <= nbins) atomicAdd(&sh_hist[c], 1u);
                if (out_counts) out_counts[pos + below + (head ? 1u : 0u) - 1u] = c;
            }
            if (head && out_keys) out_keys[pos + below] = keys[r];
            pos += __popc(hball[r]);
        }
        __syncthreads();
    }
    __syncthreads();
    for (u32 i = tid; i <= nbins; i += CP_BLOCK) {
        const u32 c = sh_hist[i];
        if (c) atomicAdd(&hist[i], (u64)c);
    }
    if (tid == 0 && my_runs) atomicAdd(d_runs, my_runs);
}

// ---- prefix-run resolution (fused path) -------------------------------------------------------------------
// The fused path sorts hashed keys by a PREFIX only (bits >= pshift, see khb_prefix_plan), which costs 3-4 radix
// passes instead of ceil(2k/8).  Keys with equal prefix are adjacent ("prefix run") but not ordered among
// themselves, so equality inside a run is resolved by comparison.  Sentinels never share a prefix with a real key
// (the plan covers the spare bit above 2k) and every genome segment ends with at least one sentinel, so a scan
// never leaves its segment.
// K4 / K5 / K6 on prefix-sorted input.  Blocked arrangement: every thread owns CQ_ITEMS CONSECUTIVE keys plus its
// predecessor and successor, so heads / tails of adjacent runs and the prefix checks are register-to-register
// compares; ONE warp max-scan per thread (not per key) carries the position of the last run head across threads,
// and warp 0 finds the head of the run that is open at the tile start (run_head_before).  Every adjacent run
// [h, t] of equal keys is accounted at its TAIL t in O(1):
//   len    = t - h + 1
//   first  = no key equal to in[t] precedes h inside the prefix run   (one compare unless the run is mixed)
//   extra  = equal keys after t inside the prefix run                  (nothing to do unless the run is mixed)
// A prefix run is "mixed" when it holds more than one distinct value -- rare, because the prefix of a hashed key is
// uniform and the plan gives it more slots than there are keys.  Runs that are `first` are emitted (the distinct
// keys) and, if COUNT, add hist[min(len + extra, cs)].  A tile reserves its output range with ONE atomicAdd, so no
// tile waits for another: the distinct keys come out in no particular order, which is all a set needs (the next
// stage re-sorts them); histograms and counts are exact and deterministic.
#define CQ_ITEMS 8

// Blocked tile load: thread `tid` gets keys [l0-1, l0+ITEMS] of the tile at `base` (predecessor, own, successor).
// The own keys are 64 / 128 contiguous, 16-byte aligned bytes -> 128-bit loads.
__device__ __forceinline__ void load_blocked(const Key64 *__restrict__ base, u32 l0, Key64 (&k)[CQ_ITEMS + 2])
{
    const ulonglong2 *v = (const ulonglong2 *)(base + l0);
#pragma unroll
    for (int j = 0; j < CQ_ITEMS / 2; j++) {
        const ulonglong2 x = v[j];
        k[1 + 2 * j].v = x.x;
        k[2 + 2 * j].v = x.y;
    }
    k[0] = base[(int)l0 - 1];
    k[CQ_ITEMS + 1] = base[l0 + CQ_ITEMS];
}
__device__ __forceinline__ void load_blocked(const Key128 *__restrict__ base, u32 l0, Key128 (&k)[CQ_ITEMS + 2])
{
#pragma unroll
    for (int j = 0; j < CQ_ITEMS + 2; j++) k[j] = base[(int)l0 - 1 + j];
}

template <typename Key, bool COUNT>
__global__ void __launch_bounds__(CQ_BLOCK, (sizeof(Key) == 8 ? 1024 : 512) / CQ_BLOCK)
runs_kernel(const Key *__restrict__ in, u64 n, int pshift, u32 cs, u32 nbins, u64 *__restrict__ hist,
            Key *__restrict__ out_keys, u64 *__restrict__ d_cursor)
{
    constexpr int TILE = CQ_BLOCK * CQ_ITEMS;
    extern __shared__ u32 sh_hist[];  // [nbins+1] when COUNT
    __shared__ int s_head0;           // head of the run open at the tile start, relative to the tile (<= 0)
    __shared__ u32 s_wlast[CQ_WARPS];
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u64 ntiles = (n + TILE - 1) / TILE;
    if (COUNT) {
        for (u32 i = tid; i <= nbins; i += CQ_BLOCK) sh_hist[i] = 0;
        __syncthreads();
    }
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const u64 begin = tile * TILE;
        const u32 nloc = (u32)(n - begin < (u64)TILE ? n - begin : (u64)TILE);  // keys of this tile
        const bool last_tile = begin + TILE >= n;
        const u32 l0 = tid * CQ_ITEMS;  // local index of this thread's first key
        const Key *base = in + begin;
        Key k[CQ_ITEMS + 2];            // k[0] = predecessor, k[1..ITEMS] = own keys, k[ITEMS+1] = successor
        if (begin > 0 && !last_tile) {
            load_blocked(base, l0, k);
        } else {
#pragma unroll
            for (int j = 0; j < CQ_ITEMS + 2; j++) {
                const u64 g = begin + l0 + j;  // index + 1
                k[j] = (g >= 1 && g - 1 < n) ? in[g - 1] : sentinel_key<Key>();
            }
        }
        if (warp == 0) {
            // head of the run that is open at the tile start (only matters if the first key continues it).  The 32
            // keys in front of the tile are fetched together with the tile itself, so the common case (runs shorter
            // than 32 keys) costs no extra memory round trip.
            Key back = sentinel_key<Key>();
            if (begin > lane) back = in[begin - 1 - lane];
            const Key k0 = shfl_key(k[1], 0);
            int h0 = 0;
            if (begin > 0 && !key_is_sentinel(k0)) {
                const u32 diff = __ballot_sync(0xffffffffu, !(begin > lane && key_eq(back, k0)));
                if (diff) h0 = -(int)(__ffs(diff) - 1);
                else h0 = (int)((long long)run_head_before(in, begin - 32, k0, lane) - (long long)begin);
            }
            if (lane == 0) s_head0 = h0;
        }
        u32 headm = 0, tailm = 0, lasth = 0;  // lasth = local index + 1 of the thread's last head
#pragma unroll
        for (int j = 0; j < CQ_ITEMS; j++) {
            const u32 l = l0 + j;
            const bool valid = l < nloc && !key_is_sentinel(k[j + 1]);
            const bool head = valid && ((begin == 0 && l == 0) || !key_eq(k[j + 1], k[j]));
            const bool tail = valid && (l + 1 >= nloc ? (last_tile || !key_eq(k[j + 1], k[j + 2])) : !key_eq(k[j + 1], k[j + 2]));
            headm |= (head ? 1u : 0u) << j;
            tailm |= (tail ? 1u : 0u) << j;
            if (head) lasth = l + 1;
        }
        // last head before this thread: exclusive max-scan over the CTA
        const u32 inc = warp_incl_max<u32>(lasth);
        u32 carry = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) carry = 0;
        if (lane == 31) s_wlast[warp] = inc;
        __syncthreads();  // s_wlast, s_head0
        u32 wprefix = 0;
        for (u32 w = 0; w < warp; w++) wprefix = s_wlast[w] > wprefix ? s_wlast[w] : wprefix;
        carry = carry > wprefix ? carry : wprefix;
        const int head0 = s_head0;
        // resolve every tail
        u32 emitm = 0;
        int cur = carry ? (int)carry - 1 : head0;  // local index of the head of the run open at key j
        bool own = false;                          // that head belongs to this thread -> its predecessor is in `pred`
        Key pred = k[0];
#pragma unroll
        for (int j = 0; j < CQ_ITEMS; j++) {
            if ((headm >> j) & 1u) { cur = (int)(l0 + j); own = true; pred = k[j]; }
            if ((tailm >> j) & 1u) {
                const Key key = k[j + 1];
                bool first = true;
                const long long hg = (long long)begin + cur;  // global index of the run's head
                if (hg > 0) {
                    const Key pk = own ? pred : in[hg - 1];
                    if (!key_is_sentinel(pk) && same_prefix(pk, key, pshift)) {
                        u64 q = (u64)hg - 1;  // in[q] differs from key but shares its prefix: mixed run, scan back
                        while (q > 0) {
                            --q;
                            const Key kq = in[q];
                            if (!same_prefix(kq, key, pshift)) break;
                            if (key_eq(kq, key)) { first = false; break; }
                        }
                    }
                }
                if (first) {
                    emitm |= 1u << j;
                    if (COUNT) {
                        u32 len = (u32)((int)(l0 + j) - cur + 1);
                        const Key nk = k[j + 2];
                        const u64 g = begin + l0 + j;
                        if (g + 1 < n && !key_is_sentinel(nk) && same_prefix(nk, key, pshift)) {
                            for (u64 r = g + 2; r < n; r++) {  // later occurrences inside the prefix run
                                const Key kr = in[r];
                                if (!same_prefix(kr, key, pshift)) break;
                                len += key_eq(kr, key) ? 1u : 0u;
                            }
                        }
                        const u32 c = len > cs ? cs : len;
                        if (c <= nbins) atomicAdd(&sh_hist[c], 1u);
                    }
                }
            }
        }
        // every WARP reserves its own output range with one atomicAdd (no CTA-wide scan, no CTA-wide wait)
        const u32 mine = __popc(emitm);
        const u32 incl = warp_incl_sum<u32>(mine);
        const u32 wtotal = __shfl_sync(0xffffffffu, incl, 31);
        u64 wbase = 0;
        if (lane == 31 && wtotal) wbase = atomicAdd(d_cursor, (u64)wtotal);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        if (out_keys != nullptr && mine) {
            u64 pos = wbase + (incl - mine);
#pragma unroll
            for (int j = 0; j < CQ_ITEMS; j++)
                if ((emitm >> j) & 1u) out_keys[pos++] = k[j + 1];
        }
        __syncthreads();  // s_wlast / s_head0 are rewritten by the next iteration
    }
    if (COUNT) {
        __syncthreads();
        for (u32 i = tid; i <= nbins; i += CQ_BLOCK) {
            const u32 c = sh_hist[i];
            if (c) atomicAdd(&hist[i], (u64)c);
        }
    }
}

// ---- single-sort group path: keys with a genome-id payload ------------------------------------------------------
// Instead of a per-genome sort + dedup followed by a group sort, the windows of ALL genomes of a group are sorted
// once (stable, prefix only) together with a 16-bit genome id.  Equal keys then sit next to each other in genome
// order, so the number of genomes containing a k-mer is the number of (key, genome) CHANGES inside its run:
//   F[i] = key[i] != key[i-1]  ||  gid[i] != gid[i-1]          ("new pair")
//   c(x) = sum of F over the run of x          (= kmc per genome, set_counts 1, union-sum -- exp_type_1.smk:156-182)
//
// Two kernels:
//   pairs_kernel       handles every CLEAN run -- a run of equal keys whose neighbours on both sides have a different
//                      prefix -- with nothing but register compares: blocked arrangement (CQ_ITEMS consecutive keys per
//                      thread), relations between neighbours collected as bit masks, one segmented warp scan per thread
//                      carries (pair count, mixed flag) of the run that is open at a thread boundary, and the run open
//                      at the chunk start is recovered from the keys in front of the chunk (cooperative look-back, 32
//                      keys per step), so every warp is independent.  Keys of MIXED prefix runs (more than one distinct
//                      value under one prefix: ~1-2 % of the runs, because the hashed prefix has more slots than there
//                      are keys) are skipped; the tail of the first run of every mixed prefix run is marked in a bitmap.
//   mixed_runs_kernel  compacts the bitmap per CTA and resolves one mixed prefix run per thread with a small register
//                      table of its distinct keys (count, last genome); more than MIXED_MAXD distinct keys under one
//                      prefix take an exact quadratic fallback.
// Both add hist[min(c, cs)], emit the distinct keys (unordered) and return sum c = the sum of the per-genome set sizes.
__device__ __forceinline__ u32 seg_combine(u32 a, u32 b)
{
    // state = bit31: a run head was seen | bit30: that head's predecessor shares its prefix | low 30 bits: pair count
    return (b >> 31) ? b : ((a & 0xC0000000u) | ((a + b) & 0x3FFFFFFFu));
}

// Relation bits between two neighbouring keys: bit 0 = they differ, bit 1 = they differ but share the prefix.
__device__ __forceinline__ u32 key_relation(const Key64 &a, const Key64 &b, int pshift)
{
    const u64 x = a.v ^ b.v;
    const u32 ne = x != 0 ? 1u : 0u;
    const u32 sp = (x >> pshift) == 0 ? 2u : 0u;
    return ne | (ne ? sp : 0u);
}
__device__ __forceinline__ u32 key_relation(const Key128 &a, const Key128 &b, int pshift)
{
    const u32 ne = key_eq(a, b) ? 0u : 1u;
    return ne | ((ne && same_prefix(a, b, pshift)) ? 2u : 0u);
}

// One warp chunk in registers: CQ_ITEMS consecutive keys per lane with both neighbours, their genome ids (two per
// word), and the first step of the look-back in front of the chunk (lane i: key begin-1-i and its predecessor).
template <typename Key> struct PairChunk {
    Key k[CQ_ITEMS + 2];          // k[0] = predecessor, k[1..ITEMS] = own keys, k[ITEMS+1] = successor
    u32 gw[CQ_ITEMS / 2], g0w;    // own genome ids, packed; the word in front of them (predecessor's id on top)
    Key kp, kq;                   // in[begin-1-lane], in[begin-2-lane]   (sentinel where there is none)
    unsigned short gp, gq;        // kept as loaded: widening at the point of use keeps the loads free of consumers
};

template <typename Key>
__device__ __forceinline__ void pair_chunk_load(PairChunk<Key> &c, const Key *__restrict__ in, const unsigned short *__restrict__ gid,
                                                u64 n, u64 tile, u32 lane)
{
    constexpr int TILE = 32 * CQ_ITEMS;
    const u64 begin = tile * TILE;
    const u32 l0 = lane * CQ_ITEMS;
    if (begin > 0 && begin + TILE < n) {
        const unsigned short *gb = gid + begin;
        load_blocked(in + begin, l0, c.k);
        const uint4 gv = *(const uint4 *)(gb + l0);  // 8 own genome ids: 16 aligned bytes
        c.g0w = *(const u32 *)(gb + (int)l0 - 2);     // 4-byte aligned: begin and l0 are multiples of 8
        c.gw[0] = gv.x; c.gw[1] = gv.y; c.gw[2] = gv.z; c.gw[3] = gv.w;
    } else {
        // first / last chunk: indices outside [0, n) read as sentinels, which never equal a real key
        u32 g[CQ_ITEMS + 1];
#pragma unroll
        for (int j = 0; j < CQ_ITEMS + 2; j++) {
            const u64 x = begin + l0 + j;  // index + 1
            const bool ok = x >= 1 && x - 1 < n;
            c.k[j] = ok ? in[x - 1] : sentinel_key<Key>();
            if (j <= CQ_ITEMS) g[j] = ok ? (u32)gid[x - 1] : 0u;
        }
        c.g0w = g[0] << 16;
#pragma unroll
        for (int j = 0; j < CQ_ITEMS / 2; j++) c.gw[j] = g[1 + 2 * j] | (g[2 + 2 * j] << 16);
    }
    c.kp = sentinel_key<Key>();
    c.kq = sentinel_key<Key>();
    c.gp = 0;
    c.gq = 0;
    if (begin > lane) {
        const u64 p = begin - 1 - lane;
        c.kp = in[p];
        c.gp = gid[p];
        if (p > 0) { c.kq = in[p - 1]; c.gq = gid[p - 1]; }
    }
}

// State (seg_combine encoding) of the run of `k0` that ends right in front of index `begin` (> 0): number of
// (key, genome) pairs in it and whether the predecessor of its head shares its prefix.  0 if in[begin-1] != k0.
// Warp-cooperative, 32 keys per step; the first step's keys come preloaded with the chunk, further steps (only for
// runs that reach back more than 32 keys) load their own.
template <typename Key>
__device__ __forceinline__ u32 open_run_state(const Key *__restrict__ in, const unsigned short *__restrict__ gid, u64 begin,
                                              const Key &k0, int pshift, u32 lane, Key kp, Key kq, u32 gp, u32 gq)
{
    u32 cnt = 0;
    for (u64 top = begin;;) {            // lanes look at p = top-1-lane (kp, gp) and its predecessor p-1 (kq, gq)
        const bool have = top > lane;
        const u64 p = top - 1 - lane;
        const bool eq = have && key_eq(kp, k0);
        const u32 neq = __ballot_sync(0xffffffffu, !eq);
        const u32 inrun = neq ? ((neq & (0u - neq)) - 1u) : 0xffffffffu;  // lanes before the first mismatch
        const bool mine = (inrun >> lane) & 1u;
        const bool head = mine && (p == 0 || !key_eq(kq, k0));
        cnt += __popc(__ballot_sync(0xffffffffu, mine && (head || gp != gq)));
        const u32 headb = __ballot_sync(0xffffffffu, head);
        if (headb || neq) {
            if (cnt == 0) return 0u;
            const u32 mix = __ballot_sync(0xffffffffu, head && p > 0 && !key_is_sentinel(kq) && same_prefix(kq, k0, pshift));
            return 0x80000000u | (mix ? 0x40000000u : 0u) | cnt;
        }
        top -= 32;  // all 32 keys belong to the run and none is its head: keep walking
        kp = sentinel_key<Key>();
        kq = sentinel_key<Key>();
        gp = 0;
        gq = 0xffffffffu;
        if (top > lane) {
            const u64 q = top - 1 - lane;
            kp = in[q];
            gp = gid[q];
            if (q > 0) { kq = in[q - 1]; gq = gid[q - 1]; }
        }
    }
}

// PF: the next chunk's loads are issued before the current chunk is processed (more registers, fewer resident warps).
template <typename Key, bool PF> struct PairsCfg { static constexpr int MINB = sizeof(Key) == 8 ? (PF ? 5 : 8) : (PF ? 3 : 4); };
// PIVOT (experiment type 2, exp_type_2.smk:354-380): genome `pivot_gid` -- the LAST of the group -- is the pivot and the
// other genomes are the rest of the set.  A run whose last pair belongs to the pivot is a pivot k-mer: hist[c] counts
// it with c = 1 + (rest genomes containing it), i.e. c = 1 is `kmc_tools simple pivot rest kmers_subtract` and c >= 2 the
// counter of `... intersect -ocsum`; its key goes to out_pivot.  out_keys receives the union of the REST only (k-mers
// seen in nothing but the pivot are left out).
template <typename Key, bool PF, bool PIVOT>
__global__ void __launch_bounds__(CQ_BLOCK, (PairsCfg<Key, PF>::MINB))
pairs_kernel(const Key *__restrict__ in, const unsigned short *__restrict__ gid, u64 n, int pshift, u32 cs, u32 nbins, u32 hot_bin,
             u64 *__restrict__ hist, Key *__restrict__ out_keys, u64 *__restrict__ d_cursor, u64 *__restrict__ d_pairs,
             unsigned char *__restrict__ mixed_map /* one byte per lane chunk: bit j = key j ends the first run of a mixed prefix run */,
             u32 pivot_gid, Key *__restrict__ out_pivot, u64 *__restrict__ d_pcursor)
{
    constexpr int TILE = 32 * CQ_ITEMS;
    extern __shared__ u32 sh_hist[];  // [nbins+1]
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const u64 ntiles = (n + TILE - 1) / TILE;
    const u64 stride = (u64)gridDim.x * CQ_WARPS;
    for (u32 i = tid; i <= nbins; i += CQ_BLOCK) sh_hist[i] = 0;
    __syncthreads();
    u32 my_pairs = 0, my_ones = 0, my_hots = 0;  // multiplicities 1 and hot_bin are counted in registers
    u64 tile = (u64)blockIdx.x * CQ_WARPS + warp;
    PairChunk<Key> c;
    if (PF && tile < ntiles) pair_chunk_load(c, in, gid, n, tile, lane);
    while (tile < ntiles) {
        // PF: the next chunk's loads are in flight while this one is processed
        PairChunk<Key> nx;
        const u64 next = tile + stride;
        if (PF) {
            if (next < ntiles) pair_chunk_load(nx, in, gid, n, next, lane);
        } else {
            pair_chunk_load(c, in, gid, n, tile, lane);
        }
        const u64 begin = tile * TILE;
        u32 g[CQ_ITEMS + 1];
        g[0] = c.g0w >> 16;
#pragma unroll
        for (int j = 0; j < CQ_ITEMS / 2; j++) { g[1 + 2 * j] = c.gw[j] & 0xffffu; g[2 + 2 * j] = c.gw[j] >> 16; }
        // run open at the chunk start (lane 0's first key continues it)
        u32 open_state = 0;
        if (begin > 0) open_state = open_run_state(in, gid, begin, shfl_key(c.k[1], 0), pshift, lane, c.kp, c.kq, c.gp, c.gq);
        // relations between neighbours, one bit per own key:
        //   headm: differs from its predecessor       hmix: ... which shares its prefix
        //   tailm: differs from its successor         tmix: ... which shares its prefix
        //   fm   : starts a new (key, genome) pair
        u32 nem = 0, spm = 0, gdm = 0;
#pragma unroll
        for (int j = 0; j <= CQ_ITEMS; j++) {
            const u32 r = key_relation(c.k[j], c.k[j + 1], pshift);
            nem |= (r & 1u) << j;
            spm |= (r >> 1) << j;
            if (j < CQ_ITEMS) gdm |= (g[j] != g[j + 1] ? 1u : 0u) << j;
        }
        constexpr u32 OWN = (1u << CQ_ITEMS) - 1u;
        const u32 headm = nem & OWN, hmix = spm & OWN, tmix = (spm >> 1) & OWN;
        u32 tailm = (nem >> 1) & OWN;
        const u32 fm = headm | gdm;
        if (key_is_sentinel(c.k[CQ_ITEMS])) {  // the padding at the very end of the sorted array: never a result
#pragma unroll
            for (int j = 0; j < CQ_ITEMS; j++)
                if (key_is_sentinel(c.k[j + 1])) tailm &= ~(1u << j);
        }
        // segmented scan of (count, mixed) over threads
        u32 st;
        if (headm == 0) {
            st = __popc(fm);
        } else {
            const int hl = 31 - __clz(headm);
            st = 0x80000000u | (((hmix >> hl) & 1u) << 30) | (u32)__popc(fm >> hl);
        }
        u32 inc = st;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 up = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= (u32)o) inc = seg_combine(up, inc);
        }
        u32 excl = __shfl_up_sync(0xffffffffu, inc, 1);
        if (lane == 0) excl = 0;
        const u32 carry = seg_combine(open_state, excl);
        // resolve every tail
        u32 emitm = 0, itemm = 0, pemitm = 0;
        u32 open_cnt = carry & 0x3FFFFFFFu;
        u32 open_mix = (carry >> 30) & 1u;
#pragma unroll
        for (int j = 0; j < CQ_ITEMS; j++) {
            if ((headm >> j) & 1u) { open_cnt = 0; open_mix = (hmix >> j) & 1u; }
            open_cnt += (fm >> j) & 1u;
            if ((tailm >> j) & 1u) {
                const u32 after = (tmix >> j) & 1u;
                if ((open_mix | after) == 0) {
                    const bool pv = PIVOT && g[j + 1] == pivot_gid;
                    if (!PIVOT || !(pv && open_cnt == 1u)) emitm |= 1u << j;
                    if (pv) pemitm |= 1u << j;
                    my_pairs += open_cnt - (pv ? 1u : 0u);
                    if (!PIVOT || pv) {
                        if (open_cnt == 1u) my_ones++;
                        else if (open_cnt == hot_bin) my_hots++;
                        else {
                            const u32 cc = open_cnt > cs ? cs : open_cnt;
                            if (cc <= nbins) atomicAdd(&sh_hist[cc], 1u);
                        }
                    }
                } else if (open_mix == 0) {
                    itemm |= 1u << j;
                }
            }
        }
        mixed_map[tile * 32 + lane] = (unsigned char)itemm;
        // per warp: one atomicAdd reserves the output range
        const u32 mine = __popc(emitm);
        const u32 incl = warp_incl_sum<u32>(mine);
        const u32 wtotal = __shfl_sync(0xffffffffu, incl, 31);
        u64 wbase = 0;
        if (lane == 31 && wtotal) wbase = atomicAdd(d_cursor, (u64)wtotal);
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        if (out_keys != nullptr && mine) {
            Key *dst = out_keys + wbase + (incl - mine);
#pragma unroll
            for (int j = 0; j < CQ_ITEMS; j++)
                if ((emitm >> j) & 1u) *dst++ = c.k[j + 1];
        }
        if (PIVOT) {
            const u32 pmine = __popc(pemitm);
            const u32 pincl = warp_incl_sum<u32>(pmine);
            const u32 ptotal = __shfl_sync(0xffffffffu, pincl, 31);
            u64 pbase = 0;
            if (lane == 31 && ptotal) pbase = atomicAdd(d_pcursor, (u64)ptotal);
            pbase = __shfl_sync(0xffffffffu, pbase, 31);
            if (out_pivot != nullptr && pmine) {
                Key *dst = out_pivot + pbase + (pincl - pmine);
#pragma unroll
                for (int j = 0; j < CQ_ITEMS; j++)
                    if ((pemitm >> j) & 1u) *dst++ = c.k[j + 1];
            }
        }
        if (PF) c = nx;
        tile = next;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
        my_ones += __shfl_xor_sync(0xffffffffu, my_ones, o);
        my_hots += __shfl_xor_sync(0xffffffffu, my_hots, o);
    }
    if (lane == 0) {
        if (my_pairs) atomicAdd(d_pairs, (u64)my_pairs);
        if (my_ones && 1u <= nbins) atomicAdd(&sh_hist[1], my_ones);   // cs >= 1
        if (my_hots) atomicAdd(&sh_hist[hot_bin], my_hots);            // hot_bin <= min(cs, nbins) or 0 (never matched)
    }
    __syncthreads();
    for (u32 i = tid; i <= nbins; i += CQ_BLOCK) {
        const u32 cc = sh_hist[i];
        if (cc) atomicAdd(&hist[i], (u64)cc);
    }
}

#define MIXED_BLOCK 256
#define MIXED_MAXD 6
// Bitmap -> work list of the marked tails (positions), one list slot reservation per CTA and span.
__global__ void __launch_bounds__(MIXED_BLOCK)
mixed_collect_kernel(const u32 *__restrict__ mixed_map, u64 nwords, u32 *__restrict__ items, u64 *__restrict__ d_nitems)
{
    __shared__ u32 s_count;
    __shared__ u64 s_base;
    const u32 tid = threadIdx.x;
    for (u64 w0 = (u64)blockIdx.x * MIXED_BLOCK; w0 < nwords; w0 += (u64)gridDim.x * MIXED_BLOCK) {
        if (tid == 0) s_count = 0;
        __syncthreads();
        u32 bits = (w0 + tid < nwords) ? mixed_map[w0 + tid] : 0u;  // bit i <-> key 32 * (w0 + tid) + i
        u32 at = 0;
        if (bits) at = atomicAdd(&s_count, (u32)__popc(bits));
        __syncthreads();
        if (tid == 0 && s_count) s_base = atomicAdd(d_nitems, (u64)s_count);
        __syncthreads();
        if (bits) {
            u64 pos = s_base + at;
            while (bits) {
                const u32 b = __ffs(bits) - 1;
                bits &= bits - 1;
                items[pos++] = (u32)((w0 + tid) * 32 + b);
            }
        }
        __syncthreads();
    }
}

// Exact fallback for a mixed prefix run [s, e) with more than MIXED_MAXD distinct keys: every first occurrence counts
// its own pairs by scanning forward.  Quadratic in the run length; practically never taken for hashed keys.
template <typename Key, bool PIVOT>
__device__ __noinline__ void mixed_run_slow(const Key *__restrict__ in, const unsigned short *__restrict__ gid, u64 s, u64 e, u32 cs,
                                            u32 nbins, u32 *sh_hist, Key *__restrict__ out_keys, u64 *__restrict__ d_cursor,
                                            u64 *__restrict__ d_pairs, u32 pivot_gid, Key *__restrict__ out_pivot,
                                            u64 *__restrict__ d_pcursor)
{
    for (u64 i = s; i < e; i++) {
        const Key ki = in[i];
        bool first = true;
        for (u64 q = s; q < i && first; q++) first = !key_eq(in[q], ki);
        if (!first) continue;
        u32 cnt = 1, lastg = gid[i];
        for (u64 r = i + 1; r < e; r++) {
            if (!key_eq(in[r], ki)) continue;
            const u32 gr = gid[r];
            if (gr != lastg) { cnt++; lastg = gr; }
        }
        const u32 c = cnt > cs ? cs : cnt;
        const bool pv = PIVOT && lastg == pivot_gid;  // the pivot is the last genome, so it is the last one seen
        if ((!PIVOT || pv) && c <= nbins) atomicAdd(&sh_hist[c], 1u);
        atomicAdd(d_pairs, (u64)(cnt - (pv ? 1u : 0u)));
        if (!PIVOT || !(pv && cnt == 1u)) {
            const u64 pos = atomicAdd(d_cursor, 1ull);
            if (out_keys != nullptr) out_keys[pos] = ki;
        }
        if (pv) {
            const u64 pos = atomicAdd(d_pcursor, 1ull);
            if (out_pivot != nullptr) out_pivot[pos] = ki;
        }
    }
}

// One mixed prefix run per thread: a small register table of its distinct keys (pair count, last genome seen).
template <typename Key, bool PIVOT>
__global__ void __launch_bounds__(MIXED_BLOCK)
mixed_runs_kernel(const Key *__restrict__ in, const unsigned short *__restrict__ gid, u64 n, int pshift, u32 cs, u32 nbins,
                  u64 *__restrict__ hist, Key *__restrict__ out_keys, u64 *__restrict__ d_cursor, u64 *__restrict__ d_pairs,
                  const u32 *__restrict__ items, const u64 *__restrict__ d_nitems, u32 pivot_gid, Key *__restrict__ out_pivot,
                  u64 *__restrict__ d_pcursor)
{
    extern __shared__ u32 sh_hist[];            // [nbins+1]
    const u32 tid = threadIdx.x, lane = lane_id();
    const u64 count = *d_nitems;
    if ((u64)blockIdx.x * MIXED_BLOCK >= count) return;
    for (u32 i = tid; i <= nbins; i += MIXED_BLOCK) sh_hist[i] = 0;
    __syncthreads();
    for (u64 i0 = (u64)blockIdx.x * MIXED_BLOCK; i0 < count; i0 += (u64)gridDim.x * MIXED_BLOCK) {
        const u64 i = i0 + tid;
        Key lk[MIXED_MAXD];
        u32 lc[MIXED_MAXD], lg[MIXED_MAXD];
        u32 d = 0, pairs = 0;
        u32 gm = 0, pm = 0;  // table entries that go to the group set / to the pivot set
        if (i < count) {
            const u64 t = items[i];
            const Key key0 = in[t];
            u64 s = t;  // the first run of the prefix run is all key0 and starts the prefix run
            while (s > 0 && key_eq(in[s - 1], key0)) --s;
            bool overflow = false;
            u64 e = s;
            for (; e < n; e++) {
                const Key ke = in[e];
                if (key_is_sentinel(ke) || !same_prefix(ke, key0, pshift)) break;
                if (overflow) continue;  // only looking for the end of the prefix run
                const u32 ge = gid[e];
                bool found = false;
#pragma unroll
                for (int q = 0; q < MIXED_MAXD; q++) {
                    if ((u32)q < d && key_eq(lk[q], ke)) {
                        found = true;
                        if (lg[q] != ge) { lc[q]++; lg[q] = ge; }
                    }
                }
                if (!found) {
                    if (d < MIXED_MAXD) {
#pragma unroll
                        for (int q = 0; q < MIXED_MAXD; q++)
                            if ((u32)q == d) { lk[q] = ke; lc[q] = 1; lg[q] = ge; }
                        d++;
                    } else {
                        overflow = true;
                    }
                }
            }
            if (overflow) {
                mixed_run_slow<Key, PIVOT>(in, gid, s, e, cs, nbins, sh_hist, out_keys, d_cursor, d_pairs, pivot_gid, out_pivot, d_pcursor);
            } else {
#pragma unroll
                for (int q = 0; q < MIXED_MAXD; q++) {
                    if ((u32)q < d) {
                        const bool pv = PIVOT && lg[q] == pivot_gid;  // the pivot is the last genome of the group
                        pairs += lc[q] - (pv ? 1u : 0u);
                        if (!PIVOT || !(pv && lc[q] == 1u)) gm |= 1u << q;
                        if (pv) pm |= 1u << q;
                        const u32 c = lc[q] > cs ? cs : lc[q];
                        if ((!PIVOT || pv) && c <= nbins) atomicAdd(&sh_hist[c], 1u);
                    }
                }
            }
        }
        // per warp: one atomicAdd reserves each output range, one adds the pair count
        const u32 mine = (u32)__popc(gm) | ((u32)__popc(pm) << 16);
        const u32 incl = warp_incl_sum<u32>(mine);
        const u32 wtotal = __shfl_sync(0xffffffffu, incl, 31);
        u32 wpairs = pairs;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wpairs += __shfl_xor_sync(0xffffffffu, wpairs, o);
        u64 wbase = 0, pbase = 0;
        if (lane == 31) {
            if (wtotal & 0xffffu) wbase = atomicAdd(d_cursor, (u64)(wtotal & 0xffffu));
            if (PIVOT && (wtotal >> 16)) pbase = atomicAdd(d_pcursor, (u64)(wtotal >> 16));
            if (wpairs) atomicAdd(d_pairs, (u64)wpairs);
        }
        wbase = __shfl_sync(0xffffffffu, wbase, 31);
        if (out_keys != nullptr) {
            u64 pos = wbase + ((incl - mine) & 0xffffu);
#pragma unroll
            for (int q = 0; q < MIXED_MAXD; q++)
                if ((gm >> q) & 1u) out_keys[pos++] = lk[q];
        }
        if (PIVOT) {
            pbase = __shfl_sync(0xffffffffu, pbase, 31);
            if (out_pivot != nullptr) {
                u64 pos = pbase + ((incl - mine) >> 16);
#pragma unroll
                for (int q = 0; q < MIXED_MAXD; q++)
                    if ((pm >> q) & 1u) out_pivot[pos++] = lk[q];
            }
        }
    }
    __syncthreads();
    for (u32 i = tid; i <= nbins; i += MIXED_BLOCK) {
        const u32 c = sh_hist[i];
        if (c) atomicAdd(&hist[i], (u64)c);
    }
}

// ---- host side -----------------------------------------------------------------------------------------
static int compact_scratch(khb_ctx *ctx, u64 ntiles, u64 **d_lb, u32 **d_ticket)
{
    void *p;
    int rc = khb_scratch_get(ctx, SCR_FLAGS, (ntiles + 8) * sizeof(u64), &p);
    if (rc) return rc;
    *d_lb = (u64 *)p + 1;
    *d_ticket = (u32 *)p;
    KHB_CUDA(ctx, cudaMemsetAsync(p, 0, (ntiles + 8) * sizeof(u64), ctx->stream));
    return KHB_OK;
}

int khb_unique_impl(khb_ctx *ctx, const void *d_sorted, size_t n, int k, void *d_out, u64 *d_count)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_unique: k=%d outside 1..64", k);
    KHB_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(u64), ctx->stream));
    if (n == 0) return KHB_OK;
    const u64 ntiles = div_up(n, CP_TILE);
    u64 *d_lb;
    u32 *d_ticket;
    int rc = compact_scratch(ctx, ntiles, &d_lb, &d_ticket);
    if (rc) return rc;
    u64 grid = (u64)ctx->num_sms * 3;
    if (grid > ntiles) grid = ntiles;
    khb_prof_begin(ctx, KHB_K_UNIQUE);
    if (k <= 32)
        unique_kernel<Key64><<<(unsigned)grid, CP_BLOCK, 0, ctx->stream>>>((const Key64 *)d_sorted, n, (Key64 *)d_out, d_lb, d_ticket, 1u, d_count);
    else
        unique_kernel<Key128><<<(unsigned)grid, CP_BLOCK, 0, ctx->stream>>>((const Key128 *)d_sorted, n, (Key128 *)d_out, d_lb, d_ticket, 1u, d_count);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_UNIQUE, 2 * (u64)n * (k <= 32 ? 8 : 16));  // upper bound: every key kept
    return KHB_OK;
}

int khb_count_runs_impl(khb_ctx *ctx, const void *d_sorted, size_t n, int k, u32 cs, u32 nbins, u64 *d_hist,
                        void *d_out_keys, u32 *d_out_counts, u64 *d_runs)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_count_runs: k=%d outside 1..64", k);
    if (nbins < 1 || nbins > 8192) return khb_fail(ctx, KHB_ERR_ARG, "khb_count_runs: nbins=%u outside 1..8192", nbins);
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_runs, 0, sizeof(u64), ctx->stream));
    if (n == 0) return KHB_OK;
    const u64 ntiles = div_up(n, CP_TILE);
    u64 *d_lb;
    u32 *d_ticket;
    int rc = compact_scratch(ctx, ntiles, &d_lb, &d_ticket);
    if (rc) return rc;
    u64 grid = (u64)ctx->num_sms * 3;
    if (grid > ntiles) grid = ntiles;
    const size_t shm = ((size_t)nbins + 1) * sizeof(u32);
    khb_prof_begin(ctx, KHB_K_RLE);
    if (k <= 32)
        rle_hist_kernel<Key64><<<(unsigned)grid, CP_BLOCK, shm, ctx->stream>>>((const Key64 *)d_sorted, n, cs, nbins, d_hist, (Key64 *)d_out_keys, d_out_counts, d_lb, d_ticket, 1u, d_runs);
    else
        rle_hist_kernel<Key128><<<(unsigned)grid, CP_BLOCK, shm, ctx->stream>>>((const Key128 *)d_sorted, n, cs, nbins, d_hist, (Key128 *)d_out_keys, d_out_counts, d_lb, d_ticket, 1u, d_runs);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_RLE, (u64)n * (k <= 32 ? 8 : 16));  // input read; emitted keys are data dependent
    return KHB_OK;
}

// K4 on prefix-sorted input: d_count <- number of distinct keys; d_out <- the distinct keys (unordered).
int khb_resolve_unique_impl(khb_ctx *ctx, const void *d_sorted, size_t n, int k, int pshift, void *d_out, u64 *d_count)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_resolve_unique: k=%d outside 1..64", k);
    KHB_CUDA(ctx, cudaMemsetAsync(d_count, 0, sizeof(u64), ctx->stream));
    if (n == 0) return KHB_OK;
    const u64 ntiles = div_up(n, CQ_BLOCK * CQ_ITEMS);
    u64 grid = (u64)ctx->num_sms * (k <= 32 ? 2048 : 1024) / CQ_BLOCK;
    if (grid > ntiles) grid = ntiles;
    khb_prof_begin(ctx, KHB_K_UNIQUE);
    if (k <= 32)
        runs_kernel<Key64, false><<<(unsigned)grid, CQ_BLOCK, 0, ctx->stream>>>((const Key64 *)d_sorted, n, pshift, 0, 0, nullptr, (Key64 *)d_out, d_count);
    else
        runs_kernel<Key128, false><<<(unsigned)grid, CQ_BLOCK, 0, ctx->stream>>>((const Key128 *)d_sorted, n, pshift, 0, 0, nullptr, (Key128 *)d_out, d_count);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_UNIQUE, 2 * (u64)n * (k <= 32 ? 8 : 16));
    return KHB_OK;
}

// K5/K6 on prefix-sorted input: histogram of multiplicities (+ optional distinct keys, unordered).
int khb_resolve_count_impl(khb_ctx *ctx, const void *d_sorted, size_t n, int k, int pshift, u32 cs, u32 nbins, u64 *d_hist,
                           void *d_out_keys, u64 *d_runs)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_resolve_count: k=%d outside 1..64", k);
    if (nbins < 1 || nbins > 8192) return khb_fail(ctx, KHB_ERR_ARG, "khb_resolve_count: nbins=%u outside 1..8192", nbins);
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_runs, 0, sizeof(u64), ctx->stream));
    if (n == 0) return KHB_OK;
    const u64 ntiles = div_up(n, CQ_BLOCK * CQ_ITEMS);
    u64 grid = (u64)ctx->num_sms * (k <= 32 ? 2048 : 1024) / CQ_BLOCK;
    if (grid > ntiles) grid = ntiles;
    const size_t shm = ((size_t)nbins + 1) * sizeof(u32);
    khb_prof_begin(ctx, KHB_K_RLE);
    if (k <= 32)
        runs_kernel<Key64, true><<<(unsigned)grid, CQ_BLOCK, shm, ctx->stream>>>((const Key64 *)d_sorted, n, pshift, cs, nbins, d_hist, (Key64 *)d_out_keys, d_runs);
    else
        runs_kernel<Key128, true><<<(unsigned)grid, CQ_BLOCK, shm, ctx->stream>>>((const Key128 *)d_sorted, n, pshift, cs, nbins, d_hist, (Key128 *)d_out_keys, d_runs);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_RLE, (u64)n * (k <= 32 ? 8 : 16));
    return KHB_OK;
}

// Single-sort group path: histogram of genomes-per-k-mer from prefix-sorted (key, genome id) pairs.
int khb_pairs_count_impl(khb_ctx *ctx, const void *d_sorted, const unsigned short *d_gid, size_t n, int k, int pshift, u32 cs, u32 nbins,
                         u32 n_genomes, u64 *d_hist, void *d_out_keys, u64 *d_runs, u64 *d_pairs, int pivot, void *d_out_pivot,
                         u64 *d_pruns)
{
    // pivot != 0 (experiment type 2): the last genome (id n_genomes - 1) is the pivot, see pairs_kernel
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_pairs_count: k=%d outside 1..64", k);
    if (nbins < 1 || nbins > 8192) return khb_fail(ctx, KHB_ERR_ARG, "khb_pairs_count: nbins=%u outside 1..8192", nbins);
    if ((u64)n >= (1ull << 32)) return khb_fail(ctx, KHB_ERR_ARG, "khb_pairs_count: %zu keys in one group (limit 2^32 - 1)", n);
    if (pivot && (n_genomes < 1 || !d_pruns)) return khb_fail(ctx, KHB_ERR_ARG, "khb_pairs_count: pivot mode needs the genome count");
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_runs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_pairs, 0, sizeof(u64), ctx->stream));
    if (pivot) KHB_CUDA(ctx, cudaMemsetAsync(d_pruns, 0, sizeof(u64), ctx->stream));
    if (n == 0) return KHB_OK;
    const u64 nchunks = div_up(n, 32 * CQ_ITEMS);  // warp chunks
    static int pf = -1;
    if (pf < 0) {
        const char *e = getenv("KHB_PAIRS_PREFETCH");
        pf = e ? atoi(e) : 0;  // measured on config 2: 2.1 ms per group without, 2.3 ms with (fewer resident warps)
    }
    const int minb = k <= 32 ? (pf ? PairsCfg<Key64, true>::MINB : PairsCfg<Key64, false>::MINB)
                             : (pf ? PairsCfg<Key128, true>::MINB : PairsCfg<Key128, false>::MINB);  // resident CTAs per SM: one persistent wave
    u64 grid = (u64)ctx->num_sms * minb;
    if (grid > div_up(nchunks, CQ_WARPS)) grid = div_up(nchunks, CQ_WARPS);
    const size_t shm = ((size_t)nbins + 1) * sizeof(u32);
    // multiplicity counted in registers besides 1: "in every genome of the group"
    const u32 hot = (n_genomes >= 2 && n_genomes <= nbins && n_genomes <= cs) ? n_genomes : 0u;
    const u32 pgid = pivot ? n_genomes - 1u : 0xffffffffu;
    // scratch: [item count][bitmap of the mixed prefix runs: one byte per 8 keys, written completely by pairs_kernel]
    //          [work list: at most one entry per two keys]
    const u64 nwords = nchunks * (32 * CQ_ITEMS / 32);
    void *p;
    int rc = khb_scratch_get(ctx, SCR_FLAGS, 64 + nwords * sizeof(u32) + (n / 2 + 2) * sizeof(u32), &p);
    if (rc) return rc;
    u64 *d_nitems = (u64 *)p;
    u32 *d_map = (u32 *)((char *)p + 64);
    u32 *d_items = d_map + nwords;
    KHB_CUDA(ctx, cudaMemsetAsync(d_nitems, 0, sizeof(u64), ctx->stream));
    u64 grid2 = div_up(nwords, MIXED_BLOCK);
    if (grid2 > (u64)ctx->num_sms * 8) grid2 = (u64)ctx->num_sms * 8;
    const u64 grid3 = (u64)ctx->num_sms * 4;  // CTAs past the item count exit at once
    khb_prof_begin(ctx, KHB_K_RLE);
#define PAIRS_LAUNCH(KEY, PFV, PIV)                                                                                                      \
    pairs_kernel<KEY, PFV, PIV><<<(unsigned)grid, CQ_BLOCK, shm, ctx->stream>>>((const KEY *)d_sorted, d_gid, n, pshift, cs, nbins, hot, d_hist, \
                                                                                (KEY *)d_out_keys, d_runs, d_pairs, (unsigned char *)d_map, pgid, \
                                                                                (KEY *)d_out_pivot, d_pruns)
#define PAIRS_DISPATCH(KEY)                                           \
    do {                                                              \
        if (pivot) { if (pf) PAIRS_LAUNCH(KEY, true, true); else PAIRS_LAUNCH(KEY, false, true); }   \
        else { if (pf) PAIRS_LAUNCH(KEY, true, false); else PAIRS_LAUNCH(KEY, false, false); }       \
    } while (0)
    if (k <= 32) PAIRS_DISPATCH(Key64); else PAIRS_DISPATCH(Key128);
#undef PAIRS_DISPATCH
#undef PAIRS_LAUNCH
    KHB_LAUNCH_CHECK(ctx);
    mixed_collect_kernel<<<(unsigned)grid2, MIXED_BLOCK, 0, ctx->stream>>>(d_map, nwords, d_items, d_nitems);
    KHB_LAUNCH_CHECK(ctx);
#define MIXED_LAUNCH(KEY, PIV)                                                                                                           \
    mixed_runs_kernel<KEY, PIV><<<(unsigned)grid3, MIXED_BLOCK, shm, ctx->stream>>>((const KEY *)d_sorted, d_gid, n, pshift, cs, nbins, d_hist, \
                                                                                    (KEY *)d_out_keys, d_runs, d_pairs, d_items, d_nitems, pgid, \
                                                                                    (KEY *)d_out_pivot, d_pruns)
    if (k <= 32) { if (pivot) MIXED_LAUNCH(Key64, true); else MIXED_LAUNCH(Key64, false); }
    else { if (pivot) MIXED_LAUNCH(Key128, true); else MIXED_LAUNCH(Key128, false); }
#undef MIXED_LAUNCH
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_RLE, (u64)n * ((k <= 32 ? 8 : 16) + 2));
    return KHB_OK;
}

// ---- experiment type 2, across groups (exp_type_2.smk:440-496) ---------------------------------------------------------
// Input: the distinct k-mer sets U_1 .. U_G of the groups' rest-of-set unions followed by the pivot sets P_1 .. P_G,
// sorted (stable, by prefix) with payload = set index (U_i: i-1, P_j: G+j-1).  Inside a prefix run the order of the
// payloads is preserved, so every U element precedes every P element.  For a pivot element x of P_j:
//   g = #{ i != j : x in U_i },  hist[j][1 + g]++     (1 = kmers_subtract, >= 2 = counter of intersect -ocsum)
// One thread per element; pivot elements walk back over their prefix run (at most a few times 2G keys).
template <typename Key>
__global__ void __launch_bounds__(256)
pivot_across_kernel(const Key *__restrict__ in, const unsigned short *__restrict__ pay, u64 n, int pshift, u32 n_groups, u32 cs, u32 nbins,
                    u64 *__restrict__ hist /* [n_groups][nbins+1] */, u32 sh_bins /* columns of the shared histogram, 0 = none */)
{
    extern __shared__ u32 sh_hist[];  // [n_groups][sh_bins]
    const u32 tid = threadIdx.x;
    for (u32 i = tid; i < n_groups * sh_bins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    for (u64 i = (u64)blockIdx.x * blockDim.x + tid; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 p = pay[i];
        if (p < n_groups) continue;
        const u32 j = p - n_groups;
        const Key key = in[i];
        if (key_is_sentinel(key)) continue;
        u32 n_u = 0, own = 0;
        for (u64 q = i; q > 0;) {
            --q;
            const Key kq = in[q];
            if (key_is_sentinel(kq) || !same_prefix(kq, key, pshift)) break;
            if (!key_eq(kq, key)) continue;
            const u32 pq = pay[q];
            if (pq < n_groups) { n_u++; own |= (pq == j) ? 1u : 0u; }
        }
        u32 c = 1u + n_u - own;
        c = c > cs ? cs : c;
        if (c > nbins) continue;
        if (c < sh_bins) atomicAdd(&sh_hist[j * sh_bins + c], 1u);
        else atomicAdd(&hist[(size_t)j * (nbins + 1) + c], 1ull);
    }
    __syncthreads();
    for (u32 i = tid; i < n_groups * sh_bins; i += blockDim.x) {
        const u32 v = sh_hist[i];
        if (v) atomicAdd(&hist[(size_t)(i / sh_bins) * (nbins + 1) + (i % sh_bins)], (u64)v);
    }
}

// pay[i] = s for seg_off[s] <= i < seg_off[s+1]  (blockIdx.y = s)
__global__ void fill_segment_ids_kernel(unsigned short *__restrict__ pay, const u64 *__restrict__ seg_off)
{
    const u32 s = blockIdx.y;
    const u64 b = seg_off[s], e = seg_off[s + 1];
    for (u64 i = b + (u64)blockIdx.x * blockDim.x + threadIdx.x; i < e; i += (u64)gridDim.x * blockDim.x) pay[i] = (unsigned short)s;
}

int khb_fill_segment_ids_impl(khb_ctx *ctx, unsigned short *d_pay, const u64 *d_seg_off, int nseg, u64 max_len)
{
    if (nseg <= 0) return KHB_OK;
    u64 bx = div_up(max_len, 256 * 8);
    if (bx < 1) bx = 1;
    if (bx > 1024) bx = 1024;
    fill_segment_ids_kernel<<<dim3((unsigned)bx, (unsigned)nseg), 256, 0, ctx->stream>>>(d_pay, d_seg_off);
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}

int khb_pivot_across_impl(khb_ctx *ctx, const void *d_sorted, const unsigned short *d_pay, size_t n, int k, int pshift, u32 n_groups,
                          u32 cs, u32 nbins, u64 *d_hist /* [n_groups][nbins+1], zeroed here */)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_pivot_across: k=%d outside 1..64", k);
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, (size_t)n_groups * (nbins + 1) * sizeof(u64), ctx->stream));
    if (n == 0) return KHB_OK;
    // every count is at most n_groups: a [G][G+1] shared histogram takes all updates when it fits
    u32 sh_bins = (n_groups + 1 < nbins + 1) ? n_groups + 1 : nbins + 1;
    if ((size_t)n_groups * sh_bins * sizeof(u32) > 96 * 1024) sh_bins = 0;
    const size_t shm = (size_t)n_groups * sh_bins * sizeof(u32);
    // per-device function attribute: set on every call so that every context on every GPU has it
    cudaFuncSetAttribute(pivot_across_kernel<Key64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(pivot_across_kernel<Key128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    u64 grid = div_up(n, 256);
    if (grid > (u64)ctx->num_sms * 8) grid = (u64)ctx->num_sms * 8;
    khb_prof_begin(ctx, KHB_K_RLE);
    if (k <= 32)
        pivot_across_kernel<Key64><<<(unsigned)grid, 256, shm, ctx->stream>>>((const Key64 *)d_sorted, d_pay, n, pshift, n_groups, cs, nbins, d_hist, sh_bins);
    else
        pivot_across_kernel<Key128><<<(unsigned)grid, 256, shm, ctx->stream>>>((const Key128 *)d_sorted, d_pay, n, pshift, n_groups, cs, nbins, d_hist, sh_bins);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_RLE, (u64)n * ((k <= 32 ? 8 : 16) + 2));
    return KHB_OK;
}

// ---- sorted-set lookup (rule-compatible `kmc_tools simple A B intersect / kmers_subtract`, exp_type_2.smk:354-380) ------
// idx[i] = position of a[i] in the sorted, duplicate-free array b, or ~0 if it is not there.
template <typename Key> __device__ __forceinline__ bool key_less(const Key &a, const Key &b);
template <> __device__ __forceinline__ bool key_less<Key64>(const Key64 &a, const Key64 &b) { return a.v < b.v; }
template <> __device__ __forceinline__ bool key_less<Key128>(const Key128 &a, const Key128 &b) { return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo); }

template <typename Key>
__global__ void __launch_bounds__(256)
sorted_lookup_kernel(const Key *__restrict__ a, u64 na, const Key *__restrict__ b, u64 nb, u64 *__restrict__ idx)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < na; i += (u64)gridDim.x * blockDim.x) {
        const Key x = a[i];
        u64 lo = 0, hi = nb;  // first position with b[pos] >= x
        while (lo < hi) {
            const u64 mid = lo + ((hi - lo) >> 1);
            if (key_less(b[mid], x)) lo = mid + 1; else hi = mid;
        }
        idx[i] = (lo < nb && key_eq(b[lo], x)) ? lo : ~0ull;
    }
}

int khb_sorted_lookup_impl(khb_ctx *ctx, const void *d_a, u64 na, const void *d_b, u64 nb, int k, u64 *d_idx)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_sorted_lookup: k=%d outside 1..64", k);
    if (na == 0) return KHB_OK;
    u64 grid = div_up(na, 256);
    if (grid > (u64)ctx->num_sms * 16) grid = (u64)ctx->num_sms * 16;
    if (k <= 32)
        sorted_lookup_kernel<Key64><<<(unsigned)grid, 256, 0, ctx->stream>>>((const Key64 *)d_a, na, (const Key64 *)d_b, nb, d_idx);
    else
        sorted_lookup_kernel<Key128><<<(unsigned)grid, 256, 0, ctx->stream>>>((const Key128 *)d_a, na, (const Key128 *)d_b, nb, d_idx);
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}

// ---- group membership of query k-mers (experiment type 4: src/merge_lists.py:14-33, exp_type_4.smk:217-270) -------------
// Input: the retained group sets U_1 .. U_G followed by the query sets Q_1 .. Q_P (here: the pivots' k-mers), sorted
// (stable, by prefix) with payload = set index.  Inside a prefix run every U element precedes every Q element, so a Q
// element finds all groups that hold its k-mer by walking back over its prefix run.  The result goes to the position the
// k-mer has in its query set's own (canonical, ascending) order, found by binary search of the un-mixed key:
//   mask[(q_off[p] + rank) * words + w] bit b  <=>  group 64 w + b contains the k-mer
template <typename Key> __device__ __forceinline__ Key key_unmix(const Key &x, int k);
template <> __device__ __forceinline__ Key64 key_unmix<Key64>(const Key64 &x, int k) { return Key64{kmer_unmix64(x.v, k)}; }
template <> __device__ __forceinline__ Key128 key_unmix<Key128>(const Key128 &x, int k)
{
    u64 hi = x.hi, lo = x.lo;
    kmer_unmix128(hi, lo, k);
    return Key128{lo, hi};
}

#define MEMBER_MAXW 4
template <typename Key>
__global__ void __launch_bounds__(256)
membership_kernel(const Key *__restrict__ in, const unsigned short *__restrict__ pay, u64 n, int pshift, u32 n_groups, int k, int hashed,
                  const Key *__restrict__ q_sorted /* canonical, ascending per query set */, const u64 *__restrict__ q_off, int words,
                  u64 *__restrict__ mask_out)
{
    for (u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (u64)gridDim.x * blockDim.x) {
        const u32 p = pay[i];
        if (p < n_groups) continue;
        const u32 q = p - n_groups;
        const Key key = in[i];
        if (key_is_sentinel(key)) continue;
        u64 m[MEMBER_MAXW] = {0, 0, 0, 0};
        for (u64 j = i; j > 0;) {
            --j;
            const Key kj = in[j];
            if (key_is_sentinel(kj) || !same_prefix(kj, key, pshift)) break;
            if (!key_eq(kj, key)) continue;
            const u32 pj = pay[j];
            if (pj < n_groups) {
#pragma unroll
                for (int w = 0; w < MEMBER_MAXW; w++)
                    if ((int)(pj >> 6) == w) m[w] |= 1ull << (pj & 63u);
            }
        }
        const Key c = hashed ? key_unmix(key, k) : key;
        u64 lo = q_off[q], hi = q_off[q + 1];  // first position with q_sorted[pos] >= c
        while (lo < hi) {
            const u64 mid = lo + ((hi - lo) >> 1);
            if (key_less(q_sorted[mid], c)) lo = mid + 1; else hi = mid;
        }
#pragma unroll
        for (int w = 0; w < MEMBER_MAXW; w++)
            if (w < words) mask_out[lo * (u64)words + w] = m[w];
    }
}

int khb_membership_impl(khb_ctx *ctx, const void *d_sorted, const unsigned short *d_pay, size_t n, int k, int pshift, u32 n_groups, int hashed,
                        const void *d_q_sorted, const u64 *d_q_off, int words, u64 *d_mask_out)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_membership: k=%d outside 1..64", k);
    if (words < 1 || words > MEMBER_MAXW) return khb_fail(ctx, KHB_ERR_ARG, "khb_membership: %u groups (limit %d)", n_groups, 64 * MEMBER_MAXW);
    if (n == 0) return KHB_OK;
    u64 grid = div_up(n, 256);
    if (grid > (u64)ctx->num_sms * 8) grid = (u64)ctx->num_sms * 8;
    khb_prof_begin(ctx, KHB_K_RLE);
    if (k <= 32)
        membership_kernel<Key64><<<(unsigned)grid, 256, 0, ctx->stream>>>((const Key64 *)d_sorted, d_pay, n, pshift, n_groups, k, hashed, (const Key64 *)d_q_sorted, d_q_off, words, d_mask_out);
    else
        membership_kernel<Key128><<<(unsigned)grid, 256, 0, ctx->stream>>>((const Key128 *)d_sorted, d_pay, n, pshift, n_groups, k, hashed, (const Key128 *)d_q_sorted, d_q_off, words, d_mask_out);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_RLE, (u64)n * ((k <= 32 ? 8 : 16) + 2));
    return KHB_OK;
}
