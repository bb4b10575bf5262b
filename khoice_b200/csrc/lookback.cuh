// lookback.cuh -- decoupled look-back (single-pass chained prefix sum over tiles).
//
// One 64-bit word per (tile, column): bits 0..39 value, bits 40..41 status, bits 48..63 epoch tag.
// The value and its status travel in ONE word, so a relaxed 64-bit store/load pair is enough: a reader
// either sees a complete (status, value) or an entry from an older epoch, which it treats as "not
// ready".  The epoch tag lets several passes reuse the array with a single memset at the start.
// Tiles take their index from an atomic ticket, so every predecessor of a running tile is already
// running or finished: spinning on a predecessor cannot deadlock.
#pragma once
#include "khb_common.cuh"

#define LB_AGG 1ull     // tile-local aggregate available
#define LB_PREFIX 2ull  // inclusive prefix available
#define LB_VALUE_MASK 0xffffffffffull

__device__ __forceinline__ u64 lb_pack(u64 status, u64 value, u32 epoch)
{
    return (value & LB_VALUE_MASK) | (status << 40) | ((u64)epoch << 48);
}
__device__ __forceinline__ u64 lb_load(const u64 *p)
{
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void lb_store(u64 *p, u64 v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u32 lb_status(u64 e, u32 epoch) { return ((u32)(e >> 48) == epoch) ? (u32)((e >> 40) & 3u) : 0u; }

// Serial walk by ONE thread over column `col` of a [tile][ncols] array (used by the radix passes, where
// 256 threads each own a digit column, so a hop is one coalesced 2 KiB read for the CTA).
// `first` is the first tile of the chain (it publishes LB_PREFIX itself).  Returns the exclusive prefix.
__device__ __forceinline__ u64 lb_walk_serial(const u64 *lb, size_t tile, size_t first, u32 ncols, u32 col, u32 epoch)
{
    u64 excl = 0;
    size_t t = tile;
    while (t > first) {
        --t;
        u64 e;
        u32 st;
        do {
            e = lb_load(lb + t * ncols + col);
            st = lb_status(e, epoch);
        } while (st == 0);
        excl += e & LB_VALUE_MASK;
        if (st == LB_PREFIX) break;
    }
    return excl;
}

// Warp-parallel walk over a single-column array (used by the compaction kernels): 32 predecessors per
// hop.  Must be called by a full warp; every lane returns the exclusive prefix.
__device__ __forceinline__ u64 lb_walk_warp(const u64 *lb, size_t tile, size_t first, u32 epoch)
{
    u64 excl = 0;
    size_t hi = tile;  // entries [first, hi) remain to be examined
    const u32 lane = lane_id();
    while (hi > first) {
        const bool have = hi - first > lane;
        const size_t t = have ? hi - 1 - lane : first;
        u64 e = 0;
        u32 st = 0;
        for (;;) {
            if (have) {
                e = lb_load(lb + t);
                st = lb_status(e, epoch);
            } else {
                st = 3;  // beyond the chain start: neutral
                e = 0;
            }
            const u32 pfx = __ballot_sync(0xffffffffu, have && st == LB_PREFIX);
            const u32 notready = __ballot_sync(0xffffffffu, st == 0);
            // lanes nearer than the first PREFIX lane must all be ready
            const u32 upto = pfx ? ((pfx & (0u - pfx)) << 1) - 1u : 0xffffffffu;  // mask of lanes <= first prefix lane
            if ((notready & upto) == 0) {
                u64 v = ((upto >> lane) & 1u) ? (e & LB_VALUE_MASK) : 0ull;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                excl += v;
                if (pfx) return excl;
                break;
            }
        }
        hi = hi - first > 32 ? hi - 32 : first;
    }
    return excl;
}
