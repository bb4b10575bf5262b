// khb_common.cuh -- shared device/host helpers of libkhoice_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/khoice_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;

// ---- k-mer words ---------------------------------------------------------------------------
// k <= 32 : one 64-bit word.  k <= 64 : 128-bit, stored little-endian (lo word first, 16-byte records).
struct __align__(8) Key64 {
    u64 v;
};
struct __align__(16) Key128 {
    u64 lo, hi;
};

__host__ __device__ __forceinline__ bool key_eq(const Key64 &a, const Key64 &b) { return a.v == b.v; }
__host__ __device__ __forceinline__ bool key_eq(const Key128 &a, const Key128 &b) { return a.lo == b.lo && a.hi == b.hi; }
__host__ __device__ __forceinline__ bool key_is_sentinel(const Key64 &a) { return a.v == ~0ull; }
__host__ __device__ __forceinline__ bool key_is_sentinel(const Key128 &a) { return (a.lo & a.hi) == ~0ull; }
// 8-bit digit starting at bit `shift` (any alignment; bits past the top read as zero)
__host__ __device__ __forceinline__ u32 key_digit(const Key64 &a, int shift) { return (u32)(a.v >> shift) & 0xffu; }
__host__ __device__ __forceinline__ u32 key_digit(const Key128 &a, int shift)
{
    if (shift >= 64) return (u32)(a.hi >> (shift - 64)) & 0xffu;
    if (shift <= 56) return (u32)(a.lo >> shift) & 0xffu;
    return (u32)((a.lo >> shift) | (a.hi << (64 - shift))) & 0xffu;
}
// bits [shift, top) of a key: two keys are in the same prefix run iff these are equal (shift = 0: the whole key)
__host__ __device__ __forceinline__ bool same_prefix(const Key64 &a, const Key64 &b, int shift) { return ((a.v ^ b.v) >> shift) == 0; }
__host__ __device__ __forceinline__ bool same_prefix(const Key128 &a, const Key128 &b, int shift)
{
    const u64 xl = a.lo ^ b.lo, xh = a.hi ^ b.hi;
    if (shift >= 64) return (xh >> (shift - 64)) == 0;
    return xh == 0 && (xl >> shift) == 0;
}

// ---- bijective k-mer mixer -------------------------------------------------------------------
// The fused path sorts only a PREFIX (top 8*P bits) of each key and resolves the rest by comparison
// (compact.cu: resolve kernels).  For that the top bits must be uniformly distributed, so K2 can emit
// h(c) instead of the canonical value c, where h is a bijection of [0, 4^k):
//     y = ~c;  y *= C1;  y ^= y >> k;  y *= C2;  y ^= y >> k;  y *= C1;  h = ~y      (all mod 4^k)
// Odd multiplications and the half-width xor-shift are invertible (the xor-shift is its own inverse), so
// equality, set sizes and multiplicities are unchanged and kmer_unmix recovers c exactly.  The two
// complements make h(all ones) = all ones, which keeps the 64/128-bit sentinel a sentinel for k = 32 / 64.
#define KHB_MIX_C1 0x9E3779B97F4A7C15ull
#define KHB_MIX_C2 0xBF58476D1CE4E5B9ull

__host__ __device__ __forceinline__ u64 mix_mask64(int k) { return k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull); }
__host__ __device__ __forceinline__ u64 odd_inverse64(u64 c)
{
    u64 x = c;  // Newton: x <- x * (2 - c x), doubles the number of correct low bits
    for (int i = 0; i < 6; i++) x *= 2ull - c * x;
    return x;
}
__host__ __device__ __forceinline__ u64 kmer_mix64(u64 c, int k)
{
    const u64 M = mix_mask64(k);
    u64 y = ~c & M;
    y = (y * KHB_MIX_C1) & M;
    y ^= y >> k;
    y = (y * KHB_MIX_C2) & M;
    y ^= y >> k;
    y = (y * KHB_MIX_C1) & M;
    return ~y & M;
}
__host__ __device__ __forceinline__ u64 kmer_unmix64(u64 h, int k)
{
    const u64 M = mix_mask64(k);
    const u64 i1 = odd_inverse64(KHB_MIX_C1), i2 = odd_inverse64(KHB_MIX_C2);
    u64 y = ~h & M;
    y = (y * i1) & M;
    y ^= y >> k;
    y = (y * i2) & M;
    y ^= y >> k;
    y = (y * i1) & M;
    return ~y & M;
}

// 128-bit variant for 33 <= k <= 64: (hi:lo) arithmetic mod 4^k, same recipe.
#ifdef __CUDA_ARCH__
#define KHB_MULHI64(a, b) __umul64hi((a), (b))
#else
#define KHB_MULHI64(a, b) ((u64)(((unsigned __int128)(a) * (unsigned __int128)(b)) >> 64))
#endif
__host__ __device__ __forceinline__ void mul128(u64 &hi, u64 &lo, u64 ch, u64 cl)
{
    const u64 nlo = lo * cl;
    const u64 nhi = KHB_MULHI64(lo, cl) + lo * ch + hi * cl;
    hi = nhi;
    lo = nlo;
}
__host__ __device__ __forceinline__ void xorshift128(u64 &hi, u64 &lo, int k)
{
    // (hi:lo) ^= (hi:lo) >> k with 33 <= k <= 64 and hi < 2^(2k-64): only lo changes
    lo ^= k >= 64 ? hi : ((lo >> k) | (hi << (64 - k)));
}
__host__ __device__ __forceinline__ void mix_steps128(u64 &hi, u64 &lo, int k, u64 a_hi, u64 a_lo, u64 b_hi, u64 b_lo)
{
    const u64 Mh = k >= 64 ? ~0ull : ((1ull << (2 * k - 64)) - 1ull);
    hi = ~hi & Mh;
    lo = ~lo;
    mul128(hi, lo, a_hi, a_lo);
    hi &= Mh;
    xorshift128(hi, lo, k);
    mul128(hi, lo, b_hi, b_lo);
    hi &= Mh;
    xorshift128(hi, lo, k);
    mul128(hi, lo, a_hi, a_lo);
    hi &= Mh;
    hi = ~hi & Mh;
    lo = ~lo;
}
__host__ __device__ __forceinline__ void kmer_mix128(u64 &hi, u64 &lo, int k)
{
    mix_steps128(hi, lo, k, 0ull, KHB_MIX_C1, 0ull, KHB_MIX_C2);
}
__host__ __device__ __forceinline__ void odd_inverse128(u64 c, u64 &ih, u64 &il)
{
    // inverse of the 64-bit odd constant c modulo 2^128, by Newton iteration in 128-bit arithmetic
    u64 xh = 0, xl = c;
    for (int i = 0; i < 7; i++) {
        // t = 2 - c * x
        u64 th = xh, tl = xl;
        mul128(th, tl, 0ull, c);
        const u64 nl = 2ull - tl;
        const u64 nh = ~th + (tl <= 2ull ? 1ull : 0ull);  // (0:2) - (th:tl)
        mul128(xh, xl, nh, nl);
    }
    ih = xh;
    il = xl;
}
__host__ __device__ __forceinline__ void kmer_unmix128(u64 &hi, u64 &lo, int k)
{
    u64 a_h, a_l, b_h, b_l;
    odd_inverse128(KHB_MIX_C1, a_h, a_l);
    odd_inverse128(KHB_MIX_C2, b_h, b_l);
    mix_steps128(hi, lo, k, a_h, a_l, b_h, b_l);
}

// ---- owner of a k-mer in the multi-GPU exchange: hash-range partition of the key space (K7, peer.cu) ----------
__device__ __forceinline__ u64 mix64(u64 x)
{
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ u32 part_of(const Key64 &k, u32 nparts) { return (u32)__umul64hi(mix64(k.v), (u64)nparts); }
__device__ __forceinline__ u32 part_of(const Key128 &k, u32 nparts)
{
    return (u32)__umul64hi(mix64(k.lo ^ mix64(k.hi)), (u64)nparts);
}

// ---- context ------------------------------------------------------------------------------
struct khb_scratch {
    void *ptr;
    size_t bytes;
};

struct khb_ctx {
    int device;
    int num_sms;
    cudaStream_t stream;
    cudaEvent_t ev0, ev1;
    char err[512];
    int sticky;  // first CUDA error code seen (0 = none); CUDA errors are sticky per ctx
    khb_scratch scratch[12];
    // pinned host mailbox for small device->host results
    u64 *h_mail;
    u64 *d_mail;
    // kernel launch counter (bench.py's gpu_launches)
    u64 launches;
    // accumulated group sets for the across-group stage (fused mode)
    void *gs_buf;       // device buffer holding the concatenation of kept group sets
    size_t gs_cap;      // capacity in BYTES
    size_t gs_len;      // keys stored
    int gs_k;           // k the stored sets were built with (0 = none)
    int gs_groups;      // number of groups stored
    int gs_hashed;      // 1: stored keys are kmer_mix(canonical), 0: canonical values
    // staging (fused host entry points)
    uint8_t *stage_dev;
    size_t stage_dev_cap;
    // double-buffered ingestion (khb_group_prefetch_fasta): the next group's text is copied on copy_stream
    // into stage_next while the current group is processed on `stream`
    uint8_t *stage_next;
    size_t stage_next_cap;
    cudaStream_t copy_stream;
    cudaEvent_t copy_done;
    u64 *pf_tab;            // device table for the separator kernel of the prefetch
    size_t pf_tab_cap;
    int pf_valid, pf_n;
    const void *pf_first;   // identity of the prefetched request: first file pointer, count, staged size
    size_t pf_bytes;
    struct khb_hostvec *pf_begin;
    // timing of the last fused call (ms, CUDA events on ctx->stream)
    float last_ms[8];
    // optional per-kernel timing (khb_profile_enable): CUDA event pairs around every launch
    struct khb_prof *prof;
    int prof_on;
    // experiment type 2: retained pivot sets (api.cu)
    struct khb_pivot_store *pv;
    // hash group stage (hashset.cu): the table is all zero between calls unless hs_dirty
    u32 *hs_tab;
    size_t hs_bytes;
    int hs_dirty;
    u64 hs_overflows;  // groups that fell back to the sort path because a probe sequence hit the limit
    struct khb_peer *peer;  // multi-GPU exchange over peer memory (peer.cu)
    cudaStream_t prof_stream;  // stream the next khb_prof_begin/end pair records on (null: `stream`)
    double bins_rho;   // distinct k-mers per window in the last group the minimizer-bin path counted (sizes the next group's passes)
    u64 bins_fallbacks, bins_bigbins, bins_repartitions;
    int bins_hint_k, bins_hint_genomes;            // shape of the last group the minimizer-bin path counted ...
    u64 bins_hint_regions, bins_hint_max, bins_last_regions;  // ... its regions and the records of its fullest one  // groups the minimizer-bin path handed to the sort path / bins redone in hash classes (bins.cu)
    // segment events of the group-set store (bins.cu: mb_event): while every retained group came through the minimizer bins with the same
    // number of bins, the across-group stage counts bin by bin instead of sorting the store
    void *ev_buf;       // device, ev_cap events of 16 bytes
    u64 *ev_count;      // device counter
    u64 ev_cap, ev_len; // capacity; events logged so far (host copy, read back with every group's mailbox)
    int ev_ok;          // 1: the events describe the whole store
    u32 ev_nb;          // bins of the groups in the store
    u64 across_by_bins, across_by_sort;   // performance counters
    int group_mode;    // KHB_GROUP_* (khb_set_group_mode; initial value from the environment variable KHB_GROUP_MODE)
    struct khb_team *team;  // one group sharded over several GPUs (team.cu)
};

// ---- a group sharded over the members of a team (team.cu, bins.cu) ----------------------------------------------------------------
#define KHB_TEAM_MAX 8
// what the sharded partition pass needs to know: the owner of bin b is member b / bpo; its record buffer and its table of region sizes
struct mb_shard {
    u32 team;            // members (0: not sharded)
    u32 bpo;             // bins per owner
    u32 nchunks_total;   // chunks of 64 genome ids of the whole group
    u32 chunk_base;      // first chunk of this member's slice (its genome ids start at 64 * chunk_base)
    u32 member;          // this member: it packs its regions into area `member` of every owner's buffer
    u64 area;            // records one sender's area holds
    u64 *rec[KHB_TEAM_MAX];   // every member's record buffer (own entry included) ...
    u64 *roff[KHB_TEAM_MAX];  // ... its table of region starts (in records)
    u32 *cur[KHB_TEAM_MAX];   // ... and of region sizes
};
struct khb_team {
    int size = 0, member = 0;
    size_t half_bytes = 0;             // bytes of ONE of the two receive buffers
    void *recv = nullptr;              // this member's two receive buffers
    void *peer_base[KHB_TEAM_MAX];     // every member's (own entry = recv)
    bool opened = false;
    u64 slice_sym = 0, slice_bases = 0, slice_fasta_bytes = 0;   // this member's slice of the group in flight (for the stats of the count call)
};


// kernel ids for khb_profile_read
enum { KHB_K_PACK = 0, KHB_K_EXTRACT = 1, KHB_K_RADIX_HIST = 2, KHB_K_ONESWEEP = 3, KHB_K_UNIQUE = 4, KHB_K_RLE = 5, KHB_K_PARTITION = 6, KHB_K_HASH_INSERT = 7, KHB_K_HASH_COUNT = 8, KHB_K_BIN_PARTITION = 9, KHB_K_BIN_COUNT = 10, KHB_K_BIN_ACROSS = 11, KHB_K_COUNT = 12 };
void khb_prof_begin(khb_ctx *ctx, int id);
void khb_prof_end(khb_ctx *ctx, int id, u64 alg_bytes);
void khb_prof_patch(khb_ctx *ctx, int id, u64 alg_bytes);  // algorithmic bytes of the LAST record of that kernel, once they are known

enum { SCR_TILE = 0, SCR_LOOKBACK = 1, SCR_HIST = 2, SCR_MISC = 3, SCR_KEYS_A = 4, SCR_KEYS_B = 5, SCR_PACK = 6, SCR_FLAGS = 7, SCR_PAY_A = 8, SCR_PAY_B = 9, SCR_AUX = 10, SCR_TEAM = 11 };  // SCR_MISC belongs to the sort (segment tables, tickets)
#define KHB_NSCRATCH 12

int khb_fail(khb_ctx *ctx, int code, const char *fmt, ...);
int khb_cuda_fail(khb_ctx *ctx, cudaError_t e, const char *what, const char *file, int line);
int khb_scratch_get(khb_ctx *ctx, int slot, size_t bytes, void **out);
int khb_peer_wait(khb_ctx *ctx);  // wait for pushes in flight (before the group-set store moves)
// Where the multi-GPU exchange wants a group's distinct keys (peer.cu): owner = part_of(key, world), region dst[owner] of `cap` keys,
// filled through the sender-side cursor[owner]; cursor[64] != 0: some region overflowed.  world = 0: no exchange is open.
struct khb_peer_route {
    u32 world;
    u32 flags;      // bit 0: store every bin's keys in owner-sorted order (KHB_PEER_SORTED=1)
    u64 cap;
    u64 *cursor;
    void *const *dst;
};
int khb_peer_route_get(khb_ctx *ctx, int key_bytes, khb_peer_route *out);  // for kernels that push while they emit (bins.cu)
void khb_peer_mark_pushed(khb_ctx *ctx);                                    // ... the store's keys so far are at their owners
int khb_peer_poison(khb_ctx *ctx);                                          // ... or not: make this round fall back to the NCCL route
// one group on several GPUs (bins.cu; the entry points are in api.cu, the buffers in team.cu)
int khb_bins_team_plan_impl(khb_ctx *ctx, int k, const khb_team_group *tg, u32 nbins_hist, u32 *nb, u32 *cap, u64 *half_bytes);
int khb_bins_team_partition_impl(khb_ctx *ctx, const u64 *d_codes, const u32 *d_valid, u64 n_sym, int k, const u64 *d_seg_off, int n_genomes,
                                 const khb_team_group *tg, u64 *d_info);
int khb_bins_team_count_impl(khb_ctx *ctx, int k, const khb_team_group *tg, u32 cs, u32 nbins_hist, u64 *d_hist, void *d_out_keys, u64 *d_runs,
                             u64 *d_pairs, u64 *d_stat, khb_peer_route route);

#define KHB_CUDA(ctx, expr)                                                           \
    do {                                                                              \
        cudaError_t e__ = (expr);                                                     \
        if (e__ != cudaSuccess) return khb_cuda_fail((ctx), e__, #expr, __FILE__, __LINE__); \
    } while (0)

#define KHB_LAUNCH_CHECK(ctx)                                                         \
    do {                                                                              \
        (ctx)->launches++;                                                            \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) return khb_cuda_fail((ctx), e__, "kernel launch", __FILE__, __LINE__); \
    } while (0)

#define KHB_CHECK_CTX(ctx)                                                            \
    do {                                                                              \
        if (!(ctx)) return KHB_ERR_ARG;                                               \
        if ((ctx)->sticky) return KHB_ERR_CUDA;                                       \
        KHB_CUDA((ctx), cudaSetDevice((ctx)->device));                                \
    } while (0)

static inline size_t div_up(size_t a, size_t b) { return (a + b - 1) / b; }

// ---- warp / block scan helpers ---------------------------------------------------------------
__device__ __forceinline__ u32 lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ u32 lanemask_lt()
{
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

template <typename T>
__device__ __forceinline__ T warp_incl_sum(T v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (u32)o) v += n;
    }
    return v;
}
template <typename T>
__device__ __forceinline__ T warp_incl_max(T v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        T n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane_id() >= (u32)o) v = n > v ? n : v;
    }
    return v;
}

// Block-wide exclusive sum over blockDim.x threads (multiple of 32, <= 1024).
// `ws` is a 33-entry shared array.  Returns the exclusive prefix; *total = block sum.
template <typename T>
__device__ __forceinline__ T block_excl_sum(T v, T *ws, T *total)
{
    const u32 w = threadIdx.x >> 5, l = lane_id(), nw = blockDim.x >> 5;
    T inc = warp_incl_sum(v);
    if (l == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = l < nw ? ws[l] : (T)0;
        T xi = warp_incl_sum(x);
        ws[l] = xi - x;
        if (l == 31) ws[32] = xi;
    }
    __syncthreads();
    T r = ws[w] + inc - v;
    *total = ws[32];
    __syncthreads();
    return r;
}
// Block-wide exclusive max (identity 0).
template <typename T>
__device__ __forceinline__ T block_excl_max(T v, T *ws, T *total)
{
    const u32 w = threadIdx.x >> 5, l = lane_id(), nw = blockDim.x >> 5;
    T inc = warp_incl_max(v);
    T up = __shfl_up_sync(0xffffffffu, inc, 1);
    T exc = l == 0 ? (T)0 : up;
    if (l == 31) ws[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = l < nw ? ws[l] : (T)0;
        T xi = warp_incl_max(x);
        T xu = __shfl_up_sync(0xffffffffu, xi, 1);
        ws[l] = l == 0 ? (T)0 : xu;
        if (l == 31) ws[32] = xi;
    }
    __syncthreads();
    T base = ws[w];
    T r = base > exc ? base : exc;
    *total = ws[32];
    __syncthreads();
    return r;
}
