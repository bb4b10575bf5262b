// peer.cu -- the multi-GPU exchange of the across-group stage, written into peer memory by the producing GPU.
//
// The across-group union (`kmc_tools complex` over all groups, /root/reference/workflow/rules/exp_type_1.smk:243-259) needs
// every copy of a k-mer on one GPU; the k-mer space is hash-range partitioned (part_of).  The NCCL route (dist.py:
// exchange_and_count) partitions the retained group sets into a send buffer (two sweeps + two host round trips), runs a
// size all-to-all, a payload all-to-all and copies the result into the store.  Here the exchange is ONE kernel per group,
// launched right behind the group's K5 on the context's stream:
//   every rank owns a receive buffer of `world` regions (region s is written by rank s only) and has every peer's buffer
//   mapped (CUDA IPC, NVLink P2P);  push_kernel reads the group's new keys from the local store once, counts them per owner
//   in shared memory, reserves space with ONE atomicAdd per (CTA, owner) on a LOCAL cursor -- no remote atomics, the sender
//   alone writes its region -- and stores every key straight into its owner's region over NVLink (posted 8/16-byte writes).
// The transfer of group g therefore overlaps the kernels of group g+1 on the receiving side, no send buffer exists, and
// the only collective left on the path is the 8 x 8 table of counts at the end (which is also the barrier that makes
// the pushed data visible).  A region that would overflow raises a flag; the caller then redoes the step over NCCL.
#include <stdlib.h>

#include "khb_common.cuh"

#define PP_BLOCK 256
#define PP_ITEMS 16
#define PP_MAXPARTS 64

struct khb_peer {
    int world = 0, rank = 0, key_bytes = 0;
    u64 region_keys = 0;           // capacity of one sender's region, in keys
    void *recv = nullptr;          // my receive buffer: world regions
    void *peer_base[PP_MAXPARTS];  // mapped base of every rank's receive buffer (own entry = recv)
    void **d_dst = nullptr;        // device copy: where MY region starts inside every rank's buffer
    u64 *d_cursor = nullptr;       // [world] keys pushed to every rank so far + [1] overflow flag
    u64 pushed_upto = 0;           // keys of the local group-set store already pushed
    bool opened = false;
    // KHB_PEER_ASYNC=1: the push runs on its own stream behind an event of the producing K5, so that the NVLink stores of
    // group g overlap the kernels of group g+1 (no gain measured on 2 GPUs: the SMs are busy either way; default off)
    cudaStream_t push_stream = nullptr;
    cudaEvent_t produced = nullptr;
    bool in_flight = false;
};

// A CTA takes PP_BLOCK * PP_ITEMS consecutive keys of the store, groups them by owner in shared memory and writes every
// owner's keys as ONE contiguous run: full 128-byte lines instead of scattered 8-byte stores, which is what NVLink wants
// (the scattered version reached 334 GB/s outbound per GPU on 8 B200s).
template <typename Key>
__global__ void __launch_bounds__(PP_BLOCK)
push_kernel(const Key *__restrict__ in, u64 n, u32 nparts, u64 cap, u64 *__restrict__ cursor, void *const *__restrict__ dst)
{
    constexpr int TILE = PP_BLOCK * PP_ITEMS;
    extern __shared__ __align__(16) unsigned char pp_smem[];
    Key *staged = (Key *)pp_smem;                                   // [TILE] keys grouped by owner
    unsigned char *owner = (unsigned char *)(staged + TILE);        // [TILE] owner of staged[j]
    __shared__ u32 sc[PP_MAXPARTS];      // keys per owner in this tile
    __shared__ u32 sstart[PP_MAXPARTS];  // first staged slot of every owner
    __shared__ u64 sbase[PP_MAXPARTS];   // reserved offset inside my region of every owner's buffer
    const u32 tid = threadIdx.x;
    const u64 begin = (u64)blockIdx.x * TILE;
    if (tid < PP_MAXPARTS) sc[tid] = 0;
    __syncthreads();
    Key keys[PP_ITEMS];
    u32 part[PP_ITEMS], slot[PP_ITEMS];
#pragma unroll
    for (int r = 0; r < PP_ITEMS; r++) {
        const u64 g = begin + (u64)r * PP_BLOCK + tid;
        part[r] = 0xffffffffu;
        if (g < n) {
            keys[r] = in[g];
            part[r] = part_of(keys[r], nparts);
            slot[r] = atomicAdd(&sc[part[r]], 1u);
        }
    }
    __syncthreads();
    if (tid < 32) {
        // exclusive scan of the (at most 64) owner counts by one warp, two owners per lane
        const u32 c0 = tid < nparts ? sc[tid] : 0u, c1 = tid + 32 < nparts ? sc[tid + 32] : 0u;
        const u32 i0 = warp_incl_sum(c0);
        const u32 t0 = __shfl_sync(0xffffffffu, i0, 31);
        const u32 i1 = warp_incl_sum(c1);
        sstart[tid] = i0 - c0;
        sstart[tid + 32] = t0 + i1 - c1;
    }
    if (tid < nparts) {
        const u32 c = sc[tid];
        sbase[tid] = c ? atomicAdd(&cursor[tid], (u64)c) : 0ull;
        if (c && sbase[tid] + c > cap) cursor[PP_MAXPARTS] = 1ull;  // region full: the caller falls back to NCCL
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < PP_ITEMS; r++) {
        if (part[r] == 0xffffffffu) continue;
        const u32 j = sstart[part[r]] + slot[r];
        staged[j] = keys[r];
        owner[j] = (unsigned char)part[r];
    }
    __syncthreads();
    const u32 cnt = (u32)(n - begin < (u64)TILE ? n - begin : (u64)TILE);
    for (u32 j = tid; j < cnt; j += PP_BLOCK) {
        const u32 d = owner[j];
        const u64 o = sbase[d] + (j - sstart[d]);
        if (o < cap) ((Key *)dst[d])[o] = staged[j];
    }
}

extern "C" {

// Allocate this rank's receive buffer (world regions of region_keys keys of key_bytes each) and return its IPC handle.
int khb_peer_alloc(khb_ctx *ctx, int world, int rank, int key_bytes, uint64_t region_keys, unsigned char *handle_out /* 64 bytes */)
{
    KHB_CHECK_CTX(ctx);
    if (world < 1 || world > PP_MAXPARTS || rank < 0 || rank >= world || (key_bytes != 8 && key_bytes != 16) || !region_keys || !handle_out)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_peer_alloc: bad arguments (world=%d rank=%d key_bytes=%d)", world, rank, key_bytes);
    if (ctx->peer) return khb_fail(ctx, KHB_ERR_STATE, "khb_peer_alloc: a peer exchange is already set up (khb_peer_close first)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    khb_peer *pp = new khb_peer();
    pp->world = world;
    pp->rank = rank;
    pp->key_bytes = key_bytes;
    pp->region_keys = region_keys;
    const size_t bytes = (size_t)world * region_keys * key_bytes;
    cudaError_t e = cudaMalloc(&pp->recv, bytes);
    if (e == cudaSuccess) e = cudaMalloc((void **)&pp->d_dst, PP_MAXPARTS * sizeof(void *));
    if (e == cudaSuccess) e = cudaMalloc((void **)&pp->d_cursor, (PP_MAXPARTS + 1) * sizeof(u64));
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, pp->recv);
    const char *as = getenv("KHB_PEER_ASYNC");
    if (e == cudaSuccess && as && atoi(as) != 0) {  // measured on 2 B200s: 136.0 ms per step with, 135.6 ms without -- off by default
        e = cudaStreamCreateWithFlags(&pp->push_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pp->produced, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (pp->recv) cudaFree(pp->recv);
        if (pp->d_dst) cudaFree(pp->d_dst);
        if (pp->d_cursor) cudaFree(pp->d_cursor);
        if (pp->push_stream) cudaStreamDestroy(pp->push_stream);
        if (pp->produced) cudaEventDestroy(pp->produced);
        delete pp;
        return khb_fail(ctx, e == cudaErrorMemoryAllocation ? KHB_ERR_NOMEM : KHB_ERR_CUDA, "khb_peer_alloc (%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, 64);
    ctx->peer = pp;
    return KHB_OK;
}

// Map every rank's receive buffer.  handles: world x 64 bytes, in rank order (the own entry is ignored).
int khb_peer_open(khb_ctx *ctx, const unsigned char *handles)
{
    KHB_CHECK_CTX(ctx);
    khb_peer *pp = ctx->peer;
    if (!pp || pp->opened || !handles) return khb_fail(ctx, KHB_ERR_STATE, "khb_peer_open: call khb_peer_alloc first (once)");
    void *dst[PP_MAXPARTS] = {nullptr};
    for (int r = 0; r < pp->world; r++) {
        if (r == pp->rank) {
            pp->peer_base[r] = pp->recv;
        } else {
            cudaIpcMemHandle_t h;
            memcpy(&h, handles + (size_t)r * 64, 64);
            void *p = nullptr;
            cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
            if (e != cudaSuccess) {
                cudaGetLastError();
                for (int q = 0; q < r; q++)
                    if (q != pp->rank) cudaIpcCloseMemHandle(pp->peer_base[q]);
                return khb_fail(ctx, KHB_ERR_CUDA, "khb_peer_open: cudaIpcOpenMemHandle for rank %d: %s", r, cudaGetErrorString(e));
            }
            pp->peer_base[r] = p;
        }
        dst[r] = (char *)pp->peer_base[r] + (size_t)pp->rank * pp->region_keys * pp->key_bytes;  // my region inside rank r's buffer
    }
    KHB_CUDA(ctx, cudaMemcpyAsync(pp->d_dst, dst, sizeof(dst), cudaMemcpyHostToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(pp->d_cursor, 0, (PP_MAXPARTS + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pp->opened = true;
    return KHB_OK;
}

// Start a new exchange round: cursors to zero, nothing of the store pushed yet.  Every rank must have finished
// khb_peer_import of the previous round before any rank pushes again (the caller's histogram all-reduce orders that).
int khb_peer_begin(khb_ctx *ctx)
{
    KHB_CHECK_CTX(ctx);
    khb_peer *pp = ctx->peer;
    if (!pp || !pp->opened) return khb_fail(ctx, KHB_ERR_STATE, "khb_peer_begin: no peer exchange set up");
    KHB_CUDA(ctx, cudaMemsetAsync(pp->d_cursor, 0, (PP_MAXPARTS + 1) * sizeof(u64), ctx->stream));
    pp->pushed_upto = 0;
    return KHB_OK;
}

// Push the keys the group-set store gained since the last push to their owners (asynchronous, on the context's stream).
int khb_peer_push(khb_ctx *ctx)
{
    KHB_CHECK_CTX(ctx);
    khb_peer *pp = ctx->peer;
    if (!pp || !pp->opened) return khb_fail(ctx, KHB_ERR_STATE, "khb_peer_push: no peer exchange set up");
    if (ctx->gs_len < pp->pushed_upto) return khb_fail(ctx, KHB_ERR_STATE, "khb_peer_push: the group-set store shrank; call khb_peer_begin after a reset");
    const u64 n = ctx->gs_len - pp->pushed_upto;
    if (!n) return KHB_OK;
    const size_t W = (size_t)khb_key_bytes(ctx->gs_k);
    if ((int)W != pp->key_bytes) return khb_fail(ctx, KHB_ERR_STATE, "khb_peer_push: the exchange was set up for %d-byte keys", pp->key_bytes);
    const char *src = (const char *)ctx->gs_buf + pp->pushed_upto * W;
    const u64 blocks = div_up(n, PP_BLOCK * PP_ITEMS);
    cudaStream_t st = ctx->stream;
    if (pp->push_stream) {
        KHB_CUDA(ctx, cudaEventRecord(pp->produced, ctx->stream));       // behind the K5 that wrote these keys (and khb_peer_begin's memset)
        KHB_CUDA(ctx, cudaStreamWaitEvent(pp->push_stream, pp->produced, 0));
        st = pp->push_stream;
        ctx->prof_stream = st;
        pp->in_flight = true;
    }
    khb_prof_begin(ctx, KHB_K_PARTITION);
    const size_t shm = (size_t)PP_BLOCK * PP_ITEMS * (W + 1);
    if (W == 8) {
        push_kernel<Key64><<<(unsigned)blocks, PP_BLOCK, shm, st>>>((const Key64 *)src, n, (u32)pp->world, pp->region_keys, pp->d_cursor, pp->d_dst);
    } else {
        cudaFuncSetAttribute(push_kernel<Key128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);   // per device: every call
        push_kernel<Key128><<<(unsigned)blocks, PP_BLOCK, shm, st>>>((const Key128 *)src, n, (u32)pp->world, pp->region_keys, pp->d_cursor, pp->d_dst);
    }
    ctx->launches++;
    cudaError_t le = cudaGetLastError();
    khb_prof_end(ctx, KHB_K_PARTITION, 2 * n * W);
    ctx->prof_stream = nullptr;
    if (le != cudaSuccess) return khb_cuda_fail(ctx, le, "push_kernel launch", __FILE__, __LINE__);
    pp->pushed_upto = ctx->gs_len;
    return KHB_OK;
}

// Wait for this rank's pushes and report how many keys went to every rank (h_counts[world]) and whether a region overflowed.
int khb_peer_counts(khb_ctx *ctx, uint64_t *h_counts, int *overflow)
{
    KHB_CHECK_CTX(ctx);
    khb_peer *pp = ctx->peer;
    if (!pp || !pp->opened || !h_counts) return khb_fail(ctx, KHB_ERR_STATE, "khb_peer_counts: no peer exchange set up");
    u64 *h = ctx->h_mail + 20000;
    int rc = khb_peer_wait(ctx);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(h, pp->d_cursor, (PP_MAXPARTS + 1) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < pp->world; r++) h_counts[r] = h[r];
    if (overflow) *overflow = h[PP_MAXPARTS] ? 1 : 0;
    return KHB_OK;
}

// Replace the retained group sets by what the ranks pushed here: h_recv_counts[s] keys in region s.  Call only after every
// rank's khb_peer_counts returned (the count exchange between the ranks is that barrier).
int khb_peer_import(khb_ctx *ctx, const uint64_t *h_recv_counts, int k, int n_groups, int hashed);

}  // extern "C"

int khb_peer_wait(khb_ctx *ctx)
{
    khb_peer *pp = ctx->peer;
    if (pp && pp->push_stream && pp->in_flight) {
        KHB_CUDA(ctx, cudaStreamSynchronize(pp->push_stream));
        pp->in_flight = false;
    }
    return KHB_OK;
}

extern "C" {

// Drop the mappings of the other ranks' buffers (pushing is over).  Every rank must have done this before any rank frees
// its buffer with khb_peer_close: put a barrier between the two calls.
int khb_peer_unmap(khb_ctx *ctx)
{
    if (!ctx || !ctx->peer) return KHB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    khb_peer *pp = ctx->peer;
    if (pp->push_stream) cudaStreamSynchronize(pp->push_stream);
    if (pp->opened)
        for (int r = 0; r < pp->world; r++)
            if (r != pp->rank && pp->peer_base[r]) {
                cudaIpcCloseMemHandle(pp->peer_base[r]);
                pp->peer_base[r] = nullptr;
            }
    pp->opened = false;
    return KHB_OK;
}

int khb_peer_close(khb_ctx *ctx)
{
    if (!ctx || !ctx->peer) return KHB_OK;
    khb_peer_unmap(ctx);
    khb_peer *pp = ctx->peer;
    if (pp->recv) cudaFree(pp->recv);
    if (pp->d_dst) cudaFree(pp->d_dst);
    if (pp->d_cursor) cudaFree(pp->d_cursor);
    if (pp->push_stream) cudaStreamDestroy(pp->push_stream);
    if (pp->produced) cudaEventDestroy(pp->produced);
    delete pp;
    ctx->peer = nullptr;
    return KHB_OK;
}

uint64_t khb_peer_region_keys(const khb_ctx *ctx) { return ctx && ctx->peer ? ctx->peer->region_keys : 0; }

}  // extern "C"

// The minimizer-bin group stage pushes every bin's distinct keys to their owners from its own end-of-bin pass (bins.cu), so that
// khb_peer_push finds nothing left to do.  KHB_PEER_FUSE=0 keeps the separate push kernel.
int khb_peer_route_get(khb_ctx *ctx, int key_bytes, khb_peer_route *out)
{
    out->world = 0;
    out->flags = 0;
    out->cap = 0;
    out->cursor = nullptr;
    out->dst = nullptr;
    khb_peer *pp = ctx->peer;
    static int fuse = -1;
    if (fuse < 0) {
        const char *e = getenv("KHB_PEER_FUSE");
        fuse = e ? atoi(e) : 1;
    }
    if (!fuse || !pp || !pp->opened || pp->key_bytes != key_bytes || pp->push_stream) return KHB_OK;
    if (pp->pushed_upto != ctx->gs_len) return KHB_OK;   // earlier keys of the store are still waiting for khb_peer_push: keep the order simple
    static int sorted = -1;
    if (sorted < 0) {
        const char *e = getenv("KHB_PEER_SORTED");
        sorted = e ? atoi(e) : 0;
    }
    out->world = (u32)pp->world;
    out->flags = sorted ? 1u : 0u;
    out->cap = pp->region_keys;
    out->cursor = pp->d_cursor;
    out->dst = pp->d_dst;
    return KHB_OK;
}
void khb_peer_mark_pushed(khb_ctx *ctx)
{
    if (ctx->peer) ctx->peer->pushed_upto = ctx->gs_len;
}
int khb_peer_poison(khb_ctx *ctx)
{
    khb_peer *pp = ctx->peer;
    if (!pp || !pp->opened) return KHB_OK;
    static const u64 one = 1;
    KHB_CUDA(ctx, cudaMemcpyAsync(pp->d_cursor + PP_MAXPARTS, &one, sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    return KHB_OK;
}

// used by khb_peer_import (api.cu owns the group-set store)
int khb_peer_regions(khb_ctx *ctx, const void **recv, u64 *region_keys, int *world, int *key_bytes)
{
    khb_peer *pp = ctx->peer;
    if (!pp || !pp->opened) return khb_fail(ctx, KHB_ERR_STATE, "khb_peer_import: no peer exchange set up");
    *recv = pp->recv;
    *region_keys = pp->region_keys;
    *world = pp->world;
    *key_bytes = pp->key_bytes;
    return KHB_OK;
}
