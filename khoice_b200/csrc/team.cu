// team.cu -- the receive buffers of a team: ONE group of experiment type 1 sharded over several GPUs.
//
// The reference runs the per-genome and per-group rules of a group (/root/reference/workflow/rules/exp_type_1.smk:156-191) as
// independent processes on one host; dealing whole groups to GPUs (dist.py) keeps that shape but leaves GPUs idle when the groups do
// not divide evenly (20 groups on 8 GPUs: 3/3/3/3/2/2/2/2) and cannot use a second GPU for a single large group.  A team shards the
// group itself: the genomes are split over the members, the minimizer bins over owners, and the super-k-mer records of pass P cross
// NVLink once, stored by the producing GPU straight into the owner's record buffer (bins.cu: mb_partition_kernel<KW, true>).  This file
// only owns the memory: every member allocates two receive buffers (they alternate from group to group, so a member may already
// partition the next group while another still counts this one) and maps the other members' through CUDA IPC.
#include <stdlib.h>

#include "khb_common.cuh"

extern "C" {

// Allocate this member's two receive buffers of half_bytes each and return the IPC handle of the allocation.
int khb_team_alloc(khb_ctx *ctx, int team_size, int member, uint64_t half_bytes, unsigned char *handle_out /* 64 bytes */)
{
    KHB_CHECK_CTX(ctx);
    if (team_size < 2 || team_size > KHB_TEAM_MAX || member < 0 || member >= team_size || !half_bytes || !handle_out)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_team_alloc: bad arguments (team_size=%d member=%d)", team_size, member);
    if (ctx->team) return khb_fail(ctx, KHB_ERR_STATE, "khb_team_alloc: a team is already set up (khb_team_close first)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    khb_team *tm = new khb_team();
    tm->size = team_size;
    tm->member = member;
    tm->half_bytes = ((size_t)half_bytes + 255) & ~(size_t)255;
    for (int t = 0; t < KHB_TEAM_MAX; t++) tm->peer_base[t] = nullptr;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaMalloc(&tm->recv, 2 * tm->half_bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, tm->recv);
    if (e != cudaSuccess) {
        cudaGetLastError();
        if (tm->recv) cudaFree(tm->recv);
        const size_t bytes = 2 * tm->half_bytes;
        delete tm;
        return khb_fail(ctx, e == cudaErrorMemoryAllocation ? KHB_ERR_NOMEM : KHB_ERR_CUDA, "khb_team_alloc (%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    memcpy(handle_out, &h, 64);
    ctx->team = tm;
    return KHB_OK;
}

// Map every member's receive buffers.  handles: team_size x 64 bytes, in member order (the own entry is ignored).
int khb_team_open(khb_ctx *ctx, const unsigned char *handles)
{
    KHB_CHECK_CTX(ctx);
    khb_team *tm = ctx->team;
    if (!tm || tm->opened || !handles) return khb_fail(ctx, KHB_ERR_STATE, "khb_team_open: call khb_team_alloc first (once)");
    for (int t = 0; t < tm->size; t++) {
        if (t == tm->member) {
            tm->peer_base[t] = tm->recv;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)t * 64, 64);
        void *p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < t; q++)
                if (q != tm->member && tm->peer_base[q]) {
                    cudaIpcCloseMemHandle(tm->peer_base[q]);
                    tm->peer_base[q] = nullptr;
                }
            return khb_fail(ctx, KHB_ERR_CUDA, "khb_team_open: cudaIpcOpenMemHandle for member %d: %s", t, cudaGetErrorString(e));
        }
        tm->peer_base[t] = p;
    }
    tm->opened = true;
    return KHB_OK;
}

// Drop the mappings of the other members' buffers.  Every member must have done this before any member frees its buffers with
// khb_team_close: put a barrier between the two calls.
int khb_team_unmap(khb_ctx *ctx)
{
    if (!ctx || !ctx->team) return KHB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    khb_team *tm = ctx->team;
    if (tm->opened)
        for (int t = 0; t < tm->size; t++)
            if (t != tm->member && tm->peer_base[t]) {
                cudaIpcCloseMemHandle(tm->peer_base[t]);
                tm->peer_base[t] = nullptr;
            }
    tm->opened = false;
    return KHB_OK;
}

int khb_team_close(khb_ctx *ctx)
{
    if (!ctx || !ctx->team) return KHB_OK;
    khb_team_unmap(ctx);
    khb_team *tm = ctx->team;
    if (tm->recv) cudaFree(tm->recv);
    delete tm;
    ctx->team = nullptr;
    return KHB_OK;
}

}  // extern "C"
