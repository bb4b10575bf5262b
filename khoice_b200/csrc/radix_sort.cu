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
#include "lookback.cuh"

#define RS_BLOCK 512
#define RS_WARPS (RS_BLOCK / 32)

template <typename Key> struct RsCfg;
template <> struct RsCfg<Key64> { static constexpr int ITEMS = 12; };
template <> struct RsCfg<Key128> { static constexpr int ITEMS = 8; };

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
template <typename Key, int ITEMS>
__global__ void __launch_bounds__(RS_BLOCK)
radix_hist_kernel(const Key *__restrict__ in, const u64 *__restrict__ seg_off, const u64 *__restrict__ seg_tile,
                  int nseg, u64 ntiles, int npass, u32 *__restrict__ hist /* [nseg][npass][256] */)
{
    constexpr int TILE = RS_BLOCK * ITEMS;
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
        if (seg != cur) {
            if (cur >= 0) {
                __syncthreads();
                for (int i = tid; i < npass * 256; i += RS_BLOCK) {
                    const u32 c = sh[i];
                    if (c) atomicAdd(&hist[(size_t)cur * npass * 256 + i], c);
                    sh[i] = 0;
                }
                __syncthreads();
            }
            cur = seg;
        }
        const u64 begin = seg_off[seg] + (t - seg_tile[seg]) * TILE;
        const u64 end = seg_off[seg + 1];
        const u32 n = (u32)(end - begin < (u64)TILE ? end - begin : (u64)TILE);
#pragma unroll 4
        for (u32 i = tid; i < n; i += RS_BLOCK) {
            const Key key = in[begin + i];
            for (int p = 0; p < npass; p++) atomicAdd(&sh[p * 256 + key_digit(key, p)], 1u);
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
template <typename Key, int ITEMS>
__global__ void __launch_bounds__(RS_BLOCK, 2)
onesweep_kernel(const Key *__restrict__ in, Key *__restrict__ out, const u64 *__restrict__ seg_off,
                const u64 *__restrict__ seg_tile, int nseg, int pass, int npass,
                const u32 *__restrict__ bin_base /* [nseg][npass][256], exclusive */, u64 *__restrict__ lookback,
                u32 *__restrict__ ticket, u32 epoch)
{
    constexpr int TILE = RS_BLOCK * ITEMS;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Key *sorted = (Key *)smem_raw;                                          // [TILE]
    unsigned short *wcnt = (unsigned short *)(smem_raw + sizeof(Key) * TILE);  // [RS_WARPS][256]
    u32 *digit_start = (u32 *)(wcnt + RS_WARPS * 256);                      // [256]
    u64 *glob_base = (u64 *)(digit_start + 256);                            // [256]
    u64 *ws = glob_base + 256;                                              // [33] scan scratch
    __shared__ u32 s_tile;
    __shared__ int s_seg;

    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    if (tid == 0) {
        const u32 t = atomicAdd(ticket, 1u);
        s_tile = t;
        s_seg = find_segment(seg_tile, nseg, t);
    }
    for (int i = tid; i < RS_WARPS * 256 / 2; i += RS_BLOCK) ((u32 *)wcnt)[i] = 0;
    __syncthreads();
    const u64 tile = s_tile;
    const int seg = s_seg;
    const u64 first_tile = seg_tile[seg];
    const u64 seg_begin = seg_off[seg];
    const u64 begin = seg_begin + (tile - first_tile) * TILE;
    const u64 seg_end = seg_off[seg + 1];
    const u32 n = (u32)(seg_end - begin < (u64)TILE ? seg_end - begin : (u64)TILE);

    // load: warp-striped, memory order = (warp, item, lane)
    Key keys[ITEMS];
    u32 rank[ITEMS];
    const u32 wbase = warp * (32 * ITEMS);
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const u32 idx = wbase + r * 32 + lane;
        keys[r] = idx < n ? in[begin + idx] : key_max<Key>();
    }
    // rank within the warp's chunk (stable): peers = lanes holding the same digit in this round
    unsigned short *mycnt = wcnt + warp * 256;
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const u32 d = key_digit(keys[r], pass);
        const u32 peers = __match_any_sync(0xffffffffu, d);
        const u32 below = __popc(peers & lanemask_lt());
        u32 old = 0;
        if (below == 0) {
            old = mycnt[d];
            mycnt[d] = (unsigned short)(old + __popc(peers));
        }
        old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
        rank[r] = old + below;
        __syncwarp();
    }
    __syncthreads();
    // per digit: exclusive scan across warps (in place) -> tile count
    u32 count = 0;
    if (tid < 256) {
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            const u32 c = wcnt[w * 256 + tid];
            wcnt[w * 256 + tid] = (unsigned short)count;
            count += c;
        }
    }
    // exclusive scan across digits (padding keys of a partial tile carry digit 255 and stay at the end)
    u64 total;
    const u32 dstart = (u32)block_excl_sum<u64>((u64)count, ws, &total);
    u64 *lb = lookback + tile * 256 + tid;
    u64 first_hop = 0;
    if (tid < 256) {
        digit_start[tid] = dstart;
        if (tid == 255) count -= (u32)(TILE - n);  // padding is not data
        if (tile == first_tile) {
            lb_store(lb, lb_pack(LB_PREFIX, count, epoch));
        } else {
            lb_store(lb, lb_pack(LB_AGG, count, epoch));
            first_hop = lb_load(lb - 256);  // issued now, consumed after the reorder below
        }
    }
    __syncthreads();
    // reorder through shared memory
#pragma unroll
    for (int r = 0; r < ITEMS; r++) {
        const u32 d = key_digit(keys[r], pass);
        sorted[digit_start[d] + wcnt[warp * 256 + d] + rank[r]] = keys[r];
    }
    // finish the look-back (256 digit threads)
    if (tid < 256) {
        u64 excl = 0;
        if (tile != first_tile) {
            u64 t = tile - 1;
            u64 e = first_hop;
            for (;;) {
                u32 st = lb_status(e, epoch);
                while (st == 0) {
                    e = lb_load(lookback + t * 256 + tid);
                    st = lb_status(e, epoch);
                }
                excl += e & LB_VALUE_MASK;
                if (st == LB_PREFIX || t == first_tile) break;
                --t;
                e = lb_load(lookback + t * 256 + tid);
            }
            lb_store(lb, lb_pack(LB_PREFIX, excl + count, epoch));
        }
        glob_base[tid] = seg_begin + (u64)bin_base[((size_t)seg * npass + pass) * 256 + tid] + excl - (u64)dstart;
    }
    __syncthreads();
    // coalesced store: position j of the sorted tile goes to glob_base[digit] + j
#pragma unroll 4
    for (u32 j = tid; j < n; j += RS_BLOCK) {
        const Key key = sorted[j];
        out[glob_base[key_digit(key, pass)] + j] = key;
    }
}

// ---- host side ---------------------------------------------------------------------------------------
template <typename Key>
static int sort_impl(khb_ctx *ctx, Key *d_keys, Key *d_tmp, const u64 *h_seg_off, int nseg, int k, int *result_in_tmp)
{
    constexpr int ITEMS = RsCfg<Key>::ITEMS;
    constexpr int TILE = RS_BLOCK * ITEMS;
    const int npass = (2 * k + 7) / 8;
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
    int rc = khb_scratch_get(ctx, SCR_MISC, sizeof(u64) * 2 * ((size_t)nseg + 1) + 256, &p);
    if (rc) { free(h_tab); return rc; }
    u64 *d_off = (u64 *)p, *d_tile = d_off + nseg + 1;
    u32 *d_ticket = (u32 *)(d_tile + nseg + 1);  // 64 tickets max (one per pass)
    // synchronous small copy: h_tab is pageable and freed right after
    KHB_CUDA(ctx, cudaMemcpyAsync(d_off, h_tab, sizeof(u64) * 2 * ((size_t)nseg + 1), cudaMemcpyHostToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free(h_tab);
    KHB_CUDA(ctx, cudaMemsetAsync(d_ticket, 0, 64 * sizeof(u32), ctx->stream));

    const size_t hist_bytes = (size_t)nseg * npass * 256 * sizeof(u32);
    rc = khb_scratch_get(ctx, SCR_HIST, hist_bytes, &p);
    if (rc) return rc;
    u32 *d_hist = (u32 *)p;
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, hist_bytes, ctx->stream));
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
        radix_hist_kernel<Key, ITEMS><<<(unsigned)grid, RS_BLOCK, shm, ctx->stream>>>(d_keys, d_off, d_tile, nseg, ntiles, npass, d_hist);
        KHB_LAUNCH_CHECK(ctx);
        const int nhist = nseg * npass;
        radix_scan_kernel<<<(unsigned)div_up(nhist, 8), 256, 0, ctx->stream>>>(d_hist, nhist);
        KHB_LAUNCH_CHECK(ctx);
        khb_prof_end(ctx, KHB_K_RADIX_HIST, n_keys * sizeof(Key));
    }
    const size_t shm = sizeof(Key) * TILE + RS_WARPS * 256 * sizeof(unsigned short) + 256 * sizeof(u32) + 256 * sizeof(u64) + 33 * sizeof(u64);
    static bool attr_set[2] = {false, false};
    const int which = sizeof(Key) == 8 ? 0 : 1;
    if (!attr_set[which]) {
        KHB_CUDA(ctx, cudaFuncSetAttribute(onesweep_kernel<Key, ITEMS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
        attr_set[which] = true;
    }
    Key *src = d_keys, *dst = d_tmp;
    for (int pass = 0; pass < npass; pass++) {
        khb_prof_begin(ctx, KHB_K_ONESWEEP);
        onesweep_kernel<Key, ITEMS><<<(unsigned)ntiles, RS_BLOCK, shm, ctx->stream>>>(
            src, dst, d_off, d_tile, nseg, pass, npass, d_hist, d_lb, d_ticket + pass, (u32)(pass + 1));
        KHB_LAUNCH_CHECK(ctx);
        khb_prof_end(ctx, KHB_K_ONESWEEP, 2 * n_keys * sizeof(Key));  // read + write every key once
        Key *t = src; src = dst; dst = t;
    }
    *result_in_tmp = (npass & 1);
    return KHB_OK;
}

int khb_sort_keys_impl(khb_ctx *ctx, void *d_keys, void *d_tmp, const u64 *h_seg_off, int nseg, int k, int *result_in_tmp)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "khb_sort_keys: k=%d outside 1..64", k);
    return k <= 32 ? sort_impl<Key64>(ctx, (Key64 *)d_keys, (Key64 *)d_tmp, h_seg_off, nseg, k, result_in_tmp)
                   : sort_impl<Key128>(ctx, (Key128 *)d_keys, (Key128 *)d_tmp, h_seg_off, nseg, k, result_in_tmp);
}
