// hashset.cu -- group stage without a sort: K2 fused with an open-addressing table of (k-mer, genome bit set) records.
//
// Replaces, for one group, `kmc` per genome + `kmc_tools transform set_counts 1` + `kmc_tools complex (set1 + ... + setN)`
// + `transform histogram` (reference call sites /root/reference/workflow/rules/exp_type_1.smk:156-191): for every distinct
// canonical k-mer x of the group, c(x) = number of genomes that contain x, h[c] = #{x : c(x) = c}.
//
// The single-sort path moves every window four times through a radix pass (10 bytes each way) although 85 % of a group's
// windows are k-mers another genome of the group already delivered.  Here a window is looked up where it is produced:
//   record (R u32 words, R = 4 / 8 / 16):  words 0-1 = h(canonical) + 1 (0 = empty), words 2.. = one bit per genome
//   hash_insert_kernel   canonical k-mer of a window (same bit arithmetic as extract64_kernel), mixed with the bijective
//                        mixer, slot = top bits of the mixed key, linear probing; ONE 16-byte load answers "is the key
//                        there, is my genome's bit there"; an empty slot is claimed with a 64-bit compare-and-swap, a
//                        missing bit is set with a non-returning atomicOr (RED).  CTAs take 4096-window chunks in
//                        GENOME-MINOR order (chunk v = position block v / N of genome v % N): genomes of one group are
//                        largely collinear, so the records a chunk touches were touched microseconds earlier by the
//                        same position block of the other genomes and are still in L2 (126 MB); without collinearity
//                        the kernel runs at the DRAM random-sector rate instead (KHB_HASH_ORDER=genome measures that).
//   hash_count_kernel    streams the table once: popcount of a record = c(x) -> histogram (shared memory), the key goes
//                        to the group-set store (warp-aggregated reservation, unordered: the across-group stage sorts),
//                        sum of c = sum of the per-genome set sizes; every non-empty record is zeroed again, so the
//                        table is clean for the next group without a memset.
//                        PIVOT (exp_type_2.smk:354-380): records whose last genome bit is set are the pivot's k-mers.
// Results are exact (full keys are compared); the table holds at least as many slots as the group has windows, and a probe
// sequence longer than the limit raises a flag on which the caller falls back to the sort path (only reachable when nearly
// every window of a group is a distinct k-mer AND the slot count is within a few percent of the window count).
// Measured on config 2 (profiles/r1s3_hash_group_stage.md): insert 11.8 ms + count 1.6 ms per group against 11.0 ms for the
// single-sort path -- every (k-mer, genome) pair costs one write to a random 32-byte sector, which is as expensive as two
// coalesced radix passes; the path is therefore opt-in (KHB_GROUP_MODE=hash / khb_set_group_mode), not the default.
// Algorithmic bytes: insert B/4 + B/8 (packed stream) + 16 per window (its record), count: R*4 per slot + 8 per distinct key.
#include <stdlib.h>

#include "khb_common.cuh"

#define HS_BLOCK 256
#define HS_CHUNK 4096  // windows per CTA

__device__ __forceinline__ u64 hs_swap_pairs(u64 r)
{
    return ((r >> 1) & 0x5555555555555555ull) | ((r & 0x5555555555555555ull) << 1);
}

// INTERLEAVE: chunk v -> (genome v % nseg, position block v / nseg); otherwise (genome v / n_pb, block v % n_pb)
template <int R>
__global__ void __launch_bounds__(HS_BLOCK, 8)
hash_insert_kernel(const u64 *__restrict__ codes, const u32 *__restrict__ valid, int k, const u64 *__restrict__ seg_off, u32 nseg, u32 n_pb,
                   int interleave, u32 *__restrict__ table, int shift, u64 mask, u32 max_probe, u32 *__restrict__ overflow, int diag)
{
    const u32 v = blockIdx.x;
    const u32 g = interleave ? v % nseg : v / n_pb;
    const u32 pb = interleave ? v / nseg : v % n_pb;
    const u64 seg_lo = __ldg(seg_off + g), seg_hi = __ldg(seg_off + g + 1);
    const u64 start = seg_lo + (u64)pb * HS_CHUNK;
    if (start >= seg_hi) return;
    const u64 end = start + HS_CHUNK < seg_hi ? start + HS_CHUNK : seg_hi;
    const u64 ones_k = (1ull << k) - 1ull;  // k <= 31 here
    const int rs = 64 - 2 * k;
    const u32 bit = 1u << (g & 31u);
    const u32 bw = g >> 5;  // word of the genome's bit inside the record's bit area
    for (u64 i = start + threadIdx.x; i < end; i += HS_BLOCK) {
        const u64 m = i >> 5;
        const u32 o = (u32)(i & 31);
        const u64 c0 = __ldg(codes + m), c1 = __ldg(codes + m + 1);
        const u64 vv = ((u64)__ldg(valid + m) << 32) | (u64)__ldg(valid + m + 1);
        if (((vv << o) >> (64 - k)) != ones_k) continue;
        const u64 x = o ? ((c0 << (2 * o)) | (c1 >> (64 - 2 * o))) : c0;
        const u64 fwd = x >> rs;
        const u64 rc = hs_swap_pairs(__brevll(~x) << rs >> rs);
        const u64 mixed = kmer_mix64(fwd < rc ? fwd : rc, k);
        const u64 key = mixed + 1ull;  // never 0
        u64 slot = (mixed >> shift) & mask;
        u32 probes = 0;
        for (;;) {
            u32 *rec = table + slot * R;
            u64 cur;
            u32 bits;
            if (R == 4) {
                const ulonglong2 q = __ldcg((const ulonglong2 *)rec);
                cur = q.x;
                bits = bw ? (u32)(q.y >> 32) : (u32)q.y;
            } else {
                cur = __ldcg((const u64 *)rec);
                bits = 0;
            }
            bool mine = cur == key;
            if (diag == 2) break;
            if (cur == 0ull) {
                const u64 old = atomicCAS((u64 *)rec, 0ull, key);
                mine = old == 0ull || old == key;
                bits = 0;  // a stale "bit missing" only costs a redundant atomicOr
            } else if (R != 4 && mine) {
                bits = __ldcg(rec + 2 + bw);
            }
            if (mine) {
                if (diag == 1) break;
                if (diag == 3) {
                    if (!(bits & bit)) rec[2 + bw] = bits | bit;
                    break;
                }
                if (!(bits & bit)) atomicOr(rec + 2 + bw, bit);
                break;
            }
            slot = (slot + 1) & mask;
            if (++probes > max_probe) {
                *overflow = 1u;
                break;
            }
        }
    }
}

#define HS_PER 8  // records per thread and round of hash_count_kernel

template <int R, bool PIVOT>
__global__ void __launch_bounds__(HS_BLOCK)
hash_count_kernel(u32 *__restrict__ table, u64 n_slots, u32 cs, u32 nbins, u32 pivot_gid, u64 *__restrict__ hist, u64 *__restrict__ out_keys,
                  u64 *__restrict__ d_cursor, u64 *__restrict__ d_pairs, u64 *__restrict__ out_pivot, u64 *__restrict__ d_pcursor)
{
    extern __shared__ u32 sh_hist[];  // [nbins+1]
    const u32 tid = threadIdx.x, lane = lane_id();
    for (u32 i = tid; i <= nbins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    u64 my_pairs = 0;
    // a warp takes 32 * HS_PER consecutive records per round (n_slots is a power of two >= 1024) and reserves its output
    // range with ONE atomicAdd per round: the cursor is a single address, and same-address atomics serialise in L2
    const u64 n_warps = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 base = ((u64)blockIdx.x * (blockDim.x >> 5) + (tid >> 5)) * (32 * HS_PER); base < n_slots; base += n_warps * (32 * HS_PER)) {
        uint4 q[HS_PER];
#pragma unroll
        for (int j = 0; j < HS_PER; j++) q[j] = __ldcs((const uint4 *)(table + (base + j * 32 + lane) * R));
        u32 emit_m = 0, piv_m = 0;
#pragma unroll
        for (int j = 0; j < HS_PER; j++) {
            if ((q[j].x | q[j].y) == 0u) continue;
            uint4 *rec = (uint4 *)(table + (base + j * 32 + lane) * R);
            u32 c = __popc(q[j].z) + __popc(q[j].w);
            bool pv = false;
            if (PIVOT && (pivot_gid >> 5) < 2) pv = ((pivot_gid >> 5 ? q[j].w : q[j].z) >> (pivot_gid & 31u)) & 1u;
#pragma unroll
            for (int t = 1; t < R / 4; t++) {
                const uint4 e = __ldcs(rec + t);
                c += __popc(e.x) + __popc(e.y) + __popc(e.z) + __popc(e.w);
                if (PIVOT) {
                    const u32 w0 = 4 * t - 2;  // index (in the bit area) of e.x
                    const u32 pw = pivot_gid >> 5;
                    if (pw >= w0 && pw < w0 + 4) {
                        const u32 word = pw == w0 ? e.x : pw == w0 + 1 ? e.y : pw == w0 + 2 ? e.z : e.w;
                        pv = (word >> (pivot_gid & 31u)) & 1u;
                    }
                }
            }
#pragma unroll
            for (int t = 0; t < R / 4; t++) rec[t] = make_uint4(0u, 0u, 0u, 0u);  // clean for the next group
            my_pairs += c - (pv ? 1u : 0u);
            if (!PIVOT || pv) {
                const u32 cc = c > cs ? cs : c;
                if (cc <= nbins) atomicAdd(&sh_hist[cc], 1u);
            }
            if (!(PIVOT && pv && c == 1)) emit_m |= 1u << j;
            if (PIVOT && pv) piv_m |= 1u << j;
        }
        {
            const u32 n = __popc(emit_m);
            const u32 inc = warp_incl_sum(n);
            const u32 total = __shfl_sync(0xffffffffu, inc, 31);
            if (total) {
                u64 ob = 0;
                if (lane == 31) ob = atomicAdd(d_cursor, (u64)total);
                ob = __shfl_sync(0xffffffffu, ob, 31) + (inc - n);
                if (out_keys != nullptr) {
#pragma unroll
                    for (int j = 0; j < HS_PER; j++)
                        if (emit_m & (1u << j)) out_keys[ob++] = (((u64)q[j].y << 32) | q[j].x) - 1ull;
                }
            }
        }
        if (PIVOT) {
            const u32 n = __popc(piv_m);
            const u32 inc = warp_incl_sum(n);
            const u32 total = __shfl_sync(0xffffffffu, inc, 31);
            if (total) {
                u64 ob = 0;
                if (lane == 31) ob = atomicAdd(d_pcursor, (u64)total);
                ob = __shfl_sync(0xffffffffu, ob, 31) + (inc - n);
                if (out_pivot != nullptr) {
#pragma unroll
                    for (int j = 0; j < HS_PER; j++)
                        if (piv_m & (1u << j)) out_pivot[ob++] = (((u64)q[j].y << 32) | q[j].x) - 1ull;
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o);
    if (lane == 0 && my_pairs) atomicAdd(d_pairs, my_pairs);
    __syncthreads();
    for (u32 i = tid; i <= nbins; i += blockDim.x) {
        const u32 v = sh_hist[i];
        if (v) atomicAdd(&hist[i], (u64)v);
    }
}

// Table geometry for a group of n_genomes genomes with n_sym windows, or 0 bytes if the hash path does not apply
// (k > 31: the stored key needs the value 4^k to be free and the slot needs mixed keys; more than 448 genomes).
size_t khb_hash_table_bytes(int k, int n_genomes, u64 n_sym, int *rec_words, int *log2_slots)
{
    if (k < 1 || k > 31 || n_genomes < 1 || n_genomes > 448 || n_sym == 0) return 0;
    const int R = n_genomes <= 64 ? 4 : n_genomes <= 192 ? 8 : 16;
    int L = 10;
    while ((1ull << L) < n_sym) L++;
    if (rec_words) *rec_words = R;
    if (log2_slots) *log2_slots = L;
    return ((size_t)1 << L) * (size_t)R * sizeof(u32);
}

template <int R>
static int hash_count_launch(khb_ctx *ctx, const u64 *d_codes, const u32 *d_valid, int k, const u64 *d_seg_off, int n_genomes, u32 n_pb,
                             int interleave, u32 *d_table, int L, u64 n_sym, u32 cs, u32 nbins, u64 *d_hist, void *d_out_keys, u64 *d_runs,
                             u64 *d_pairs, int pivot, void *d_out_pivot, u64 *d_pruns, u32 *d_overflow)
{
    const u64 n_slots = 1ull << L;
    const int shift = 2 * k > L ? 2 * k - L : 0;
    const u64 total = (u64)n_pb * (u64)n_genomes;
    if (total > 0x7fffffffull) return khb_fail(ctx, KHB_ERR_ARG, "hash group stage: %llu chunks exceed the grid limit", total);
    static long long probe_limit = -1;
    if (probe_limit < 0) {
        const char *e = getenv("KHB_HASH_MAX_PROBE");  // tests force the fall-back to the sort path with 0
        probe_limit = e ? atoll(e) : 8192;
    }
#ifdef KHB_EXPERIMENTS
    static int diag = -1;  // timing experiments only (WRONG results): 1 no bit update, 2 load only, 3 plain store instead of RED.  Not in the product build.
    if (diag < 0) {
        const char *e = getenv("KHB_HASH_DIAG");
        diag = e ? atoi(e) : 0;
    }
#else
    const int diag = 0;
#endif
    const u32 max_probe = n_slots > (u64)probe_limit ? (u32)probe_limit : (u32)n_slots;
    khb_prof_begin(ctx, KHB_K_HASH_INSERT);
    hash_insert_kernel<R><<<(unsigned)total, HS_BLOCK, 0, ctx->stream>>>(d_codes, d_valid, k, d_seg_off, (u32)n_genomes, n_pb, interleave, d_table,
                                                                        shift, n_slots - 1, max_probe, d_overflow, diag);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_HASH_INSERT, n_sym / 4 + n_sym / 8 + n_sym * 16);
    u64 grid = div_up(n_slots, (u64)HS_BLOCK * HS_PER);
    if (grid > (u64)ctx->num_sms * 16) grid = (u64)ctx->num_sms * 16;
    const size_t shm = ((size_t)nbins + 1) * sizeof(u32);
    khb_prof_begin(ctx, KHB_K_HASH_COUNT);
    if (pivot)
        hash_count_kernel<R, true><<<(unsigned)grid, HS_BLOCK, shm, ctx->stream>>>(d_table, n_slots, cs, nbins, (u32)(n_genomes - 1), d_hist, (u64 *)d_out_keys,
                                                                                   d_runs, d_pairs, (u64 *)d_out_pivot, d_pruns);
    else
        hash_count_kernel<R, false><<<(unsigned)grid, HS_BLOCK, shm, ctx->stream>>>(d_table, n_slots, cs, nbins, 0u, d_hist, (u64 *)d_out_keys, d_runs,
                                                                                    d_pairs, nullptr, nullptr);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_HASH_COUNT, n_slots * R * 4);
    return KHB_OK;
}

// K2..K5 of one group through the hash table.  Outputs as khb_pairs_count_impl; *d_overflow (u32, zeroed here) becomes
// non-zero when a probe sequence hit the limit: the outputs are then incomplete and the table is dirty.
// d_table: the context's table (api.cu: hash_table_get), all zero on entry, all zero again on exit unless overflowed.
int khb_hash_count_impl(khb_ctx *ctx, const u64 *d_codes, const u32 *d_valid, u64 n_sym, int k, const u64 *d_seg_off, const u64 *h_seg_off,
                        int n_genomes, u32 *d_table, int rec_words, int log2_slots, u32 cs, u32 nbins, u64 *d_hist, void *d_out_keys, u64 *d_runs,
                        u64 *d_pairs, int pivot, void *d_out_pivot, u64 *d_pruns, u32 *d_overflow)
{
    KHB_CUDA(ctx, cudaMemsetAsync(d_hist, 0, ((size_t)nbins + 1) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_runs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_pairs, 0, sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_overflow, 0, sizeof(u64), ctx->stream));
    if (pivot) KHB_CUDA(ctx, cudaMemsetAsync(d_pruns, 0, sizeof(u64), ctx->stream));
    u64 max_len = 0;
    for (int g = 0; g < n_genomes; g++) max_len = h_seg_off[g + 1] - h_seg_off[g] > max_len ? h_seg_off[g + 1] - h_seg_off[g] : max_len;
    if (n_sym == 0 || max_len == 0) return KHB_OK;
    const u32 n_pb = (u32)div_up(max_len, HS_CHUNK);
    static int interleave = -1;
    if (interleave < 0) {
        const char *e = getenv("KHB_HASH_ORDER");  // "genome": one genome after the other (no reuse across genomes in L2)
        interleave = (e && strcmp(e, "genome") == 0) ? 0 : 1;
    }
#define HS_ARGS ctx, d_codes, d_valid, k, d_seg_off, n_genomes, n_pb, interleave, d_table, log2_slots, n_sym, cs, nbins, d_hist, d_out_keys, d_runs, d_pairs, \
                pivot, d_out_pivot, d_pruns, d_overflow
    if (rec_words == 4) return hash_count_launch<4>(HS_ARGS);
    if (rec_words == 8) return hash_count_launch<8>(HS_ARGS);
    if (rec_words == 16) return hash_count_launch<16>(HS_ARGS);
#undef HS_ARGS
    return khb_fail(ctx, KHB_ERR_ARG, "hash group stage: record of %d words", rec_words);
}
