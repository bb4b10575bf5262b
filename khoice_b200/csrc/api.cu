// api.cu -- C ABI of libkhoice_b200.so (see include/khoice_b200.h): context, memory, thin wrappers
// around the kernels, K7 hash partition, and the fused per-group / across-group stages.
#include <stdarg.h>
#include <stdlib.h>

#include <vector>

#include "khb_common.cuh"

// implemented in the kernel translation units
int khb_pack_fasta_impl(khb_ctx *, const uint8_t *, size_t, u64 *, u32 *, size_t, u64 *, u64 *);
int khb_fasta_separators_impl(khb_ctx *, uint8_t *, const u64 *, const u64 *, int, u64, cudaStream_t);
struct khb_hostvec {
    std::vector<u64> v;                    // begin offsets of the prefetched group
    std::vector<const uint8_t *> d_files;  // deferred prefetch request (issued while another one is pending)
    std::vector<size_t> d_sizes;
    bool deferred = false;
};
int khb_extract_kmers_impl(khb_ctx *, const u64 *, const u32 *, size_t, int, int, void *, unsigned short *, const u64 *, int, u32 *, int, int);
int khb_sort_hist_buffer(khb_ctx *, int, u32 **);
int khb_remix_impl(khb_ctx *, void *, size_t, int, int);
int khb_sort_keys_impl(khb_ctx *, void *, void *, const u64 *, int, int, int *);
int khb_sort_bits_impl(khb_ctx *, void *, void *, const u64 *, int, int, int, int, int *, unsigned short *, unsigned short *, int hist_ready = 0);
int khb_sort_gathered_impl(khb_ctx *, void *, void *, const u64 *, const u64 *, int, int, int, int, int *);
int khb_resolve_unique_impl(khb_ctx *, const void *, size_t, int, int, void *, u64 *);
int khb_resolve_count_impl(khb_ctx *, const void *, size_t, int, int, u32, u32, u64 *, void *, u64 *);
int khb_pairs_count_impl(khb_ctx *, const void *, const unsigned short *, size_t, int, int, u32, u32, u32, u64 *, void *, u64 *, u64 *, int, void *, u64 *);
int khb_fill_segment_ids_impl(khb_ctx *, unsigned short *, const u64 *, int, u64);
size_t khb_presence_table_bytes(int, int);
int khb_presence_count_impl(khb_ctx *, const u64 *, const u32 *, size_t, int, int, const u64 *, int, u32 *, u32, u32, u64 *, void *, u64 *, u64 *, int,
                            void *, u64 *);
size_t khb_hash_table_bytes(int, int, u64, int *, int *);
int khb_bins_eligible(int, int, u64);
int khb_bins_count_impl(khb_ctx *, const u64 *, const u32 *, u64, int, const u64 *, int, u32, u32, u64 *, void *, u64 *, u64 *, u64 *, int, u32, u32 *, void *, u64 *, u64,
                        u64, khb_peer_route);
int khb_bins_across_impl(khb_ctx *, int, const void *, const void *, u64, u32, int, u32, u32, u64 *, u64 *, u64 *);
int khb_hash_count_impl(khb_ctx *, const u64 *, const u32 *, u64, int, const u64 *, const u64 *, int, u32 *, int, int, u32, u32, u64 *, void *, u64 *, u64 *, int,
                        void *, u64 *, u32 *);
int khb_peer_regions(khb_ctx *, const void **, u64 *, int *, int *);
extern "C" int khb_peer_close(khb_ctx *);
extern "C" int khb_team_close(khb_ctx *);
int khb_pivot_across_impl(khb_ctx *, const void *, const unsigned short *, size_t, int, int, u32, u32, u32, u64 *);
int khb_sorted_lookup_impl(khb_ctx *, const void *, u64, const void *, u64, int, u64 *);
int khb_membership_impl(khb_ctx *, const void *, const unsigned short *, size_t, int, int, u32, int, const void *, const u64 *, int, u64 *);
// Experiment type 2: the pivot k-mer sets P_1 .. P_G (device, concatenated) and, for every pivot group call, the range its
// rest-of-set union occupies in the group-set store.
struct khb_pivot_store {
    void *buf = nullptr;
    size_t cap = 0;            // bytes
    u64 len = 0;               // keys
    std::vector<u64> p_off;    // pivot set j = [p_off[j], p_off[j+1])
    std::vector<u64> u_off;    // union set j = gs_buf[u_off[j], u_off[j+1])
};
int khb_unique_impl(khb_ctx *, const void *, size_t, int, void *, u64 *);
int khb_count_runs_impl(khb_ctx *, const void *, size_t, int, u32, u32, u64 *, void *, u32 *, u64 *);

static char g_init_error[512] = "";

// ---- per-kernel timing ---------------------------------------------------------------------------
struct khb_prof_rec {
    int id;
    cudaEvent_t a, b;
    u64 bytes;
};
struct khb_prof {
    std::vector<khb_prof_rec> recs;   // used records (in launch order)
    std::vector<cudaEvent_t> pool;    // spare events
    cudaEvent_t open_a;
    bool open;
};

static cudaEvent_t prof_event(khb_prof *p)
{
    cudaEvent_t e = nullptr;
    if (!p->pool.empty()) {
        e = p->pool.back();
        p->pool.pop_back();
    } else {
        cudaEventCreate(&e);
    }
    return e;
}

void khb_prof_begin(khb_ctx *ctx, int id)
{
    (void)id;
    if (!ctx->prof_on || !ctx->prof) return;
    khb_prof *p = ctx->prof;
    p->open_a = prof_event(p);
    p->open = true;
    cudaEventRecord(p->open_a, ctx->prof_stream ? ctx->prof_stream : ctx->stream);
}

void khb_prof_end(khb_ctx *ctx, int id, u64 alg_bytes)
{
    if (!ctx->prof_on || !ctx->prof || !ctx->prof->open) return;
    khb_prof *p = ctx->prof;
    khb_prof_rec r;
    r.id = id;
    r.a = p->open_a;
    r.b = prof_event(p);
    r.bytes = alg_bytes;
    cudaEventRecord(r.b, ctx->prof_stream ? ctx->prof_stream : ctx->stream);
    p->recs.push_back(r);
    p->open = false;
}

void khb_prof_patch(khb_ctx *ctx, int id, u64 alg_bytes)
{
    if (!ctx->prof_on || !ctx->prof) return;
    for (auto it = ctx->prof->recs.rbegin(); it != ctx->prof->recs.rend(); ++it)
        if (it->id == id) {
            it->bytes = alg_bytes;
            return;
        }
}

int khb_fail(khb_ctx *ctx, int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx ? ctx->err : g_init_error, 512, fmt, ap);
    va_end(ap);
    return code;
}

int khb_cuda_fail(khb_ctx *ctx, cudaError_t e, const char *what, const char *file, int line)
{
    if (ctx && !ctx->sticky) ctx->sticky = (int)e;
    const char *base = strrchr(file, '/');
    return khb_fail(ctx, e == cudaErrorMemoryAllocation ? KHB_ERR_NOMEM : KHB_ERR_CUDA, "CUDA error %d (%s) at %s:%d: %s",
                    (int)e, cudaGetErrorString(e), base ? base + 1 : file, line, what);
}

int khb_scratch_get(khb_ctx *ctx, int slot, size_t bytes, void **out)
{
    khb_scratch *s = &ctx->scratch[slot];
    if (s->bytes < bytes) {
        if (s->ptr) {
            KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            KHB_CUDA(ctx, cudaFree(s->ptr));
            s->ptr = nullptr;
            s->bytes = 0;
        }
        size_t want = bytes + bytes / 8 + 4096;
        cudaError_t e = cudaMalloc(&s->ptr, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&s->ptr, want);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            s->ptr = nullptr;
            return khb_fail(ctx, KHB_ERR_NOMEM, "device allocation of %zu bytes failed (scratch slot %d)", bytes, slot);
        }
        s->bytes = want;
    }
    *out = s->ptr;
    return KHB_OK;
}

extern "C" {

int khb_abi_version(void) { return KHB_ABI_VERSION; }

const char *khb_last_error(const khb_ctx *ctx) { return ctx ? ctx->err : g_init_error; }

int khb_init(int device, khb_ctx **out)
{
    if (!out) return KHB_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return khb_fail(nullptr, KHB_ERR_NODEV, "no CUDA device (%s); libkhoice_b200 has no CPU fallback",
                        e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= ndev) return khb_fail(nullptr, KHB_ERR_ARG, "device %d out of range (0..%d)", device, ndev - 1);
    cudaDeviceProp prop;
    if ((e = cudaSetDevice(device)) != cudaSuccess || (e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return khb_fail(nullptr, KHB_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    if (prop.major != 10)
        return khb_fail(nullptr, KHB_ERR_NODEV, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                        prop.major, prop.minor);
    khb_ctx *ctx = (khb_ctx *)calloc(1, sizeof(khb_ctx));
    if (!ctx) return khb_fail(nullptr, KHB_ERR_NOMEM, "host allocation failed");
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    {
        const char *gm = getenv("KHB_GROUP_MODE");
        ctx->group_mode = !gm ? KHB_GROUP_AUTO : strcmp(gm, "two-sort") == 0 ? KHB_GROUP_TWO_SORT : strcmp(gm, "single-sort") == 0 ? KHB_GROUP_SINGLE_SORT
                          : strcmp(gm, "hash") == 0 ? KHB_GROUP_HASH : strcmp(gm, "bins") == 0 ? KHB_GROUP_BINS : KHB_GROUP_AUTO;
    }
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess ||
        (e = cudaMallocHost((void **)&ctx->h_mail, 1 << 20)) != cudaSuccess ||
        (e = cudaMalloc((void **)&ctx->d_mail, 1 << 20)) != cudaSuccess) {
        khb_fail(nullptr, KHB_ERR_CUDA, "context setup: %s", cudaGetErrorString(e));
        free(ctx);
        return KHB_ERR_CUDA;
    }
    *out = ctx;
    return KHB_OK;
}

int khb_destroy(khb_ctx *ctx)
{
    if (!ctx) return KHB_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (int i = 0; i < KHB_NSCRATCH; i++)
        if (ctx->scratch[i].ptr) cudaFree(ctx->scratch[i].ptr);
    if (ctx->gs_buf) cudaFree(ctx->gs_buf);
    if (ctx->hs_tab) cudaFree(ctx->hs_tab);
    khb_peer_close(ctx);
    khb_team_close(ctx);
    if (ctx->stage_dev) cudaFree(ctx->stage_dev);
    if (ctx->stage_next) cudaFree(ctx->stage_next);
    if (ctx->pf_tab) cudaFree(ctx->pf_tab);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->copy_done) cudaEventDestroy(ctx->copy_done);
    delete ctx->pf_begin;
    if (ctx->pv) {
        if (ctx->pv->buf) cudaFree(ctx->pv->buf);
        delete ctx->pv;
    }
    if (ctx->h_mail) cudaFreeHost(ctx->h_mail);
    if (ctx->d_mail) cudaFree(ctx->d_mail);
    cudaEventDestroy(ctx->ev0);
    cudaEventDestroy(ctx->ev1);
    cudaStreamDestroy(ctx->stream);
    free(ctx);
    return KHB_OK;
}

int khb_device_info(khb_ctx *ctx, int *num_sms, size_t *free_bytes, size_t *total_bytes)
{
    KHB_CHECK_CTX(ctx);
    size_t f = 0, t = 0;
    KHB_CUDA(ctx, cudaMemGetInfo(&f, &t));
    if (num_sms) *num_sms = ctx->num_sms;
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    return KHB_OK;
}

uint64_t khb_launch_count(const khb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int khb_profile_enable(khb_ctx *ctx, int on)
{
    KHB_CHECK_CTX(ctx);
    if (!ctx->prof) ctx->prof = new khb_prof();
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof->recs) {
        ctx->prof->pool.push_back(r.a);
        ctx->prof->pool.push_back(r.b);
    }
    ctx->prof->recs.clear();
    ctx->prof->open = false;
    ctx->prof_on = on ? 1 : 0;
    return KHB_OK;
}

int khb_profile_read(khb_ctx *ctx, int kernel_id, uint64_t *launches, double *ms, uint64_t *alg_bytes)
{
    KHB_CHECK_CTX(ctx);
    uint64_t n = 0, bytes = 0;
    double t = 0.0;
    if (ctx->prof) {
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (auto &r : ctx->prof->recs) {
            if (r.id != kernel_id) continue;
            float f = 0.f;
            KHB_CUDA(ctx, cudaEventElapsedTime(&f, r.a, r.b));
            t += f;
            n++;
            bytes += r.bytes;
        }
    }
    if (launches) *launches = n;
    if (ms) *ms = t;
    if (alg_bytes) *alg_bytes = bytes;
    return KHB_OK;
}
void *khb_stream(khb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int khb_alloc(khb_ctx *ctx, size_t bytes, void **d_ptr)
{
    KHB_CHECK_CTX(ctx);
    if (!d_ptr) return khb_fail(ctx, KHB_ERR_ARG, "khb_alloc: null out pointer");
    *d_ptr = nullptr;
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return khb_fail(ctx, KHB_ERR_NOMEM, "khb_alloc: device allocation of %zu bytes failed", bytes);
    }
    return KHB_OK;
}
int khb_free(khb_ctx *ctx, void *d_ptr)
{
    KHB_CHECK_CTX(ctx);
    if (!d_ptr) return KHB_OK;
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    KHB_CUDA(ctx, cudaFree(d_ptr));
    return KHB_OK;
}
int khb_alloc_host(khb_ctx *ctx, size_t bytes, void **h_ptr)
{
    KHB_CHECK_CTX(ctx);
    if (!h_ptr) return khb_fail(ctx, KHB_ERR_ARG, "khb_alloc_host: null out pointer");
    cudaError_t e = cudaMallocHost(h_ptr, bytes ? bytes : 16);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return khb_fail(ctx, KHB_ERR_NOMEM, "khb_alloc_host: pinned allocation of %zu bytes failed", bytes);
    }
    return KHB_OK;
}
int khb_free_host(khb_ctx *ctx, void *h_ptr)
{
    KHB_CHECK_CTX(ctx);
    if (h_ptr) KHB_CUDA(ctx, cudaFreeHost(h_ptr));
    return KHB_OK;
}
int khb_memcpy_h2d(khb_ctx *ctx, void *d, const void *h, size_t bytes)
{
    KHB_CHECK_CTX(ctx);
    if (bytes) KHB_CUDA(ctx, cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return KHB_OK;
}
int khb_memcpy_d2h(khb_ctx *ctx, void *h, const void *d, size_t bytes)
{
    KHB_CHECK_CTX(ctx);
    if (bytes) KHB_CUDA(ctx, cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return KHB_OK;
}
int khb_memcpy_d2d(khb_ctx *ctx, void *d_dst, const void *d_src, size_t bytes)
{
    KHB_CHECK_CTX(ctx);
    if (bytes) KHB_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return KHB_OK;
}
int khb_memset(khb_ctx *ctx, void *d, int value, size_t bytes)
{
    KHB_CHECK_CTX(ctx);
    if (bytes) KHB_CUDA(ctx, cudaMemsetAsync(d, value, bytes, ctx->stream));
    return KHB_OK;
}
int khb_sync(khb_ctx *ctx)
{
    KHB_CHECK_CTX(ctx);
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KHB_OK;
}

// ---- staging ---------------------------------------------------------------------------------------
static inline size_t staged_len(size_t n) { return (n + 3 + KHB_FASTA_TILE - 1) / KHB_FASTA_TILE * KHB_FASTA_TILE; }

size_t khb_staged_size(int n_files, const size_t *h_sizes)
{
    size_t t = 0;
    for (int i = 0; i < n_files; i++) t += staged_len(h_sizes[i]);
    return t;
}

int khb_stage_fasta(khb_ctx *ctx, int n_files, const uint8_t *const *h_files, const size_t *h_sizes, uint8_t *d_fasta,
                    size_t d_capacity, uint64_t *h_begin)
{
    KHB_CHECK_CTX(ctx);
    if (n_files < 0 || (n_files > 0 && (!h_files || !h_sizes)) || !h_begin) return khb_fail(ctx, KHB_ERR_ARG, "khb_stage_fasta: bad arguments");
    std::vector<u64> tab(2 * (size_t)n_files + 2);
    u64 off = 0;
    for (int i = 0; i < n_files; i++) {
        h_begin[i] = off;
        tab[i] = off;
        tab[n_files + i] = h_sizes[i];
        off += staged_len(h_sizes[i]);
    }
    h_begin[n_files] = off;
    if (off > d_capacity) return khb_fail(ctx, KHB_ERR_CAPACITY, "khb_stage_fasta: need %llu bytes, buffer has %zu", off, d_capacity);
    if (n_files == 0) return KHB_OK;
    for (int i = 0; i < n_files; i++)
        if (h_sizes[i]) KHB_CUDA(ctx, cudaMemcpyAsync(d_fasta + h_begin[i], h_files[i], h_sizes[i], cudaMemcpyHostToDevice, ctx->stream));
    void *p;
    int rc = khb_scratch_get(ctx, SCR_MISC, tab.size() * sizeof(u64), &p);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(p, tab.data(), 2 * (size_t)n_files * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // tab is pageable and dies with this frame
    return khb_fasta_separators_impl(ctx, d_fasta, (const u64 *)p, (const u64 *)p + n_files, n_files, off, ctx->stream);
}

// ---- thin kernel wrappers ------------------------------------------------------------------------
int khb_pack_fasta(khb_ctx *ctx, const uint8_t *d_fasta, size_t nbytes, uint64_t *d_codes, uint32_t *d_valid,
                   size_t cap_symbols, uint64_t *d_tile_base, uint64_t *d_counts)
{
    KHB_CHECK_CTX(ctx);
    return khb_pack_fasta_impl(ctx, d_fasta, nbytes, (u64 *)d_codes, d_valid, cap_symbols, (u64 *)d_tile_base, (u64 *)d_counts);
}

int khb_extract_kmers(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, size_t n_symbols, int k, void *d_keys)
{
    KHB_CHECK_CTX(ctx);
    return khb_extract_kmers_impl(ctx, (const u64 *)d_codes, d_valid, n_symbols, k, 0, d_keys, nullptr, nullptr, 0, nullptr, 0, 0);
}

int khb_extract_kmers_hashed(khb_ctx *ctx, const uint64_t *d_codes, const uint32_t *d_valid, size_t n_symbols, int k, void *d_keys)
{
    KHB_CHECK_CTX(ctx);
    return khb_extract_kmers_impl(ctx, (const u64 *)d_codes, d_valid, n_symbols, k, 1, d_keys, nullptr, nullptr, 0, nullptr, 0, 0);
}

int khb_remix_keys(khb_ctx *ctx, void *d_keys, size_t n, int k, int inverse)
{
    KHB_CHECK_CTX(ctx);
    return khb_remix_impl(ctx, d_keys, n, k, inverse);
}

int khb_prefix_plan(int k, uint64_t n_max, int *first_bit, int *npass)
{
    if (k < 1 || k > 64 || !first_bit || !npass) return KHB_ERR_ARG;
    const int full = (2 * k + 7) / 8;
    *first_bit = 0;
    *npass = full;
    if (k == 32 || k == 64) return KHB_OK;  // no spare bit above the key: full sort
    static int slack = -1;
    if (slack < 0) {
        const char *e = getenv("KHB_PREFIX_SLACK");
        slack = e ? atoi(e) : 0;
    }
    int lg = 1;
    while (lg < 63 && (1ull << lg) < n_max) lg++;
    const int np = (lg + slack + 1 + 7) / 8;  // prefix key bits + the spare bit that separates sentinels
    if (np < 1 || 8 * np >= 2 * k + 1) return KHB_OK;
    *first_bit = 2 * k + 1 - 8 * np;
    *npass = np;
    return KHB_OK;
}

int khb_sort_key_bits(khb_ctx *ctx, void *d_keys, void *d_tmp, const uint64_t *h_seg_off, int n_segments, int key_bytes,
                      int first_bit, int npass, int *result_in_tmp)
{
    KHB_CHECK_CTX(ctx);
    int dummy;
    return khb_sort_bits_impl(ctx, d_keys, d_tmp, (const u64 *)h_seg_off, n_segments, key_bytes, first_bit, npass,
                              result_in_tmp ? result_in_tmp : &dummy, nullptr, nullptr);
}

int khb_resolve_unique(khb_ctx *ctx, const void *d_sorted, size_t n, int k, int prefix_shift, void *d_out, uint64_t *h_count)
{
    KHB_CHECK_CTX(ctx);
    int rc = khb_resolve_unique_impl(ctx, d_sorted, n, k, prefix_shift, d_out, ctx->d_mail);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_count) *h_count = ctx->h_mail[0];
    return KHB_OK;
}

int khb_resolve_count(khb_ctx *ctx, const void *d_sorted, size_t n, int k, int prefix_shift, uint32_t cs, uint32_t nbins,
                      uint64_t *h_hist, void *d_out_keys, uint64_t *h_runs)
{
    KHB_CHECK_CTX(ctx);
    if (!h_hist) return khb_fail(ctx, KHB_ERR_ARG, "khb_resolve_count: null histogram");
    u64 *d_hist = ctx->d_mail + 8, *d_runs = ctx->d_mail;
    int rc = khb_resolve_count_impl(ctx, d_sorted, n, k, prefix_shift, cs, nbins, d_hist, d_out_keys, d_runs);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, (nbins + 9) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, ctx->h_mail + 8, ((size_t)nbins + 1) * sizeof(u64));
    if (h_runs) *h_runs = ctx->h_mail[0];
    return KHB_OK;
}

int khb_sort_keys(khb_ctx *ctx, void *d_keys, void *d_tmp, const uint64_t *h_seg_off, int n_segments, int k, int *result_in_tmp)
{
    KHB_CHECK_CTX(ctx);
    int dummy;
    return khb_sort_keys_impl(ctx, d_keys, d_tmp, (const u64 *)h_seg_off, n_segments, k, result_in_tmp ? result_in_tmp : &dummy);
}

int khb_unique(khb_ctx *ctx, const void *d_sorted, size_t n, int k, void *d_out, uint64_t *h_count)
{
    KHB_CHECK_CTX(ctx);
    int rc = khb_unique_impl(ctx, d_sorted, n, k, d_out, ctx->d_mail);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h_count) *h_count = ctx->h_mail[0];
    return KHB_OK;
}

int khb_count_runs(khb_ctx *ctx, const void *d_sorted, size_t n, int k, uint32_t cs, uint32_t nbins, uint64_t *h_hist,
                   void *d_out_keys, uint32_t *d_out_counts, uint64_t *h_runs)
{
    KHB_CHECK_CTX(ctx);
    if (!h_hist) return khb_fail(ctx, KHB_ERR_ARG, "khb_count_runs: null histogram");
    u64 *d_hist = ctx->d_mail + 8, *d_runs = ctx->d_mail;
    int rc = khb_count_runs_impl(ctx, d_sorted, n, k, cs, nbins, d_hist, d_out_keys, d_out_counts, d_runs);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, (nbins + 9) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, ctx->h_mail + 8, ((size_t)nbins + 1) * sizeof(u64));
    if (h_runs) *h_runs = ctx->h_mail[0];
    return KHB_OK;
}

}  // extern "C"

// ---- K7: hash partition ------------------------------------------------------------------------------
// (mix64 / part_of: khb_common.cuh, shared with peer.cu)

#define PT_BLOCK 256
#define PT_ITEMS 16
#define PT_MAXPARTS 64

// counts[p] += number of keys of bucket p  (mode 0)   or   scatter (mode 1: cursor[p] holds the next free slot)
template <typename Key, int MODE>
__global__ void __launch_bounds__(PT_BLOCK)
partition_kernel(const Key *__restrict__ in, u64 n, u32 nparts, u64 *__restrict__ counts_or_cursor, Key *__restrict__ out)
{
    __shared__ u32 sc[PT_MAXPARTS];
    __shared__ u64 sbase[PT_MAXPARTS];
    const u32 tid = threadIdx.x;
    const u64 begin = (u64)blockIdx.x * (PT_BLOCK * PT_ITEMS);
    if (tid < PT_MAXPARTS) sc[tid] = 0;
    __syncthreads();
    Key keys[PT_ITEMS];
    u32 part[PT_ITEMS], slot[PT_ITEMS];
#pragma unroll
    for (int r = 0; r < PT_ITEMS; r++) {
        const u64 g = begin + (u64)r * PT_BLOCK + tid;
        part[r] = 0xffffffffu;
        if (g < n) {
            keys[r] = in[g];
            part[r] = part_of(keys[r], nparts);
            slot[r] = atomicAdd(&sc[part[r]], 1u);
        }
    }
    __syncthreads();
    if (tid < nparts) {
        const u32 c = sc[tid];
        if (MODE == 0) {
            if (c) atomicAdd(&counts_or_cursor[tid], (u64)c);
        } else {
            sbase[tid] = c ? atomicAdd(&counts_or_cursor[tid], (u64)c) : 0ull;
        }
    }
    if (MODE == 1) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < PT_ITEMS; r++)
            if (part[r] != 0xffffffffu) out[sbase[part[r]] + slot[r]] = keys[r];
    }
}

template <typename Key>
static int partition_impl(khb_ctx *ctx, const Key *d_keys, u64 n, int nparts, Key *d_out, u64 *h_off)
{
    u64 *d_cnt = ctx->d_mail + 16384;  // 64 counters + 64 cursors, away from the histogram mailbox
    KHB_CUDA(ctx, cudaMemsetAsync(d_cnt, 0, 2 * PT_MAXPARTS * sizeof(u64), ctx->stream));
    const u64 blocks = div_up(n, PT_BLOCK * PT_ITEMS);
    if (n) {
        partition_kernel<Key, 0><<<(unsigned)blocks, PT_BLOCK, 0, ctx->stream>>>(d_keys, n, (u32)nparts, d_cnt, nullptr);
        KHB_LAUNCH_CHECK(ctx);
    }
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 16384, d_cnt, PT_MAXPARTS * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    u64 cursor[PT_MAXPARTS];
    u64 off = 0;
    for (int p = 0; p < nparts; p++) {
        h_off[p] = off;
        cursor[p] = off;
        off += ctx->h_mail[16384 + p];
    }
    h_off[nparts] = off;
    if (n) {
        KHB_CUDA(ctx, cudaMemcpyAsync(d_cnt + PT_MAXPARTS, cursor, nparts * sizeof(u64), cudaMemcpyHostToDevice, ctx->stream));
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        partition_kernel<Key, 1><<<(unsigned)blocks, PT_BLOCK, 0, ctx->stream>>>(d_keys, n, (u32)nparts, d_cnt + PT_MAXPARTS, d_out);
        KHB_LAUNCH_CHECK(ctx);
    }
    return KHB_OK;
}

// ---- fused stages ----------------------------------------------------------------------------------------
static int gs_reserve(khb_ctx *ctx, int k, u64 extra)
{
    const size_t W = (size_t)khb_key_bytes(k);
    if (ctx->gs_k && ctx->gs_k != k) return khb_fail(ctx, KHB_ERR_STATE, "retained group sets were built with k=%d, not k=%d; call khb_group_sets_reset", ctx->gs_k, k);
    const u64 need = (ctx->gs_len + extra + 2) * W;  // gs_cap is in BYTES (W changes with k)
    if (need > ctx->gs_cap) {
        int prc = khb_peer_wait(ctx);  // a push may still be reading the store that is about to move
        if (prc) return prc;
        u64 cap = ctx->gs_cap ? ctx->gs_cap : (8u << 20);
        while (cap < need) cap += cap / 2 + 16;
        void *nb = nullptr;
        cudaError_t e = cudaMalloc(&nb, cap);
        if (e != cudaSuccess) {
            cudaGetLastError();
            cap = need;
            e = cudaMalloc(&nb, cap);
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            return khb_fail(ctx, KHB_ERR_NOMEM, "group set store: device allocation of %llu bytes failed", cap);
        }
        if (ctx->gs_len) KHB_CUDA(ctx, cudaMemcpyAsync(nb, ctx->gs_buf, ctx->gs_len * W, cudaMemcpyDeviceToDevice, ctx->stream));
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->gs_buf) KHB_CUDA(ctx, cudaFree(ctx->gs_buf));
        ctx->gs_buf = nb;
        ctx->gs_cap = cap;
    }
    ctx->gs_k = k;
    return KHB_OK;
}

static int pv_reserve(khb_ctx *ctx, int k, u64 extra)
{
    if (!ctx->pv) ctx->pv = new khb_pivot_store();
    khb_pivot_store *pv = ctx->pv;
    const size_t W = (size_t)khb_key_bytes(k);
    const u64 need = (pv->len + extra + 2) * W;
    if (need > pv->cap) {
        u64 cap = pv->cap ? pv->cap : (8u << 20);
        while (cap < need) cap += cap / 2 + 16;
        void *nb = nullptr;
        if (cudaMalloc(&nb, cap) != cudaSuccess) {
            cudaGetLastError();
            return khb_fail(ctx, KHB_ERR_NOMEM, "pivot set store: device allocation of %llu bytes failed", cap);
        }
        if (pv->len) KHB_CUDA(ctx, cudaMemcpyAsync(nb, pv->buf, pv->len * W, cudaMemcpyDeviceToDevice, ctx->stream));
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (pv->buf) KHB_CUDA(ctx, cudaFree(pv->buf));
        pv->buf = nb;
        pv->cap = cap;
    }
    return KHB_OK;
}

// Room for `extra` more segment events (bins.cu: mb_event, 16 bytes each) behind the ev_len logged so far.
static int ev_reserve(khb_ctx *ctx, u64 extra)
{
    if (!ctx->ev_count) {
        KHB_CUDA(ctx, cudaMalloc((void **)&ctx->ev_count, 64));
        KHB_CUDA(ctx, cudaMemsetAsync(ctx->ev_count, 0, 64, ctx->stream));
    }
    const u64 need = ctx->ev_len + extra;
    if (need <= ctx->ev_cap) return KHB_OK;
    u64 cap = ctx->ev_cap ? ctx->ev_cap : (1u << 16);
    while (cap < need) cap += cap / 2;
    void *nb = nullptr;
    if (cudaMalloc(&nb, cap * 16) != cudaSuccess) {
        cudaGetLastError();
        return khb_fail(ctx, KHB_ERR_NOMEM, "segment events: device allocation of %llu bytes failed", (unsigned long long)(cap * 16));
    }
    if (ctx->ev_len) KHB_CUDA(ctx, cudaMemcpyAsync(nb, ctx->ev_buf, ctx->ev_len * 16, cudaMemcpyDeviceToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->ev_buf) KHB_CUDA(ctx, cudaFree(ctx->ev_buf));
    ctx->ev_buf = nb;
    ctx->ev_cap = cap;
    return KHB_OK;
}
// KHB_ACROSS_MODE=bins: the across-group stage counts bin by bin where the store allows it (measured slower than the sort: opt-in)
static bool across_by_bins_wanted()
{
    static int v = -1;
    if (v < 0) {
        const char *e = getenv("KHB_ACROSS_MODE");
        v = e && strcmp(e, "bins") == 0 ? 1 : 0;
    }
    return v == 1;
}
static void ev_forget(khb_ctx *ctx)
{
    ctx->ev_ok = 0;
    ctx->ev_len = 0;
    ctx->ev_nb = 0;
    if (ctx->ev_count) cudaMemsetAsync(ctx->ev_count, 0, 8, ctx->stream);
}

struct PhaseTimer {
    khb_ctx *ctx;
    cudaEvent_t ev[12];
    int n;
    bool ok;
    explicit PhaseTimer(khb_ctx *c) : ctx(c), n(0), ok(true)
    {
        for (int i = 0; i < 12; i++)
            if (cudaEventCreate(&ev[i]) != cudaSuccess) ok = false;
    }
    ~PhaseTimer()
    {
        for (int i = 0; i < 12; i++) cudaEventDestroy(ev[i]);
    }
    void mark()
    {
        if (n < 12) cudaEventRecord(ev[n++], ctx->stream);
    }
    float ms(int a, int b)
    {
        float t = 0.f;
        if (a < n && b < n) cudaEventElapsedTime(&t, ev[a], ev[b]);
        return t;
    }
};

// A group after K1: the packed symbol stream plus the symbol offset of every genome.  Either a transient view of the
// context's scratch (fused calls) or a persistent allocation (khb_pack_group: pack once, sweep k).
struct khb_packed {
    u64 *d_codes;
    u32 *d_valid;
    u64 n_sym, n_breaks, fasta_bytes;
    int n_genomes;
    std::vector<u64> seg;  // n_genomes + 1 symbol offsets
    bool owned;
};

// K1 into the context's scratch.
static int pack_stage(khb_ctx *ctx, int n_genomes, const uint8_t *d_fasta, const u64 *h_begin, khb_packed &pk, PhaseTimer &tm)
{
    if (n_genomes < 1 || !h_begin) return khb_fail(ctx, KHB_ERR_ARG, "group stage: bad arguments");
    const size_t nbytes = h_begin[n_genomes];
    const size_t ntiles = nbytes / KHB_FASTA_TILE;
    int rc;
    void *p;
    const size_t cw = khb_codes_words(nbytes), vw = khb_valid_words(nbytes);
    const size_t pack_bytes = cw * 8 + vw * 4 + (ntiles + 1) * 8 + 64;
    if ((rc = khb_scratch_get(ctx, SCR_PACK, pack_bytes, &p))) return rc;
    u64 *d_codes = (u64 *)p;
    u64 *d_tile_base = d_codes + cw;
    u64 *d_counts = d_tile_base + ntiles + 1;
    u32 *d_valid = (u32 *)(d_counts + 4);
    if ((rc = khb_pack_fasta_impl(ctx, d_fasta, nbytes, d_codes, d_valid, nbytes, d_tile_base, d_counts))) return rc;
    std::vector<u64> tile_base(ntiles + 1);
    u64 counts[2];
    KHB_CUDA(ctx, cudaMemcpyAsync(tile_base.data(), d_tile_base, (ntiles + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, 16, cudaMemcpyDeviceToHost, ctx->stream));
    tm.mark();  // 2: pack done
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    pk.d_codes = d_codes;
    pk.d_valid = d_valid;
    pk.n_sym = counts[0];
    pk.n_breaks = counts[1];
    pk.fasta_bytes = nbytes;
    pk.n_genomes = n_genomes;
    pk.owned = false;
    pk.seg.resize((size_t)n_genomes + 1);
    for (int g = 0; g < n_genomes; g++) pk.seg[g] = tile_base[h_begin[g] / KHB_FASTA_TILE];
    pk.seg[0] = 0;
    pk.seg[n_genomes] = pk.n_sym;
    return KHB_OK;
}

// K2 .. K5 for one k on a packed group.
static int pv_reserve(khb_ctx *ctx, int k, u64 extra);

// The context's hash table for the sort-free group stage: at least `bytes`, all zero.
static int hash_table_get(khb_ctx *ctx, size_t bytes, u32 **out)
{
    if (ctx->hs_bytes < bytes) {
        if (ctx->hs_tab) {
            KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            KHB_CUDA(ctx, cudaFree(ctx->hs_tab));
            ctx->hs_tab = nullptr;
            ctx->hs_bytes = 0;
        }
        cudaError_t e = cudaMalloc((void **)&ctx->hs_tab, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            ctx->hs_tab = nullptr;
            return khb_fail(ctx, KHB_ERR_NOMEM, "device allocation of %zu bytes failed (group hash table)", bytes);
        }
        ctx->hs_bytes = bytes;
        ctx->hs_dirty = 1;
    }
    if (ctx->hs_dirty) {
        KHB_CUDA(ctx, cudaMemsetAsync(ctx->hs_tab, 0, ctx->hs_bytes, ctx->stream));
        ctx->hs_dirty = 0;
    }
    *out = ctx->hs_tab;
    return KHB_OK;
}

// pivot != 0: experiment type 2 -- the last genome of the group is the pivot (khb_pivot_group_from_packed).
static int count_stage(khb_ctx *ctx, int k, const khb_packed &pk, u32 nbins, u64 *h_hist, int keep_set, khb_stats *stats, PhaseTimer &tm,
                       int pivot = 0)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "k=%d outside 1..64", k);
    if (!h_hist) return khb_fail(ctx, KHB_ERR_ARG, "group stage: null histogram");
    if (nbins < 1 || nbins > 8192) return khb_fail(ctx, KHB_ERR_ARG, "nbins=%u outside 1..8192", nbins);
    const size_t W = (size_t)khb_key_bytes(k);
    const int n_genomes = pk.n_genomes;
    const u64 n_sym = pk.n_sym;
    const size_t nbytes = pk.fasta_bytes;
    const u64 counts[2] = {pk.n_sym, pk.n_breaks};
    u64 *d_codes = pk.d_codes;
    u32 *d_valid = pk.d_valid;
    int rc;
    void *p;
    // K2 (hashed unless k = 32 / 64, see khb_prefix_plan)
    const int hashed = (k != 32 && k != 64) ? 1 : 0;
    if (ctx->gs_k && keep_set && ctx->gs_hashed != hashed) return khb_fail(ctx, KHB_ERR_STATE, "retained group sets use a different key encoding");
    const size_t key_bytes = (n_sym + 4) * W;
    void *bufA = nullptr, *bufB = nullptr;  // sort buffers, taken only by the branches that sort
    const int single_sort = ctx->group_mode != KHB_GROUP_TWO_SORT;
    if (pivot && !(single_sort && n_genomes <= 65535)) return khb_fail(ctx, KHB_ERR_STATE, "pivot analysis needs the single-sort group path");
    if (single_sort && n_genomes <= 65535) {
        // ---- single-sort path: ONE prefix sort of all windows of the group with the genome id as payload ----
        const std::vector<u64> &seg = pk.seg;
        u64 *d_seg = ctx->d_mail + 32768;  // up to 65536 offsets fit the 1 MiB mailbox behind the histogram area
        if ((size_t)(n_genomes + 1) > 65536) return khb_fail(ctx, KHB_ERR_ARG, "too many genomes in one group");
        KHB_CUDA(ctx, cudaMemcpyAsync(d_seg, seg.data(), (size_t)(n_genomes + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // seg is pageable
        // small k: the canonical k-mer space is smaller than the data -> direct-address presence table (presence.cu)
        static long long smallk_budget = -1;
        if (smallk_budget < 0) {
            const char *e = getenv("KHB_SMALLK_TABLE_MB");  // 0 disables the path
            smallk_budget = (e ? atoll(e) : 1024) << 20;
        }
        const size_t table_bytes = khb_presence_table_bytes(k, n_genomes);
        // ... and when a table word is hit at least four times on average: below that most updates are first-time atomics on
        // random DRAM sectors and sorting is as fast (measured: k = 13 with 50 genomes, 1.9 hits per word, 106 ms either way)
        const bool small_k = table_bytes > 0 && (long long)table_bytes <= smallk_budget && table_bytes <= (size_t)n_sym;
        int fb = 0, np = 0;
        void *out_keys = nullptr;
        const u64 set_bound = small_k ? ((1ull << (2 * k)) < n_sym ? (1ull << (2 * k)) : n_sym) : n_sym;  // distinct keys of the group at most
        if (keep_set) {
            if ((rc = gs_reserve(ctx, k, set_bound))) return rc;
            ctx->gs_hashed = hashed;
            out_keys = (char *)ctx->gs_buf + ctx->gs_len * W;
        }
        u64 *d_hist = ctx->d_mail + 8, *d_runs = ctx->d_mail, *d_pairs = ctx->d_mail + 1, *d_pruns = ctx->d_mail + 2;
        void *out_pivot = nullptr;
        if (pivot && keep_set) {
            // upper bound for the pivot's distinct keys: its windows
            if ((rc = pv_reserve(ctx, k, seg[n_genomes] - seg[n_genomes - 1]))) return rc;
            out_pivot = (char *)ctx->pv->buf + ctx->pv->len * W;
        }
        // group stage without a sort (hashset.cu): K2 fused with an open-addressing table, one record per distinct k-mer
        int hs_R = 0, hs_L = 0;
        const size_t hs_bytes = (!small_k && hashed && ctx->group_mode == KHB_GROUP_HASH) ? khb_hash_table_bytes(k, n_genomes, n_sym, &hs_R, &hs_L) : 0;
        static long long hash_budget = -1;
        if (hash_budget < 0) {
            const char *e = getenv("KHB_HASH_TABLE_MB");  // 0 disables the path
            hash_budget = (e ? atoll(e) : 98304) << 20;
        }
        bool use_hash = hs_bytes > 0 && (long long)hs_bytes <= hash_budget;
        // minimizer bins + per-bin shared-memory counting (bins.cu): the default wherever it applies
        bool use_bins = !small_k && !use_hash && hashed && !pivot && (ctx->group_mode == KHB_GROUP_AUTO || ctx->group_mode == KHB_GROUP_BINS) &&
                        khb_bins_eligible(k, n_genomes, n_sym);
        int bins_exact = 0;
        khb_peer_route bins_route = {0u, 0u, 0ull, nullptr, nullptr};
        // the store's segment events (across-group stage bin by bin): kept while every retained group comes through the bins
        const bool log_events = use_bins && keep_set && across_by_bins_wanted() && (ctx->gs_len == 0 || ctx->ev_ok);
        u32 bins_nb = 0;
        if (keep_set && ctx->gs_len == 0) {
            ev_forget(ctx);
            ctx->ev_ok = log_events ? 1 : 0;
        }
        if (keep_set && !log_events) ctx->ev_ok = 0;
      again:
        if (use_bins) {
            tm.mark();  // 3
            tm.mark();  // 4
            tm.mark();  // 5
            tm.mark();  // 6
            const bool ev = log_events && ctx->ev_ok;
            if (ev) {
                if ((rc = ev_reserve(ctx, (u64)(ctx->ev_nb ? ctx->ev_nb : n_sym / 1024 + 16) * 8 + 4096))) return rc;
                if (bins_exact) KHB_CUDA(ctx, cudaMemcpyAsync(ctx->ev_count, &ctx->ev_len, 8, cudaMemcpyHostToDevice, ctx->stream));   // forget the failed attempt's events
            }
            // multi-GPU: the end-of-bin passes store every distinct key straight into its owner's region (peer.cu), no separate push pass
            if (keep_set && (rc = khb_peer_route_get(ctx, (int)W, &bins_route))) return rc;
            if ((rc = khb_bins_count_impl(ctx, d_codes, d_valid, n_sym, k, d_seg, n_genomes, KHB_COUNTER_MAX, nbins, d_hist, out_keys, d_runs, d_pairs,
                                          ctx->d_mail + 3, bins_exact, ev ? ctx->ev_nb : 0u, &bins_nb, ev ? ctx->ev_buf : nullptr, ctx->ev_count, ctx->ev_cap,
                                          ctx->gs_len, bins_route))) return rc;
            if (ev) KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 130000, ctx->ev_count, 8, cudaMemcpyDeviceToHost, ctx->stream));
        } else if (use_hash) {
            u32 *tab = nullptr;
            if ((rc = hash_table_get(ctx, hs_bytes, &tab))) return rc;
            tm.mark();  // 3
            tm.mark();  // 4
            tm.mark();  // 5
            tm.mark();  // 6
            if ((rc = khb_hash_count_impl(ctx, d_codes, d_valid, n_sym, k, d_seg, seg.data(), n_genomes, tab, hs_R, hs_L, KHB_COUNTER_MAX, nbins, d_hist,
                                          out_keys, d_runs, d_pairs, pivot, out_pivot, d_pruns, (u32 *)(ctx->d_mail + 3)))) return rc;
        } else if (small_k) {
            if ((rc = khb_scratch_get(ctx, SCR_AUX, table_bytes + 64, &p))) return rc;
            tm.mark();  // 3
            tm.mark();  // 4
            tm.mark();  // 5
            tm.mark();  // 6
            if ((rc = khb_presence_count_impl(ctx, d_codes, d_valid, n_sym, k, hashed, d_seg, n_genomes, (u32 *)p, KHB_COUNTER_MAX, nbins, d_hist, out_keys,
                                              d_runs, d_pairs, pivot, out_pivot, d_pruns))) return rc;
        } else {
            if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, key_bytes, &bufA))) return rc;
            if ((rc = khb_scratch_get(ctx, SCR_KEYS_B, key_bytes, &bufB))) return rc;
            if ((rc = khb_scratch_get(ctx, SCR_PAY_A, (n_sym + 8) * 2, &p))) return rc;
            unsigned short *payA = (unsigned short *)p;
            if ((rc = khb_scratch_get(ctx, SCR_PAY_B, (n_sym + 8) * 2, &p))) return rc;
            unsigned short *payB = (unsigned short *)p;
            // K2 also counts the digits of the prefix passes while the keys are in registers (the sort skips its histogram sweep)
            khb_prefix_plan(k, n_sym, &fb, &np);
            static int fuse_hist = -1;
            if (fuse_hist < 0) {
                const char *e = getenv("KHB_FUSE_HIST");
                fuse_hist = e ? atoi(e) : 1;
            }
            u32 *d_dig = nullptr;
            if (fuse_hist && (rc = khb_sort_hist_buffer(ctx, np, &d_dig))) return rc;
            if ((rc = khb_extract_kmers_impl(ctx, d_codes, d_valid, n_sym, k, hashed, bufA, payA, d_seg, n_genomes, d_dig, np, fb))) return rc;
            tm.mark();  // 3: extract done
            u64 one_seg[2] = {0, n_sym};
            int in_tmp = 0;
            if ((rc = khb_sort_bits_impl(ctx, bufA, bufB, one_seg, 1, (int)W, fb, np, &in_tmp, payA, payB, d_dig ? 1 : 0))) return rc;
            void *sorted = in_tmp ? bufB : bufA;
            unsigned short *spay = in_tmp ? payB : payA;
            tm.mark();  // 4: sort done
            tm.mark();  // 5
            tm.mark();  // 6
            if ((rc = khb_pairs_count_impl(ctx, sorted, spay, n_sym, k, fb, KHB_COUNTER_MAX, nbins, (u32)n_genomes, d_hist, out_keys, d_runs, d_pairs,
                                           pivot, out_pivot, d_pruns))) return rc;
        }
        KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, (nbins + 9) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        tm.mark();  // 7: count done
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (use_bins) {
            if ((ctx->h_mail[3] & 3) == 1 && !bins_exact) {
                // a region outgrew its share of the record buffer (uneven minimizers): partition again with the sizes just counted
                ctx->bins_repartitions++;
                bins_exact = 1;
                tm.n = 3;
                goto again;
            }
            if (ctx->h_mail[3] & 3) {
                // a bin could not be counted in shared memory at all: redo this group by sorting
                ctx->bins_fallbacks++;
                use_bins = false;
                if (keep_set) ctx->ev_ok = 0;
                if (bins_route.world && (rc = khb_peer_poison(ctx))) return rc;   // some of this group's keys are at their owners already: redo the round over NCCL
                tm.n = 3;
                goto again;
            }
            if (log_events && ctx->ev_ok) {
                if (ctx->h_mail[3] & 4) {
                    ctx->ev_ok = 0;                      // more segments than reserved: the across stage sorts
                } else {
                    ctx->ev_len = ctx->h_mail[130000];
                    ctx->ev_nb = bins_nb;
                }
            }
            ctx->bins_bigbins += ctx->h_mail[5];
            ctx->bins_hint_k = k;
            ctx->bins_hint_genomes = n_genomes;
            ctx->bins_hint_regions = ctx->bins_last_regions;
            ctx->bins_hint_max = ctx->h_mail[6];
            if (getenv("KHB_BINS_VERBOSE"))
                fprintf(stderr, "[bins] records=%llu distinct_records=%llu fullest_region=%llu distinct_kmers=%llu bigbins=%llu\n", (unsigned long long)ctx->h_mail[4],
                        (unsigned long long)ctx->h_mail[7], (unsigned long long)ctx->h_mail[6], (unsigned long long)ctx->h_mail[0], (unsigned long long)ctx->h_mail[5]);
            if (n_sym) ctx->bins_rho = (double)ctx->h_mail[0] / (double)n_sym;
            khb_prof_patch(ctx, KHB_K_BIN_PARTITION, n_sym * 3 / 8 + ctx->h_mail[4] * (k <= 32 ? 24 : 32));
            khb_prof_patch(ctx, KHB_K_BIN_COUNT, ctx->h_mail[4] * (k <= 32 ? 24 : 32) + ctx->h_mail[0] * W);
        }
        if (use_hash && ctx->h_mail[3]) {
            // a probe sequence hit the limit (table nearly full of distinct k-mers): redo this group by sorting
            ctx->hs_dirty = 1;
            ctx->hs_overflows++;
            use_hash = false;
            tm.n = 3;
            goto again;
        }
        memcpy(h_hist, ctx->h_mail + 8, ((size_t)nbins + 1) * sizeof(u64));
        const u64 d_g = ctx->h_mail[0];
        if (keep_set) {
            if (pivot) {
                if (ctx->pv->u_off.empty()) ctx->pv->u_off.push_back(ctx->gs_len);
                if (ctx->pv->u_off.back() != ctx->gs_len) return khb_fail(ctx, KHB_ERR_STATE, "pivot groups and plain groups were mixed in the group-set store");
                ctx->pv->u_off.push_back(ctx->gs_len + d_g);
                if (ctx->pv->p_off.empty()) ctx->pv->p_off.push_back(0);
                ctx->pv->len += ctx->h_mail[2];
                ctx->pv->p_off.push_back(ctx->pv->len);
            }
            ctx->gs_len += d_g;
            ctx->gs_groups += 1;
            if (use_bins && bins_route.world) khb_peer_mark_pushed(ctx);
        }
        if (stats) {
            stats->fasta_bytes = nbytes;
            stats->bases = n_sym - counts[1];
            stats->windows = n_sym;
            stats->genome_distinct = ctx->h_mail[1];
            stats->distinct = d_g;
            stats->passes_genome = 0;
            stats->passes_group = use_bins ? 0 : np;
        }
        return KHB_OK;
    }
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, key_bytes, &bufA))) return rc;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_B, key_bytes, &bufB))) return rc;
    if ((rc = khb_extract_kmers_impl(ctx, d_codes, d_valid, n_sym, k, hashed, bufA, nullptr, nullptr, 0, nullptr, 0, 0))) return rc;
    tm.mark();  // 3: extract done
    // K3 per genome (segmented), prefix only
    const std::vector<u64> &seg = pk.seg;
    u64 max_seg = 1;
    for (int g = 0; g < n_genomes; g++) max_seg = seg[g + 1] - seg[g] > max_seg ? seg[g + 1] - seg[g] : max_seg;
    int fb1, np1, fb2, np2;
    khb_prefix_plan(k, max_seg, &fb1, &np1);
    int in_tmp = 0;
    if ((rc = khb_sort_bits_impl(ctx, bufA, bufB, seg.data(), n_genomes, (int)W, fb1, np1, &in_tmp, nullptr, nullptr))) return rc;
    void *sorted = in_tmp ? bufB : bufA, *other = in_tmp ? bufA : bufB;
    tm.mark();  // 4: sort1 done
    // K4
    if ((rc = khb_resolve_unique_impl(ctx, sorted, n_sym, k, fb1, other, ctx->d_mail))) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, 8, cudaMemcpyDeviceToHost, ctx->stream));
    tm.mark();  // 5: unique done
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    const u64 s_g = ctx->h_mail[0];
    // K3' group sort, prefix only
    khb_prefix_plan(k, s_g, &fb2, &np2);
    u64 one_seg[2] = {0, s_g};
    if ((rc = khb_sort_bits_impl(ctx, other, sorted, one_seg, 1, (int)W, fb2, np2, &in_tmp, nullptr, nullptr))) return rc;
    void *gsorted = in_tmp ? sorted : other;
    tm.mark();  // 6: sort2 done
    // K5
    void *out_keys = nullptr;
    if (keep_set) {
        if ((rc = gs_reserve(ctx, k, s_g))) return rc;
        ctx->gs_hashed = hashed;
        out_keys = (char *)ctx->gs_buf + ctx->gs_len * W;
    }
    u64 *d_hist = ctx->d_mail + 8, *d_runs = ctx->d_mail;
    if ((rc = khb_resolve_count_impl(ctx, gsorted, s_g, k, fb2, KHB_COUNTER_MAX, nbins, d_hist, out_keys, d_runs))) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, (nbins + 9) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    tm.mark();  // 7: count done
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, ctx->h_mail + 8, ((size_t)nbins + 1) * sizeof(u64));
    const u64 d_g = ctx->h_mail[0];
    if (keep_set) {
        ctx->gs_len += d_g;
        ctx->gs_groups += 1;
    }
    if (stats) {
        stats->passes_genome = np1;
        stats->passes_group = np2;
    }
    if (stats) {
        stats->fasta_bytes = nbytes;
        stats->bases = n_sym - counts[1];
        stats->windows = n_sym;
        stats->genome_distinct = s_g;
        stats->distinct = d_g;
    }
    return KHB_OK;
}

static int group_from_staged_impl(khb_ctx *ctx, int k, int n_genomes, const uint8_t *d_fasta, const u64 *h_begin, u32 nbins,
                                  u64 *h_hist, int keep_set, khb_stats *stats, PhaseTimer &tm)
{
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "k=%d outside 1..64", k);
    khb_packed pk;
    int rc = pack_stage(ctx, n_genomes, d_fasta, h_begin, pk, tm);
    if (rc) return rc;
    return count_stage(ctx, k, pk, nbins, h_hist, keep_set, stats, tm);
}

static void fill_times(khb_stats *stats, PhaseTimer &tm)
{
    if (!stats) return;
    // marks: 0 start, 1 h2d done, 2 pack, 3 extract, 4 sort1, 5 unique, 6 sort2, 7 count
    stats->ms_h2d = tm.ms(0, 1);
    stats->ms_pack = tm.ms(1, 2);
    stats->ms_extract = tm.ms(2, 3);
    stats->ms_sort1 = tm.ms(3, 4);
    stats->ms_unique = tm.ms(4, 5);
    stats->ms_sort2 = tm.ms(5, 6);
    stats->ms_count = tm.ms(6, 7);
    stats->ms_total = tm.ms(0, 7);
}

extern "C" {

int khb_group_from_staged(khb_ctx *ctx, int k, int n_genomes, const uint8_t *d_fasta, const uint64_t *h_begin, uint32_t nbins,
                          uint64_t *h_hist, int keep_set, khb_stats *stats)
{
    KHB_CHECK_CTX(ctx);
    if (stats) memset(stats, 0, sizeof(*stats));
    PhaseTimer tm(ctx);
    tm.mark();
    tm.mark();
    int rc = group_from_staged_impl(ctx, k, n_genomes, d_fasta, (const u64 *)h_begin, nbins, (u64 *)h_hist, keep_set, stats, tm);
    if (rc == KHB_OK) fill_times(stats, tm);
    return rc;
}

static int ensure_stage(khb_ctx *ctx, uint8_t **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return KHB_OK;
    if (*buf) {
        KHB_CUDA(ctx, cudaDeviceSynchronize());
        KHB_CUDA(ctx, cudaFree(*buf));
        *buf = nullptr;
        *cap = 0;
    }
    const size_t want = need + need / 8;
    cudaError_t e = cudaMalloc((void **)buf, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return khb_fail(ctx, KHB_ERR_NOMEM, "staging buffer: device allocation of %zu bytes failed", want);
    }
    *cap = want;
    return KHB_OK;
}

int khb_group_prefetch_fasta(khb_ctx *ctx, int n_genomes, const uint8_t *const *h_files, const size_t *h_sizes)
{
    KHB_CHECK_CTX(ctx);
    if (n_genomes < 1 || !h_files || !h_sizes) return khb_fail(ctx, KHB_ERR_ARG, "khb_group_prefetch_fasta: bad arguments");
    if (!ctx->copy_stream) {
        KHB_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        KHB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->copy_done, cudaEventDisableTiming));
        ctx->pf_begin = new khb_hostvec();
    }
    if (ctx->pf_valid) {
        // a prefetched group is still waiting to be consumed: its buffer cannot be overwritten.  Remember the
        // request; khb_group_from_fasta starts it as soon as it has taken over the pending buffer.
        ctx->pf_begin->d_files.assign(h_files, h_files + n_genomes);
        ctx->pf_begin->d_sizes.assign(h_sizes, h_sizes + n_genomes);
        ctx->pf_begin->deferred = true;
        return KHB_OK;
    }
    ctx->pf_valid = 0;
    const size_t need = khb_staged_size(n_genomes, h_sizes);
    int rc = ensure_stage(ctx, &ctx->stage_next, &ctx->stage_next_cap, need);
    if (rc) return rc;
    const size_t tab_bytes = 2 * (size_t)n_genomes * sizeof(u64);
    if (tab_bytes > ctx->pf_tab_cap) {
        if (ctx->pf_tab) {
            KHB_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
            KHB_CUDA(ctx, cudaFree(ctx->pf_tab));
        }
        KHB_CUDA(ctx, cudaMalloc((void **)&ctx->pf_tab, tab_bytes * 2));
        ctx->pf_tab_cap = tab_bytes * 2;
    }
    std::vector<u64> &begin = ctx->pf_begin->v;
    begin.assign((size_t)n_genomes + 1, 0);
    std::vector<u64> tab(2 * (size_t)n_genomes);
    u64 off = 0;
    for (int i = 0; i < n_genomes; i++) {
        begin[i] = off;
        tab[i] = off;
        tab[n_genomes + i] = h_sizes[i];
        off += staged_len(h_sizes[i]);
    }
    begin[n_genomes] = off;
    for (int i = 0; i < n_genomes; i++)
        if (h_sizes[i]) KHB_CUDA(ctx, cudaMemcpyAsync(ctx->stage_next + begin[i], h_files[i], h_sizes[i], cudaMemcpyHostToDevice, ctx->copy_stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->pf_tab, tab.data(), tab_bytes, cudaMemcpyHostToDevice, ctx->copy_stream));  // pageable: staged before return
    rc = khb_fasta_separators_impl(ctx, ctx->stage_next, ctx->pf_tab, ctx->pf_tab + n_genomes, n_genomes, off, ctx->copy_stream);
    if (rc) return rc;
    ctx->launches++;
    KHB_CUDA(ctx, cudaEventRecord(ctx->copy_done, ctx->copy_stream));
    ctx->pf_valid = 1;
    ctx->pf_n = n_genomes;
    ctx->pf_first = h_files[0];
    ctx->pf_bytes = off;
    return KHB_OK;
}

// The text of a group in ctx->stage_dev: the prefetched copy when it is this group's (the buffers are swapped, the stream waits for the
// copy on the device, a deferred prefetch request is started), else a copy made now.  Records marks 0 and 1 of `tm`.
static int fasta_acquire(khb_ctx *ctx, int n_genomes, const uint8_t *const *h_files, const size_t *h_sizes, PhaseTimer &tm, std::vector<u64> &begin)
{
    if (ctx->pf_valid && ctx->pf_n == n_genomes && ctx->pf_first == (const void *)h_files[0] &&
        ctx->pf_bytes == khb_staged_size(n_genomes, h_sizes)) {
        // the text of this group was prefetched: swap staging buffers and wait for the copy on the device
        ctx->pf_valid = 0;
        uint8_t *tb = ctx->stage_dev; ctx->stage_dev = ctx->stage_next; ctx->stage_next = tb;
        size_t tc = ctx->stage_dev_cap; ctx->stage_dev_cap = ctx->stage_next_cap; ctx->stage_next_cap = tc;
        tm.mark();
        KHB_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_done, 0));
        tm.mark();
        begin = ctx->pf_begin->v;
        if (ctx->pf_begin->deferred) {  // start copying the next group now; it overlaps this group's kernels
            ctx->pf_begin->deferred = false;
            std::vector<const uint8_t *> nf = ctx->pf_begin->d_files;
            std::vector<size_t> ns = ctx->pf_begin->d_sizes;
            int prc = khb_group_prefetch_fasta(ctx, (int)nf.size(), nf.data(), ns.data());
            if (prc) return prc;
        }
        return KHB_OK;
    }
    ctx->pf_valid = 0;
    const size_t need = khb_staged_size(n_genomes, h_sizes);
    if (need > ctx->stage_dev_cap) {
        if (ctx->stage_dev) {
            KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            KHB_CUDA(ctx, cudaFree(ctx->stage_dev));
            ctx->stage_dev = nullptr;
            ctx->stage_dev_cap = 0;
        }
        const size_t cap = need + need / 8;
        cudaError_t e = cudaMalloc((void **)&ctx->stage_dev, cap);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return khb_fail(ctx, KHB_ERR_NOMEM, "staging buffer: device allocation of %zu bytes failed", cap);
        }
        ctx->stage_dev_cap = cap;
    }
    tm.mark();
    begin.assign((size_t)n_genomes + 1, 0);
    int rc = khb_stage_fasta(ctx, n_genomes, h_files, h_sizes, ctx->stage_dev, ctx->stage_dev_cap, (uint64_t *)begin.data());
    if (rc) return rc;
    tm.mark();
    return KHB_OK;
}

int khb_group_from_fasta(khb_ctx *ctx, int k, int n_genomes, const uint8_t *const *h_files, const size_t *h_sizes,
                         uint32_t nbins, uint64_t *h_hist, int keep_set, khb_stats *stats)
{
    KHB_CHECK_CTX(ctx);
    if (n_genomes < 1 || !h_files || !h_sizes) return khb_fail(ctx, KHB_ERR_ARG, "khb_group_from_fasta: bad arguments");
    if (stats) memset(stats, 0, sizeof(*stats));
    PhaseTimer tm(ctx);
    std::vector<u64> begin;
    int rc = fasta_acquire(ctx, n_genomes, h_files, h_sizes, tm, begin);
    if (rc) return rc;
    rc = group_from_staged_impl(ctx, k, n_genomes, ctx->stage_dev, begin.data(), nbins, (u64 *)h_hist, keep_set, stats, tm);
    if (rc == KHB_OK) fill_times(stats, tm);
    return rc;
}

// ---- one group on several GPUs (include/khoice_b200.h: khb_team_*; buffers: team.cu; kernels: bins.cu) ---------------------------------
static int team_partition(khb_ctx *ctx, int k, const khb_packed &pk, const khb_team_group *tg, uint64_t *h_info)
{
    khb_team *tmm = ctx->team;
    if (!tmm || !tmm->opened) return khb_fail(ctx, KHB_ERR_STATE, "khb_team_partition: khb_team_alloc / khb_team_open first");
    if (!tg || !h_info) return khb_fail(ctx, KHB_ERR_ARG, "khb_team_partition: null argument");
    if (pk.n_genomes < 1 || pk.n_genomes > 8191) return khb_fail(ctx, KHB_ERR_ARG, "khb_team_partition: %d genomes in the slice", pk.n_genomes);
    u64 *d_seg = ctx->d_mail + 32768;
    KHB_CUDA(ctx, cudaMemcpyAsync(d_seg, pk.seg.data(), (size_t)(pk.n_genomes + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // seg is pageable
    u64 *d_info = ctx->d_mail + 16640;   // away from the histogram mailbox and the partition counters
    int rc = khb_bins_team_partition_impl(ctx, pk.d_codes, pk.d_valid, pk.n_sym, k, d_seg, pk.n_genomes, tg, d_info);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail + 16640, d_info, 4 * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));   // the kernel is complete: its stores into the owners' buffers are visible to them
    h_info[0] = ctx->h_mail[16640] & 3ull;
    h_info[1] = ctx->h_mail[16643];
    h_info[2] = pk.n_sym;
    h_info[3] = pk.n_sym - pk.n_breaks;
    tmm->slice_sym = pk.n_sym;
    tmm->slice_bases = pk.n_sym - pk.n_breaks;
    tmm->slice_fasta_bytes = pk.fasta_bytes;
    const u64 rec_bytes = ctx->h_mail[16641] * (k <= 32 ? 24 : 32);
    khb_prof_patch(ctx, KHB_K_BIN_PARTITION, pk.n_sym * 3 / 8 + rec_bytes);
    khb_prof_patch(ctx, KHB_K_PARTITION, 2 * rec_bytes);
    return KHB_OK;
}

int khb_team_partition_staged(khb_ctx *ctx, int k, int n_genomes, const uint8_t *d_fasta, const uint64_t *h_begin, const khb_team_group *tg, uint64_t *h_info)
{
    KHB_CHECK_CTX(ctx);
    PhaseTimer tm(ctx);
    tm.mark();
    tm.mark();
    khb_packed pk;
    int rc = pack_stage(ctx, n_genomes, d_fasta, (const u64 *)h_begin, pk, tm);
    if (rc) return rc;
    return team_partition(ctx, k, pk, tg, h_info);
}

int khb_team_partition_fasta(khb_ctx *ctx, int k, int n_genomes, const uint8_t *const *h_files, const size_t *h_sizes, const khb_team_group *tg,
                             uint64_t *h_info)
{
    KHB_CHECK_CTX(ctx);
    if (n_genomes < 1 || !h_files || !h_sizes) return khb_fail(ctx, KHB_ERR_ARG, "khb_team_partition_fasta: bad arguments");
    PhaseTimer tm(ctx);
    std::vector<u64> begin;
    int rc = fasta_acquire(ctx, n_genomes, h_files, h_sizes, tm, begin);
    if (rc) return rc;
    khb_packed pk;
    if ((rc = pack_stage(ctx, n_genomes, ctx->stage_dev, begin.data(), pk, tm))) return rc;
    return team_partition(ctx, k, pk, tg, h_info);
}

int khb_team_partition_packed(khb_ctx *ctx, int k, const khb_packed *pk, const khb_team_group *tg, uint64_t *h_info)
{
    KHB_CHECK_CTX(ctx);
    if (!pk) return khb_fail(ctx, KHB_ERR_ARG, "khb_team_partition_packed: null handle");
    return team_partition(ctx, k, *pk, tg, h_info);
}

int khb_team_plan(khb_ctx *ctx, int k, const khb_team_group *tg, uint32_t *n_bins, uint32_t *region_cap, uint64_t *half_bytes)
{
    KHB_CHECK_CTX(ctx);
    u64 hb = 0;
    int rc = khb_bins_team_plan_impl(ctx, k, tg, KHB_COUNTER_MAX, n_bins, region_cap, &hb);
    if (rc == KHB_OK && half_bytes) *half_bytes = hb;
    return rc;
}

int khb_team_count(khb_ctx *ctx, int k, const khb_team_group *tg, uint32_t nbins, uint64_t *h_hist, int keep_set, khb_stats *stats)
{
    KHB_CHECK_CTX(ctx);
    khb_team *tmm = ctx->team;
    if (!tmm || !tmm->opened) return khb_fail(ctx, KHB_ERR_STATE, "khb_team_count: khb_team_alloc / khb_team_open first");
    if (!tg || !h_hist) return khb_fail(ctx, KHB_ERR_ARG, "khb_team_count: null argument");
    if (nbins < 1 || nbins > 8192) return khb_fail(ctx, KHB_ERR_ARG, "nbins=%u outside 1..8192", nbins);
    if (stats) memset(stats, 0, sizeof(*stats));
    const size_t W = (size_t)khb_key_bytes(k);
    int rc;
    void *out_keys = nullptr;
    khb_peer_route route = {0u, 0u, 0ull, nullptr, nullptr};
    if (keep_set) {
        if (ctx->gs_k && ctx->gs_hashed != 1) return khb_fail(ctx, KHB_ERR_STATE, "retained group sets use a different key encoding");
        if ((rc = gs_reserve(ctx, k, tg->n_sym_total))) return rc;   // this member's bins hold at most every window of the group
        ctx->gs_hashed = 1;
        out_keys = (char *)ctx->gs_buf + ctx->gs_len * W;
        if (ctx->gs_len == 0) ev_forget(ctx);
        ctx->ev_ok = 0;
        if ((rc = khb_peer_route_get(ctx, (int)W, &route))) return rc;
    }
    PhaseTimer tm(ctx);
    tm.mark();
    u64 *d_hist = ctx->d_mail + 8, *d_runs = ctx->d_mail, *d_pairs = ctx->d_mail + 1;
    if ((rc = khb_bins_team_count_impl(ctx, k, tg, KHB_COUNTER_MAX, nbins, d_hist, out_keys, d_runs, d_pairs, ctx->d_mail + 3, route))) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, (nbins + 9) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    tm.mark();
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_mail[3] & 3) {
        // a class of a bin that cannot be split small enough for the table: a whole group falls back to the sort here, a sharded one cannot
        if (route.world) khb_peer_poison(ctx);
        return khb_fail(ctx, KHB_ERR_STATE, "khb_team_count: a bin of this group cannot be counted in shared memory; run the group unsharded");
    }
    ctx->bins_bigbins += ctx->h_mail[5];
    memcpy(h_hist, ctx->h_mail + 8, ((size_t)nbins + 1) * sizeof(u64));
    const u64 d_g = ctx->h_mail[0];
    khb_prof_patch(ctx, KHB_K_BIN_COUNT, ctx->h_mail[4] * (k <= 32 ? 24 : 32) + d_g * W);
    if (keep_set) {
        ctx->gs_len += d_g;
        ctx->gs_groups += 1;
        if (route.world) khb_peer_mark_pushed(ctx);
    }
    if (stats) {
        stats->fasta_bytes = tmm->slice_fasta_bytes;
        stats->bases = tmm->slice_bases;
        stats->windows = tmm->slice_sym;
        stats->genome_distinct = ctx->h_mail[1];
        stats->distinct = d_g;
        stats->ms_count = tm.ms(0, 1);
        stats->ms_total = tm.ms(0, 1);
    }
    return KHB_OK;
}

int khb_pack_group(khb_ctx *ctx, int n_genomes, const uint8_t *const *h_files, const size_t *h_sizes, khb_packed **out)
{
    KHB_CHECK_CTX(ctx);
    if (n_genomes < 1 || !h_files || !h_sizes || !out) return khb_fail(ctx, KHB_ERR_ARG, "khb_pack_group: bad arguments");
    *out = nullptr;
    const size_t need = khb_staged_size(n_genomes, h_sizes);
    int rc = ensure_stage(ctx, &ctx->stage_dev, &ctx->stage_dev_cap, need);
    if (rc) return rc;
    ctx->pf_valid = 0;
    std::vector<u64> begin((size_t)n_genomes + 1);
    rc = khb_stage_fasta(ctx, n_genomes, h_files, h_sizes, ctx->stage_dev, ctx->stage_dev_cap, (uint64_t *)begin.data());
    if (rc) return rc;
    PhaseTimer tm(ctx);
    khb_packed tmp;
    if ((rc = pack_stage(ctx, n_genomes, ctx->stage_dev, begin.data(), tmp, tm))) return rc;
    // keep a compact private copy: the scratch is reused by the next call
    khb_packed *pk = new khb_packed(tmp);
    const size_t cw = khb_codes_words(tmp.n_sym), vw = khb_valid_words(tmp.n_sym);
    void *mem = nullptr;
    cudaError_t e = cudaMalloc(&mem, cw * 8 + vw * 4);
    if (e != cudaSuccess) {
        cudaGetLastError();
        delete pk;
        return khb_fail(ctx, KHB_ERR_NOMEM, "khb_pack_group: device allocation of %zu bytes failed", cw * 8 + vw * 4);
    }
    pk->d_codes = (u64 *)mem;
    pk->d_valid = (u32 *)(pk->d_codes + cw);
    pk->owned = true;
    KHB_CUDA(ctx, cudaMemcpyAsync(pk->d_codes, tmp.d_codes, cw * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(pk->d_valid, tmp.d_valid, vw * 4, cudaMemcpyDeviceToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *out = pk;
    return KHB_OK;
}

int khb_packed_info(const khb_packed *pk, uint64_t *n_symbols, uint64_t *bases, uint64_t *device_bytes)
{
    if (!pk) return KHB_ERR_ARG;
    if (n_symbols) *n_symbols = pk->n_sym;
    if (bases) *bases = pk->n_sym - pk->n_breaks;
    if (device_bytes) *device_bytes = khb_codes_words(pk->n_sym) * 8 + khb_valid_words(pk->n_sym) * 4;
    return KHB_OK;
}

int khb_packed_free(khb_ctx *ctx, khb_packed *pk)
{
    KHB_CHECK_CTX(ctx);
    if (!pk) return KHB_OK;
    if (pk->owned && pk->d_codes) {
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        KHB_CUDA(ctx, cudaFree(pk->d_codes));
    }
    delete pk;
    return KHB_OK;
}

int khb_group_from_packed(khb_ctx *ctx, int k, const khb_packed *pk, uint32_t nbins, uint64_t *h_hist, int keep_set, khb_stats *stats)
{
    KHB_CHECK_CTX(ctx);
    if (!pk) return khb_fail(ctx, KHB_ERR_ARG, "khb_group_from_packed: null handle");
    if (stats) memset(stats, 0, sizeof(*stats));
    PhaseTimer tm(ctx);
    tm.mark();
    tm.mark();
    tm.mark();  // 2: nothing to pack
    int rc = count_stage(ctx, k, *pk, nbins, (u64 *)h_hist, keep_set, stats, tm);
    if (rc == KHB_OK) fill_times(stats, tm);
    return rc;
}

int khb_set_group_mode(khb_ctx *ctx, int mode)
{
    KHB_CHECK_CTX(ctx);
    if (mode < KHB_GROUP_AUTO || mode > KHB_GROUP_BINS) return khb_fail(ctx, KHB_ERR_ARG, "khb_set_group_mode: mode %d", mode);
    ctx->group_mode = mode;
    return KHB_OK;
}

uint64_t khb_hash_overflows(const khb_ctx *ctx) { return ctx ? ctx->hs_overflows : 0; }
void khb_across_counters(const khb_ctx *ctx, uint64_t *by_bins, uint64_t *by_sort)
{
    if (by_bins) *by_bins = ctx ? ctx->across_by_bins : 0;
    if (by_sort) *by_sort = ctx ? ctx->across_by_sort : 0;
}
void khb_bins_counters(const khb_ctx *ctx, uint64_t *fallbacks, uint64_t *big_bins, uint64_t *repartitions)
{
    if (fallbacks) *fallbacks = ctx ? ctx->bins_fallbacks : 0;
    if (big_bins) *big_bins = ctx ? ctx->bins_bigbins : 0;
    if (repartitions) *repartitions = ctx ? ctx->bins_repartitions : 0;
}

int khb_across_groups(khb_ctx *ctx, uint32_t nbins, uint64_t *h_hist, khb_stats *stats)
{
    KHB_CHECK_CTX(ctx);
    if (!h_hist) return khb_fail(ctx, KHB_ERR_ARG, "khb_across_groups: null histogram");
    if (nbins < 1 || nbins > 8192) return khb_fail(ctx, KHB_ERR_ARG, "nbins=%u outside 1..8192", nbins);
    if (!ctx->gs_k) return khb_fail(ctx, KHB_ERR_STATE, "khb_across_groups: no group set retained");
    if (stats) memset(stats, 0, sizeof(*stats));
    const int k = ctx->gs_k;
    const size_t W = (size_t)khb_key_bytes(k);
    const u64 n = ctx->gs_len;
    void *p;
    int rc = khb_scratch_get(ctx, SCR_KEYS_A, (n + 4) * W, &p);
    if (rc) return rc;
    PhaseTimer tm(ctx);
    tm.mark();
    if (across_by_bins_wanted() && ctx->ev_ok && ctx->ev_len && ctx->ev_nb && ctx->gs_hashed) {
        // every group came through the minimizer bins: count bin by bin over the segments they left (bins.cu), no sort -- 15.3 ms against
        // the sort's 12.8 ms at config 2 (profiles/r2_bench_history.md), so only on request
        u64 *d_hist = ctx->d_mail + 8, *d_runs = ctx->d_mail;
        if ((rc = khb_bins_across_impl(ctx, k, ctx->gs_buf, ctx->ev_buf, ctx->ev_len, ctx->ev_nb, ctx->gs_groups, KHB_COUNTER_MAX, nbins, d_hist, d_runs,
                                       ctx->d_mail + 3))) return rc;
        KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, (nbins + 9) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
        tm.mark();
        tm.mark();
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (!ctx->h_mail[3]) {
            ctx->across_by_bins++;
            memcpy(h_hist, ctx->h_mail + 8, ((size_t)nbins + 1) * sizeof(u64));
            if (stats) {
                stats->windows = n;
                stats->genome_distinct = n;
                stats->distinct = ctx->h_mail[0];
                stats->ms_count = tm.ms(0, 1);
                stats->ms_total = tm.ms(0, 1);
                stats->passes_group = 0;
            }
            return KHB_OK;
        }
        if (getenv("KHB_BINS_VERBOSE")) fprintf(stderr, "[bins] across-group stage by bins gave up (flags %llu); sorting\n", (unsigned long long)ctx->h_mail[3]);
        tm.n = 1;   // a table filled up: sort after all
    }
    ctx->across_by_sort++;
    u64 one_seg[2] = {0, n};
    int in_tmp = 0, fb, np;
    khb_prefix_plan(k, n, &fb, &np);
    if (!ctx->gs_hashed) { fb = 0; np = (2 * k + 7) / 8; }  // unhashed keys (k = 32 / 64 or appended raw sets): full sort
    if ((rc = khb_sort_bits_impl(ctx, ctx->gs_buf, p, one_seg, 1, (int)W, fb, np, &in_tmp, nullptr, nullptr))) return rc;
    tm.mark();
    u64 *d_hist = ctx->d_mail + 8, *d_runs = ctx->d_mail;
    if ((rc = khb_resolve_count_impl(ctx, in_tmp ? p : ctx->gs_buf, n, k, fb, KHB_COUNTER_MAX, nbins, d_hist, nullptr, d_runs))) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, (nbins + 9) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    tm.mark();
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, ctx->h_mail + 8, ((size_t)nbins + 1) * sizeof(u64));
    if (stats) {
        stats->windows = n;
        stats->genome_distinct = n;
        stats->distinct = ctx->h_mail[0];
        stats->ms_sort2 = tm.ms(0, 1);
        stats->ms_count = tm.ms(1, 2);
        stats->ms_total = tm.ms(0, 2);
        stats->passes_group = np;
    }
    return KHB_OK;
}

int khb_group_sets_hashed(khb_ctx *ctx) { return ctx ? ctx->gs_hashed : 0; }

int khb_group_sets_export(khb_ctx *ctx, void *h_out)
{
    KHB_CHECK_CTX(ctx);
    if (!ctx->gs_len) return KHB_OK;
    const size_t W = (size_t)khb_key_bytes(ctx->gs_k);
    void *p;
    int rc = khb_scratch_get(ctx, SCR_KEYS_B, (ctx->gs_len + 4) * W, &p);
    if (rc) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(p, ctx->gs_buf, ctx->gs_len * W, cudaMemcpyDeviceToDevice, ctx->stream));
    if (ctx->gs_hashed && (rc = khb_remix_impl(ctx, p, ctx->gs_len, ctx->gs_k, 1))) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(h_out, p, ctx->gs_len * W, cudaMemcpyDeviceToHost, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KHB_OK;
}

int khb_group_sets_info(khb_ctx *ctx, int *k, int *n_groups, uint64_t *n_keys)
{
    if (!ctx) return KHB_ERR_ARG;
    if (k) *k = ctx->gs_k;
    if (n_groups) *n_groups = ctx->gs_groups;
    if (n_keys) *n_keys = ctx->gs_len;
    return KHB_OK;
}

int khb_group_sets_device(khb_ctx *ctx, void **d_keys, uint64_t *n_keys)
{
    KHB_CHECK_CTX(ctx);
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (d_keys) *d_keys = ctx->gs_buf;
    if (n_keys) *n_keys = ctx->gs_len;
    return KHB_OK;
}

int khb_group_sets_append_device(khb_ctx *ctx, int k, const void *d_keys, uint64_t n_keys, int n_groups, int hashed)
{
    KHB_CHECK_CTX(ctx);
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "k=%d outside 1..64", k);
    if (ctx->gs_k && ctx->gs_len && ctx->gs_hashed != (hashed ? 1 : 0)) return khb_fail(ctx, KHB_ERR_STATE, "mixing hashed and raw group sets");
    int rc = gs_reserve(ctx, k, n_keys);
    if (rc) return rc;
    ctx->gs_hashed = hashed ? 1 : 0;
    const size_t W = (size_t)khb_key_bytes(k);
    if (n_keys) KHB_CUDA(ctx, cudaMemcpyAsync((char *)ctx->gs_buf + ctx->gs_len * W, d_keys, n_keys * W, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->gs_len += n_keys;
    ctx->gs_groups += n_groups;
    ctx->ev_ok = 0;
    return KHB_OK;
}

int khb_peer_import(khb_ctx *ctx, const uint64_t *h_recv_counts, int k, int n_groups, int hashed)
{
    KHB_CHECK_CTX(ctx);
    const void *recv;
    u64 region;
    int world, kb;
    int rc = khb_peer_regions(ctx, &recv, &region, &world, &kb);
    if (rc) return rc;
    if (!h_recv_counts || k < 1 || k > 64 || khb_key_bytes(k) != kb) return khb_fail(ctx, KHB_ERR_ARG, "khb_peer_import: bad arguments");
    u64 total = 0;
    for (int s = 0; s < world; s++) {
        if (h_recv_counts[s] > region) return khb_fail(ctx, KHB_ERR_ARG, "khb_peer_import: rank %d reports %llu keys, a region holds %llu", s, (u64)h_recv_counts[s], region);
        total += h_recv_counts[s];
    }
    ctx->gs_len = 0;
    ctx->gs_groups = 0;
    ctx->gs_k = 0;
    if ((rc = gs_reserve(ctx, k, total))) return rc;
    ctx->gs_hashed = hashed ? 1 : 0;
    const size_t W = (size_t)kb;
    for (int s = 0; s < world; s++) {
        if (!h_recv_counts[s]) continue;
        KHB_CUDA(ctx, cudaMemcpyAsync((char *)ctx->gs_buf + ctx->gs_len * W, (const char *)recv + (size_t)s * region * W, h_recv_counts[s] * W,
                                      cudaMemcpyDeviceToDevice, ctx->stream));
        ctx->gs_len += h_recv_counts[s];
    }
    ctx->gs_groups = n_groups;
    ctx->ev_ok = 0;
    return KHB_OK;
}

// khb_peer_import + khb_across_groups in one call, without the copy in between: the regions the ranks pushed here are sorted where they
// lie -- the first radix pass reads them piece by piece (radix_sort.cu: gathered input) and writes one contiguous array, the following
// passes ping-pong between that array and the receive buffer itself -- and the run-length count gives the step_8 histogram of this rank's
// hash range.  The retained group sets are forgotten (they were pushed).  Same barrier rules as khb_peer_import.
int khb_peer_across(khb_ctx *ctx, const uint64_t *h_recv_counts, int k, int n_groups, int hashed, uint32_t nbins, uint64_t *h_hist, khb_stats *stats)
{
    KHB_CHECK_CTX(ctx);
    const void *recv;
    u64 region;
    int world, kb;
    int rc = khb_peer_regions(ctx, &recv, &region, &world, &kb);
    if (rc) return rc;
    if (!h_recv_counts || !h_hist || k < 1 || k > 64 || khb_key_bytes(k) != kb || nbins < 1 || nbins > 8192)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_peer_across: bad arguments");
    if (stats) memset(stats, 0, sizeof(*stats));
    std::vector<u64> pos((size_t)world), len((size_t)world);
    u64 n = 0;
    for (int s = 0; s < world; s++) {
        if (h_recv_counts[s] > region) return khb_fail(ctx, KHB_ERR_ARG, "khb_peer_across: rank %d reports %llu keys, a region holds %llu", s, (u64)h_recv_counts[s], region);
        pos[s] = (u64)s * region;
        len[s] = h_recv_counts[s];
        n += h_recv_counts[s];
    }
    ctx->gs_len = 0;
    ctx->gs_groups = 0;
    ctx->gs_k = 0;
    ctx->ev_ok = 0;
    const size_t W = (size_t)kb;
    void *p;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, (n + 4) * W, &p))) return rc;
    PhaseTimer tm(ctx);
    tm.mark();
    ctx->across_by_sort++;
    int in_tmp = 0, fb, np;
    khb_prefix_plan(k, n, &fb, &np);
    if (!hashed) { fb = 0; np = (2 * k + 7) / 8; }
    u64 *d_hist = ctx->d_mail + 8, *d_runs = ctx->d_mail;
    if (n) {
        if ((rc = khb_sort_gathered_impl(ctx, (void *)recv, p, pos.data(), len.data(), world, (int)W, fb, np, &in_tmp))) return rc;
        tm.mark();
        if ((rc = khb_resolve_count_impl(ctx, in_tmp ? p : recv, n, k, fb, KHB_COUNTER_MAX, nbins, d_hist, nullptr, d_runs))) return rc;
    } else {
        tm.mark();
        KHB_CUDA(ctx, cudaMemsetAsync(ctx->d_mail, 0, (nbins + 9) * sizeof(u64), ctx->stream));
    }
    KHB_CUDA(ctx, cudaMemcpyAsync(ctx->h_mail, ctx->d_mail, (nbins + 9) * sizeof(u64), cudaMemcpyDeviceToHost, ctx->stream));
    tm.mark();
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    memcpy(h_hist, ctx->h_mail + 8, ((size_t)nbins + 1) * sizeof(u64));
    if (stats) {
        stats->windows = n;
        stats->genome_distinct = n;
        stats->distinct = ctx->h_mail[0];
        stats->ms_sort2 = tm.ms(0, 1);
        stats->ms_count = tm.ms(1, 2);
        stats->ms_total = tm.ms(0, 2);
        stats->passes_group = np;
    }
    return KHB_OK;
}

int khb_group_sets_append_host(khb_ctx *ctx, int k, const void *h_keys, uint64_t n_keys, int n_groups)
{
    KHB_CHECK_CTX(ctx);
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "k=%d outside 1..64", k);
    if (ctx->gs_k && ctx->gs_len && ctx->gs_hashed) return khb_fail(ctx, KHB_ERR_STATE, "mixing hashed and raw group sets");
    int rc = gs_reserve(ctx, k, n_keys);
    if (rc) return rc;
    ctx->gs_hashed = 0;  // host-provided sets are canonical k-mer values
    const size_t W = (size_t)khb_key_bytes(k);
    if (n_keys) {
        KHB_CUDA(ctx, cudaMemcpyAsync((char *)ctx->gs_buf + ctx->gs_len * W, h_keys, n_keys * W, cudaMemcpyHostToDevice, ctx->stream));
        KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->gs_len += n_keys;
    ctx->gs_groups += n_groups;
    ctx->ev_ok = 0;
    return KHB_OK;
}

int khb_group_sets_reset(khb_ctx *ctx)
{
    if (!ctx) return KHB_ERR_ARG;
    ev_forget(ctx);
    ctx->gs_len = 0;
    ctx->gs_groups = 0;
    ctx->gs_k = 0;
    ctx->gs_hashed = 0;
    if (ctx->pv) {
        ctx->pv->len = 0;
        ctx->pv->p_off.clear();
        ctx->pv->u_off.clear();
    }
    return KHB_OK;
}

// ---- experiment type 2 (pivot analysis) -------------------------------------------------------------------------------
int khb_pivot_group_from_packed(khb_ctx *ctx, int k, const khb_packed *pk, uint32_t nbins, uint64_t *h_hist, int keep_sets, khb_stats *stats)
{
    KHB_CHECK_CTX(ctx);
    if (!pk) return khb_fail(ctx, KHB_ERR_ARG, "khb_pivot_group_from_packed: null handle");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (!ctx->pv) ctx->pv = new khb_pivot_store();
    if (keep_sets && ctx->gs_groups != (int)(ctx->pv->p_off.empty() ? 0 : ctx->pv->p_off.size() - 1))
        return khb_fail(ctx, KHB_ERR_STATE, "pivot groups and plain groups were mixed in the group-set store; call khb_group_sets_reset");
    PhaseTimer tm(ctx);
    tm.mark();
    tm.mark();
    tm.mark();  // 2: nothing to pack
    int rc = count_stage(ctx, k, *pk, nbins, (u64 *)h_hist, keep_sets, stats, tm, 1);
    if (rc == KHB_OK) fill_times(stats, tm);
    return rc;
}

int khb_pivot_sets_info(khb_ctx *ctx, int *n_pivots, uint64_t *n_pivot_keys, uint64_t *n_union_keys)
{
    if (!ctx) return KHB_ERR_ARG;
    const khb_pivot_store *pv = ctx->pv;
    if (n_pivots) *n_pivots = (pv && !pv->p_off.empty()) ? (int)pv->p_off.size() - 1 : 0;
    if (n_pivot_keys) *n_pivot_keys = pv ? pv->len : 0;
    if (n_union_keys) *n_union_keys = ctx->gs_len;
    return KHB_OK;
}

int khb_pivot_across(khb_ctx *ctx, uint32_t nbins, uint64_t *h_hists, khb_stats *stats)
{
    KHB_CHECK_CTX(ctx);
    if (!h_hists) return khb_fail(ctx, KHB_ERR_ARG, "khb_pivot_across: null histogram");
    if (nbins < 1 || nbins > 8192) return khb_fail(ctx, KHB_ERR_ARG, "nbins=%u outside 1..8192", nbins);
    khb_pivot_store *pv = ctx->pv;
    const int G = (pv && !pv->p_off.empty()) ? (int)pv->p_off.size() - 1 : 0;
    if (G < 1 || !ctx->gs_k) return khb_fail(ctx, KHB_ERR_STATE, "khb_pivot_across: no pivot group retained");
    if (2 * G > 65535) return khb_fail(ctx, KHB_ERR_ARG, "khb_pivot_across: %d groups (limit 32767)", G);
    if (pv->u_off.front() != 0 || pv->u_off.back() != ctx->gs_len) return khb_fail(ctx, KHB_ERR_STATE, "khb_pivot_across: group-set store holds other sets");
    if (stats) memset(stats, 0, sizeof(*stats));
    const int k = ctx->gs_k;
    const size_t W = (size_t)khb_key_bytes(k);
    const u64 nu = ctx->gs_len, np_keys = pv->len, n = nu + np_keys;
    if (n >= (1ull << 32)) return khb_fail(ctx, KHB_ERR_ARG, "khb_pivot_across: %llu keys (limit 2^32 - 1)", n);
    int rc;
    void *p;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, (n + 4) * W, &p))) return rc;
    void *bufA = p;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_B, (n + 4) * W, &p))) return rc;
    void *bufB = p;
    if ((rc = khb_scratch_get(ctx, SCR_PAY_A, (n + 8) * 2, &p))) return rc;
    unsigned short *payA = (unsigned short *)p;
    if ((rc = khb_scratch_get(ctx, SCR_PAY_B, (n + 8) * 2, &p))) return rc;
    unsigned short *payB = (unsigned short *)p;
    const size_t hist_bytes = (size_t)G * (nbins + 1) * sizeof(u64);
    if ((rc = khb_scratch_get(ctx, SCR_AUX, hist_bytes + (size_t)(2 * G + 1) * 8 + 64, &p))) return rc;
    u64 *d_hist = (u64 *)p;
    u64 *d_seg = d_hist + (size_t)G * (nbins + 1);
    PhaseTimer tm(ctx);
    tm.mark();
    // concatenate U_1 .. U_G, P_1 .. P_G; payload = set index
    std::vector<u64> seg((size_t)2 * G + 1);
    u64 max_len = 0;
    for (int j = 0; j <= G; j++) seg[j] = pv->u_off[j];
    for (int j = 1; j <= G; j++) seg[G + j] = nu + pv->p_off[j];
    for (int j = 0; j < 2 * G; j++) max_len = seg[j + 1] - seg[j] > max_len ? seg[j + 1] - seg[j] : max_len;
    KHB_CUDA(ctx, cudaMemcpyAsync(bufA, ctx->gs_buf, nu * W, cudaMemcpyDeviceToDevice, ctx->stream));
    if (np_keys) KHB_CUDA(ctx, cudaMemcpyAsync((char *)bufA + nu * W, pv->buf, np_keys * W, cudaMemcpyDeviceToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(d_seg, seg.data(), seg.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // seg is pageable
    if ((rc = khb_fill_segment_ids_impl(ctx, payA, d_seg, 2 * G, max_len))) return rc;
    int in_tmp = 0, fb, npass;
    khb_prefix_plan(k, n, &fb, &npass);
    if (!ctx->gs_hashed) { fb = 0; npass = (2 * k + 7) / 8; }
    u64 one_seg[2] = {0, n};
    if ((rc = khb_sort_bits_impl(ctx, bufA, bufB, one_seg, 1, (int)W, fb, npass, &in_tmp, payA, payB))) return rc;
    tm.mark();
    if ((rc = khb_pivot_across_impl(ctx, in_tmp ? bufB : bufA, in_tmp ? payB : payA, n, k, fb, (u32)G, KHB_COUNTER_MAX, nbins, d_hist))) return rc;
    KHB_CUDA(ctx, cudaMemcpyAsync(h_hists, d_hist, hist_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    tm.mark();
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (stats) {
        stats->windows = n;
        stats->genome_distinct = nu;
        stats->distinct = np_keys;
        stats->ms_sort2 = tm.ms(0, 1);
        stats->ms_count = tm.ms(1, 2);
        stats->ms_total = tm.ms(0, 2);
        stats->passes_group = npass;
    }
    return KHB_OK;
}

int khb_sorted_lookup(khb_ctx *ctx, const void *d_a, uint64_t n_a, const void *d_b, uint64_t n_b, int k, uint64_t *d_index)
{
    KHB_CHECK_CTX(ctx);
    return khb_sorted_lookup_impl(ctx, d_a, n_a, d_b, n_b, k, (u64 *)d_index);
}

// ---- group membership of query k-mers (experiment type 4) ---------------------------------------------------------
int khb_group_membership(khb_ctx *ctx, int n_groups, const uint64_t *h_group_off, const void *d_queries, int n_query_sets,
                         const uint64_t *h_query_off, uint64_t *d_mask, int mask_words)
{
    KHB_CHECK_CTX(ctx);
    if (n_groups < 1 || n_query_sets < 1 || !h_group_off || !h_query_off || !d_queries || !d_mask)
        return khb_fail(ctx, KHB_ERR_ARG, "khb_group_membership: bad arguments");
    if (!ctx->gs_k) return khb_fail(ctx, KHB_ERR_STATE, "khb_group_membership: no group set retained");
    if (h_group_off[0] != 0 || h_group_off[n_groups] != ctx->gs_len) return khb_fail(ctx, KHB_ERR_ARG, "khb_group_membership: group offsets do not cover the retained sets");
    if (mask_words < (n_groups + 63) / 64 || mask_words > 4) return khb_fail(ctx, KHB_ERR_ARG, "khb_group_membership: %d groups need %d mask words (limit 4)", n_groups, (n_groups + 63) / 64);
    if (n_groups + n_query_sets > 65535) return khb_fail(ctx, KHB_ERR_ARG, "khb_group_membership: too many sets");
    const int k = ctx->gs_k;
    const size_t W = (size_t)khb_key_bytes(k);
    const u64 nu = ctx->gs_len, nq = h_query_off[n_query_sets], n = nu + nq;
    if (n >= (1ull << 32)) return khb_fail(ctx, KHB_ERR_ARG, "khb_group_membership: %llu keys (limit 2^32 - 1)", n);
    int rc;
    void *p;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_A, (n + 4) * W, &p))) return rc;
    void *bufA = p;
    if ((rc = khb_scratch_get(ctx, SCR_KEYS_B, (n + 4) * W, &p))) return rc;
    void *bufB = p;
    if ((rc = khb_scratch_get(ctx, SCR_PAY_A, (n + 8) * 2, &p))) return rc;
    unsigned short *payA = (unsigned short *)p;
    if ((rc = khb_scratch_get(ctx, SCR_PAY_B, (n + 8) * 2, &p))) return rc;
    unsigned short *payB = (unsigned short *)p;
    const int nsets = n_groups + n_query_sets;
    if ((rc = khb_scratch_get(ctx, SCR_AUX, (size_t)(nsets + 1 + n_query_sets + 1) * 8 + 64, &p))) return rc;
    u64 *d_seg = (u64 *)p, *d_qoff = d_seg + nsets + 1;
    std::vector<u64> seg((size_t)nsets + 1);
    u64 max_len = 0;
    for (int j = 0; j <= n_groups; j++) seg[j] = h_group_off[j];
    for (int j = 1; j <= n_query_sets; j++) seg[n_groups + j] = nu + h_query_off[j];
    for (int j = 0; j < nsets; j++) {
        if (seg[j + 1] < seg[j]) return khb_fail(ctx, KHB_ERR_ARG, "khb_group_membership: offsets not monotone");
        max_len = seg[j + 1] - seg[j] > max_len ? seg[j + 1] - seg[j] : max_len;
    }
    KHB_CUDA(ctx, cudaMemcpyAsync(bufA, ctx->gs_buf, nu * W, cudaMemcpyDeviceToDevice, ctx->stream));
    if (nq) KHB_CUDA(ctx, cudaMemcpyAsync((char *)bufA + nu * W, d_queries, nq * W, cudaMemcpyDeviceToDevice, ctx->stream));
    if (nq && ctx->gs_hashed && (rc = khb_remix_impl(ctx, (char *)bufA + nu * W, nq, k, 0))) return rc;  // queries are canonical values
    KHB_CUDA(ctx, cudaMemcpyAsync(d_seg, seg.data(), seg.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaMemcpyAsync(d_qoff, h_query_off, (size_t)(n_query_sets + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // pageable sources
    if ((rc = khb_fill_segment_ids_impl(ctx, payA, d_seg, nsets, max_len))) return rc;
    KHB_CUDA(ctx, cudaMemsetAsync(d_mask, 0, nq * (size_t)mask_words * 8, ctx->stream));
    int in_tmp = 0, fb, npass;
    khb_prefix_plan(k, n, &fb, &npass);
    if (!ctx->gs_hashed) { fb = 0; npass = (2 * k + 7) / 8; }
    u64 one_seg[2] = {0, n};
    if ((rc = khb_sort_bits_impl(ctx, bufA, bufB, one_seg, 1, (int)W, fb, npass, &in_tmp, payA, payB))) return rc;
    if ((rc = khb_membership_impl(ctx, in_tmp ? bufB : bufA, in_tmp ? payB : payA, n, k, fb, (u32)n_groups, ctx->gs_hashed, d_queries, d_qoff,
                                  mask_words, (u64 *)d_mask))) return rc;
    KHB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KHB_OK;
}

int khb_partition_by_hash(khb_ctx *ctx, const void *d_keys, uint64_t n, int k, int n_parts, void *d_out, uint64_t *h_part_off)
{
    KHB_CHECK_CTX(ctx);
    if (k < 1 || k > 64) return khb_fail(ctx, KHB_ERR_ARG, "k=%d outside 1..64", k);
    if (n_parts < 1 || n_parts > PT_MAXPARTS || !h_part_off) return khb_fail(ctx, KHB_ERR_ARG, "khb_partition_by_hash: n_parts=%d outside 1..%d", n_parts, PT_MAXPARTS);
    return k <= 32 ? partition_impl<Key64>(ctx, (const Key64 *)d_keys, n, n_parts, (Key64 *)d_out, (u64 *)h_part_off)
                   : partition_impl<Key128>(ctx, (const Key128 *)d_keys, n, n_parts, (Key128 *)d_out, (u64 *)h_part_off);
}

}  // extern "C"
