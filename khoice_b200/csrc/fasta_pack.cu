// fasta_pack.cu -- K1: FASTA text -> 2-bit symbol stream + validity bitmap.
//
// Replaces the FASTA reader + splitter front end of `kmc -fm` (reference call site
// /root/reference/workflow/rules/exp_type_1.smk:163; KMC semantics R1-R4 of SURVEY.md section 8c):
//   R1  '>' .. end of line is a header (dropped); it contributes ONE invalid "break" symbol so that no
//       window spans two records (R4);
//   R2  '\n' and '\r' are skipped;
//   R3  ACGTacgt -> 0..3; every other byte is an invalid symbol.
//
// Layout in HBM (MSB-first so that a k-mer is a contiguous bit range, see kmer_extract.cu):
//   codes : u64 words, 32 symbols per word, symbol j at bits [63-2(j%32)-1, 63-2(j%32)] of word j/32
//   valid : u32 words, 32 symbols per word, symbol j at bit 31-(j%32) of word j/32
// Both arrays must be zero-filled by the caller (tile boundaries are merged with atomicOr).
//
// Three launches over tiles of KHB_FASTA_TILE = 16 KiB (1024 threads x one 16-byte vector load):
//   fasta_summary_kernel : per tile, symbol counts for both possible entry states + last line event
//   fasta_scan_kernel    : one CTA; resolves every tile's entry state and exclusive symbol offset
//   fasta_pack_kernel    : re-reads the tile, builds the bit stream in shared memory (atomicOr of
//                          per-thread fragments), copies whole words out coalesced
// A thread's 16 bytes are handled four at a time with byte-parallel (SWAR) arithmetic: inside a sequence
// line the only special byte that normally occurs is '\n', so a chunk whose bytes below 0x40 are all '\n'
// takes a branch-free path (2-bit codes = ((c >> 1) ^ (c >> 2)) & 3, validity = byte compare against
// "ACGT"[code] selected with PRMT, bit gathers by multiplication, newline slots squeezed out of the
// fragment); chunks with '>' / '\r' / other control or punctuation bytes, and the chunk that ends a
// header line, take the byte-serial state machine.
// Algorithmic bytes: F (text) + B/4 (codes) + B/8 (valid); the text is read twice (2F) by design --
// it is 1 byte/base against >100 bytes/base for the sorts that follow.
#include "khb_common.cuh"

#define TILE_BYTES KHB_FASTA_TILE
#define TILE_THREADS 1024
#define TILE_WARPS (TILE_THREADS / 32)

enum { ST_SEQ = 0, ST_HDR = 1, ST_UNK = 2 };
enum { EV_NONE = 0, EV_NL = 1, EV_GT = 2 };

struct __align__(16) TileSummary {
    u32 cnt_common;  // symbols emitted after the tile's first line event (entry-state independent)
    u32 cnt_pre;     // symbols emitted up to and including the first event IF the tile is entered in SEQ
    u32 brk;         // break symbols: low 16 bits common, high 16 bits pre
    u32 last_event;  // 0 none, 1 '\n', 2 '>'
};

__device__ __forceinline__ u32 byte_of(const uint4 &d, int j)
{
    u32 w = j < 4 ? d.x : j < 8 ? d.y : j < 12 ? d.z : d.w;
    return (w >> (8 * (j & 3))) & 0xffu;
}

// 2-bit code of a valid symbol, or 4 for an invalid one.
__device__ __forceinline__ u32 base_code(u32 c)
{
    c &= 0xdfu;  // fold case
    return c == 'A' ? 0u : c == 'C' ? 1u : c == 'G' ? 2u : c == 'T' ? 3u : 4u;
}

// ---- byte-parallel helpers (four text bytes per 32-bit word) -------------------------------------------
// 0x80 in every byte of w that equals the corresponding byte of pat
__device__ __forceinline__ u32 eq80(u32 w, u32 pat)
{
    const u32 t = w ^ pat;
    return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu);
}
// 0x80 in every byte of w that is below 0x40 (control characters, digits, punctuation: '\n' '\r' '>' ...)
__device__ __forceinline__ u32 low80(u32 w) { return ~(w | (w << 1)) & 0x80808080u; }

struct Chunk {
    u32 w[4];   // the 16 text bytes
    u32 nl[4];  // 0x80 per '\n' byte
    bool odd;   // some byte below 0x40 is not '\n' ('>' '\r' digits ...): byte-serial path
    bool has_nl;
};
__device__ __forceinline__ Chunk chunk_classify(const uint4 &d)
{
    Chunk c;
    c.w[0] = d.x; c.w[1] = d.y; c.w[2] = d.z; c.w[3] = d.w;
    u32 odd = 0, any = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        c.nl[j] = eq80(c.w[j], 0x0a0a0a0au);
        odd |= low80(c.w[j]) & ~c.nl[j];
        any |= c.nl[j];
    }
    c.odd = odd != 0;
    c.has_nl = any != 0;
    return c;
}

// Type of the last line event ('\n' or '>') in the chunk.
__device__ __forceinline__ u32 chunk_last_event(const uint4 &d, const Chunk &c)
{
    if (!c.odd) return c.has_nl ? EV_NL : EV_NONE;
    u32 ev = EV_NONE;
#pragma unroll
    for (int j = 0; j < 16; j++) {
        const u32 b = byte_of(d, j);
        if (b == '>') ev = EV_GT;
        else if (b == '\n') ev = EV_NL;
    }
    return ev;
}

// State in which this thread's chunk is entered: decided by the nearest preceding event in the tile (earlier
// lanes of the warp, then earlier warps), else `tile_entry`.  One barrier.  *tile_last = last event of the tile.
__device__ __forceinline__ int entry_state(u32 my_ev, int tile_entry, u32 *s_wev, u32 *tile_last)
{
    const u32 lane = lane_id(), warp = threadIdx.x >> 5;
    const u32 has = __ballot_sync(0xffffffffu, my_ev != EV_NONE);
    const u32 before = has & lanemask_lt();
    const u32 ev_in_warp = __shfl_sync(0xffffffffu, my_ev, before ? 31 - __clz(before) : 0);
    const u32 warp_last = __shfl_sync(0xffffffffu, my_ev, has ? 31 - __clz(has) : 0);
    if (lane == 0) s_wev[warp] = has ? warp_last : (u32)EV_NONE;
    __syncthreads();
    const u32 wev = s_wev[lane];  // TILE_WARPS == 32
    const u32 whas = __ballot_sync(0xffffffffu, wev != EV_NONE);
    const u32 wbefore = whas & ((1u << warp) - 1u);
    const u32 ev_warps = __shfl_sync(0xffffffffu, wev, wbefore ? 31 - __clz(wbefore) : 0);
    *tile_last = whas ? __shfl_sync(0xffffffffu, wev, 31 - __clz(whas)) : (u32)EV_NONE;
    const u32 ev = before ? ev_in_warp : (wbefore ? ev_warps : (u32)EV_NONE);
    return ev == EV_NONE ? tile_entry : (ev == EV_GT ? ST_HDR : ST_SEQ);
}

// Byte-serial state machine over one chunk (the reference semantics R1-R3, one byte at a time): used for chunks
// that hold '>' / '\r' / other bytes below 0x40, that end a header line, or whose entry state is unknown.
// Summary flavour: symbols / breaks split into "common" (after the first event) and "pre" (entered in ST_UNK).
__device__ __noinline__ u64 chunk_count_serial(const uint4 d, int st)
{
    u32 cc = 0, cp = 0, bc = 0, bp = 0;
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        const u32 c = byte_of(d, j);
        const bool gt = c == '>', nl = c == '\n', cr = c == '\r';
        if (st == ST_HDR) {
            if (nl) st = ST_SEQ;
        } else if (st == ST_SEQ) {
            if (gt) { cc++; bc++; st = ST_HDR; }
            else if (!nl && !cr) cc++;
        } else {
            if (gt) { cp++; bp++; st = ST_HDR; }
            else if (nl) st = ST_SEQ;
            else if (!cr) cp++;
        }
    }
    return (u64)cc | ((u64)cp << 16) | ((u64)bc << 32) | ((u64)bp << 48);
}
// Pack flavour: right-aligned code / validity fragments of the n symbols the chunk emits.
__device__ __noinline__ u32 chunk_pack_serial(const uint4 d, int st, u32 *frag_c_out, u32 *frag_v_out)
{
    u32 frag_c = 0, frag_v = 0, n = 0;
#pragma unroll 1
    for (int j = 0; j < 16; j++) {
        const u32 c = byte_of(d, j);
        const bool gt = c == '>', nl = c == '\n', cr = c == '\r';
        if (st == ST_HDR) {
            if (nl) st = ST_SEQ;
        } else {
            if (gt) {
                frag_c <<= 2; frag_v <<= 1; n++;      // break symbol: code 0, invalid
                st = ST_HDR;
            } else if (!nl && !cr) {
                const u32 code = base_code(c);
                frag_c = (frag_c << 2) | (code & 3u);
                frag_v = (frag_v << 1) | (code < 4u ? 1u : 0u);
                n++;
            }
        }
    }
    *frag_c_out = frag_c;
    *frag_v_out = frag_v;
    return n;
}

__device__ __forceinline__ u32 count_nl(const Chunk &c)
{
    return (u32)__popc(c.nl[0]) + (u32)__popc(c.nl[1]) + (u32)__popc(c.nl[2]) + (u32)__popc(c.nl[3]);
}

__global__ void __launch_bounds__(TILE_THREADS)
fasta_summary_kernel(const uint4 *__restrict__ fasta, size_t ntiles, TileSummary *__restrict__ out)
{
    __shared__ u32 s_wev[TILE_WARPS];
    __shared__ u64 s_wsum[TILE_WARPS];
    const size_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    const uint4 d = __ldg(fasta + tile * TILE_THREADS + tid);
    const Chunk c = chunk_classify(d);
    const u32 my_ev = chunk_last_event(d, c);
    u32 tile_last;
    const int st = entry_state(my_ev, ST_UNK, s_wev, &tile_last);
    u64 packed;  // cc | cp << 16 | bc << 32 | bp << 48
    if (st == ST_SEQ && !c.odd) packed = 16u - count_nl(c);
    else if (st == ST_HDR && !c.has_nl) packed = 0;
    else if (st == ST_UNK && !c.odd && !c.has_nl) packed = (u64)16u << 16;
    else packed = chunk_count_serial(d, st);
    // block sum of the four 16-bit lanes (a tile emits at most 16384 symbols, so no lane overflows)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) packed += __shfl_xor_sync(0xffffffffu, packed, o);
    if (lane == 0) s_wsum[warp] = packed;
    __syncthreads();
    if (warp == 0) {
        u64 total = s_wsum[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
        if (lane == 0) {
            TileSummary s;
            s.cnt_common = (u32)(total & 0xffffu);
            s.cnt_pre = (u32)((total >> 16) & 0xffffu);
            s.brk = (u32)((total >> 32) & 0xffffu) | ((u32)((total >> 48) & 0xffffu) << 16);
            s.last_event = tile_last;  // EV_NONE / EV_NL / EV_GT = 0 / 1 / 2
            out[tile] = s;
        }
    }
}
// NOTE on the 16-bit fields: the per-thread counters are summed over the block in 16-bit lanes of one u64;
// 16384 < 65536 so no lane overflows into its neighbour.

__global__ void __launch_bounds__(TILE_THREADS)
fasta_scan_kernel(const TileSummary *__restrict__ summ, size_t ntiles, u64 *__restrict__ tile_base,
                  uint8_t *__restrict__ tile_state, u64 *__restrict__ counts)
{
    __shared__ u64 ws[33];
    const u32 tid = threadIdx.x;
    int carry_state = ST_SEQ;
    u64 carry_sym = 0, carry_brk = 0;
    for (size_t base = 0; base < ntiles; base += TILE_THREADS) {
        const size_t t = base + tid;
        TileSummary s = {0, 0, 0, 0};
        if (t < ntiles) s = summ[t];
        const u64 e = s.last_event ? (((u64)(tid + 1) << 2) | s.last_event) : 0ull;
        u64 tot_e;
        const u64 prev = block_excl_max<u64>(e, ws, &tot_e);
        const int s0 = prev ? (((prev & 3) == 2) ? ST_HDR : ST_SEQ) : carry_state;
        const u32 cnt = s.cnt_common + (s0 == ST_SEQ ? s.cnt_pre : 0u);
        const u32 brk = (s.brk & 0xffffu) + (s0 == ST_SEQ ? (s.brk >> 16) : 0u);
        u64 total;
        const u64 ex = block_excl_sum<u64>((u64)cnt | ((u64)brk << 32), ws, &total);
        if (t < ntiles) {
            tile_base[t] = carry_sym + (ex & 0xffffffffull);
            tile_state[t] = (uint8_t)s0;
        }
        carry_sym += total & 0xffffffffull;
        carry_brk += total >> 32;
        if (tot_e) carry_state = ((tot_e & 3) == 2) ? ST_HDR : ST_SEQ;
    }
    if (tid == 0) {
        tile_base[ntiles] = carry_sym;
        counts[0] = carry_sym;              // stream symbols (bases + one break per header)
        counts[1] = carry_brk;              // break symbols (= header lines)
    }
}

#define CW_WORDS (TILE_BYTES / 16 + 4)   // u32 code words in shared memory (16 symbols each) + shift slack
#define VW_WORDS (TILE_BYTES / 32 + 2)   // u32 validity words (32 symbols each)

// Branch-free fragments of a chunk that lies inside a sequence line and holds nothing below 0x40 but '\n':
// returns n = symbols; *fc = their 2-bit codes, *fv = their validity bits, both right-aligned, first symbol on top.
__device__ __forceinline__ u32 chunk_pack_fast(Chunk &c, u32 *fc, u32 *fv)
{
    u32 frag = 0, vfrag = 0;  // 16 symbols, MSB first: codes fill 32 bits, validity the low 16 bits
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const u32 w = c.w[j];
        u32 x = (w >> 1) & 0x03030303u;           // A 0, C 1, T 2, G 3
        x ^= (x >> 1) & 0x01010101u;              // A 0, C 1, G 2, T 3
        const u32 t = x | (x >> 4);
        const u32 sel = (t & 0x33u) | ((t >> 8) & 0x3300u);              // PRMT selector: nibble i = code of byte i
        const u32 expect = __byte_perm(0x54474341u, 0u, sel);            // "ACGT"[code] per byte
        const u32 v = eq80(w & 0xdfdfdfdfu, expect) >> 7;                // 0x01 per valid byte (case folded)
        x &= v * 3u;                                                     // invalid symbols carry code 0
        frag = (frag << 8) | ((x * 0x40100401u) >> 24);                  // byte i -> bits [7-2i, 6-2i]
        vfrag = (vfrag << 4) | (((v * 0x08040201u) >> 24) & 0xfu);       // byte i -> bit 3-i
    }
    u32 n = 16;
    // squeeze the newline slots out, last one first so that earlier positions stay put
#pragma unroll
    for (int j = 3; j >= 0; j--) {
        u32 m = c.nl[j];
        while (m) {
            const u32 b = (31u - (u32)__clz(m)) >> 3;
            m &= ~(0x80u << (8 * b));
            const u32 p = 4u * (u32)j + b;                 // symbol index of the newline
            const u32 sh = 32u - 2u * p;                   // bits from the top of frag down to and including slot p
            const u32 hi = (u32)(((u64)frag >> sh) << sh); // symbols before p
            const u32 lo = frag & ((1u << (sh - 2u)) - 1u);  // symbols after p
            frag = hi | (lo << 2);
            const u32 vhi = vfrag & ~((1u << (16u - p)) - 1u);
            const u32 vlo = vfrag & ((1u << (15u - p)) - 1u);
            vfrag = (vhi | (vlo << 1)) & 0xffffu;
            n--;
        }
    }
    *fc = n ? frag >> (32u - 2u * n) : 0u;
    *fv = vfrag >> (16u - n);
    return n;
}

__global__ void __launch_bounds__(TILE_THREADS)
fasta_pack_kernel(const uint4 *__restrict__ fasta, size_t ntiles, const u64 *__restrict__ tile_base,
                  const uint8_t *__restrict__ tile_state, u64 *__restrict__ codes, u32 *__restrict__ valid)
{
    __shared__ u32 s_wev[TILE_WARPS];
    __shared__ u32 s_wsum[TILE_WARPS];
    __shared__ u32 cw[CW_WORDS];
    __shared__ u32 vw[VW_WORDS];
    const size_t tile = blockIdx.x;
    if (tile >= ntiles) return;
    const u32 tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    for (u32 i = tid; i < CW_WORDS; i += TILE_THREADS) cw[i] = 0;
    for (u32 i = tid; i < VW_WORDS; i += TILE_THREADS) vw[i] = 0;
    const uint4 d = __ldg(fasta + tile * TILE_THREADS + tid);
    Chunk c = chunk_classify(d);
    const u32 my_ev = chunk_last_event(d, c);
    u32 tile_last;
    const int st = entry_state(my_ev, (int)tile_state[tile], s_wev, &tile_last);  // its barrier also orders the zero-fill
    u32 frag_c = 0, frag_v = 0, n = 0;
    if (st == ST_SEQ && !c.odd) n = chunk_pack_fast(c, &frag_c, &frag_v);
    else if (st == ST_HDR && !c.has_nl) n = 0;
    else n = chunk_pack_serial(d, st, &frag_c, &frag_v);
    // exclusive sum of n over the block
    const u32 inc = warp_incl_sum<u32>(n);
    if (lane == 31) s_wsum[warp] = inc;
    __syncthreads();
    const u32 wtot = s_wsum[lane];
    const u32 winc = warp_incl_sum<u32>(wtot);
    const u32 cnt = __shfl_sync(0xffffffffu, winc, 31);
    const u32 wexc = __shfl_sync(0xffffffffu, winc - wtot, warp);
    const u32 pos = wexc + inc - n;
    const u64 base = tile_base[tile];
    const u32 shift = (u32)(base & 31ull);
    if (n) {
        const u32 p = pos + shift;
        {   // codes: 2n bits at bit offset 2p (MSB-first)
            const u32 b = 2u * p, w = b >> 5, s = b & 31u;
            const u64 x = (u64)frag_c << (64u - 2u * n - s);
            atomicOr(&cw[w], (u32)(x >> 32));
            if ((u32)x) atomicOr(&cw[w + 1], (u32)x);
        }
        {   // validity: n bits at bit offset p
            const u32 w = p >> 5, s = p & 31u;
            const u64 y = (u64)frag_v << (64u - n - s);
            if ((u32)(y >> 32)) atomicOr(&vw[w], (u32)(y >> 32));
            if ((u32)y) atomicOr(&vw[w + 1], (u32)y);
        }
    }
    __syncthreads();
    if (cnt == 0) return;
    const u32 n_units = (shift + cnt + 31u) >> 5;   // 32-symbol units touched by this tile (<= 513)
    for (u32 u = tid; u < n_units; u += TILE_THREADS) {
        const u64 g = (base >> 5) + u;
        const u64 c64 = ((u64)cw[2 * u] << 32) | (u64)cw[2 * u + 1];
        const u32 v32 = vw[u];
        const bool full = (u > 0 || shift == 0) && ((u + 1) * 32u <= shift + cnt);
        if (full) {
            codes[g] = c64;
            valid[g] = v32;
        } else {
            if (c64) atomicOr(&codes[g], c64);
            if (v32) atomicOr(&valid[g], v32);
        }
    }
}

// Fill the gap between two staged files: '\n' (ends an unterminated header line), '>' (break symbol),
// then '\n' up to the next tile-aligned file start.  One thread per gap byte, gaps are tiny.
__global__ void fasta_separator_kernel(uint8_t *__restrict__ fasta, const u64 *__restrict__ file_begin,
                                       const u64 *__restrict__ file_len, int nfiles, u64 total_bytes)
{
    const int f = blockIdx.x;
    if (f >= nfiles) return;
    const u64 from = file_begin[f] + file_len[f];
    const u64 to = (f + 1 < nfiles) ? file_begin[f + 1] : total_bytes;
    for (u64 i = from + threadIdx.x; i < to; i += blockDim.x) fasta[i] = (i == from + 1) ? '>' : '\n';
}

// ---- host side ---------------------------------------------------------------------------------

int khb_pack_fasta_impl(khb_ctx *ctx, const uint8_t *d_fasta, size_t nbytes, u64 *d_codes, u32 *d_valid,
                        size_t cap_symbols, u64 *d_tile_base, u64 *d_counts)
{
    if (nbytes % TILE_BYTES) return khb_fail(ctx, KHB_ERR_ARG, "khb_pack_fasta: nbytes %zu is not a multiple of %d", nbytes, TILE_BYTES);
    if (cap_symbols < nbytes) return khb_fail(ctx, KHB_ERR_CAPACITY, "khb_pack_fasta: cap_symbols %zu < nbytes %zu", cap_symbols, nbytes);
    const size_t ntiles = nbytes / TILE_BYTES;
    KHB_CUDA(ctx, cudaMemsetAsync(d_codes, 0, khb_codes_words(cap_symbols) * sizeof(u64), ctx->stream));
    KHB_CUDA(ctx, cudaMemsetAsync(d_valid, 0, khb_valid_words(cap_symbols) * sizeof(u32), ctx->stream));
    if (ntiles == 0) {
        KHB_CUDA(ctx, cudaMemsetAsync(d_tile_base, 0, sizeof(u64), ctx->stream));
        KHB_CUDA(ctx, cudaMemsetAsync(d_counts, 0, 2 * sizeof(u64), ctx->stream));
        return KHB_OK;
    }
    void *scr;
    int rc = khb_scratch_get(ctx, SCR_TILE, ntiles * (sizeof(TileSummary) + 1) + 64, &scr);
    if (rc) return rc;
    TileSummary *summ = (TileSummary *)scr;
    uint8_t *state = (uint8_t *)(summ + ntiles);
    khb_prof_begin(ctx, KHB_K_PACK);
    fasta_summary_kernel<<<(unsigned)ntiles, TILE_THREADS, 0, ctx->stream>>>((const uint4 *)d_fasta, ntiles, summ);
    KHB_LAUNCH_CHECK(ctx);
    fasta_scan_kernel<<<1, TILE_THREADS, 0, ctx->stream>>>(summ, ntiles, d_tile_base, state, d_counts);
    KHB_LAUNCH_CHECK(ctx);
    fasta_pack_kernel<<<(unsigned)ntiles, TILE_THREADS, 0, ctx->stream>>>((const uint4 *)d_fasta, ntiles, d_tile_base, state, d_codes, d_valid);
    KHB_LAUNCH_CHECK(ctx);
    khb_prof_end(ctx, KHB_K_PACK, (u64)nbytes + nbytes / 4 + nbytes / 8);  // F + B/4 + B/8 with B ~ F
    return KHB_OK;
}

int khb_fasta_separators_impl(khb_ctx *ctx, uint8_t *d_fasta, const u64 *d_file_begin, const u64 *d_file_len,
                              int nfiles, u64 total_bytes, cudaStream_t stream)
{
    if (nfiles <= 0) return KHB_OK;
    fasta_separator_kernel<<<nfiles, 256, 0, stream>>>(d_fasta, d_file_begin, d_file_len, nfiles, total_bytes);
    KHB_LAUNCH_CHECK(ctx);
    return KHB_OK;
}
