// decode_kernels.cuh -- sm_100a kernels of the decode path.
//
//   D0 wah_tile_sum / wah_tile_base / wah_find_lines   where does every WAH line start in its matrix
//                        (the reference finds out by walking: wah2_extract consumes words until >= N bits,
//                        include/wah.hpp:177-223; here: prefix sum of the group length of every word)
//   D1 wah_expand        WAH2-16 words -> bit-row (+ popcount), one warp per line   (wah.hpp:177-235)
//   D2 pbwt_unpermute    per block, sequential over its WAH lines: x[a[j]] = y[j], a <- stable partition
//                        (accessor_internals_new.hpp:221-231, 548-589; gt_block.hpp:124-136)
//   D3 sparse_index      start of every sparse / missing / end-of-vector list (accessor_internals_new.hpp:639-653)
//   D4 compose_records   bit-rows + sparse lists + overlays -> int32 genotype rows
//                        (fill_genotype_array_advance, accessor_internals_new.hpp:198-384); HBM-bound, 4 B/genotype
#pragma once
#include "common.cuh"

namespace xsi {

// dline_flags bits (one byte per binary line of the loaded set)
#define DL_WAH 1u
#define DL_HAPLOID 2u
#define DL_MISSING 4u
#define DL_EOV 8u
#define DL_PHASE 16u
#define DL_WEIRD_WAH 32u  // missing / end-of-vector lines of this record are WAH rows (WS_WAH): mord / eord are job ordinals

struct DecSeg {       // one WAH matrix (GT lines or phase lines) of one block
    uint64_t byte_off;  // in blob
    uint32_t n_words;
    uint32_t job0, njobs;
    uint32_t tile0;     // first tile of this segment
};
struct DecBlock {
    uint64_t blob_off;
    uint64_t sparse_off, miss_off, eov_off;  // byte offsets in blob of the matrices (or ~0)
    uint64_t sp_end, ms_end, ev_end;          // entries (aet bytes each) each matrix may hold: bound of the list walk
    uint32_t line0, n_lines;                  // binary lines (global index base)
    uint32_t wah0, n_wah;                     // GT WAH jobs
    uint32_t sp0, n_sp, ms0, n_ms, ev0, n_ev; // ordinals of sparse / missing / eov lists
    uint32_t default_phasing;
    // lazy chain (D2 v3): WAH lines [0, wah_done) of the block are back in sample order; a launch of the inverse-PBWT
    // kernel processes [wah_done, wah_todo) and leaves its positions in pos_state for the next one
    uint32_t wah_done, wah_todo, pad2;
};

struct DecDev {
    const uint8_t* blob;
    const DecSeg* segs; uint32_t nseg;
    const DecBlock* blocks; uint32_t nb;
    const uint32_t* tile_seg;   // [ntiles]
    const uint32_t* tile_word0; // [ntiles]
    uint32_t ntiles;
    uint32_t* tile_sum;         // [ntiles] groups in tile, then exclusive base within the segment
    const uint32_t* job_gcum;   // [NJ] groups before this line within its segment
    const uint32_t* job_nbits;  // [NJ]
    const uint32_t* job_seg;    // [NJ]
    uint32_t* job_word0;        // [NJ] word index within the segment (0xFFFFFFFF = not found)
    uint32_t* job_ones;         // [NJ]
    uint32_t NJ;
    uint32_t* rows;             // [NJ][WS] expanded rows; GT rows become natural order after D2
    uint32_t WS;
    uint32_t* tabs;             // [n_gt_jobs][TW] or NULL: per 16 positions (zeros before << 16 | y bits), then at
                                // [2*WS] the line's total zeros; TW = 2*WS + 4; for D2 v2
    uint32_t TW, n_gt_jobs;
    uint32_t tab_inv;           // 0xFFFF: table entries keep the zero positions as set bits (D2 v3), 0: the y bits (D2 v2)
    uint32_t tab_wide;          // 1: one entry per 32 positions, {zeros before (u32), zero positions as set bits} (D2 wide, > 65534 haplotypes)
    uint32_t n_samples;         // header.num_samples
    uint32_t aet;               // 2 or 4
    const uint8_t* dline_flags; // [Lt]
    const uint32_t* dline_ord;  // [Lt] WAH job ordinal (DL_WAH) or sparse list ordinal
    const uint32_t* dline_mord; // [Lt] ordinal of the missing list (if DL_MISSING)
    const uint32_t* dline_eord; // [Lt]
    const uint32_t* dline_pord; // [Lt] job ordinal of the phase row (if DL_PHASE)
    uint64_t* sp_off;           // [nsp] entry offset of each sparse list within its matrix
    uint64_t* ms_off;           // [nms]
    uint64_t* ev_off;           // [nev]
    uint32_t* err;              // device error word
};
#define DERR_WAH_STREAM 1u   // a WAH line does not start on a word boundary / wrong total
#define DERR_INDEX 2u

constexpr int D0_TILE = 2048;
constexpr int D0_THREADS = 256;

__device__ __forceinline__ uint32_t wah_word_groups(uint32_t w) { return (w & 0x8000u) ? (w & 0x3FFFu) : 1u; }

__global__ void __launch_bounds__(D0_THREADS) wah_tile_sum_kernel(DecDev d) {
    __shared__ uint32_t s_sum;
    const uint32_t t = blockIdx.x;
    const DecSeg sg = d.segs[d.tile_seg[t]];
    const uint16_t* w = reinterpret_cast<const uint16_t*>(d.blob + sg.byte_off);
    const uint32_t w0 = d.tile_word0[t], w1 = min(sg.n_words, w0 + D0_TILE);
    if (threadIdx.x == 0) s_sum = 0;
    __syncthreads();
    uint32_t acc = 0;
    for (uint32_t i = w0 + threadIdx.x; i < w1; i += D0_THREADS) acc += wah_word_groups(w[i]);
    acc = __reduce_add_sync(XSI_FULL, acc);
    if (lane_id() == 0) atomicAdd(&s_sum, acc);
    __syncthreads();
    if (threadIdx.x == 0) d.tile_sum[t] = s_sum;
}

// one warp per segment: exclusive scan of its tile sums
__global__ void __launch_bounds__(128) wah_tile_base_kernel(DecDev d, uint32_t* seg_total) {
    const uint32_t s = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (s >= d.nseg) return;
    const uint32_t lane = lane_id();
    const DecSeg sg = d.segs[s];
    const uint32_t nt = (sg.n_words + D0_TILE - 1) / D0_TILE;
    uint32_t carry = 0;
    for (uint32_t b = 0; b < nt; b += 32) {
        const uint32_t i = b + lane;
        const uint32_t v = i < nt ? d.tile_sum[sg.tile0 + i] : 0u;
        uint32_t incl = v;
#pragma unroll
        for (int q = 1; q < 32; q <<= 1) { const uint32_t o = __shfl_up_sync(XSI_FULL, incl, q); if (lane >= (uint32_t)q) incl += o; }
        if (i < nt) d.tile_sum[sg.tile0 + i] = carry + incl - v;
        carry += __shfl_sync(XSI_FULL, incl, 31);
    }
    if (lane == 0) seg_total[s] = carry;
}

// per tile: local scan, then locate the lines that start inside this tile
__global__ void __launch_bounds__(D0_THREADS) wah_find_lines_kernel(DecDev d) {
    __shared__ uint32_t s_g[D0_TILE + 1];
    __shared__ uint32_t s_w[D0_THREADS / 32];
    const uint32_t t = blockIdx.x;
    const DecSeg sg = d.segs[d.tile_seg[t]];
    const uint16_t* w = reinterpret_cast<const uint16_t*>(d.blob + sg.byte_off);
    const uint32_t w0 = d.tile_word0[t], w1 = min(sg.n_words, w0 + D0_TILE);
    const uint32_t n = w1 - w0;
    const uint32_t base = d.tile_sum[t];
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    constexpr int PER = D0_TILE / D0_THREADS;  // 8 consecutive words per thread
    uint32_t v[PER], tot = 0;
#pragma unroll
    for (int q = 0; q < PER; ++q) { const uint32_t i = tid * PER + q; v[q] = i < n ? wah_word_groups(w[w0 + i]) : 0u; tot += v[q]; }
    uint32_t incl = tot;
#pragma unroll
    for (int q = 1; q < 32; q <<= 1) { const uint32_t o = __shfl_up_sync(XSI_FULL, incl, q); if (lane >= (uint32_t)q) incl += o; }
    if (lane == 31) s_w[warp] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t q = 0; q < warp; ++q) wbase += s_w[q];
    uint32_t ex = base + wbase + incl - tot;
#pragma unroll
    for (int q = 0; q < PER; ++q) { s_g[tid * PER + q] = ex; ex += v[q]; }
    if (tid == D0_THREADS - 1) s_g[D0_TILE] = ex;
    __syncthreads();
    const uint32_t g_lo = base, g_hi = s_g[D0_TILE];  // groups [g_lo, g_hi) start in this tile
    // jobs of this segment whose gcum lies in [g_lo, g_hi) (an empty tile range finds nothing)
    const uint32_t* gc = d.job_gcum + sg.job0;
    uint32_t lo = 0, hi = sg.njobs;
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (gc[m] < g_lo) lo = m + 1; else hi = m; }
    const uint32_t j_lo = lo;
    hi = sg.njobs;
    while (lo < hi) { const uint32_t m = (lo + hi) >> 1; if (gc[m] < g_hi) lo = m + 1; else hi = m; }
    const uint32_t j_hi = lo;
    for (uint32_t j = j_lo + tid; j < j_hi; j += D0_THREADS) {
        const uint32_t target = gc[j];
        uint32_t a = 0, b = n;  // first word with s_g >= target
        while (a < b) { const uint32_t m = (a + b) >> 1; if (s_g[m] < target) a = m + 1; else b = m; }
        if (a < n && s_g[a] == target) d.job_word0[sg.job0 + j] = w0 + a;
        else atomicOr(d.err, DERR_WAH_STREAM);
    }
}

// =============================================================================================
// D1: expand one WAH line per warp into a bit-row
// =============================================================================================
// dynamic smem per warp: g15[Gpad] (u16) | tog[Tpad] (u32)
__global__ void wah_expand_kernel(DecDev d, uint32_t warps_per_cta, uint32_t Gpad, uint32_t Tpad) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t wi = threadIdx.x >> 5, lane = lane_id();
    const uint32_t job = blockIdx.x * warps_per_cta + wi;
    if (job >= d.NJ) return;
    uint16_t* g15 = reinterpret_cast<uint16_t*>(smem_raw + (size_t)wi * (Gpad * 2 + Tpad * 4));
    uint32_t* tog = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(g15) + Gpad * 2);
    const DecSeg sg = d.segs[d.job_seg[job]];
    const uint16_t* w = reinterpret_cast<const uint16_t*>(d.blob + sg.byte_off);
    const uint32_t nbits = d.job_nbits[job];
    const uint32_t G = (nbits + 14) / 15;
    const uint32_t ws = d.job_word0[job];
    const uint32_t we = (job + 1 < sg.job0 + sg.njobs) ? d.job_word0[job + 1] : sg.n_words;
    uint32_t* out = d.rows + (size_t)job * d.WS;
    if (ws == 0xFFFFFFFFu || we == 0xFFFFFFFFu || we < ws) {
        if (lane == 0) atomicOr(d.err, DERR_WAH_STREAM);
        for (uint32_t m = lane; m < d.WS; m += 32) out[m] = 0;
        return;
    }
    for (uint32_t i = lane; i < Gpad; i += 32) g15[i] = 0;
    for (uint32_t i = lane; i < Tpad; i += 32) tog[i] = 0;
    __syncwarp();
    uint32_t gbase = 0;
    uint32_t wpre = ws + lane < we ? w[ws + lane] : 0u;  // the next 32 WAH words are in flight while these are placed
    for (uint32_t b = ws; b < we; b += 32) {
        const uint32_t i = b + lane;
        const uint32_t word = wpre;
        wpre = i + 32 < we ? w[i + 32] : 0u;
        const uint32_t ng = i < we ? wah_word_groups(word) : 0u;
        uint32_t incl = ng;
#pragma unroll
        for (int q = 1; q < 32; q <<= 1) { const uint32_t o = __shfl_up_sync(XSI_FULL, incl, q); if (lane >= (uint32_t)q) incl += o; }
        const uint32_t gs = gbase + incl - ng;
        // the last line of a matrix may be followed by alignment bytes: words past G groups are not ours
        if (i < we && gs < G) {
            if (!(word & 0x8000u)) { if (gs < G) g15[gs] = (uint16_t)word; }
            else if ((word & 0x4000u) && ng) {  // run of all-one groups: toggle marks, filled by the prefix-xor below
                const uint32_t s0 = min(gs, G), e0 = min(gs + ng, G);
                atomicXor(&tog[s0 >> 5], 1u << (s0 & 31));
                atomicXor(&tog[e0 >> 5], 1u << (e0 & 31));
            }
        }
        const uint32_t endg = (i < we && gs < G) ? gs + ng : 0u;  // end of the last word that starts inside the line
        gbase = max(gbase, __reduce_max_sync(XSI_FULL, endg));
        if (gbase >= G) break;
    }
    if (gbase != G && lane == 0) atomicOr(d.err, DERR_WAH_STREAM);
    __syncwarp();
    // prefix-xor over the toggle bits -> bit g set iff group g lies inside a ones-run
    uint32_t carry = 0;
    for (uint32_t b = 0; b < Tpad; b += 32) {
        const uint32_t i = b + lane;
        uint32_t tw = i < Tpad ? tog[i] : 0u;
        tw ^= tw << 1; tw ^= tw << 2; tw ^= tw << 4; tw ^= tw << 8; tw ^= tw << 16;
        const uint32_t pm = __ballot_sync(XSI_FULL, tw >> 31);
        const uint32_t cin = (__popc(pm & lanemask_lt()) & 1u) ^ carry;
        if (cin) tw = ~tw;
        if (i < Tpad) tog[i] = tw;
        carry ^= (__popc(pm) & 1u);
    }
    __syncwarp();
    uint32_t ones = 0;
    const uint32_t nwords = (nbits + 31) >> 5;
    uint32_t* tab = (d.tabs && job < d.n_gt_jobs) ? d.tabs + (size_t)job * d.TW : nullptr;
    uint32_t zcarry = 0;
    for (uint32_t m0 = 0; m0 < d.WS; m0 += 32) {
        const uint32_t m = m0 + lane;
        uint32_t o = 0;
        if (m < nwords) {
            const uint32_t b0 = m * 32, g0 = b0 / 15, sh = b0 - g0 * 15;
            uint64_t acc = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t g = g0 + q;
                uint32_t val = 0;
                if (g < G) val = ((tog[g >> 5] >> (g & 31)) & 1u) ? 0x7FFFu : g15[g];
                acc |= (uint64_t)val << (15 * q);
            }
            o = (uint32_t)(acc >> sh);
            if (m == nwords - 1 && (nbits & 31)) o &= (1u << (nbits & 31)) - 1u;
        }
        if (m < d.WS) out[m] = o;
        ones += __popc(o);
        if (tab) {  // zeros before every 16-position chunk (positions past nbits are never looked up)
            const uint32_t nz = 32u - __popc(o);
            uint32_t incl = nz;
#pragma unroll
            for (int q = 1; q < 32; q <<= 1) { const uint32_t t = __shfl_up_sync(XSI_FULL, incl, q); if (lane >= (uint32_t)q) incl += t; }
            const uint32_t zp = zcarry + incl - nz;
            // (lanes past the row must not store: their entry would land on the Z slot at [2*WS], which lane 0
            // writes below -- two lanes, one address, no ordering between them)
            if (m < d.WS) {
                *reinterpret_cast<uint2*>(tab + 2 * m) = d.tab_wide
                    ? make_uint2(zp, ~o)
                    : make_uint2((zp << 16) | ((o & 0xFFFFu) ^ d.tab_inv), ((zp + 16u - __popc(o & 0xFFFFu)) << 16) | ((o >> 16) ^ d.tab_inv));
            }
            zcarry += __shfl_sync(XSI_FULL, incl, 31);
        }
    }
    ones = __reduce_add_sync(XSI_FULL, ones);
    if (tab && lane == 0) tab[2 * d.WS] = nbits - ones;  // Z, total zeros of the line
    if (lane == 0) d.job_ones[job] = ones;
}

// =============================================================================================
// D1 wide: one WAH line per CTA for long lines (> 65,534 haplotypes).  The warp-per-line kernel above needs
// 2 bytes of shared memory per 15-bit group (133 KB at a million haplotypes: one warp per SM, 27 ms for 7,045
// lines).  Here the NSEG warps of a CTA each own a segment of the OUTPUT (a multiple of 15 row words = 32 whole
// groups, so word and group boundaries coincide): every warp walks the line's WAH words from the start (a few
// thousand: the stream is compressed, the output is not), keeps the literals and the clipped one-runs that fall
// into its segment, and expands only that.  Zero prefixes for the table are segment-relative first and get
// their base after one block barrier.
// dynamic smem per warp: g15[SEGG + 8] (u16) | tog[SEGT] (u32)
// =============================================================================================
constexpr int D1W_WARPS = 16;
__global__ void __launch_bounds__(D1W_WARPS * 32) wah_expand_wide_kernel(DecDev d, uint32_t SEGW, uint32_t SEGT) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_zeros[D1W_WARPS], s_ones[D1W_WARPS];
    const uint32_t wi = threadIdx.x >> 5, lane = lane_id();
    const uint32_t job = blockIdx.x;
    const uint32_t SEGG = SEGW * 32 / 15;  // groups per segment (SEGW is a multiple of 15)
    const size_t per_warp = ((size_t)(SEGG + 8) * 2 + 3) / 4 * 4 + (size_t)SEGT * 4;
    uint16_t* g15 = reinterpret_cast<uint16_t*>(smem_raw + (size_t)wi * per_warp);
    uint32_t* tog = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(g15) + ((size_t)(SEGG + 8) * 2 + 3) / 4 * 4);
    const DecSeg sg = d.segs[d.job_seg[job]];
    const uint16_t* w = reinterpret_cast<const uint16_t*>(d.blob + sg.byte_off);
    const uint32_t nbits = d.job_nbits[job];
    const uint32_t G = (nbits + 14) / 15;
    const uint32_t ws = d.job_word0[job];
    const uint32_t we = (job + 1 < sg.job0 + sg.njobs) ? d.job_word0[job + 1] : sg.n_words;
    uint32_t* out = d.rows + (size_t)job * d.WS;
    const bool bad = ws == 0xFFFFFFFFu || we == 0xFFFFFFFFu || we < ws;  // uniform over the CTA
    if (bad) {
        if (threadIdx.x == 0) atomicOr(d.err, DERR_WAH_STREAM);
        for (uint32_t m = threadIdx.x; m < d.WS; m += blockDim.x) out[m] = 0;
        return;
    }
    const uint32_t g_lo = wi * SEGG, g_hi = min(G, g_lo + SEGG);  // groups of this warp (empty when g_lo >= G)
    for (uint32_t i = lane; i < SEGG + 8; i += 32) g15[i] = 0;
    for (uint32_t i = lane; i < SEGT; i += 32) tog[i] = 0;
    __syncwarp();
    // ---- 1: walk the words; keep what falls into [g_lo, g_hi) ----
    uint32_t gbase = 0;
    if (g_lo < G) {
        for (uint32_t b = ws; b < we; b += 32) {
            const uint32_t i = b + lane;
            const uint32_t word = i < we ? w[i] : 0u;
            const uint32_t ng = i < we ? wah_word_groups(word) : 0u;
            uint32_t incl = ng;
#pragma unroll
            for (int q = 1; q < 32; q <<= 1) { const uint32_t o = __shfl_up_sync(XSI_FULL, incl, q); if (lane >= (uint32_t)q) incl += o; }
            const uint32_t gs = gbase + incl - ng;
            if (i < we && gs < G) {  // words past G groups belong to the next line / the alignment bytes
                if (!(word & 0x8000u)) { if (gs >= g_lo && gs < g_hi) g15[gs - g_lo] = (uint16_t)word; }
                else if ((word & 0x4000u) && ng) {  // run of all-one groups, clipped to the segment: toggle marks
                    const uint32_t s0 = max(gs, g_lo), e0 = min(gs + ng, g_hi);
                    if (s0 < e0) {
                        atomicXor(&tog[(s0 - g_lo) >> 5], 1u << ((s0 - g_lo) & 31));
                        atomicXor(&tog[(e0 - g_lo) >> 5], 1u << ((e0 - g_lo) & 31));
                    }
                }
            }
            const uint32_t endg = (i < we && gs < G) ? gs + ng : 0u;
            gbase = max(gbase, __reduce_max_sync(XSI_FULL, endg));
            if (gbase >= g_hi) break;
        }
        // the warp of the last segment has seen the whole line: it must end exactly at G groups
        if (g_hi == G && gbase != G && lane == 0) atomicOr(d.err, DERR_WAH_STREAM);
    }
    __syncwarp();
    // ---- 2: prefix-xor over the toggle bits -> bit g set iff group g_lo+g lies inside a ones-run ----
    uint32_t carry = 0;
    for (uint32_t b = 0; b < SEGT; b += 32) {
        const uint32_t i = b + lane;
        uint32_t tw = i < SEGT ? tog[i] : 0u;
        tw ^= tw << 1; tw ^= tw << 2; tw ^= tw << 4; tw ^= tw << 8; tw ^= tw << 16;
        const uint32_t pm = __ballot_sync(XSI_FULL, tw >> 31);
        const uint32_t cin = (__popc(pm & lanemask_lt()) & 1u) ^ carry;
        if (cin) tw = ~tw;
        if (i < SEGT) tog[i] = tw;
        carry ^= (__popc(pm) & 1u);
    }
    __syncwarp();
    // ---- 3a: the row words of the segment, zero prefixes relative to the segment ----
    const uint32_t nwords = (nbits + 31) >> 5;
    const uint32_t m_lo = wi * SEGW, m_hi = min(d.WS, m_lo + SEGW);
    uint32_t* tab = (d.tabs && job < d.n_gt_jobs) ? d.tabs + (size_t)job * d.TW : nullptr;
    uint32_t ones = 0, zcarry = 0;
    for (uint32_t m0 = m_lo; m0 < m_hi; m0 += 32) {
        const uint32_t m = m0 + lane;
        uint32_t o = 0;
        if (m < m_hi && m < nwords) {
            const uint32_t b0 = m * 32, g0 = b0 / 15, sh = b0 - g0 * 15;
            uint64_t acc = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t g = g0 + q;
                uint32_t val = 0;
                if (g < g_hi) { const uint32_t gr = g - g_lo; val = ((tog[gr >> 5] >> (gr & 31)) & 1u) ? 0x7FFFu : g15[gr]; }
                acc |= (uint64_t)val << (15 * q);
            }
            o = (uint32_t)(acc >> sh);
            if (m == nwords - 1 && (nbits & 31)) o &= (1u << (nbits & 31)) - 1u;
        }
        if (m < m_hi) out[m] = o;
        ones += __popc(o);
        const uint32_t nz = m < m_hi ? 32u - __popc(o) : 0u;
        uint32_t incl = nz;
#pragma unroll
        for (int q = 1; q < 32; q <<= 1) { const uint32_t t = __shfl_up_sync(XSI_FULL, incl, q); if (lane >= (uint32_t)q) incl += t; }
        if (tab && m < m_hi) *reinterpret_cast<uint2*>(tab + 2 * m) = make_uint2(zcarry + incl - nz, ~o);  // wide entries only
        zcarry += __shfl_sync(XSI_FULL, incl, 31);
    }
    ones = __reduce_add_sync(XSI_FULL, ones);
    if (lane == 0) { s_zeros[wi] = zcarry; s_ones[wi] = ones; }
    __syncthreads();
    // ---- 3b: segment bases ----
    uint32_t base = 0, total_ones = 0;
    for (uint32_t q = 0; q < (uint32_t)D1W_WARPS; ++q) { if (q < wi) base += s_zeros[q]; total_ones += s_ones[q]; }
    if (tab && base) for (uint32_t m = m_lo + lane; m < m_hi; m += 32) tab[2 * m] += base;
    if (threadIdx.x == 0) {
        if (tab) tab[2 * d.WS] = nbits - total_ones;  // Z, total zeros of the line
        d.job_ones[job] = total_ones;
    }
}

// =============================================================================================
// D2: undo the PBWT order, a[] in shared memory as uint16 (2*num_samples <= 65536)
// =============================================================================================
// dynamic smem: a[N] u16 | ybuf[2][WS] | xb[N+32] u8 | zc[64] | mbar[2]
template <int WPW, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) pbwt_unpermute_smem_kernel(DecDev d, const uint8_t* __restrict__ job_hap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t S = d.n_samples, N = 2 * S;
    const uint32_t W = (N + 31) >> 5, WS = d.WS;
    const uint32_t a_bytes = ((N * 2 + 15) / 16) * 16;
    uint16_t* a = reinterpret_cast<uint16_t*>(smem_raw);
    uint32_t* ybuf = reinterpret_cast<uint32_t*>(smem_raw + a_bytes);
    uint8_t* xb = reinterpret_cast<uint8_t*>(ybuf + 2 * WS);
    const uint32_t xb_bytes = ((N + 32 + 15) / 16) * 16;
    uint32_t* zc = reinterpret_cast<uint32_t*>(xb + xb_bytes);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(zc + 64);
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5, NW = blockDim.x >> 5;
    const DecBlock blk = d.blocks[blockIdx.x];
    const uint32_t nwah = blk.n_wah, row_bytes = WS * 4;
    for (uint32_t i = tid; i < N; i += blockDim.x) a[i] = (uint16_t)i;
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); fence_proxy_async(); }
    __syncthreads();
    if (nwah == 0) return;
    if (tid == 0) { mbar_expect_tx(&mbar[0], row_bytes); bulk_g2s(ybuf, d.rows + (size_t)blk.wah0 * WS, row_bytes, &mbar[0]); }
    uint32_t par0 = 0, par1 = 0;
    const uint32_t w0 = warp * WPW, ltm = lanemask_lt();
    for (uint32_t k = 0; k < nwah; ++k) {
        const uint32_t cur = k & 1, job = blk.wah0 + k;
        const bool hap = job_hap[job] != 0;
        if (k + 1 < nwah && tid == 0) {
            mbar_expect_tx(&mbar[cur ^ 1], row_bytes);
            bulk_g2s(ybuf + (cur ^ 1) * WS, d.rows + (size_t)(job + 1) * WS, row_bytes, &mbar[cur ^ 1]);
        }
        if (cur == 0) { mbar_wait(&mbar[0], par0); par0 ^= 1; } else { mbar_wait(&mbar[1], par1); par1 ^= 1; }
        uint32_t* yrow = ybuf + cur * WS;
        uint32_t* grow = d.rows + (size_t)job * WS;
        uint32_t av[WPW / 2 > 0 ? WPW / 2 : 1];
        uint32_t zeros = 0, evens = 0;
        if (!hap) {
            // ---- A: x[a[j]] = y[j] ----
#pragma unroll
            for (int q = 0; q < WPW; ++q) {
                const uint32_t widx = w0 + q, j = widx * 32 + lane;
                const bool valid = j < N;
                const uint32_t aj = valid ? a[j] : 0u;
                if (q & 1) av[q >> 1] |= aj << 16; else av[q >> 1] = aj;
                const uint32_t yk = widx < W ? yrow[widx] : 0u;
                if (valid) xb[aj] = (uint8_t)((yk >> lane) & 1u);
                zeros += __popc(~yk & __ballot_sync(XSI_FULL, valid));
            }
        } else {
            // haploid line: y is over a1 = even entries of a (interfaces.hpp:318-333); x[sample] = y[rank among evens]
#pragma unroll
            for (int q = 0; q < WPW; ++q) {
                const uint32_t widx = w0 + q, j = widx * 32 + lane;
                const bool valid = j < N;
                const uint32_t aj = valid ? a[j] : 0u;
                if (q & 1) av[q >> 1] |= aj << 16; else av[q >> 1] = aj;
                evens += __popc(__ballot_sync(XSI_FULL, valid && !(aj & 1u)));
            }
            if (lane == 0) zc[32 + warp] = evens;
            __syncthreads();
            const uint32_t ev = lane < NW ? zc[32 + lane] : 0u;
            uint32_t ebase = __reduce_add_sync(XSI_FULL, lane < warp ? ev : 0u);
#pragma unroll
            for (int q = 0; q < WPW; ++q) {
                const uint32_t widx = w0 + q, j = widx * 32 + lane;
                const bool valid = j < N;
                const uint32_t aj = (q & 1) ? (av[q >> 1] >> 16) : (av[q >> 1] & 0xFFFFu);
                const bool even = valid && !(aj & 1u);
                const uint32_t ek = __ballot_sync(XSI_FULL, even);
                if (even) {
                    const uint32_t i = ebase + __popc(ek & ltm);
                    xb[aj >> 1] = i < S ? (uint8_t)((yrow[i >> 5] >> (i & 31)) & 1u) : (uint8_t)0;
                }
                ebase += __popc(ek);
            }
            __syncthreads();
            // y2[j] = x[a[j]/2] for all j, kept in ybuf (the consumed row) for the partition below
            __syncthreads();
#pragma unroll
            for (int q = 0; q < WPW; ++q) {
                const uint32_t widx = w0 + q, j = widx * 32 + lane;
                const bool valid = j < N;
                const uint32_t aj = (q & 1) ? (av[q >> 1] >> 16) : (av[q >> 1] & 0xFFFFu);
                const uint32_t bit = valid ? xb[aj >> 1] : 0u;
                const uint32_t yk = __ballot_sync(XSI_FULL, bit);
                zeros += __popc(~yk & __ballot_sync(XSI_FULL, valid));
                if (widx < WS && lane == 0) yrow[widx] = yk;
            }
        }
        if (lane == 0) zc[warp] = zeros;
        __syncthreads();  // #1
        // ---- natural-order row back to global (in place) ----
        {
            const uint32_t nx = hap ? S : N;
            const uint32_t nxw = (nx + 31) >> 5;
            for (uint32_t m = warp; m * 32 < WS; m += NW) {  // 32 words per warp-iteration
                uint32_t keep = 0;
                for (uint32_t s = 0; s < 32; ++s) {
                    const uint32_t widx = m * 32 + s;
                    if (widx >= nxw) break;
                    const uint32_t i = widx * 32 + lane;
                    const uint32_t bw = __ballot_sync(XSI_FULL, i < nx && xb[i]);
                    if (lane == s) keep = bw;
                }
                const uint32_t widx = m * 32 + lane;
                if (widx < WS) grow[widx] = keep;
            }
        }
        // ---- B/C: stable partition of a by y ----
        const uint32_t zv = lane < NW ? zc[lane] : 0u;
        const uint32_t Z = __reduce_add_sync(XSI_FULL, zv);
        uint32_t zbase = __reduce_add_sync(XSI_FULL, lane < warp ? zv : 0u);
        uint32_t obase = Z + (min(w0 * 32, N) - zbase);
#pragma unroll
        for (int q = 0; q < WPW; ++q) {
            const uint32_t widx = w0 + q, j = widx * 32 + lane;
            const bool valid = j < N;
            const uint32_t vm = __ballot_sync(XSI_FULL, valid);
            const uint32_t yk = (widx < W ? yrow[widx] : 0u) & vm;
            const uint32_t nz = ~yk & vm;
            const uint32_t bit = (yk >> lane) & 1u;
            const uint32_t aj = (q & 1) ? (av[q >> 1] >> 16) : (av[q >> 1] & 0xFFFFu);
            const uint32_t dest = bit ? obase + __popc(yk & ltm) : zbase + __popc(nz & ltm);
            if (valid) a[dest] = (uint16_t)aj;
            zbase += __popc(nz);
            obase += __popc(yk);
        }
        if (hap) fence_proxy_async();
        __syncthreads();  // #2
    }
}

// =============================================================================================
// D2 v3: inverse-permutation formulation (state = pos[i]; given the line's table every haplotype is independent:
// j = pos[i]; x[i] = y[j]; pos[i] = y[j] ? Z + j - zb(j) : zb(j)), with the state in REGISTERS.  Thread t owns KH
// consecutive haplotypes: their positions never touch shared memory (v2: one LDS + one STS per
// haplotype and line), and the KH decoded bits of a line are assembled inside the thread (v2: one
// ballot per 32 haplotypes), so the only shared-memory traffic left is the random table lookup
// itself.  A dedicated producer warp keeps the TMA table ring full; consumer warps only wait on
// `full` and arrive on `empty`.  The KH lookups of a thread are independent (ILP hides the LDS latency).
// grid = (PBWT block, haplotype slice of NC*KH), block = NC consumer threads + 32 producer threads
// dynamic smem: ring[D][TW] u32 | full[D], empty[D] u64
// =============================================================================================
constexpr int D3_STAGES = 4;
// The chain may stop early and be continued later (xsi_decode_load_blocks_lazy / xsi_decode_extend): a launch covers the
// WAH lines [wah_done, wah_todo) of every block [b0, b0 + gridDim.x) and parks the positions in pos_state (uint16 per
// haplotype slot) so that a record near the start of a block does not pay for the whole block (seek, :154-196, in reverse).
template <int KH>  // haplotypes per thread: 8, 16 or 32 (one uint8 / uint16 / uint32 store per thread and line)
__global__ void __launch_bounds__(544) pbwt_unpermute_v3_kernel(DecDev d, uint32_t b0, uint16_t* __restrict__ pos_state, uint32_t ps_stride, int fence) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t N = 2 * d.n_samples, TW = d.TW, WS = d.WS;
    const uint32_t tid = threadIdx.x, lane = tid & 31u;
    const uint32_t NC = blockDim.x - 32;  // consumer threads
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)D3_STAGES * TW);
    uint64_t* empty = full + D3_STAGES;
    const DecBlock blk = d.blocks[blockIdx.x + b0];
    const uint32_t k0 = blk.wah_done, k1 = min(blk.wah_todo, blk.n_wah);
    const uint32_t tab_bytes = TW * 4;
    if (tid == 0) {
        for (int s = 0; s < D3_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NC >> 5); }
        fence_proxy_async();
    }
    __syncthreads();
    if (k0 >= k1 || blockIdx.y * NC * KH >= N) return;
    const uint32_t* tabs = d.tabs + (size_t)blk.wah0 * TW;
    if (tid >= NC) {  // ---- producer warp ----
        if (lane == 0) {
            for (uint32_t k = k0; k < k1; ++k) {
                const uint32_t st = (k - k0) % D3_STAGES, use = (k - k0) / D3_STAGES;
                if (use > 0) mbar_wait(&empty[st], (use - 1) & 1u);
                mbar_expect_tx(&full[st], tab_bytes);
                bulk_g2s(ring + (size_t)st * TW, tabs + (size_t)k * TW, tab_bytes, &full[st]);
            }
        }
        return;
    }
    // ---- consumers ----
    // Table entries for v3 hold the ZERO positions as set bits (DecDev::tab_inv): with w = e << (31 - s), s = j & 15,
    // the sign of w says "position j holds a zero" and popc(w & 0x7FFFFFFF) counts the zeros in [16c, j).
    const uint32_t hb = (blockIdx.y * NC + tid) * KH;  // first haplotype of this thread
    const uint32_t nvalid = hb >= N ? 0u : (N - hb >= (uint32_t)KH ? (uint32_t)KH : N - hb);
    uint32_t pk[KH];
    // identity at block start (gt_block.hpp:179).  Haplotypes past N start at 0 and wander inside [0, N] (their
    // lookups stay inside the table); their bits are masked off before the store.
    uint16_t* ps = pos_state ? pos_state + (size_t)(blockIdx.x + b0) * ps_stride + hb : nullptr;
    if (k0 == 0 || !ps) {
#pragma unroll
        for (int q = 0; q < KH; ++q) pk[q] = (uint32_t)q < nvalid ? hb + q : 0u;
    } else {  // continue the chain where the previous launch stopped
#pragma unroll
        for (int q = 0; q < KH; ++q) pk[q] = ps[q];
    }
    const uint32_t vmask = nvalid >= 32 ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);
    const bool store = hb < WS * 32;
    const uint32_t ring_sa = smem_u32(ring);
    for (uint32_t k = k0; k < k1; ++k) {
        const uint32_t st = (k - k0) % D3_STAGES, use = (k - k0) / D3_STAGES;
        mbar_wait(&full[st], use & 1u);
        const uint32_t stw = st * TW;  // word offset of this stage in the ring
        uint32_t e[KH];
#pragma unroll
        for (int q = 0; q < KH; ++q) e[q] = lds_u32(ring_sa + (((pk[q] >> 4) + stw) << 2));
        const uint32_t Z = lds_u32(ring_sa + ((2 * WS + stw) << 2));
        uint32_t xinv = 0;
#pragma unroll
        for (int q = KH - 1; q >= 0; --q) {
            const uint32_t j = pk[q];
            const uint32_t w = __funnelshift_l(0u, e[q], ~j | 16u);  // e << (31 - (j & 15)): the shift wraps mod 32
            const uint32_t zb = (e[q] >> 16) + __popc(w & 0x7FFFFFFFu);
            pk[q] = ((int32_t)w < 0) ? zb : Z + j - zb;
            xinv = __funnelshift_l(w, xinv, 1);  // (xinv << 1) | sign(w)
        }
        // Release the stage only when this warp's table reads have PERFORMED, not merely issued: ptxas would otherwise put
        // the arrive right after the LDS issue (their results are consumed later), and with consumers that were starved on
        // `full` all waking at once the shared-memory pipe is backed up far enough for the producer's next TMA fill of the
        // stage to overtake reads still queued (seen as wrong rows / wild positions when several contexts decode
        // concurrently; XSI_UNPERM_NC=128 made it near-certain).  The update chain above has CONSUMED every loaded entry by
        // the time xinv and pk[] are complete, so tying the arrive to those registers orders it behind the reads without a
        // memory fence (round 1 used __threadfence_block here, which also waited for the row store of the previous line:
        // `fence` keeps that variant selectable, XSI_UNPERM_FENCE=1).
        if (fence) __threadfence_block();
        else {
#pragma unroll
            for (int q = 0; q < KH; ++q) asm volatile("" ::"r"(pk[q]) : "memory");
            asm volatile("" ::"r"(xinv) : "memory");
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);
        if (store) {  // natural-order row, in place
            const uint32_t x = ~xinv & vmask;
            const size_t row = (size_t)(blk.wah0 + k) * WS;
            if (KH == 32) d.rows[row + (hb >> 5)] = x;
            else if (KH == 16) reinterpret_cast<uint16_t*>(d.rows + row)[hb >> 4] = (uint16_t)x;
            else reinterpret_cast<uint8_t*>(d.rows + row)[hb >> 3] = (uint8_t)x;
        }
    }
    if (ps) {  // also after the last line: xsi_decode_internal_access reads the arrangement from here
#pragma unroll
        for (int q = 0; q < KH; ++q) ps[q] = (uint16_t)pk[q];
    }
}

// =============================================================================================
// D2 wide: the v3 formulation for MORE than 65,534 haplotypes (biobank scale, uint32 indices).  A line's
// table (one {zeros before, zero-position bits} pair per 32 positions, N/4 bytes: 250 KB at a million
// haplotypes) no longer fits shared memory, and no ring could hold several of them; it stays in
// global memory, where the tables of the lines in flight are L2-resident (126 MB), and every lookup is
// one 8-byte read-only load.  State = KH positions per thread in registers, haplotypes independent
// given the tables, so there is no barrier inside the kernel and the whole GPU works on one PBWT block:
// grid = (PBWT block, haplotype slice of blockDim.x*KH), one launch per window of lines.
// =============================================================================================
template <int KH>  // haplotypes per thread: 8, 16 or 32
__global__ void __launch_bounds__(256) pbwt_unpermute_wide_kernel(DecDev d, uint32_t k0, uint32_t k1, uint32_t* __restrict__ pos_state) {
    const uint32_t N = 2 * d.n_samples, TW = d.TW, WS = d.WS;
    const DecBlock blk = d.blocks[blockIdx.x];
    const uint32_t nwah = blk.n_wah;
    const uint32_t t = blockIdx.y * blockDim.x + threadIdx.x;
    const uint32_t hb = t * KH;  // first haplotype of this thread
    if (k0 >= nwah || hb >= WS * 32) return;
    const uint32_t nvalid = hb >= N ? 0u : (N - hb >= (uint32_t)KH ? (uint32_t)KH : N - hb);
    // Lines [k0, k1) of the block per launch: the launch boundary keeps every thread of the GPU within k1-k0 lines
    // of each other, i.e. the tables in flight inside L2 (free-running threads drift apart by hundreds of lines and
    // every lookup becomes a DRAM access: 486 GB of DRAM reads for 1.8 GB of tables, ncu r01o).  Positions travel
    // between launches through pos_state, laid out [block][q][thread] (coalesced).
    const uint32_t TT = gridDim.y * blockDim.x;
    uint32_t* ps = pos_state + ((size_t)blockIdx.x * KH) * TT + t;
    uint32_t pk[KH];
    if (k0 == 0) {
        // identity at block start (gt_block.hpp:179); slots past N wander inside [0, N] and are masked off at the store
#pragma unroll
        for (int q = 0; q < KH; ++q) pk[q] = (uint32_t)q < nvalid ? hb + q : 0u;
    } else {
#pragma unroll
        for (int q = 0; q < KH; ++q) pk[q] = ps[(size_t)q * TT];
    }
    const uint32_t vmask = nvalid >= 32 ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);
    const uint32_t kend = min(k1, nwah);
    const uint32_t* tab = d.tabs + (size_t)(blk.wah0 + k0) * TW;
    uint32_t* row = d.rows + (size_t)(blk.wah0 + k0) * WS;
    for (uint32_t k = k0; k < kend; ++k, tab += TW, row += WS) {
        const uint2* T = reinterpret_cast<const uint2*>(tab);
        uint2 e[KH];
#pragma unroll
        for (int q = 0; q < KH; ++q) e[q] = __ldg(T + (pk[q] >> 5));
        const uint32_t Z = __ldg(tab + 2 * WS);
        uint32_t xinv = 0;
#pragma unroll
        for (int q = KH - 1; q >= 0; --q) {
            const uint32_t j = pk[q];
            const uint32_t w = __funnelshift_l(0u, e[q].y, ~j);  // e.y << (31 - (j & 31)): sign = "position j holds a zero"
            const uint32_t zb = e[q].x + __popc(w & 0x7FFFFFFFu);
            pk[q] = ((int32_t)w < 0) ? zb : Z + j - zb;
            xinv = __funnelshift_l(w, xinv, 1);
        }
        const uint32_t x = ~xinv & vmask;  // natural-order row, in place
        if (KH == 32) row[hb >> 5] = x;
        else if (KH == 16) reinterpret_cast<uint16_t*>(row)[hb >> 4] = (uint16_t)x;
        else reinterpret_cast<uint8_t*>(row)[hb >> 3] = (uint8_t)x;
    }
    if (kend < nwah) {
#pragma unroll
        for (int q = 0; q < KH; ++q) ps[(size_t)q * TT] = pk[q];
    }
}

// generic fallback (all-haploid lines above 65,534 haplotypes): a[] and x[] in global memory
__global__ void __launch_bounds__(1024, 1) pbwt_unpermute_gmem_kernel(DecDev d, const uint8_t* __restrict__ job_hap,
                                                                      uint32_t* a_pool, uint8_t* x_pool) {
    __shared__ uint32_t zc[64];
    const uint32_t S = d.n_samples, N = 2 * S;
    const uint32_t W = (N + 31) >> 5, WS = d.WS;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5, NW = blockDim.x >> 5;
    const DecBlock blk = d.blocks[blockIdx.x];
    uint32_t* abuf[2] = {a_pool + (size_t)blockIdx.x * 2 * N, a_pool + (size_t)blockIdx.x * 2 * N + N};
    uint8_t* xb = x_pool + (size_t)blockIdx.x * (N + 32);
    uint32_t* y2 = a_pool + (size_t)gridDim.x * 2 * N + (size_t)blockIdx.x * WS;
    for (uint32_t i = tid; i < N; i += blockDim.x) abuf[0][i] = i;
    __syncthreads();
    const uint32_t wpw = (W + NW - 1) / NW;
    const uint32_t w0 = warp * wpw, w1 = min(W, w0 + wpw), ltm = lanemask_lt();
    uint32_t cur = 0;
    for (uint32_t k = 0; k < blk.n_wah; ++k) {
        const uint32_t job = blk.wah0 + k;
        const bool hap = job_hap[job] != 0;
        uint32_t* row = d.rows + (size_t)job * WS;
        const uint32_t* a = abuf[cur];
        uint32_t* an = abuf[cur ^ 1];
        uint32_t zeros = 0, evens = 0;
        if (!hap) {
            for (uint32_t widx = w0; widx < w1; ++widx) {
                const uint32_t j = widx * 32 + lane;
                const bool valid = j < N;
                const uint32_t yk = row[widx];
                if (valid) xb[a[j]] = (uint8_t)((yk >> lane) & 1u);
                zeros += __popc(~yk & __ballot_sync(XSI_FULL, valid));
                if (lane == 0) y2[widx] = yk;
            }
        } else {
            for (uint32_t widx = w0; widx < w1; ++widx) {
                const uint32_t j = widx * 32 + lane;
                evens += __popc(__ballot_sync(XSI_FULL, j < N && !(a[j] & 1u)));
            }
            if (lane == 0) zc[32 + warp] = evens;
            __syncthreads();
            const uint32_t ev = lane < NW ? zc[32 + lane] : 0u;
            uint32_t ebase = __reduce_add_sync(XSI_FULL, lane < warp ? ev : 0u);
            for (uint32_t widx = w0; widx < w1; ++widx) {
                const uint32_t j = widx * 32 + lane;
                const bool even = j < N && !(a[j] & 1u);
                const uint32_t ek = __ballot_sync(XSI_FULL, even);
                if (even) { const uint32_t i = ebase + __popc(ek & ltm); xb[a[j] >> 1] = i < S ? (uint8_t)((row[i >> 5] >> (i & 31)) & 1u) : (uint8_t)0; }
                ebase += __popc(ek);
            }
            __syncthreads();
            for (uint32_t widx = w0; widx < w1; ++widx) {
                const uint32_t j = widx * 32 + lane;
                const bool valid = j < N;
                const uint32_t yk = __ballot_sync(XSI_FULL, valid && xb[a[j] >> 1]);
                zeros += __popc(~yk & __ballot_sync(XSI_FULL, valid));
                if (lane == 0) y2[widx] = yk;
            }
        }
        if (lane == 0) zc[warp] = zeros;
        __syncthreads();
        const uint32_t nx = hap ? S : N, nxw = (nx + 31) >> 5;
        for (uint32_t widx = warp; widx < WS; widx += NW) {
            const uint32_t i = widx * 32 + lane;
            const uint32_t bw = (widx < nxw) ? __ballot_sync(XSI_FULL, i < nx && xb[i]) : 0u;
            if (lane == 0) row[widx] = bw;
        }
        const uint32_t zv = lane < NW ? zc[lane] : 0u;
        const uint32_t Z = __reduce_add_sync(XSI_FULL, zv);
        uint32_t zbase = __reduce_add_sync(XSI_FULL, lane < warp ? zv : 0u);
        uint32_t obase = Z + (min(w0 * 32, N) - zbase);
        for (uint32_t widx = w0; widx < w1; ++widx) {
            const uint32_t j = widx * 32 + lane;
            const bool valid = j < N;
            const uint32_t vm = __ballot_sync(XSI_FULL, valid);
            const uint32_t yk = y2[widx] & vm, nz = ~yk & vm;
            const uint32_t bit = (yk >> lane) & 1u;
            const uint32_t dest = bit ? obase + __popc(yk & ltm) : zbase + __popc(nz & ltm);
            if (valid) an[dest] = a[j];
            zbase += __popc(nz);
            obase += __popc(yk);
        }
        __syncthreads();
        cur ^= 1;
    }
}

// =============================================================================================
// D3: start offsets of the variable-length index lists (one thread per matrix)
// =============================================================================================
__global__ void sparse_index_kernel(DecDev d) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.nb * 3) return;
    const DecBlock blk = d.blocks[t / 3];
    const uint32_t kind = t % 3;
    const uint64_t moff = kind == 0 ? blk.sparse_off : kind == 1 ? blk.miss_off : blk.eov_off;
    const uint32_t n = kind == 0 ? blk.n_sp : kind == 1 ? blk.n_ms : blk.n_ev;
    uint64_t* off = kind == 0 ? d.sp_off + blk.sp0 : kind == 1 ? d.ms_off + blk.ms0 : d.ev_off + blk.ev0;
    if (!n) return;
    if (moff == ~0ull) { atomicOr(d.err, DERR_INDEX); return; }
    // a truncated or corrupt file must end in XSI_E_FORMAT, not in a wild read: every list (count word + entries) has to
    // lie inside its matrix; a list that does not raises DERR_INDEX (the load then fails and nothing reads the lists)
    const uint64_t end = kind == 0 ? blk.sp_end : kind == 1 ? blk.ms_end : blk.ev_end;
    uint64_t e = 0;
    if (d.aet == 2) {
        const uint16_t* m = reinterpret_cast<const uint16_t*>(d.blob + moff);
        for (uint32_t i = 0; i < n; ++i) {
            off[i] = 0;
            if (e >= end) { atomicOr(d.err, DERR_INDEX); continue; }
            const uint64_t nx = e + 1 + (m[e] & 0x7FFFu);
            if (nx > end) { atomicOr(d.err, DERR_INDEX); e = end; continue; }
            off[i] = e; e = nx;
        }
    } else {
        const uint8_t* m = d.blob + moff;  // only 2-byte aligned in the file: assemble from halves
        for (uint32_t i = 0; i < n; ++i) {
            off[i] = 0;
            if (e >= end) { atomicOr(d.err, DERR_INDEX); continue; }
            const uint16_t* h = reinterpret_cast<const uint16_t*>(m + e * 4);
            const uint32_t c = (uint32_t)h[0] | ((uint32_t)h[1] << 16);
            const uint64_t nx = e + 1 + (c & 0x7FFFFFFFu);
            if (nx > end) { atomicOr(d.err, DERR_INDEX); e = end; continue; }
            off[i] = e; e = nx;
        }
    }
}

// =============================================================================================
// D4: compose genotype rows
// =============================================================================================
struct ReqDev {
    const uint32_t* blk;       // [n] loaded-block index
    const uint32_t* line;      // [n] first binary line within the block
    const uint32_t* nall;      // [n]
    uint32_t n;
    void* out; uint64_t out_stride;  // int32 (bcf_get_genotypes) or int8 (raw BCF FORMAT/GT) rows, stride in elements
    uint32_t* filled;          // [n]
    uint32_t* counts; uint32_t counts_stride;  // [n][counts_stride]
    uint8_t* scratch;          // [grid][2][Npad]  allele code / phase-by-index flag (general path)
    uint32_t Npad;
};

__device__ __forceinline__ uint32_t rd_entry(const uint8_t* m, uint64_t e, uint32_t aet) {
    if (aet == 2) return reinterpret_cast<const uint16_t*>(m)[e];
    const uint16_t* h = reinterpret_cast<const uint16_t*>(m + e * 4);
    return (uint32_t)h[0] | ((uint32_t)h[1] << 16);
}

constexpr int D4_THREADS = 256;
// Output element: int32 = what bcf_get_genotypes / fill_genotype_array hand out; int8 = the BCF record's own
// FORMAT/GT payload (htslib vcf.h:152-158, what bcf_update_genotypes narrows to when every value fits):
// same (allele+1)<<1|phased code, vector end = 0x81 (bcf_int8_vector_end).
template <typename OT> __device__ __forceinline__ OT gt_out(int32_t v);
template <> __device__ __forceinline__ int32_t gt_out<int32_t>(int32_t v) { return v; }
template <> __device__ __forceinline__ int8_t gt_out<int8_t>(int32_t v) { return v == XSI_I32_VECTOR_END ? (int8_t)0x81 : (int8_t)v; }
template <typename OT> __device__ __forceinline__ void store4(OT* p, int32_t a, int32_t b, int32_t c, int32_t e);
template <> __device__ __forceinline__ void store4<int32_t>(int32_t* p, int32_t a, int32_t b, int32_t c, int32_t e) { *reinterpret_cast<int4*>(p) = make_int4(a, b, c, e); }
template <> __device__ __forceinline__ void store4<int8_t>(int8_t* p, int32_t a, int32_t b, int32_t c, int32_t e) {
    *reinterpret_cast<uint32_t*>(p) = (uint32_t)(a & 0xFF) | ((uint32_t)(b & 0xFF) << 8) | ((uint32_t)(c & 0xFF) << 16) | ((uint32_t)(e & 0xFF) << 24);
}
constexpr uint8_t CODE_MISSING = 254, CODE_EOV = 255;

// a record goes through compose_simple when it has one ALT line, no overlay, and its row is a whole number of 16-byte units
template <typename OT>
__device__ __forceinline__ bool compose_is_simple(uint32_t nall, uint8_t f0, uint32_t n) {
    return nall == 2 && !(f0 & (DL_MISSING | DL_EOV | DL_PHASE)) && (n * (uint32_t)sizeof(OT)) % 16u == 0;
}

template <typename OT>
__global__ void __launch_bounds__(D4_THREADS) compose_records_kernel(DecDev d, ReqDev q, int skip_simple, int scratch_in_smem) {
    extern __shared__ __align__(128) unsigned char smem_raw[];  // [2][Npad] when scratch_in_smem (rows up to 24 K genotypes)
    const uint32_t tid = threadIdx.x;
    const uint32_t S = d.n_samples, NH = 2 * S;
    const uint32_t msb = d.aet == 2 ? 0x8000u : 0x80000000u;
    for (uint32_t ri = blockIdx.x; ri < q.n; ri += gridDim.x) {
        const DecBlock blk = d.blocks[q.blk[ri]];
        const uint32_t nall = q.nall[ri];
        const uint32_t gl0 = blk.line0 + q.line[ri];
        const uint8_t f0 = d.dline_flags[gl0];
        const uint32_t n = (f0 & DL_HAPLOID) ? S : NH;
        const int32_t DP = (int32_t)(blk.default_phasing & 1u);
        OT* out = static_cast<OT*>(q.out) + (size_t)ri * q.out_stride;
        uint32_t* cnts = q.counts ? q.counts + (size_t)ri * q.counts_stride : nullptr;
        const uint8_t* spm = d.blob + blk.sparse_off;
        const bool weird = (f0 & (DL_MISSING | DL_EOV | DL_PHASE)) != 0;
        if (skip_simple && compose_is_simple<OT>(nall, f0, n)) continue;  // written by compose_simple_kernel
        if (tid == 0 && q.filled) q.filled[ri] = n;

        if (nall == 2 && !weird) {
            // -------- fast path: one ALT, no overlays --------
            if (f0 & DL_WAH) {
                const uint32_t job = d.dline_ord[gl0];
                const uint32_t* row = d.rows + (size_t)job * d.WS;
                const bool hap = (f0 & DL_HAPLOID) != 0;
                for (uint32_t i4 = tid * 4; i4 < n; i4 += D4_THREADS * 4) {
                    const uint32_t bits = row[i4 >> 5] >> (i4 & 31);
                    int32_t v[4];
#pragma unroll
                    for (int k = 0; k < 4; ++k) v[k] = (int32_t)((((bits >> k) & 1u) + 1u) << 1) | (hap ? 0 : ((int32_t)((i4 + k) & 1u) & DP));
                    if (i4 + 3 < n && ((reinterpret_cast<uintptr_t>(out + i4) & (4 * sizeof(OT) - 1)) == 0)) store4<OT>(out + i4, v[0], v[1], v[2], v[3]);
                    else for (int k = 0; k < 4; ++k) if (i4 + k < n) out[i4 + k] = (OT)v[k];
                }
                if (cnts && tid == 0) { const uint32_t ones = d.job_ones[job]; cnts[1] = ones; cnts[0] = n - ones; }
            } else {
                const uint64_t e0 = d.sp_off[d.dline_ord[gl0]];
                const uint32_t hdr = rd_entry(spm, e0, d.aet);
                const bool neg = (hdr & msb) != 0;
                const uint32_t cnt = hdr & ~msb;
                const int32_t dflt = neg ? 4 : 2, spv = neg ? 2 : 4;  // bcf_gt_unphased(1) = 4, (0) = 2
                for (uint32_t i = tid; i < n; i += D4_THREADS) out[i] = (OT)(dflt | ((int32_t)(i & 1u) & DP));
                __syncthreads();
                for (uint32_t k = tid; k < cnt; k += D4_THREADS) {
                    const uint32_t i = rd_entry(spm, e0 + 1 + k, d.aet);
                    if (i < q.out_stride) out[i] = (OT)(spv | ((int32_t)(i & 1u) & DP));
                }
                if (cnts && tid == 0) { const uint32_t ones = neg ? n - cnt : cnt; cnts[1] = ones; cnts[0] = n - ones; }
            }
            __syncthreads();
            continue;
        }

        // -------- general path (accessor_internals_new.hpp:207-384) --------
        // allele code / phase-by-index flag per genotype: shared memory when the row fits (a chrX-shaped file sends every
        // record through here: 36 us per record with the scratch in global memory), else the per-CTA global scratch.
        // The passes that cover the whole row (the first ALT line, WAH lines, the final compose) work on QUADS: four
        // genotypes = one 32-bit word of codes and one of flags, four row bits spread with one multiply, 16-byte stores of
        // the result; only the index lists (sparse / missing / end-of-vector) write single bytes.
        uint8_t* val = scratch_in_smem ? smem_raw : q.scratch + (size_t)blockIdx.x * 2 * q.Npad;
        uint8_t* pf = val + q.Npad;
        uint32_t* val4 = reinterpret_cast<uint32_t*>(val);
        uint32_t* pf4 = reinterpret_cast<uint32_t*>(pf);
        const uint32_t nq = (n + 3) >> 2;
        auto spread4 = [](uint32_t bits) { return ((bits & 0xFu) * 0x00204081u) & 0x01010101u; };  // bit k -> byte k
        uint32_t total_alt = 0;
        for (uint32_t alt = 1; alt < nall; ++alt) {
            const uint32_t gl = gl0 + alt - 1;
            const uint8_t fl = d.dline_flags[gl];
            const bool hapl = (fl & DL_HAPLOID) != 0;
            const uint32_t pfw = hapl ? 0u : 0x01010101u;
            uint32_t ones;
            if (fl & DL_WAH) {
                const uint32_t job = d.dline_ord[gl];
                const uint32_t* __restrict__ row = d.rows + (size_t)job * d.WS;
                ones = d.job_ones[job];
                if (alt == 1) {
                    for (uint32_t iq = tid; iq < nq; iq += D4_THREADS) {
                        const uint32_t i = iq << 2;
                        val4[iq] = spread4(__ldg(row + (i >> 5)) >> (i & 31u));
                        pf4[iq] = pfw;
                    }
                } else {
                    const uint32_t code = hapl ? 1u : alt;  // sic: haploid WAH ALT>=2 writes allele 1 (:269)
                    for (uint32_t iq = tid; iq < nq; iq += D4_THREADS) {
                        const uint32_t i = iq << 2;
                        const uint32_t m = spread4(__ldg(row + (i >> 5)) >> (i & 31u));
                        if (m) {
                            const uint32_t mask = m * 0xFFu;
                            val4[iq] = (val4[iq] & ~mask) | (m * code);
                            pf4[iq] = (pf4[iq] & ~mask) | (m & pfw);
                        }
                    }
                }
            } else {
                const uint64_t e0 = d.sp_off[d.dline_ord[gl]];
                const uint32_t hdr = rd_entry(spm, e0, d.aet);
                const bool neg = (hdr & msb) != 0;
                const uint32_t cnt = hdr & ~msb;
                ones = neg ? n - cnt : cnt;
                if (alt == 1) {
                    for (uint32_t iq = tid; iq < nq; iq += D4_THREADS) { val4[iq] = neg ? 0x01010101u : 0u; pf4[iq] = 0x01010101u; }
                    __syncthreads();
                    for (uint32_t k = tid; k < cnt; k += D4_THREADS) { const uint32_t i = rd_entry(spm, e0 + 1 + k, d.aet); if (i < q.Npad) { val[i] = neg ? 0 : 1; pf[i] = 1; } }
                } else if (neg) {
                    for (uint32_t iq = tid; iq < nq; iq += D4_THREADS) {
                        const uint32_t v = val4[iq];
                        const uint32_t m = __vcmpeq4(v, 0u) & 0x01010101u;  // the genotypes that still carry REF
                        if (m) { val4[iq] = v | (m * alt); pf4[iq] |= m; }
                    }
                    __syncthreads();
                    for (uint32_t k = tid; k < cnt; k += D4_THREADS) { const uint32_t i = rd_entry(spm, e0 + 1 + k, d.aet); if (i < q.Npad && val[i] == (uint8_t)alt) { val[i] = 0; pf[i] = 1; } }
                } else {
                    for (uint32_t k = tid; k < cnt; k += D4_THREADS) { const uint32_t i = rd_entry(spm, e0 + 1 + k, d.aet); if (i < q.Npad) { val[i] = (uint8_t)alt; pf[i] = 1; } }
                }
            }
            if (cnts && tid == 0) cnts[alt] = ones;
            total_alt += ones;
            __syncthreads();
        }
        uint32_t n_missing = 0, n_eov = 0;
        if ((f0 & DL_MISSING) && (f0 & DL_WEIRD_WAH)) {
            // --wah-encode-missing files (accessor_internals_new.hpp:307-321): the line was expanded like any WAH row,
            // natural order (a_weird is the identity under WS_WAH)
            const uint32_t job = d.dline_mord[gl0];
            const uint32_t* __restrict__ mrow = d.rows + (size_t)job * d.WS;
            n_missing = d.job_ones[job];
            for (uint32_t iq = tid; iq < nq; iq += D4_THREADS) {
                const uint32_t i = iq << 2;
                const uint32_t m = spread4(__ldg(mrow + (i >> 5)) >> (i & 31u));
                if (m) { val4[iq] = (val4[iq] & ~(m * 0xFFu)) | (m * CODE_MISSING); pf4[iq] |= m; }
            }
            __syncthreads();
        } else if (f0 & DL_MISSING) {
            const uint8_t* mm = d.blob + blk.miss_off;
            const uint64_t e0 = d.ms_off[d.dline_mord[gl0]];
            n_missing = rd_entry(mm, e0, d.aet) & ~msb;
            for (uint32_t k = tid; k < n_missing; k += D4_THREADS) { const uint32_t i = rd_entry(mm, e0 + 1 + k, d.aet); if (i < q.Npad) { val[i] = CODE_MISSING; pf[i] = 1; } }
            __syncthreads();
        }
        if ((f0 & DL_EOV) && (f0 & DL_WEIRD_WAH)) {
            const uint32_t job = d.dline_eord[gl0];
            const uint32_t* __restrict__ erow = d.rows + (size_t)job * d.WS;
            n_eov = d.job_ones[job];
            for (uint32_t iq = tid; iq < nq; iq += D4_THREADS) {
                const uint32_t i = iq << 2;
                const uint32_t m = spread4(__ldg(erow + (i >> 5)) >> (i & 31u));
                if (m) val4[iq] |= m * 0xFFu;  // CODE_EOV = 255
            }
            __syncthreads();
        } else if (f0 & DL_EOV) {
            const uint8_t* mm = d.blob + blk.eov_off;
            const uint64_t e0 = d.ev_off[d.dline_eord[gl0]];
            n_eov = rd_entry(mm, e0, d.aet) & ~msb;
            for (uint32_t k = tid; k < n_eov; k += D4_THREADS) { const uint32_t i = rd_entry(mm, e0 + 1 + k, d.aet); if (i < q.Npad) val[i] = CODE_EOV; }
            __syncthreads();
        }
        const uint32_t* __restrict__ prow = (f0 & DL_PHASE) ? d.rows + (size_t)d.dline_pord[gl0] * d.WS : nullptr;
        const bool vec_ok = (reinterpret_cast<uintptr_t>(out) & (4 * sizeof(OT) - 1)) == 0;
        for (uint32_t iq = tid; iq < nq; iq += D4_THREADS) {
            const uint32_t i = iq << 2;
            const uint32_t c4 = val4[iq], p4 = pf4[iq];
            const uint32_t tog = prow ? (__ldg(prow + (i >> 5)) >> (i & 31u)) & 0xFu : 0u;
            int32_t v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t c = (c4 >> (8 * k)) & 0xFFu;
                const uint32_t odd = (uint32_t)k & 1u;  // i is a multiple of 4: the parity of i + k is that of k
                int32_t x = (c == CODE_MISSING ? 0 : (int32_t)((c + 1u) << 1)) | ((int32_t)((p4 >> (8 * k)) & odd) & DP);
                x ^= (int32_t)((tog >> k) & odd);
                v[k] = c == CODE_EOV ? XSI_I32_VECTOR_END : x;
            }
            if (vec_ok && i + 3 < n) store4<OT>(out + i, gt_out<OT>(v[0]), gt_out<OT>(v[1]), gt_out<OT>(v[2]), gt_out<OT>(v[3]));
            else for (int k = 0; k < 4; ++k) if (i + k < n) out[i + k] = gt_out<OT>(v[k]);
        }
        if (cnts && tid == 0) cnts[0] = n - (total_alt + n_missing + n_eov);
        __syncthreads();
    }
}

// =============================================================================================
// D6: sample subset (-s/-S of the extractor: fill_selected_genotypes, gt_decompressor_new.hpp:208-238).  One CTA per
// record: the selected samples' entries of the full row (ploidy 1 or 2 from the row length) are gathered into a
// compact row, and the selected carriers of every ALT allele are counted (ac_s, the AC the extractor rewrites).
// =============================================================================================
__global__ void __launch_bounds__(256) select_samples_kernel(const int32_t* __restrict__ full, uint64_t full_stride,
                                                              const uint32_t* __restrict__ filled, const uint32_t* __restrict__ nall,
                                                              uint32_t num_samples, const uint32_t* __restrict__ sel, uint32_t n_sel,
                                                              int32_t* __restrict__ out, uint64_t out_stride, uint32_t* __restrict__ out_filled,
                                                              uint32_t* __restrict__ ac, uint32_t ac_stride, uint32_t n) {
    __shared__ uint32_t s_ac[256];
    for (uint32_t r = blockIdx.x; r < n; r += gridDim.x) {
        const uint32_t pl = filled[r] / num_samples;  // CURRENT_LINE_PLOIDY
        const uint32_t na = nall[r];
        for (uint32_t a = threadIdx.x; a < 256; a += blockDim.x) s_ac[a] = 0;
        __syncthreads();
        const int32_t* src = full + (size_t)r * full_stride;
        int32_t* dst = out + (size_t)r * out_stride;
        for (uint32_t i = threadIdx.x; i < n_sel * pl; i += blockDim.x) {
            const uint32_t s = i / pl, k = i - s * pl;
            const int32_t v = src[(size_t)sel[s] * pl + k];
            dst[i] = v;
            const int32_t allele = (v >> 1) - 1;  // bcf_gt_allele; missing (-1) and end-of-vector never match an ALT
            if (ac && allele >= 1 && (uint32_t)allele < na) atomicAdd(&s_ac[allele], 1u);
        }
        __syncthreads();
        if (ac) for (uint32_t a = 1 + threadIdx.x; a < na; a += blockDim.x) ac[(size_t)r * ac_stride + a - 1] = s_ac[a];
        if (threadIdx.x == 0 && out_filled) out_filled[r] = n_sel * pl;
        __syncthreads();
    }
}

// =============================================================================================
// D5: allele counts only (fill_allele_counts_advance, accessor_internals_new.hpp:407-440): per ALT line the
// carriers counted when the line was expanded (WAH) or its list header (sparse, negated: N - count);
// allele_counts[0] = CURRENT_N_HAPS - sum, WITHOUT the missing / end-of-vector correction of fill_genotype_array.
// =============================================================================================
__global__ void __launch_bounds__(128) allele_counts_kernel(DecDev d, ReqDev q) {
    const uint32_t ri = blockIdx.x * blockDim.x + threadIdx.x;
    if (ri >= q.n) return;
    const DecBlock blk = d.blocks[q.blk[ri]];
    const uint32_t nall = q.nall[ri];
    const uint32_t gl0 = blk.line0 + q.line[ri];
    const uint32_t S = d.n_samples, NH = 2 * S;
    const uint32_t msb = d.aet == 2 ? 0x8000u : 0x80000000u;
    const uint32_t n = (d.dline_flags[gl0] & DL_HAPLOID) ? S : NH;
    const uint8_t* spm = d.blob + blk.sparse_off;
    uint32_t* cnts = q.counts + (size_t)ri * q.counts_stride;
    uint32_t total = 0;
    for (uint32_t alt = 1; alt < nall; ++alt) {
        const uint32_t gl = gl0 + alt - 1;
        const uint8_t fl = d.dline_flags[gl];
        uint32_t ones;
        if (fl & DL_WAH) ones = d.job_ones[d.dline_ord[gl]];
        else {
            const uint32_t hdr = rd_entry(spm, d.sp_off[d.dline_ord[gl]], d.aet);
            const uint32_t cnt = hdr & ~msb;
            ones = (hdr & msb) ? ((fl & DL_HAPLOID) ? S : NH) - cnt : cnt;
        }
        cnts[alt] = ones;
        total += ones;
    }
    cnts[0] = n - total;
    if (q.filled) q.filled[ri] = n;
}

// =============================================================================================
// D7: dot products on the encoded lines (the consumer of InternalGtAccess in the reference, dot_prod/dot_prod.hpp:113-245:
// Sxy = sum of y[sample] over the carriers of an ALT allele).  One warp per record and ALT line: a WAH line is read as its
// bit-row in sample order (1 bit per genotype instead of a 4-byte row), a sparse line as its index list.  Lists of REF
// carriers (negated sparse, MSB of the count) are left to the composed-row path (dot_rows_kernel), like the reference, which
// decompresses such lines (dot_prod.hpp:405-412).  Sums are accumulated in double, lane-strided then a shuffle tree.
// =============================================================================================
struct DotDev {
    const double* y;        // [num_samples]
    double* out;            // [n][out_stride], ALT allele a at column a-1
    uint32_t out_stride;
    uint32_t* fallback;     // [n] set to 1 when the record holds a line this kernel does not serve
};
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(XSI_FULL, v, o);
    return v;
}
__global__ void __launch_bounds__(128) dot_lines_kernel(DecDev d, ReqDev q, DotDev t) {
    const uint32_t ri = blockIdx.x * 4 + (threadIdx.x >> 5), lane = lane_id();
    if (ri >= q.n) return;
    const DecBlock blk = d.blocks[q.blk[ri]];
    const uint32_t nall = q.nall[ri];
    const uint32_t gl0 = blk.line0 + q.line[ri];
    const uint32_t S = d.n_samples;
    const uint32_t msb = d.aet == 2 ? 0x8000u : 0x80000000u;
    const bool hap = (d.dline_flags[gl0] & DL_HAPLOID) != 0;
    const uint32_t n = hap ? S : 2 * S, sh = hap ? 0u : 1u;
    const uint8_t* spm = d.blob + blk.sparse_off;
    bool fb = false;
    for (uint32_t alt = 1; alt < nall; ++alt) {
        const uint32_t gl = gl0 + alt - 1;
        const uint8_t fl = d.dline_flags[gl];
        double acc = 0.0;
        if (fl & DL_WAH) {
            const uint32_t* row = d.rows + (size_t)d.dline_ord[gl] * d.WS;
            const uint32_t nwords = (n + 31) >> 5;
            for (uint32_t w = lane; w < nwords; w += 32) {
                uint32_t x = row[w];
                while (x) {
                    const uint32_t b = __ffs(x) - 1;
                    x &= x - 1;
                    const uint32_t i = w * 32 + b;
                    if (i < n) acc += t.y[i >> sh];
                }
            }
        } else {
            const uint64_t e0 = d.sp_off[d.dline_ord[gl]];
            const uint32_t hdr = rd_entry(spm, e0, d.aet);
            if (hdr & msb) fb = true;
            else {
                const uint32_t cnt = hdr & ~msb;
                for (uint32_t k = lane; k < cnt; k += 32) {
                    const uint32_t i = rd_entry(spm, e0 + 1 + k, d.aet);
                    if (i < n) acc += t.y[i >> sh];
                }
            }
        }
        acc = warp_sum(acc);
        if (lane == 0) t.out[(size_t)ri * t.out_stride + alt - 1] = acc;
    }
    if (lane == 0) t.fallback[ri] = fb ? 1u : 0u;
}
// the same sums from composed int8 rows (records with a negated sparse line): row r of `rows` belongs to request idx[r]
__global__ void __launch_bounds__(128) dot_rows_kernel(const int8_t* __restrict__ rows, uint64_t stride, const uint32_t* __restrict__ filled,
                                                        const uint32_t* __restrict__ idx, const uint32_t* __restrict__ nall_all, uint32_t nrows,
                                                        uint32_t n_samples, DotDev t) {
    const uint32_t r = blockIdx.x * 4 + (threadIdx.x >> 5), lane = lane_id();
    if (r >= nrows) return;
    const uint32_t ri = idx[r], n = filled[r], sh = n == n_samples ? 0u : 1u;
    const int8_t* row = rows + (size_t)r * stride;
    for (uint32_t alt = 1; alt < nall_all[ri]; ++alt) {
        double acc = 0.0;
        for (uint32_t i = lane; i < n; i += 32) {
            const int32_t v = row[i];
            if ((v >> 1) - 1 == (int32_t)alt) acc += t.y[i >> sh];
        }
        acc = warp_sum(acc);
        if (lane == 0) t.out[(size_t)ri * t.out_stride + alt - 1] = acc;
    }
}

// =============================================================================================
// D4 fast path: records with one ALT line and no missing / end-of-vector / phase overlay (the bulk
// of any file).  Persistent CTAs build 8192-genotype int32 tiles in shared memory -- a thread
// expands its own 32-bit word of the bit-row (WAH lines) or writes the default pattern that the
// listed carriers then patch (sparse lines) -- and hand every tile to the TMA engine
// (cp.async.bulk shared->global), double buffered.  Needs 16-byte aligned output rows.
// =============================================================================================
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_PENDING) : "memory"); }
template <int N_PENDING>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N_PENDING) : "memory"); }

// CTAs as wide as the rows are long in words (rounded up to a warp, at most 256 threads): at 5,008 haplotypes a 256-thread
// CTA kept 157 threads busy; narrower CTAs also fit more per SM (two 20 KB tiles each), which covers the two barriers a
// record costs.
__host__ __device__ constexpr int d5_ctas_per_sm(int nt, int esize) { return nt >= 256 ? (esize == 1 ? 4 : 2) : (nt >= 192 ? 3 : 5); }
template <typename OT, int NT>
__global__ void __launch_bounds__(NT, d5_ctas_per_sm(NT, (int)sizeof(OT))) compose_simple_kernel(DecDev d, ReqDev q) {
    constexpr uint32_t D5_TILE = NT * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];  // 2 tiles of D5_TILE elements
    OT* tiles = reinterpret_cast<OT*>(smem_raw);
    const uint32_t tid = threadIdx.x;
    const uint32_t S = d.n_samples, NH = 2 * S;
    const uint32_t msb = d.aet == 2 ? 0x8000u : 0x80000000u;
    const uint32_t rot = tid & 7u;
    uint32_t buf = 0;
    // The descriptor of a record is a chain of dependent loads (request -> block -> line flags -> ordinal -> list
    // header): 4-5 L2 round trips, as long as the whole 20 KB row of a 1KGP3-shaped record takes.  It is fetched one
    // record ahead, so the chain of record i+1 resolves while the tiles of record i are built.
    struct Desc { uint32_t f0, ord, DP, cnt, nall; uint64_t e0, sparse_off; bool neg; };
    auto fetch = [&](uint32_t ri) -> Desc {
        Desc x;
        const DecBlock blk = d.blocks[q.blk[ri]];
        const uint32_t gl0 = blk.line0 + q.line[ri];
        x.f0 = d.dline_flags[gl0];
        x.ord = d.dline_ord[gl0];
        x.DP = blk.default_phasing & 1u;
        x.sparse_off = blk.sparse_off;
        x.nall = q.nall[ri];
        x.e0 = 0; x.cnt = 0; x.neg = false;
        if (!(x.f0 & DL_WAH) && x.nall == 2 && !(x.f0 & (DL_MISSING | DL_EOV | DL_PHASE))) {
            x.e0 = d.sp_off[x.ord];
            const uint32_t hdr = rd_entry(d.blob + blk.sparse_off, x.e0, d.aet);
            x.neg = (hdr & msb) != 0; x.cnt = hdr & ~msb;
        }
        return x;
    };
    Desc nx = {};
    if (blockIdx.x < q.n) nx = fetch(blockIdx.x);
    for (uint32_t ri = blockIdx.x; ri < q.n; ri += gridDim.x) {
        const Desc cur = nx;
        if (ri + gridDim.x < q.n) nx = fetch(ri + gridDim.x);
        const uint8_t f0 = (uint8_t)cur.f0;
        const bool hap = (f0 & DL_HAPLOID) != 0;
        const uint32_t n = hap ? S : NH;
        if (!compose_is_simple<OT>(cur.nall, f0, n)) continue;
        const int32_t DP = (int32_t)cur.DP;
        OT* out = static_cast<OT*>(q.out) + (size_t)ri * q.out_stride;
        const bool wah = (f0 & DL_WAH) != 0;
        const uint32_t ord = cur.ord;
        const uint32_t* row = d.rows + (size_t)ord * d.WS;
        const uint8_t* spm = d.blob + cur.sparse_off;
        const uint64_t e0 = cur.e0; const uint32_t cnt = cur.cnt; const bool neg = cur.neg;
        if (tid == 0) {
            if (q.filled) q.filled[ri] = n;
            if (q.counts) {
                const uint32_t ones = wah ? d.job_ones[ord] : (neg ? n - cnt : cnt);
                uint32_t* cnts = q.counts + (size_t)ri * q.counts_stride;
                cnts[1] = ones; cnts[0] = n - ones;
            }
        }
        const int32_t ph = (wah && hap) ? 0 : DP;        // phase bit of odd entries
        const int32_t base_even = (!wah && neg) ? 4 : 2;  // value of a 0 bit: REF (or ALT for negated lists)
        const uint32_t ntiles = (n + D5_TILE - 1) / D5_TILE;
        for (uint32_t tt = 0; tt < ntiles; ++tt) {
            if (tid == 0) bulk_wait_read<1>();  // the store that last used this buffer has read it
            __syncthreads();
            OT* tile = tiles + (size_t)buf * D5_TILE;
            const uint32_t elem0 = tt * D5_TILE;
            uint32_t w = 0;
            if (wah) { const uint32_t wi = tt * NT + tid; w = wi < d.WS ? row[wi] : 0u; }
            const int32_t ce = base_even, co = base_even | ph;
            if (sizeof(OT) == 4) {
                const uint32_t wr = __funnelshift_r(w, w, 4 * rot);  // chunk k of wr = chunk (k + rot) & 7 of w
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) {
                    const uint32_t c = (k + rot) & 7u;
                    int4 v;
                    v.x = ce + (int32_t)(((wr >> (4 * k)) & 1u) << 1);
                    v.y = co + (int32_t)(((wr >> (4 * k + 1)) & 1u) << 1);
                    v.z = ce + (int32_t)(((wr >> (4 * k + 2)) & 1u) << 1);
                    v.w = co + (int32_t)(((wr >> (4 * k + 3)) & 1u) << 1);
                    *reinterpret_cast<int4*>(reinterpret_cast<int32_t*>(tile) + tid * 32 + c * 4) = v;
                }
            } else {
                // 32 genotypes = 32 bytes per thread: nibble -> 4 bytes by one multiply (bit i lands in byte i), two
                // 16-byte stores; the half order alternates every 4 threads so that a quarter-warp covers all banks
                const uint32_t pat = (uint32_t)ce * 0x00010001u + (uint32_t)co * 0x01000100u;
                uint32_t o[8];
#pragma unroll
                for (uint32_t k = 0; k < 8; ++k) o[k] = pat + ((((w >> (4 * k)) & 0xFu) * 0x00204081u & 0x01010101u) << 1);
                const uint32_t flip = (tid >> 2) & 1u;
                uint4* t4 = reinterpret_cast<uint4*>(reinterpret_cast<int8_t*>(tile) + tid * 32);
                const uint4 lo = make_uint4(o[0], o[1], o[2], o[3]), hi = make_uint4(o[4], o[5], o[6], o[7]);
                t4[flip] = flip ? hi : lo;
                t4[flip ^ 1u] = flip ? lo : hi;
            }
            if (!wah && cnt) {
                __syncthreads();
                const int32_t spv = neg ? 2 : 4;
                for (uint32_t k = tid; k < cnt; k += NT) {
                    const uint32_t i = rd_entry(spm, e0 + 1 + k, d.aet);
                    const uint32_t li = i - elem0;
                    if (li < (uint32_t)D5_TILE && i < n) tile[li] = (OT)(spv | ((int32_t)(i & 1u) & DP));
                }
            }
            fence_proxy_async();
            __syncthreads();
            if (tid == 0) {
                const uint32_t bytes = min((uint32_t)D5_TILE, n - elem0) * (uint32_t)sizeof(OT);
                bulk_s2g(out + elem0, tile, bytes);
                bulk_commit();
            }
            buf ^= 1u;
        }
    }
    if (tid == 0) bulk_wait<0>();
}

}  // namespace xsi
