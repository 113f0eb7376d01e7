// encode_kernels.cuh -- sm_100a kernels of the encode path.
//
//   E1 scan_rows        gt rows -> per-ALT bit-rows (natural order), counts, flags, WAH/sparse decision
//                        replaces GtBlock::scan_genotypes + the decision in encode_line
//                        (reference include/gt_block.hpp:207-269, 292-329); HBM-bound, 4 B/genotype
//   E2 build_wah_lists  per block: ordered list of the lines that are PBWT+WAH encoded
//   E3 pbwt_permute     per block, sequential over its WAH lines: y[j] = bit[a[j]], a <- stable
//                        partition of a by y  (wah.hpp:506-578 gather + internal_gt_record.hpp:32-59)
//   E4 wah_encode_rows  bit-row -> WAH2-16 words (wah.hpp:376-429 process_wah_word rules), warp per line
//   E5 sparse_emit      bit-row -> [count|MSB][ascending indices]  (block.hpp:54-99)
//   E6 pack_wah         gather the per-line WAH words into the contiguous per-block matrix
//   scan_u32            exclusive prefix sums for the output offsets
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace xsi {

struct EncDev {
    const void* gt;
    const uint64_t* rec_goff;     // [R] element offset of the row
    const uint32_t* rec_ngt;      // [R]
    const uint32_t* rec_nallele;  // [R]
    const uint32_t* rec_line0;    // [R] first binary line (batch-global)
    const uint32_t* line_rec;     // [L]
    const uint32_t* blk_line0;    // [nb+1]
    uint32_t n_samples, R, L, nb;
    uint32_t WS;     // words per bit-row (multiple of 4)
    uint32_t SLOTW;  // u16 words per WAH slot
    uint32_t aux_cap, phase_cap;
    uint64_t mac_thr;
    int32_t default_phasing;
    uint32_t* bitrows;  // [L][WS]
    uint32_t* auxrows;  // [aux_cap][WS]   missing / end-of-vector rows
    uint32_t* phrows;   // [phase_cap][WS] non-default-phase rows
    uint32_t* counters; // [0] aux rows used, [1] phase rows used, [2] error bits
    int32_t* rec_aux;   // [R][3] slot of missing / eov / phase row, -1 = none
    uint32_t* line_cnt;       // [L] carriers of the line's ALT allele
    uint8_t* line_flags;      // [L]
    uint32_t* line_sparse_n;  // [L] A_T entries of the sparse line incl. header (0 for WAH lines)
    uint32_t* line_wah_n;     // [L] WAH words (0 for sparse lines)
    uint8_t* rec_flags;       // [R]
    uint32_t* rec_miss_n;     // [R] entries incl. header, 0 if none
    uint32_t* rec_eov_n;      // [R]
    uint32_t* rec_phase_n;    // [R] WAH words of the phase line
    uint32_t* wah_list;       // [L] per block (at blk_line0[b]) the WAH lines in order
    uint32_t* blk_nwah;       // [nb]
    uint16_t* wahslots;       // [L][SLOTW]
    uint16_t* phslots;        // [phase_cap][SLOTW]
    // --wah-encode-missing (WS_WAH): WAH of the missing / end-of-vector rows (natural order)
    uint32_t wah_missing;     // 0 / 1
    uint16_t* auxslots;       // [aux_cap][SLOTW]
    uint32_t* rec_missw_n;    // [R] WAH words of the record's missing line (0 if none)
    uint32_t* rec_eovw_n;     // [R]
    // PBWT kernels only: the blocks one launch works on (nb of them), or NULL for blocks 0..nb-1.  Blocks with an
    // all-haploid record go to the general kernel, the others to the cluster kernel, in the same batch.
    const uint32_t* blk_map;
};

#define ERR_ALLELE 1u
#define ERR_AUX_OVERFLOW 2u
#define ERR_PHASE_OVERFLOW 4u

// =============================================================================================
// E1
// =============================================================================================
constexpr int E1_THREADS = 256;
constexpr int E1_WARPS = E1_THREADS / 32;
constexpr int E1_MAXALLELE = 256;

// one predicate row: kind 0 allele==key, 1 missing, 2 end-of-vector, 3 non-default phase
template <int ELEM, int KIND>
__device__ __forceinline__ void emit_pred_row(const void* gt, uint64_t goff, uint32_t ngt, uint32_t WS, int32_t key,
                                              uint32_t* __restrict__ dst) {
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t nwords = (ngt + 31) >> 5, nchunks = (WS + 31) >> 5;
    for (uint32_t c = warp; c < nchunks; c += (blockDim.x >> 5)) {
        uint32_t keep = 0;
#pragma unroll 4
        for (uint32_t s = 0; s < 32; ++s) {
            const uint32_t widx = c * 32 + s;
            if (widx >= nwords) break;
            const uint32_t i = widx * 32 + lane;
            bool pred = false;
            if (i < ngt) {
                const int32_t v = load_gt<ELEM>(gt, goff + i);
                if (KIND == 0) pred = ((v >> 1) - 1) == key;
                else if (KIND == 1) pred = gt_is_missing(v);
                else if (KIND == 2) pred = (v == XSI_I32_VECTOR_END);
                else pred = (i & 1) && ((v & 1) != key);
            }
            const uint32_t b = __ballot_sync(XSI_FULL, pred);
            if (lane == s) keep = b;
        }
        const uint32_t w = c * 32 + lane;
        if (w < WS) dst[w] = keep;
    }
}

// Same rows when the record starts and ends on 16-byte boundaries: every thread builds whole words from its own
// 32 consecutive genotypes with 16-byte loads that are all in flight at once (the row is L2-hot).  The ballot
// version above chains 32 dependent loads per 1024 genotypes: 8 us per row at 5,008 haplotypes, which is what a
// chrX-shaped file pays three times per record (end-of-vector, missing, phase).
template <int ELEM, int KIND>
__device__ __forceinline__ void emit_pred_row_vec(const void* gt, uint64_t goff, uint32_t ngt, uint32_t WS, int32_t key,
                                                  uint32_t* __restrict__ dst) {
    auto pred_of = [&](int32_t v, uint32_t i) -> uint32_t {
        if (KIND == 0) return (uint32_t)(((v >> 1) - 1) == key);
        if (KIND == 1) return (uint32_t)gt_is_missing(v);
        if (KIND == 2) return (uint32_t)(v == XSI_I32_VECTOR_END);
        return (uint32_t)((i & 1u) && ((v & 1) != key));
    };
    for (uint32_t w = threadIdx.x; w < WS; w += blockDim.x) {
        const uint32_t i0 = w * 32;
        uint32_t word = 0;
        if (i0 < ngt) {
            if (ELEM == 4) {
                const int4* src = reinterpret_cast<const int4*>(reinterpret_cast<const int32_t*>(gt) + goff + i0);
                int4 v[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) v[k] = (i0 + 4u * k < ngt) ? __ldg(src + k) : make_int4(0, 0, 0, 0);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if (i0 + 4u * k < ngt) {  // a row that ends on a 16-byte boundary has whole quads only
                        word |= pred_of(v[k].x, 4 * k) << (4 * k);
                        word |= pred_of(v[k].y, 4 * k + 1) << (4 * k + 1);
                        word |= pred_of(v[k].z, 4 * k + 2) << (4 * k + 2);
                        word |= pred_of(v[k].w, 4 * k + 3) << (4 * k + 3);
                    }
                }
            } else {
                const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const signed char*>(gt) + goff + i0);
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (i0 + 16u * k < ngt) {
                        const uint4 q = __ldg(src + k);
                        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int b = 0; b < 16; ++b) {
                            const int32_t sb = (int32_t)(signed char)((u[b >> 2] >> (8 * (b & 3))) & 0xFFu);
                            const int32_t v = sb == -128 ? XSI_I32_MISSING : (sb == -127 ? XSI_I32_VECTOR_END : sb);
                            word |= pred_of(v, 16 * k + b) << (16 * k + b);
                        }
                    }
                }
            }
        }
        dst[w] = word;
    }
}

// Per-record decisions and the rare extra rows, shared by both scan kernels.  Expects the block's
// s_cnt[alt] / s_misc {nmiss, neov, phase bits, err} to be complete (barrier before the call).
template <int ELEM>
__device__ __forceinline__ void finish_record(const EncDev& p, uint32_t r, uint32_t ngt, uint32_t n_allele, uint32_t line0,
                                              uint64_t goff, uint32_t P, uint32_t* s_cnt, uint8_t* s_lflag,
                                              uint32_t* s_misc, int32_t* s_slot, bool aligned16 = false, bool aux_done = false) {
    // ---- per-record decisions (gt_block.hpp:292-338) ----
    if (threadIdx.x == 0) {
        const uint32_t nmiss = s_misc[0], neov = s_misc[1];
        uint8_t rf = 0;
        if (nmiss) rf |= RF_MISSING;
        if (neov) rf |= RF_EOV;
        if (s_misc[2]) rf |= RF_PHASE;
        if (P == 1) rf |= RF_HAPLOID;
        if (s_misc[3]) atomicOr(&p.counters[2], ERR_ALLELE);
        uint32_t sum_alt = 0;
        for (uint32_t a = 1; a < n_allele; ++a) sum_alt += s_cnt[a];
        const uint32_t cnt0 = ngt - sum_alt - nmiss - neov;
        for (uint32_t a = 1; a < n_allele; ++a) {
            const uint32_t c = s_cnt[a];
            const uint32_t mac = min(c, ngt - c);
            const uint32_t line = line0 + a - 1;
            uint8_t lf = (P == 1) ? LF_HAPLOID : 0;
            uint32_t sn = 0;
            if ((uint64_t)mac > p.mac_thr) lf |= LF_WAH;
            else if (c == mac) sn = c + 1;
            else { lf |= LF_NEGATED; sn = cnt0 + 1; }
            p.line_cnt[line] = c;
            p.line_flags[line] = lf;
            p.line_sparse_n[line] = sn;
            p.line_wah_n[line] = 0;
            s_lflag[a] = lf;
        }
        p.rec_flags[r] = rf;
        p.rec_miss_n[r] = nmiss ? nmiss + 1 : 0;
        p.rec_eov_n[r] = neov ? neov + 1 : 0;
        p.rec_phase_n[r] = 0;
        s_slot[0] = s_slot[1] = s_slot[2] = -1;
        if (nmiss) { const uint32_t s = atomicAdd(&p.counters[0], 1u); if (s < p.aux_cap) s_slot[0] = (int32_t)s; else atomicOr(&p.counters[2], ERR_AUX_OVERFLOW); }
        if (neov) { const uint32_t s = atomicAdd(&p.counters[0], 1u); if (s < p.aux_cap) s_slot[1] = (int32_t)s; else atomicOr(&p.counters[2], ERR_AUX_OVERFLOW); }
        if (rf & RF_PHASE) { const uint32_t s = atomicAdd(&p.counters[1], 1u); if (s < p.phase_cap) s_slot[2] = (int32_t)s; else atomicOr(&p.counters[2], ERR_PHASE_OVERFLOW); }
        p.rec_aux[r * 3 + 0] = s_slot[0];
        p.rec_aux[r * 3 + 1] = s_slot[1];
        p.rec_aux[r * 3 + 2] = s_slot[2];
    }
    __syncthreads();
    // ---- rare second passes over the (L2-hot) row ----
    for (uint32_t a = 1; a < n_allele; ++a)
        if (s_lflag[a] & LF_NEGATED)  // negated sparse lists REF carriers (block.hpp:59-65 with sparse_allele 0)
            emit_pred_row<ELEM, 0>(p.gt, goff, ngt, p.WS, 0, p.bitrows + (size_t)(line0 + a - 1) * p.WS);
    if (aux_done) return;  // the caller holds the missing / end-of-vector / phase words of a one-tile row in registers
    if (aligned16) {
        if (s_slot[0] >= 0) emit_pred_row_vec<ELEM, 1>(p.gt, goff, ngt, p.WS, 0, p.auxrows + (size_t)s_slot[0] * p.WS);
        if (s_slot[1] >= 0) emit_pred_row_vec<ELEM, 2>(p.gt, goff, ngt, p.WS, 0, p.auxrows + (size_t)s_slot[1] * p.WS);
        if (s_slot[2] >= 0) emit_pred_row_vec<ELEM, 3>(p.gt, goff, ngt, p.WS, p.default_phasing, p.phrows + (size_t)s_slot[2] * p.WS);
    } else {
        if (s_slot[0] >= 0) emit_pred_row<ELEM, 1>(p.gt, goff, ngt, p.WS, 0, p.auxrows + (size_t)s_slot[0] * p.WS);
        if (s_slot[1] >= 0) emit_pred_row<ELEM, 2>(p.gt, goff, ngt, p.WS, 0, p.auxrows + (size_t)s_slot[1] * p.WS);
        if (s_slot[2] >= 0) emit_pred_row<ELEM, 3>(p.gt, goff, ngt, p.WS, p.default_phasing, p.phrows + (size_t)s_slot[2] * p.WS);
    }
}

template <int ELEM>
__global__ void __launch_bounds__(E1_THREADS) scan_rows_kernel(EncDev p) {
    __shared__ uint32_t s_cnt[E1_MAXALLELE];
    __shared__ uint8_t s_lflag[E1_MAXALLELE];
    __shared__ uint32_t s_misc[4];  // nmiss, neov, phase bits, err
    __shared__ int32_t s_slot[3];
    const uint32_t r = blockIdx.x;
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
    const uint32_t ngt = p.rec_ngt[r], n_allele = p.rec_nallele[r], line0 = p.rec_line0[r];
    const uint64_t goff = p.rec_goff[r];
    const uint32_t P = p.n_samples ? ngt / p.n_samples : 0;
    const uint32_t nwords = (ngt + 31) >> 5, nchunks = (p.WS + 31) >> 5;
    const uint32_t dpmask = p.default_phasing ? 0xFFFFFFFFu : 0u;
    const int32_t filler = 2 | (p.default_phasing & 1);  // REF with default phase: matches nothing
    for (uint32_t i = threadIdx.x; i < E1_MAXALLELE; i += E1_THREADS) { s_cnt[i] = 0; s_lflag[i] = 0; }
    if (threadIdx.x < 4) s_misc[threadIdx.x] = 0;
    if (threadIdx.x < 3) s_slot[threadIdx.x] = -1;
    __syncthreads();

    uint32_t w_nmiss = 0, w_neov = 0, w_phase = 0, w_err = 0;
    uint32_t alt0 = 1;
    do {
        const uint32_t nal = n_allele > alt0 ? min(4u, n_allele - alt0) : 0u;
        uint32_t wc0 = 0, wc1 = 0, wc2 = 0, wc3 = 0;
        for (uint32_t c = warp; c < nchunks; c += E1_WARPS) {
            uint32_t k0 = 0, k1 = 0, k2 = 0, k3 = 0;
#pragma unroll 8
            for (uint32_t s = 0; s < 32; ++s) {
                const uint32_t widx = c * 32 + s;
                if (widx >= nwords) break;
                const uint32_t i = widx * 32 + lane;
                const bool valid = i < ngt;
                const int32_t v = valid ? load_gt<ELEM>(p.gt, goff + i) : filler;
                const int32_t h = v >> 1;
                if (alt0 == 1) {
                    const bool bad = (uint32_t)(h - 1) >= n_allele;  // missing / EOV / unknown allele
                    if (__any_sync(XSI_FULL, bad)) {
                        const bool miss = gt_is_missing(v);
                        const bool eov = (v == XSI_I32_VECTOR_END);
                        w_nmiss += __popc(__ballot_sync(XSI_FULL, miss));
                        w_neov += __popc(__ballot_sync(XSI_FULL, eov));
                        w_err |= __any_sync(XSI_FULL, bad && !miss && !eov) ? 1u : 0u;
                    }
                    if (P == 2) w_phase |= (__ballot_sync(XSI_FULL, v & 1) ^ dpmask) & 0xAAAAAAAAu;
                }
                if (nal > 0) { const uint32_t b = __ballot_sync(XSI_FULL, h == (int32_t)(alt0 + 1)); wc0 += __popc(b); if (lane == s) k0 = b; }
                if (nal > 1) { const uint32_t b = __ballot_sync(XSI_FULL, h == (int32_t)(alt0 + 2)); wc1 += __popc(b); if (lane == s) k1 = b; }
                if (nal > 2) { const uint32_t b = __ballot_sync(XSI_FULL, h == (int32_t)(alt0 + 3)); wc2 += __popc(b); if (lane == s) k2 = b; }
                if (nal > 3) { const uint32_t b = __ballot_sync(XSI_FULL, h == (int32_t)(alt0 + 4)); wc3 += __popc(b); if (lane == s) k3 = b; }
            }
            const uint32_t w = c * 32 + lane;
            if (w < p.WS) {
                uint32_t* row = p.bitrows + (size_t)(line0 + alt0 - 1) * p.WS + w;
                if (nal > 0) row[0] = k0;
                if (nal > 1) row[(size_t)p.WS] = k1;
                if (nal > 2) row[(size_t)p.WS * 2] = k2;
                if (nal > 3) row[(size_t)p.WS * 3] = k3;
            }
        }
        if (lane == 0) {
            if (nal > 0) atomicAdd(&s_cnt[alt0], wc0);
            if (nal > 1) atomicAdd(&s_cnt[alt0 + 1], wc1);
            if (nal > 2) atomicAdd(&s_cnt[alt0 + 2], wc2);
            if (nal > 3) atomicAdd(&s_cnt[alt0 + 3], wc3);
        }
        alt0 += 4;
    } while (alt0 < n_allele);
    if (lane == 0) {
        if (w_nmiss) atomicAdd(&s_misc[0], w_nmiss);
        if (w_neov) atomicAdd(&s_misc[1], w_neov);
        if (w_phase) atomicOr(&s_misc[2], w_phase);
        if (w_err) atomicOr(&s_misc[3], 1u);
    }
    __syncthreads();

    finish_record<ELEM>(p, r, ngt, n_allele, line0, goff, P, s_cnt, s_lflag, s_misc, s_slot);
}

// =============================================================================================
// E1 v2: the same scan fed by the TMA engine.  Persistent CTAs (2 per SM) walk records
// r = blockIdx.x, blockIdx.x + gridDim.x, ...; the row streams through a ring of 8192-genotype
// tiles (cp.async.bulk global->shared, full/empty mbarriers) and every thread turns its own 32
// consecutive genotypes into one word per ALT line with no cross-lane traffic: 16-byte
// conflict-free shared loads (chunk order rotated by the thread index), compile-time bit
// positions, one funnel shift to undo the rotation, coalesced 128-byte word stores.
// Needs 16-byte aligned rows (host checks; otherwise scan_rows_kernel runs).
// =============================================================================================
// genotypes per tile = NT threads x 32.  NT = 256 for long rows; rows of at most 224 words (7,168 genotypes: the 1KGP3 /
// chrX widths) run with NT = the row's words rounded up to a warp, so that no thread idles through a tile (157 of 256
// worked at 5,008 haplotypes) and more, smaller CTAs share an SM and cover each other's per-record bookkeeping.
__host__ __device__ constexpr int s2_tile(int nt) { return nt * 32; }
// ring depth: ~64-96 KB of bulk copies in flight per CTA for long rows whatever the element size (8 KB int8 tiles need a
// deeper ring); short rows are one tile per record, two stages prefetch the next record
__host__ __device__ constexpr int s2_stages(int elem, int nt = 256) { return elem == 4 ? (nt < 256 ? 2 : 3) : (nt < 256 ? 4 : 8); }
__host__ __device__ constexpr int s2_ctas_per_sm(int nt) { return nt >= 256 ? 2 : (nt >= 192 ? 3 : (nt >= 160 ? 5 : 6)); }

template <int ELEM>
__device__ __forceinline__ int32_t tile_value(const unsigned char* base, uint32_t i) {
    if (ELEM == 4) return reinterpret_cast<const int32_t*>(base)[i];
    return (int32_t)reinterpret_cast<const signed char*>(base)[i];
}
// classify one genotype that failed the allele range test: bit0 missing, bit1 end of vector, bit2 unknown allele
template <int ELEM>
__device__ __forceinline__ uint32_t classify_bad(int32_t v) {
    const bool miss = ELEM == 4 ? gt_is_missing(v) : (((v >> 1) == 0) || v == -128);
    const bool eov = ELEM == 4 ? (v == XSI_I32_VECTOR_END) : (v == -127);
    return miss ? 1u : (eov ? 2u : 4u);
}

struct ScanAcc { uint32_t nmiss, neov, phase, err; };

// One thread, one tile: the thread's 32 genotypes -> one word per ALT allele key-1 .. key-1+NAL-1
// (key = (v >> 1) of the first ALT of the group).  FIRST also runs the per-record checks
// (missing / end of vector / unknown allele, non-default phase).  `tile` points at the thread's
// 32 genotypes in shared memory; nvalid < 32 only for the last, partially filled word of a row.
template <int ELEM, int NAL, bool FIRST>
__device__ __forceinline__ void scan_words(const unsigned char* __restrict__ tile, uint32_t nvalid, uint32_t rot, int32_t key,
                                           uint32_t n_allele, int32_t dpx, uint32_t (&w)[4], ScanAcc& acc) {
    constexpr uint32_t NCH = 32 * ELEM / 16, EPC = 16 / ELEM;
    if (nvalid == 32) {
#pragma unroll
        for (uint32_t k = 0; k < NCH; ++k) {
            const uint32_t c = (k + rot) & (NCH - 1);
            const int4 q = *reinterpret_cast<const int4*>(tile + c * 16);
            int32_t v[EPC];
            if (ELEM == 4) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
            else {
                const uint32_t qq[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
#pragma unroll
                for (uint32_t e = 0; e < EPC; ++e) v[e] = (int32_t)(int8_t)(qq[e >> 2] >> (8 * (e & 3)));
            }
            bool badc = false;
#pragma unroll
            for (uint32_t e = 0; e < EPC; ++e) {
                const int32_t h = v[e] >> 1;
                const uint32_t bitv = 1u << (k * EPC + e);
                if (h == key) w[0] |= bitv;
                if (NAL > 1 && h == key + 1) w[1] |= bitv;
                if (NAL > 2 && h == key + 2) w[2] |= bitv;
                if (NAL > 3 && h == key + 3) w[3] |= bitv;
                if (FIRST) {
                    badc |= (uint32_t)(h - 1) >= n_allele;
                    if (e & 1) acc.phase |= (uint32_t)(v[e] ^ dpx);
                }
            }
            if (FIRST && badc) {
#pragma unroll
                for (uint32_t e = 0; e < EPC; ++e) {
                    if ((uint32_t)((v[e] >> 1) - 1) >= n_allele) {
                        const uint32_t cl = classify_bad<ELEM>(v[e]);
                        acc.nmiss += cl & 1u; acc.neov += (cl >> 1) & 1u; acc.err |= cl >> 2;
                    }
                }
            }
        }
        // undo the chunk rotation: chunk k of w[] is genotype chunk (k + rot)
        const uint32_t sh = rot * EPC;
#pragma unroll
        for (int a = 0; a < NAL; ++a) w[a] = __funnelshift_l(w[a], w[a], sh);
    } else if (nvalid) {  // the one partially filled word of the row
        for (uint32_t e = 0; e < nvalid; ++e) {
            const int32_t v = tile_value<ELEM>(tile, e);
            const int32_t h = v >> 1;
            const uint32_t bitv = 1u << e;
            if (h == key) w[0] |= bitv;
            if (NAL > 1 && h == key + 1) w[1] |= bitv;
            if (NAL > 2 && h == key + 2) w[2] |= bitv;
            if (NAL > 3 && h == key + 3) w[3] |= bitv;
            if (FIRST) {
                if (e & 1) acc.phase |= (uint32_t)(v ^ dpx);
                if ((uint32_t)(h - 1) >= n_allele) {
                    const uint32_t cl = classify_bad<ELEM>(v);
                    acc.nmiss += cl & 1u; acc.neov += (cl >> 1) & 1u; acc.err |= cl >> 2;
                }
            }
        }
    }
}

// The missing / end-of-vector / non-default-phase words of a thread's 32 genotypes, from the tile that is still in shared
// memory (same predicates as emit_pred_row_vec).  Rows of one tile (1KGP3 / chrX widths) use this instead of three more passes
// over the row: a chrX-shaped file has end-of-vector entries, missing alleles and unphased genotypes in EVERY record.
template <int ELEM>
__device__ __forceinline__ void aux_words(const unsigned char* __restrict__ tile, uint32_t nvalid, uint32_t rot, int32_t dpx,
                                          uint32_t& wm, uint32_t& we, uint32_t& wp) {
    constexpr uint32_t NCH = 32 * ELEM / 16, EPC = 16 / ELEM;
    wm = 0; we = 0; wp = 0;
#pragma unroll
    for (uint32_t k = 0; k < NCH; ++k) {
        const uint32_t c = (k + rot) & (NCH - 1);
        const int4 q = *reinterpret_cast<const int4*>(tile + c * 16);
        int32_t v[EPC];
        if (ELEM == 4) { v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
        else {
            const uint32_t qq[4] = {(uint32_t)q.x, (uint32_t)q.y, (uint32_t)q.z, (uint32_t)q.w};
#pragma unroll
            for (uint32_t e = 0; e < EPC; ++e) {
                const int32_t sb = (int32_t)(int8_t)(qq[e >> 2] >> (8 * (e & 3)));
                v[e] = sb == -128 ? XSI_I32_MISSING : (sb == -127 ? XSI_I32_VECTOR_END : sb);
            }
        }
#pragma unroll
        for (uint32_t e = 0; e < EPC; ++e) {
            const uint32_t idx = c * EPC + e;
            if (idx < nvalid) {
                const uint32_t bitv = 1u << idx;
                if (gt_is_missing(v[e])) wm |= bitv;
                if (v[e] == XSI_I32_VECTOR_END) we |= bitv;
                if ((e & 1u) && ((v[e] & 1) != dpx)) wp |= bitv;
            }
        }
    }
}

template <int ELEM, int NT>
__global__ void __launch_bounds__(NT, s2_ctas_per_sm(NT)) scan_rows_v2_kernel(EncDev p) {
    constexpr uint32_t S2_TILE = s2_tile(NT);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t s_cnt[E1_MAXALLELE];
    __shared__ uint8_t s_lflag[E1_MAXALLELE];
    __shared__ uint32_t s_misc[4];
    __shared__ int32_t s_slot[3];
    // narrow CTAs (one tile per record): totals of the record, triple buffered so that a common record costs ONE barrier
    // ([0..3] carriers of ALT 1..4, [4] missing, [5] end of vector, [6] non-default phase seen, [7] unknown allele seen)
    __shared__ uint32_t s_tot[3][8];
    constexpr uint32_t TILE_BYTES = S2_TILE * ELEM;
    constexpr uint32_t S2_STAGES = s2_stages(ELEM, NT);
    constexpr uint32_t TBYTES = 32 * ELEM;              // bytes of one thread's 32 genotypes
    constexpr uint32_t NCH = TBYTES / 16;               // 16-byte chunks per thread: 8 (int32) / 2 (int8)
    constexpr uint32_t EPC = 16 / ELEM;                 // genotypes per chunk: 4 / 16
    unsigned char* ring = smem_raw;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)S2_STAGES * TILE_BYTES);
    uint64_t* empty = full + S2_STAGES;
    const uint32_t tid = threadIdx.x, lane = lane_id();
    const uint32_t tiles_per_rec = (p.WS + NT - 1) / NT;
    const unsigned char* gbase = reinterpret_cast<const unsigned char*>(p.gt);
    const uint32_t rot = ELEM == 4 ? (tid & 7u) : ((tid >> 2) & 1u);  // chunk rotation of this thread
    if (tid == 0) {
        for (uint32_t s = 0; s < S2_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NT / 32); }
        fence_proxy_async();
    }
    __syncthreads();
    // ---- producer state (thread 0): next data tile to request ----
    uint32_t pr = blockIdx.x, pt = 0, issued = 0;
    auto produce = [&]() {
        while (pr < p.R) {
            const uint32_t ngt = p.rec_ngt[pr];
            const uint32_t ndt = (ngt + S2_TILE - 1) / S2_TILE;
            if (pt < ndt) {
                const uint32_t st = issued % S2_STAGES, use = issued / S2_STAGES;
                if (use > 0) mbar_wait(&empty[st], (use - 1) & 1u);
                const uint32_t bytes = min(TILE_BYTES, (ngt - pt * S2_TILE) * ELEM);
                mbar_expect_tx(&full[st], bytes);
                bulk_g2s(ring + (size_t)st * TILE_BYTES, gbase + (p.rec_goff[pr] + (uint64_t)pt * S2_TILE) * ELEM, bytes, &full[st]);
                ++issued; ++pt;
                return;
            }
            pr += gridDim.x; pt = 0;
        }
    };
    if (tid == 0) for (uint32_t s = 0; s + 1 < S2_STAGES; ++s) produce();
    uint32_t consumed = 0;
    const int32_t dpx = p.default_phasing & 1;

    // record descriptors one record ahead: their L2 latency (four loads) is otherwise exposed once per record, which
    // is most of the per-record cost for short rows (1KGP3 shape: 20 KB per record)
    uint32_t nx_ngt = 0, nx_nall = 0, nx_line0 = 0;
    uint64_t nx_goff = 0;
    if (blockIdx.x < p.R) { nx_ngt = p.rec_ngt[blockIdx.x]; nx_nall = p.rec_nallele[blockIdx.x]; nx_line0 = p.rec_line0[blockIdx.x]; nx_goff = p.rec_goff[blockIdx.x]; }
    for (uint32_t i = tid; i < E1_MAXALLELE; i += NT) { s_cnt[i] = 0; s_lflag[i] = 0; }
    if (tid < 4) s_misc[tid] = 0;
    if (tid < 24) (&s_tot[0][0])[tid] = 0;
    __syncthreads();
    uint32_t it = 0;  // records this CTA has taken
    for (uint32_t r = blockIdx.x; r < p.R; r += gridDim.x, ++it) {
        const uint32_t ngt = nx_ngt, n_allele = nx_nall, line0 = nx_line0;
        const uint64_t goff = nx_goff;
        if (r + gridDim.x < p.R) {
            const uint32_t rn = r + gridDim.x;
            nx_ngt = __ldg(p.rec_ngt + rn); nx_nall = __ldg(p.rec_nallele + rn); nx_line0 = __ldg(p.rec_line0 + rn); nx_goff = __ldg(p.rec_goff + rn);
        }
        const uint32_t P = p.n_samples ? ngt / p.n_samples : 0;
        const uint32_t ndt = (ngt + S2_TILE - 1) / S2_TILE;
        uint32_t c0 = 0, c1 = 0, c2 = 0, c3 = 0, nmiss = 0, neov = 0, phase = 0, err = 0;
        uint32_t am = 0, ae = 0, ap = 0;  // one-tile rows: this thread's missing / end-of-vector / phase words
        // rows of the narrow CTAs are one tile long by construction (the host picks NT >= words of the row); the 256-thread
        // kernel streams long rows and keeps the second passes (and its register count: fusing cost it 15% at the HRC width)
        constexpr bool fuse_aux = NT < 256;
        if (fuse_aux && n_allele > 5) {
            // a narrow CTA on a record with more than 4 ALT alleles takes the general epilogue below; the records before it may
            // have taken the one-barrier path, which does not touch s_cnt / s_misc: clear them here
            for (uint32_t i = tid; i <= n_allele && i < E1_MAXALLELE; i += NT) s_cnt[i] = 0;
            if (tid < 4) s_misc[tid] = 0;
            __syncthreads();
            if (tid < 8) s_tot[(it + 2u) % 3u][tid] = 0;  // what the one-barrier path would have cleared during this record
        }
        for (uint32_t tt = 0; tt < tiles_per_rec; ++tt) {
            const uint32_t wi = tt * NT + tid;
            const uint32_t elem0 = tt * S2_TILE + tid * 32;
            const bool data = tt < ndt;
            uint32_t st = 0;
            if (data) {
                if (tid == 0) produce();  // refill the stage the previous tile released
                st = consumed % S2_STAGES;
                mbar_wait(&full[st], (consumed / S2_STAGES) & 1u);
            }
            const unsigned char* tile = ring + (size_t)st * TILE_BYTES + tid * TBYTES;
            const uint32_t nvalid = (data && ngt > elem0) ? min(32u, ngt - elem0) : 0u;
            for (uint32_t alt0 = 1; alt0 < n_allele || alt0 == 1; alt0 += 4) {
                const uint32_t nal = n_allele > alt0 ? min(4u, n_allele - alt0) : 0u;
                uint32_t w[4] = {0, 0, 0, 0};
                ScanAcc acc = {nmiss, neov, phase, err};
                if (alt0 == 1) {
                    if (nal <= 1) scan_words<ELEM, 1, true>(tile, nvalid, rot, (int32_t)alt0 + 1, n_allele, dpx, w, acc);
                    else if (nal == 2) scan_words<ELEM, 2, true>(tile, nvalid, rot, (int32_t)alt0 + 1, n_allele, dpx, w, acc);
                    else scan_words<ELEM, 4, true>(tile, nvalid, rot, (int32_t)alt0 + 1, n_allele, dpx, w, acc);
                } else {
                    scan_words<ELEM, 4, false>(tile, nvalid, rot, (int32_t)alt0 + 1, n_allele, dpx, w, acc);
                }
                if (fuse_aux && alt0 == 1 && ((acc.nmiss | acc.neov | (acc.phase & 1u)) != 0u)) aux_words<ELEM>(tile, nvalid, rot, dpx, am, ae, ap);
                nmiss = acc.nmiss; neov = acc.neov; phase = acc.phase; err = acc.err;
                const uint32_t w0 = w[0], w1 = nal > 1 ? w[1] : 0u, w2 = nal > 2 ? w[2] : 0u, w3 = nal > 3 ? w[3] : 0u;
                if (wi < p.WS) {
                    uint32_t* row = p.bitrows + (size_t)(line0 + alt0 - 1) * p.WS + wi;
                    if (nal > 0) row[0] = w0;
                    if (nal > 1) row[(size_t)p.WS] = w1;
                    if (nal > 2) row[(size_t)p.WS * 2] = w2;
                    if (nal > 3) row[(size_t)p.WS * 3] = w3;
                }
                if (alt0 == 1) { c0 += __popc(w0); c1 += __popc(w1); c2 += __popc(w2); c3 += __popc(w3); }
                else if (nal > 0) {  // rare: more than 4 ALT alleles
                    const uint32_t a0 = __reduce_add_sync(XSI_FULL, __popc(w0)), a1 = __reduce_add_sync(XSI_FULL, __popc(w1));
                    const uint32_t a2 = __reduce_add_sync(XSI_FULL, __popc(w2)), a3 = __reduce_add_sync(XSI_FULL, __popc(w3));
                    if (lane == 0) {
                        atomicAdd(&s_cnt[alt0], a0);
                        if (nal > 1) atomicAdd(&s_cnt[alt0 + 1], a1);
                        if (nal > 2) atomicAdd(&s_cnt[alt0 + 2], a2);
                        if (nal > 3) atomicAdd(&s_cnt[alt0 + 3], a3);
                    }
                }
            }
            if (data) {
                __syncwarp();
                if (lane == 0) mbar_arrive(&empty[st]);
                ++consumed;
            }
        }
        // ---- block totals ----
        // the common record (one ALT, nothing missing, default phase) needs one sum and one "anything odd?" vote
        c0 = __reduce_add_sync(XSI_FULL, c0);
        if (n_allele > 2) { c1 = __reduce_add_sync(XSI_FULL, c1); c2 = __reduce_add_sync(XSI_FULL, c2); c3 = __reduce_add_sync(XSI_FULL, c3); }
        if (__reduce_or_sync(XSI_FULL, nmiss | neov | (phase & 1u) | err)) {
            nmiss = __reduce_add_sync(XSI_FULL, nmiss); neov = __reduce_add_sync(XSI_FULL, neov);
            phase = __reduce_or_sync(XSI_FULL, phase & 1u); err = __reduce_or_sync(XSI_FULL, err);
        } else { nmiss = 0; neov = 0; phase = 0; err = 0; }
        // ---- narrow CTAs, at most 4 ALT alleles: one barrier per record ----
        // Short rows spend their time on the record epilogue (25% barrier stalls at 5,008 haplotypes, ncu r02e): three CTA
        // barriers and thread 0's decisions between two of them.  Here the totals go to s_tot[it % 3]; after the one barrier
        // EVERY thread derives the (cheap) per-line decisions from them, thread 0 alone writes them out while the other warps
        // are already in the next record, and the buffer of the record after next is cleared (its last readers passed this
        // barrier, its next writers come after the next one).  Negated lines and missing / end-of-vector / phase rows are
        // rare and pay their own barrier.
        const uint32_t kb = it % 3u;
        if (fuse_aux && n_allele <= 5) {
            uint32_t* tot = s_tot[kb];
            if (lane == 0) {
                if (c0) atomicAdd(&tot[0], c0);
                if (n_allele > 2) { if (c1) atomicAdd(&tot[1], c1); if (c2) atomicAdd(&tot[2], c2); if (c3) atomicAdd(&tot[3], c3); }
                if (nmiss) atomicAdd(&tot[4], nmiss);
                if (neov) atomicAdd(&tot[5], neov);
                if (phase && P == 2) atomicOr(&tot[6], 1u);
                if (err) atomicOr(&tot[7], 1u);
            }
            __syncthreads();
            const uint32_t T[4] = {tot[0], tot[1], tot[2], tot[3]};
            const uint32_t tm = tot[4], te = tot[5], tp = tot[6], terr = tot[7];
            if (tid < 8) s_tot[(kb + 2u) % 3u][tid] = 0;
            const uint32_t cnt0 = ngt - (T[0] + T[1] + T[2] + T[3]) - tm - te;
            uint32_t negmask = 0;
#pragma unroll
            for (uint32_t a = 1; a < 5; ++a) {
                if (a < n_allele) {
                    const uint32_t c = T[a - 1], mac = min(c, ngt - c);
                    if (!((uint64_t)mac > p.mac_thr) && c != mac) negmask |= 1u << a;
                }
            }
            const bool aux = (tm | te | tp) != 0;
            if (tid == 0) {  // per-record decisions (gt_block.hpp:292-338), same values as finish_record
                uint8_t rf = 0;
                if (tm) rf |= RF_MISSING;
                if (te) rf |= RF_EOV;
                if (tp) rf |= RF_PHASE;
                if (P == 1) rf |= RF_HAPLOID;
                if (terr) atomicOr(&p.counters[2], ERR_ALLELE);
#pragma unroll
                for (uint32_t a = 1; a < 5; ++a) {
                    if (a >= n_allele) break;
                    const uint32_t c = T[a - 1], mac = min(c, ngt - c), line = line0 + a - 1;
                    uint8_t lf = (P == 1) ? LF_HAPLOID : 0;
                    uint32_t sn = 0;
                    if ((uint64_t)mac > p.mac_thr) lf |= LF_WAH;
                    else if (c == mac) sn = c + 1;
                    else { lf |= LF_NEGATED; sn = cnt0 + 1; }
                    p.line_cnt[line] = c;
                    p.line_flags[line] = lf;
                    p.line_sparse_n[line] = sn;
                    p.line_wah_n[line] = 0;
                }
                p.rec_flags[r] = rf;
                p.rec_miss_n[r] = tm ? tm + 1 : 0;
                p.rec_eov_n[r] = te ? te + 1 : 0;
                p.rec_phase_n[r] = 0;
                int32_t sl[3] = {-1, -1, -1};
                if (tm) { const uint32_t q = atomicAdd(&p.counters[0], 1u); if (q < p.aux_cap) sl[0] = (int32_t)q; else atomicOr(&p.counters[2], ERR_AUX_OVERFLOW); }
                if (te) { const uint32_t q = atomicAdd(&p.counters[0], 1u); if (q < p.aux_cap) sl[1] = (int32_t)q; else atomicOr(&p.counters[2], ERR_AUX_OVERFLOW); }
                if (tp) { const uint32_t q = atomicAdd(&p.counters[1], 1u); if (q < p.phase_cap) sl[2] = (int32_t)q; else atomicOr(&p.counters[2], ERR_PHASE_OVERFLOW); }
                p.rec_aux[r * 3 + 0] = sl[0];
                p.rec_aux[r * 3 + 1] = sl[1];
                p.rec_aux[r * 3 + 2] = sl[2];
                if (aux) { s_slot[0] = sl[0]; s_slot[1] = sl[1]; s_slot[2] = sl[2]; }
            }
            if (negmask) {  // rare: negated sparse lists REF carriers (block.hpp:59-65 with sparse_allele 0); second pass over the L2-hot row
                __syncthreads();
                for (uint32_t a = 1; a < n_allele; ++a)
                    if (negmask & (1u << a)) emit_pred_row<ELEM, 0>(p.gt, goff, ngt, p.WS, 0, p.bitrows + (size_t)(line0 + a - 1) * p.WS);
            }
            if (aux) {  // the slots were handed out by thread 0
                __syncthreads();
                if (tid < p.WS) {
                    if (s_slot[0] >= 0) p.auxrows[(size_t)s_slot[0] * p.WS + tid] = am;
                    if (s_slot[1] >= 0) p.auxrows[(size_t)s_slot[1] * p.WS + tid] = ae;
                    if (s_slot[2] >= 0) p.phrows[(size_t)s_slot[2] * p.WS + tid] = ap;
                }
            }
            continue;
        }
        if (lane == 0) {
            if (n_allele > 1) atomicAdd(&s_cnt[1], c0);
            if (n_allele > 2) atomicAdd(&s_cnt[2], c1);
            if (n_allele > 3) atomicAdd(&s_cnt[3], c2);
            if (n_allele > 4) atomicAdd(&s_cnt[4], c3);
            if (nmiss) atomicAdd(&s_misc[0], nmiss);
            if (neov) atomicAdd(&s_misc[1], neov);
            if (phase && P == 2) atomicOr(&s_misc[2], 1u);
            if (err) atomicOr(&s_misc[3], 1u);
        }
        __syncthreads();
        finish_record<ELEM>(p, r, ngt, n_allele, line0, goff, P, s_cnt, s_lflag, s_misc, s_slot, true, fuse_aux);
        if (fuse_aux && tid < p.WS) {  // slots were handed out by thread 0 inside finish_record (barrier there)
            if (s_slot[0] >= 0) p.auxrows[(size_t)s_slot[0] * p.WS + tid] = am;
            if (s_slot[1] >= 0) p.auxrows[(size_t)s_slot[1] * p.WS + tid] = ae;
            if (s_slot[2] >= 0) p.phrows[(size_t)s_slot[2] * p.WS + tid] = ap;
        }
        // counters of the NEXT record (nobody reads these after the barrier inside finish_record): the barrier below then
        // serves both as the end of this record and as the start of the next one (one barrier less per record, which
        // is what short rows spend their time on: 39% barrier stalls at 5,008 haplotypes, ncu r01final)
        for (uint32_t i = tid; i <= nx_nall && i < E1_MAXALLELE; i += NT) s_cnt[i] = 0;
        if (tid < 4) s_misc[tid] = 0;
        __syncthreads();
    }
}

// =============================================================================================
// E2: ordered list of WAH lines per block
// =============================================================================================
__global__ void __launch_bounds__(32) build_wah_lists_kernel(EncDev p) {
    const uint32_t b = blockIdx.x, lane = lane_id();
    const uint32_t l0 = p.blk_line0[b], l1 = p.blk_line0[b + 1];
    uint32_t n = 0;
    for (uint32_t base = l0; base < l1; base += 32) {
        const uint32_t l = base + lane;
        const bool w = l < l1 && (p.line_flags[l] & LF_WAH);
        const uint32_t m = __ballot_sync(XSI_FULL, w);
        // bit 31 carries the haploid flag so the sequential kernel needs no extra load per line
        if (w) p.wah_list[l0 + n + __popc(m & lanemask_lt())] = l | ((p.line_flags[l] & LF_HAPLOID) ? 0x80000000u : 0u);
        n += __popc(m);
    }
    if (lane == 0) p.blk_nwah[b] = n;
}

// =============================================================================================
// E3: PBWT permute, a[] resident in shared memory as uint16 (2*n_samples <= 65536)
// =============================================================================================
// dynamic smem layout (bytes): a[2*NHpad] | row[2][WS*4] | ybuf[WS*4] | ebuf[WS*4] | zc[64*4] | mbar[2*8]
template <int WPW, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) pbwt_permute_smem_kernel(EncDev p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t N = 2 * p.n_samples;  // a.size(), gt_block.hpp:171
    const uint32_t W = (N + 31) >> 5;
    const uint32_t WS = p.WS;
    const uint32_t a_bytes = ((N * 2 + 15) / 16) * 16;
    uint16_t* a = reinterpret_cast<uint16_t*>(smem_raw);
    uint32_t* rowbuf = reinterpret_cast<uint32_t*>(smem_raw + a_bytes);
    uint32_t* ybuf = rowbuf + 2 * WS;
    uint32_t* ebuf = ybuf + WS;
    uint32_t* zc = ebuf + WS;  // [0..31] zeros per warp, [32..63] evens per warp
    uint64_t* mbar = reinterpret_cast<uint64_t*>(zc + 64);

    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5, NW = blockDim.x >> 5;
    const uint32_t b = p.blk_map ? p.blk_map[blockIdx.x] : blockIdx.x;
    const uint32_t l0 = p.blk_line0[b];
    const uint32_t nwah = p.blk_nwah[b];
    const uint32_t* list = p.wah_list + l0;
    const uint32_t row_bytes = WS * 4;

    for (uint32_t i = tid; i < N; i += blockDim.x) a[i] = (uint16_t)i;  // iota, gt_block.hpp:179
    if (tid == 0) { mbar_init(&mbar[0], 1); mbar_init(&mbar[1], 1); fence_proxy_async(); }
    __syncthreads();
    if (nwah == 0) return;
    uint32_t par0 = 0, par1 = 0;
    if (tid == 0) {
        mbar_expect_tx(&mbar[0], row_bytes);
        bulk_g2s(rowbuf, p.bitrows + (size_t)(list[0] & 0x7FFFFFFFu) * WS, row_bytes, &mbar[0]);
    }
    const uint32_t w0 = warp * WPW;
    const uint32_t ltm = lanemask_lt();
    uint32_t entry_cur = list[0], entry_next = nwah > 1 ? list[1] : 0;

    for (uint32_t k = 0; k < nwah; ++k) {
        const uint32_t cur = k & 1;
        const uint32_t entry = entry_cur;
        const uint32_t line = entry & 0x7FFFFFFFu;
        const bool hap = (entry >> 31) != 0;
        // prefetch: row of line k+1 into the other buffer, id of line k+2
        if (k + 1 < nwah) {
            if (tid == 0) {
                mbar_expect_tx(&mbar[cur ^ 1], row_bytes);
                bulk_g2s(rowbuf + (cur ^ 1) * WS, p.bitrows + (size_t)(entry_next & 0x7FFFFFFFu) * WS, row_bytes, &mbar[cur ^ 1]);
            }
        }
        entry_cur = entry_next;
        entry_next = (k + 2 < nwah) ? list[k + 2] : 0;
        if (cur == 0) { mbar_wait(&mbar[0], par0); par0 ^= 1; } else { mbar_wait(&mbar[1], par1); par1 ^= 1; }
        const uint32_t* row = rowbuf + cur * WS;

        // ---- phase A: gather y through a, count zeros ----
        uint32_t av[WPW / 2 > 0 ? WPW / 2 : 1];
        uint32_t zeros = 0, evens = 0;
#pragma unroll
        for (int q = 0; q < WPW; ++q) {
            const uint32_t widx = w0 + q;
            const uint32_t j = widx * 32 + lane;
            const bool valid = j < N;
            const uint32_t aj = valid ? a[j] : 0u;
            if (q & 1) av[q >> 1] |= aj << 16; else av[q >> 1] = aj;
            const uint32_t gi = hap ? (aj >> 1) : aj;
            const uint32_t bit = valid ? ((row[gi >> 5] >> (gi & 31)) & 1u) : 0u;
            const uint32_t yk = __ballot_sync(XSI_FULL, bit);
            const uint32_t vm = __ballot_sync(XSI_FULL, valid);
            zeros += __popc(~yk & vm);
            if (widx < WS && lane == 0) ybuf[widx] = yk;
            if (hap) {
                const uint32_t ek = __ballot_sync(XSI_FULL, valid && !(aj & 1u));
                evens += __popc(ek);
                if (widx < WS && lane == 0) ebuf[widx] = ek;
            }
        }
        if (lane == 0) { zc[warp] = zeros; zc[32 + warp] = evens; }
        __syncthreads();  // #1: all reads of a[] and row[] done, ybuf / zc complete

        // ---- phase B: offsets ----
        const uint32_t zv = lane < NW ? zc[lane] : 0u;
        const uint32_t Z = __reduce_add_sync(XSI_FULL, zv);
        uint32_t zbase = __reduce_add_sync(XSI_FULL, lane < warp ? zv : 0u);
        const uint32_t pos0 = min(w0 * 32, N);
        uint32_t obase = Z + (pos0 - zbase);
        uint32_t* grow = p.bitrows + (size_t)line * WS;
        uint32_t ebase = 0;
        uint32_t* obuf = rowbuf + cur * WS;  // haploid path reuses the consumed row buffer
        if (!hap) {
            for (uint32_t i = tid; i < WS; i += blockDim.x) grow[i] = i < W ? ybuf[i] : 0u;  // permuted row, in place
        } else {
            const uint32_t ev = lane < NW ? zc[32 + lane] : 0u;
            ebase = __reduce_add_sync(XSI_FULL, lane < warp ? ev : 0u);
            for (uint32_t i = tid; i < WS; i += blockDim.x) obuf[i] = 0u;
            __syncthreads();
        }
        // ---- phase C: stable partition of a by y (zeros first) ----
#pragma unroll
        for (int q = 0; q < WPW; ++q) {
            const uint32_t widx = w0 + q;
            const uint32_t j = widx * 32 + lane;
            const bool valid = j < N;
            const uint32_t yk = widx < WS ? ybuf[widx] : 0u;
            const uint32_t vm = __ballot_sync(XSI_FULL, valid);
            const uint32_t nz = ~yk & vm;
            const uint32_t bit = (yk >> lane) & 1u;
            const uint32_t aj = (q & 1) ? (av[q >> 1] >> 16) : (av[q >> 1] & 0xFFFFu);
            const uint32_t dest = bit ? obase + __popc(yk & ltm) : zbase + __popc(nz & ltm);
            if (valid) a[dest] = (uint16_t)aj;
            zbase += __popc(nz);
            obase += __popc(yk);
            if (hap) {  // WAH row of a haploid line is y over a1 = even entries of a (interfaces.hpp:318-333)
                const uint32_t ek = widx < WS ? ebuf[widx] : 0u;
                if (((ek >> lane) & 1u) && bit) {
                    const uint32_t pp = ebase + __popc(ek & ltm);
                    atomicOr(&obuf[pp >> 5], 1u << (pp & 31));
                }
                ebase += __popc(ek);
            }
        }
        __syncthreads();  // #2: a[] updated
        if (hap) {
            for (uint32_t i = tid; i < WS; i += blockDim.x) grow[i] = obuf[i];
            fence_proxy_async();
            __syncthreads();
        }
    }
}

namespace cgx = cooperative_groups;

// =============================================================================================
// E3 v4: PBWT permute on a thread-block cluster with a fence-free exchange (diploid lines,
// <= 65534 haplotypes).  Same formulation as v3 (inverse permutation pos[i], sliced over the C
// CTAs of a cluster) with three changes that remove its stalls (ncu r01: CCTL.IVALL + MEMBAR.ALL.GPU
// of the cluster-scope release/acquire pairs were 30% of all samples, barrier stalls another 12%):
//   * every cross-CTA transfer is a PUSH with st.async (..mbarrier::complete_tx::bytes): the sender
//     stores into the receiver's shared memory and the hardware completes transaction bytes on the
//     receiver's mbarrier; the receiver waits on its own mbarrier only.  No cluster-scope fence.
//   * pos[] lives in REGISTERS: thread t owns KH consecutive haplotypes (two uint16 per register) and
//     reads its KH bits of a natural bit-row with one coalesced load, prefetched four lines ahead.
//   * the pos update of line k and the scatter of line k+1's carriers are one fused pass.
// Per WAH line k and CTA c (slice = WSL row words = WSL*32 positions):
//   0  (fused with step 3 of line k-1)  ypart[pos[i]] = 1 for the carriers i of this CTA's haplotypes
//   -- __syncthreads --
//   1  push ypart word slice d to CTA d (st.async.v4 -> ystage[c] of d, mbY of d), d != c; clear it
//   -- wait own mbY: the C-1 foreign partial slices have landed --
//   2  first WSL/WPT threads: y = OR of the C partial slices; permuted row slice -> global (in place);
//      zero prefix inside the slice (warp scan + named barrier); T[chunk] = zeros-before-in-slice<<16 |
//      16 inverted bits, stored locally and pushed to every other CTA (st.async.v2 -> T, mbT) with the slice total
//   -- every thread arrives on own mbT; wait: whole table here, everybody done with step 2 --
//   3  pos[i] <- x_k[i] ? Z + j - zb(j) : zb(j),  j = pos[i],  zb(j) = base[slice(j)] + T lookup;
//      the slice bases sit in lanes 0..C-1 of every warp and are fetched with one SHFL.
// dynamic smem (u32): ypart[WT] | ystage[C][WSL] | T[2*WT] | zs[8] | sc[32] | mbar[2] (u64),  WT = C*WSL
// =============================================================================================
struct PermV4Cfg { uint32_t WSL, SH; };  // WSL = words per slice (power of 2, >= 32), SH = log2(WSL*32)

template <int C, int KH>
// (A register cap of 48 or 40, tried so that HBM-bound CTAs of another stream could share the SM, made the pipelined leg
// SLOWER, 584 -> 502 / 483 Ggt/s: whatever shares the SM lengthens the chain, which is the critical path.  r02h.)
__global__ void __launch_bounds__(1024, 1) pbwt_permute_v4_kernel(EncDev p, PermV4Cfg cfg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int XW = KH == 64 ? 2 : 1;   // 32-bit words holding this thread's KH bits of a bit-row
    constexpr int WPT = KH == 64 ? 2 : 1;  // row words per thread in step 2 (NT = WSL*32/KH threads)
    const uint32_t N = 2 * p.n_samples, WS = p.WS;
    const uint32_t WSL = cfg.WSL, SH = cfg.SH;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, NT = blockDim.x;
    const uint32_t WT = C * WSL;
    uint32_t* ypart = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* ystage = ypart + WT;
    uint32_t* T = ystage + WT;
    uint32_t* zs = T + 2 * WT;
    uint32_t* sc = zs + 8;
    uint64_t* mb = reinterpret_cast<uint64_t*>(sc + 32);  // [0] mbY: partial slices landed, [1] mbT: table complete
    uint32_t crank = 0;
    if (C > 1) crank = cgx::this_cluster().block_rank();
    const uint32_t b = p.blk_map ? p.blk_map[blockIdx.x / C] : blockIdx.x / C;
    const uint32_t nwah = p.blk_nwah[b];
    const uint32_t* list = p.wah_list + p.blk_line0[b];
    const uint32_t sw0 = crank * WSL;         // first row word of this CTA's slice
    const uint32_t hb = sw0 * 32 + tid * KH;  // first haplotype of this thread
    const bool wlive = sw0 * 32 + (tid & ~31u) * KH < N;  // this warp holds at least one real haplotype
    const uint32_t NP = WSL / WPT;            // threads taking part in step 2
    const uint32_t ypart_sa = smem_u32(ypart), ystage_sa = smem_u32(ystage), T_sa = smem_u32(T), zs_sa = smem_u32(zs), mb_sa = smem_u32(mb);

    // pos of this thread's haplotypes (identity at block start, gt_block.hpp:179): one register each, or two
    // uint16 per register (haplotypes hb+2q low, hb+2q+1 high) when KH > 16
    constexpr bool PACK = KH > 16;
    uint32_t pk[PACK ? KH / 2 : KH];
#pragma unroll
    for (int q = 0; q < (PACK ? KH / 2 : KH); ++q) pk[q] = PACK ? ((hb + 2 * q) | ((hb + 2 * q + 1) << 16)) : (hb + q);
    for (uint32_t i = tid; i < WT; i += NT) ypart[i] = 0;
    if (tid == 0) { mbar_init(&mb[0], 1); mbar_init(&mb[1], (NP + 31) / 32); }  // mbT: one arrival per OWNER warp (+ the pushed bytes)
    if (C > 1) cgx::this_cluster().sync(); else __syncthreads();
    if (nwah == 0) return;  // uniform over the cluster

    // The row word(s) holding this thread's KH bits of a natural-order bit-row.  Loaded four lines ahead and kept
    // raw: the bits are only extracted (extract_x) when the line becomes "next", so nothing waits on the load.
    auto load_x = [&](uint32_t entry, uint32_t (&x)[XW]) {
        const uint32_t* row = p.bitrows + (size_t)(entry & 0x7FFFFFFFu) * WS;
        const uint32_t w = hb >> 5;
        x[0] = w < WS ? row[w] : 0u;
        if (KH == 64) x[XW - 1] = w + 1 < WS ? row[w + 1] : 0u;
    };
    auto extract_x = [&](uint32_t v) { return KH >= 32 ? v : ((v >> (hb & 31u)) & ((1u << (KH & 31)) - 1u)); };
    uint32_t x0[XW], x1[XW], x2[XW], x3[XW];  // lines k, k+1 (extracted bits), k+2, k+3 (raw words)
#pragma unroll
    for (int i = 0; i < XW; ++i) x0[i] = x1[i] = x2[i] = x3[i] = 0;
    load_x(list[0], x0);
    if (nwah > 1) load_x(list[1], x1);
    if (nwah > 2) load_x(list[2], x2);
    if (nwah > 3) load_x(list[3], x3);
#pragma unroll
    for (int i = 0; i < XW; ++i) { x0[i] = extract_x(x0[i]); x1[i] = extract_x(x1[i]); }
    uint32_t e0 = list[0], e1 = nwah > 1 ? list[1] : 0u, e2 = nwah > 2 ? list[2] : 0u, e3 = nwah > 3 ? list[3] : 0u,
             e4 = nwah > 4 ? list[4] : 0u;  // list[k .. k+4]
    // step 0 of the first line
#pragma unroll
    for (int q = 0; q < KH; ++q)
        red_or_shared_if(x0[q >> 5] & (1u << (q & 31)), ypart_sa + (((hb + q) >> 3) & ~3u), 1u << ((hb + q) & 31u));
    const uint32_t pad_zeros = WT * 32 - N;  // positions past N never hold a carrier; they are not real zeros

    for (uint32_t k = 0; k < nwah; ++k) {
        const uint32_t par = k & 1u;
        uint32_t xf[XW];
#pragma unroll
        for (int i = 0; i < XW; ++i) xf[i] = 0;
        if (k + 4 < nwah) load_x(e4, xf);
        const uint32_t e5 = k + 5 < nwah ? list[k + 5] : 0u;
        if (C > 1 && tid == 0) mbar_expect_tx(&mb[0], (C - 1) * WSL * 4);
        __syncthreads();  // ypart of this CTA is complete
        // ---- 1: push the foreign word slices, clear them ----
        if (C > 1) {
            for (uint32_t ch = tid; ch < WT / 4; ch += NT) {
                const uint32_t w4 = ch * 4, dest = w4 >> (SH - 5);
                if (dest != crank) {
                    uint4* src = reinterpret_cast<uint4*>(ypart) + ch;
                    const uint4 v = *src;
                    *src = make_uint4(0, 0, 0, 0);
                    st_async_v4(mapa_u32(ystage_sa + 4 * (crank * WSL + (w4 & (WSL - 1))), dest), v, mapa_u32(mb_sa, dest));
                }
            }
        }
        // ---- 2: combine my slice, publish its table entries ----
        if (tid < NP) {
            if (C > 1) mbar_wait(&mb[0], par);
            uint32_t y[WPT], run = 0;
#pragma unroll
            for (int i = 0; i < WPT; ++i) {
                const uint32_t w = tid * WPT + i;
                uint32_t yy = ypart[sw0 + w];
                ypart[sw0 + w] = 0;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (C > 1 && (uint32_t)c != crank) yy |= ystage[c * WSL + w];
                y[i] = yy;
                run += 32u - __popc(yy);
            }
            uint32_t incl = run;
#pragma unroll
            for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t o = __shfl_up_sync(XSI_FULL, incl, dd); if (lane >= (uint32_t)dd) incl += o; }
            uint32_t wbase = 0, total = __shfl_sync(XSI_FULL, incl, 31);
            if (NP > 32) {
                if (lane == 31) sc[warp] = incl;
                named_bar_sync1(NP);
                const uint32_t wv = lane < (NP >> 5) ? sc[lane] : 0u;
                wbase = __reduce_add_sync(XSI_FULL, lane < warp ? wv : 0u);
                total = __reduce_add_sync(XSI_FULL, wv);
            }
            uint32_t zp = wbase + incl - run;
            uint32_t* grow = p.bitrows + (size_t)(e0 & 0x7FFFFFFFu) * WS;
#pragma unroll
            for (int i = 0; i < WPT; ++i) {
                const uint32_t w = tid * WPT + i, gw = sw0 + w, yy = y[i];
                const uint32_t ny = ~yy;  // the table keeps the ZERO positions as set bits
                const uint32_t e0 = (zp << 16) | (ny & 0xFFFFu);
                const uint32_t zmid = zp + __popc(ny & 0xFFFFu);
                const uint32_t e1 = (zmid << 16) | (ny >> 16);
                *reinterpret_cast<uint2*>(T + 2 * gw) = make_uint2(e0, e1);
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (C > 1 && (uint32_t)c != crank) st_async_v2(mapa_u32(T_sa + 8 * gw, c), e0, e1, mapa_u32(mb_sa + 8, c));
                if (gw < WS) grow[gw] = yy;  // permuted row, in place
                zp += 32u - __popc(yy);
            }
            if (tid == 0) {
                zs[crank] = total;
#pragma unroll
                for (int c = 0; c < C; ++c)
                    if (C > 1 && (uint32_t)c != crank) st_async_b32(mapa_u32(zs_sa + 4 * crank, c), total, mapa_u32(mb_sa + 8, c));
            }
        }
        // mbT completes when every owner warp of this CTA has stored its entries and the other CTAs' entries have landed.  One
        // arrival per owner WARP (lane 0, after __syncwarp has ordered the warp's stores before it); threads that own nothing
        // only wait.  (The kernel first let all 1024 threads arrive; going to 16 arrivals per line changed nothing measurable,
        // 17.37 -> 17.30 ms: the fixed per-line cost is the two DSMEM hops, not the arrivals.)
        if (tid < NP) {
            __syncwarp(NP >= 32 ? XSI_FULL : ((1u << NP) - 1u));
            if (lane == 0) {
                if (C > 1 && tid == 0) mbar_expect_tx(&mb[1], (C - 1) * (WSL * 8 + 4));  // counts as warp 0's arrival
                else mbar_arrive(&mb[1]);
            }
        }
        mbar_wait(&mb[1], par);
        // ---- 3 (+ step 0 of line k+1) ----
        uint32_t basev = 0, Z = 0;
#pragma unroll
        for (int c = 0; c < C; ++c) { const uint32_t v = zs[c]; if (lane > (uint32_t)c) basev += v; Z += v; }
        Z -= pad_zeros;
        // warps whose haplotypes all lie past N (row padding up to the power-of-two slice: 8192 positions for 5,008
        // haplotypes) keep the barriers company but skip the update and the scatter: their positions are never read
        if (wlive) {
#pragma unroll
        for (int q = 0; q < KH; ++q) {
            const uint32_t j = PACK ? ((q & 1) ? (pk[q >> 1] >> 16) : (pk[q >> 1] & 0xFFFFu)) : pk[q];
            const uint32_t e = T[j >> 4];
            uint32_t zb = (e >> 16) + __popc(e & ~(0xFFFFFFFFu << (j & 15u)));
            if (C > 1) zb += __shfl_sync(XSI_FULL, basev, j >> SH);
            const uint32_t np = (x0[q >> 5] & (1u << (q & 31))) ? Z + j - zb : zb;
            if (!PACK) pk[q] = np;
            else if (q & 1) pk[q >> 1] = (pk[q >> 1] & 0xFFFFu) | (np << 16);
            else pk[q >> 1] = (pk[q >> 1] & 0xFFFF0000u) | np;
        }
        // step 0 of line k+1: carriers are the minority (mean allele frequency of a WAH line ~8%), so this is a
        // separate, branchy pass that leaves the update loop above branch-free
#pragma unroll
        for (int i = 0; i < XW; ++i) {
            if (x1[i] == 0) continue;
#pragma unroll
            for (int qq = 0; qq < (KH < 32 ? KH : 32); ++qq) {
                const int q = i * 32 + qq;
                if (x1[i] & (1u << qq)) {
                    const uint32_t np = PACK ? ((q & 1) ? (pk[q >> 1] >> 16) : (pk[q >> 1] & 0xFFFFu)) : pk[q];
                    atomicOr(&ypart[np >> 5], 1u << (np & 31u));
                }
            }
        }
        }  // wlive
#pragma unroll
        for (int i = 0; i < XW; ++i) { x0[i] = x1[i]; x1[i] = extract_x(x2[i]); x2[i] = x3[i]; x3[i] = xf[i]; }
        e0 = e1; e1 = e2; e2 = e3; e3 = e4; e4 = e5;
    }
    if (C > 1) cgx::this_cluster().sync();  // nobody exits while a peer may still push into its shared memory
}

// =============================================================================================
// E3 v5: TWO WAH lines per exchange round (a 4-way stable partition).  v4 spends, per line, a fixed ~1.2 us
// in its two DSMEM exchanges and ~2.6 us (C = 4) in the update itself, whose cost is the random table lookup
// plus ~15 instructions per haplotype (r02d: 52.5 / 28.5 / 17.2 ms at C = 1 / 2 / 4, i.e. issue-bound per SM,
// not latency-bound).  Taking lines k and k+1 together halves the exchange rounds AND the lookups:
//   after both lines the order is: by bit of line k+1, then by bit of line k, then by the old order, so with
//   the class c = x_k[i] | x_{k+1}[i] << 1 of haplotype i at old position j
//       pos'[i] = (# positions of a class below c) + (# positions of class c before j)
//   and the second term needs ONE lookup in a per-class table: per 16 positions and class, the count of that
//   class before the chunk (inside the slice) << 16 | the 16-bit mask of the chunk's positions of that class.
// Per pair and CTA (slice = WSL row words):
//   0  (fused with step 4 of the previous pair) carriers of line k / k+1 set their bit in ypA / ypB at pos[i]
//   -- __syncthreads --
//   1  push the foreign word slices of ypA, ypB and ypO (see 5) to their owners (st.async -> ystage, mbY)
//   2  owners (one thread per row word): y0, y1 = OR of the partial slices; row k -> global (in place); the four class
//      masks, their counts packed 4 x 16 bit in one 64-bit word -> ONE block scan; table entries stored locally
//      and pushed to every other CTA (two 16-byte st.async per word and destination) with the slice totals
//   -- every thread arrives on mbT; wait: all tables here --
//   3  per lane the base of (class, slice) = classes below + same class in lower slices (read by SHFL)
//   4  pos[i] <- base(c, slice(j)) + count(c before j in the slice)          one LDS per haplotype and PAIR
//   5  row k+1 has to come out in the order AFTER line k: the owner of a row word compresses its 32 bits of y1
//      under ~y0 and under y0 (parallel-suffix compress) and ORs the two runs into ypO at the new positions of
//      its zeros / ones of line k; ypO travels with the next round's exchange and is written one pair later.
// Needs slices of at most 32768 positions (16-bit class counts), i.e. C >= 2 for more than 32768 haplotypes.
// MEASURED (r02i / r02j, 32 HRC blocks, C = 4): byte-exact, but 21.6 ms against v4's 17.3 ms, so v4 stays the default and
// this kernel is opt-in (XSI_PBWT_V=5).  The lookups did halve (16.9 instructions per haplotype and PAIR), but they are
// only 24% of the 1,040 warp instructions a warp executes per pair: the owners' section (class masks, 64-bit scan, eight
// table entries, six 16-byte remote stores, the two compress32 of step 5: 416 instructions on half of the warps) and the
// exchange itself (48 KB of tables pushed per CTA and pair, twice v4's bytes per line) grew by as much as the lookups
// shrank, and 30% of the stall samples sit in the two mbarrier waits while the owners work.  v4 itself executes 549 warp
// instructions per warp and line of which the update is ~240: both kernels are bound by the exchange machinery around the
// lookup, not by the lookup.
// dynamic smem (u32): ypA[WT] ypB[WT] ypO[WT] | ystage[3][C][WSL] | T[2*WT][4] | zs[8][4] | sc64[32] (u64) | mbar[2] (u64)
// =============================================================================================
__device__ __forceinline__ uint32_t compress32(uint32_t x, uint32_t m) {  // bits of x under mask m, packed to the right
    x &= m;
    uint32_t mk = ~m << 1;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        uint32_t mp = mk ^ (mk << 1);
        mp ^= mp << 2; mp ^= mp << 4; mp ^= mp << 8; mp ^= mp << 16;
        const uint32_t mv = mp & m;
        m = (m ^ mv) | (mv >> (1 << i));
        const uint32_t t = x & mv;
        x = (x ^ t) | (t >> (1 << i));
        mk &= ~mp;
    }
    return x;
}

template <int C, int KH>  // KH in {8, 16, 32}
__global__ void __launch_bounds__(1024, 1) pbwt_permute_v5_kernel(EncDev p, PermV4Cfg cfg) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t N = 2 * p.n_samples, WS = p.WS;
    const uint32_t WSL = cfg.WSL, SH = cfg.SH;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5, NT = blockDim.x;
    const uint32_t WT = C * WSL;
    uint32_t* yp = reinterpret_cast<uint32_t*>(smem_raw);  // ypA | ypB | ypO
    uint32_t* ystage = yp + 3 * WT;                          // [3][C][WSL]
    uint32_t* T = ystage + 3 * WT;                           // [2*WT chunks][4 classes]
    uint32_t* zs = T + 8 * WT;                               // [C][4] class totals of every slice
    uint64_t* sc64 = reinterpret_cast<uint64_t*>(zs + 32);   // [32] warp totals of the owners' scan
    uint64_t* mb = sc64 + 32;                                // [0] mbY: partial slices landed, [1] mbT: tables complete
    uint32_t crank = 0;
    if (C > 1) crank = cgx::this_cluster().block_rank();
    const uint32_t b = p.blk_map ? p.blk_map[blockIdx.x / C] : blockIdx.x / C;
    const uint32_t nwah = p.blk_nwah[b];
    const uint32_t* list = p.wah_list + p.blk_line0[b];
    const uint32_t sw0 = crank * WSL;
    const uint32_t hb = sw0 * 32 + tid * KH;
    const bool wlive = sw0 * 32 + (tid & ~31u) * KH < N;
    const uint32_t yp_sa = smem_u32(yp), ystage_sa = smem_u32(ystage), T_sa = smem_u32(T), zs_sa = smem_u32(zs), mb_sa = smem_u32(mb);

    uint32_t pk[KH];
#pragma unroll
    for (int q = 0; q < KH; ++q) pk[q] = hb + q;
    for (uint32_t i = tid; i < 3 * WT; i += NT) yp[i] = 0;
    if (tid == 0) { mbar_init(&mb[0], 1); mbar_init(&mb[1], NT); }
    if (C > 1) cgx::this_cluster().sync(); else __syncthreads();
    if (nwah == 0) return;  // uniform over the cluster
    const uint32_t npairs = (nwah + 1) / 2;

    auto load_x = [&](uint32_t entry) -> uint32_t {  // the row word holding this thread's KH bits of a natural-order bit-row
        const uint32_t* row = p.bitrows + (size_t)(entry & 0x7FFFFFFFu) * WS;
        const uint32_t w = hb >> 5;
        return w < WS ? row[w] : 0u;
    };
    auto extract_x = [&](uint32_t v) { return KH >= 32 ? v : ((v >> (hb & 31u)) & ((1u << (KH & 31)) - 1u)); };
    auto entry_of = [&](uint32_t k) { return k < nwah ? list[k] : 0xFFFFFFFFu; };  // 0xFFFFFFFF: no such line
    auto load_line = [&](uint32_t e) { return e == 0xFFFFFFFFu ? 0u : load_x(e); };
    // lines of the current pair (x0, x1) and of the next one (x2, x3) as extracted bits; the pair after that as raw words
    uint32_t e0 = entry_of(0), e1 = entry_of(1), e2 = entry_of(2), e3 = entry_of(3), e4 = entry_of(4), e5 = entry_of(5);
    uint32_t x0 = extract_x(load_line(e0)), x1 = extract_x(load_line(e1));
    uint32_t x2 = extract_x(load_line(e2)), x3 = extract_x(load_line(e3));
    uint32_t r4 = load_line(e4), r5 = load_line(e5);
    uint32_t eo = 0xFFFFFFFFu;  // the line whose permuted row sits in ypO (second line of the previous pair)
    // step 0 of the first pair
#pragma unroll
    for (int q = 0; q < KH; ++q) {
        red_or_shared_if(x0 & (1u << q), yp_sa + (((hb + q) >> 3) & ~3u), 1u << ((hb + q) & 31u));
        red_or_shared_if(x1 & (1u << q), yp_sa + 4 * WT + (((hb + q) >> 3) & ~3u), 1u << ((hb + q) & 31u));
    }
    const uint32_t pad = WT * 32 - N;  // positions past N: always class 0, at the end of the position range

    for (uint32_t t = 0; t <= npairs; ++t) {  // the last round only flushes ypO
        const bool live = t < npairs;
        const uint32_t par = t & 1u;
        const uint32_t e6 = entry_of(2 * t + 6), e7 = entry_of(2 * t + 7);
        const uint32_t r6 = live ? load_line(e6) : 0u, r7 = live ? load_line(e7) : 0u;
        if (C > 1 && tid == 0) mbar_expect_tx(&mb[0], (C - 1) * WSL * 4 * 3);
        __syncthreads();  // the three partial bitmaps of this CTA are complete
        // ---- 1: push the foreign word slices, clear them ----
        if (C > 1) {
            for (uint32_t ch = tid; ch < 3 * (WT / 4); ch += NT) {
                const uint32_t which = ch / (WT / 4), c4 = ch - which * (WT / 4);
                const uint32_t w4 = c4 * 4, dest = w4 >> (SH - 5);
                if (dest != crank) {
                    uint4* src = reinterpret_cast<uint4*>(yp + which * WT) + c4;
                    const uint4 v = *src;
                    *src = make_uint4(0, 0, 0, 0);
                    st_async_v4(mapa_u32(ystage_sa + 4 * ((which * C + crank) * WSL + (w4 & (WSL - 1))), dest), v, mapa_u32(mb_sa, dest));
                }
            }
        }
        // ---- 2: owners: combine my slice, write rows, publish the class tables ----
        uint32_t y0 = 0, y1 = 0;
        uint64_t before = 0;
        if (tid < WSL) {
            if (C > 1) mbar_wait(&mb[0], par);
            const uint32_t gw = sw0 + tid;
            uint32_t yo = yp[2 * WT + gw];
            y0 = yp[gw]; y1 = yp[WT + gw];
            yp[gw] = 0; yp[WT + gw] = 0; yp[2 * WT + gw] = 0;
#pragma unroll
            for (int c = 0; c < C; ++c)
                if (C > 1 && (uint32_t)c != crank) {
                    y0 |= ystage[(0 * C + c) * WSL + tid];
                    y1 |= ystage[(1 * C + c) * WSL + tid];
                    yo |= ystage[(2 * C + c) * WSL + tid];
                }
            if (gw < WS) {
                if (live) p.bitrows[(size_t)(e0 & 0x7FFFFFFFu) * WS + gw] = y0;                  // row k: order before line k
                if (eo != 0xFFFFFFFFu) p.bitrows[(size_t)(eo & 0x7FFFFFFFu) * WS + gw] = yo;     // row k-1: order after line k-2
            }
            if (live) {
                const uint32_t m0 = ~y0 & ~y1, m1 = y0 & ~y1, m2 = ~y0 & y1, m3 = y0 & y1;
                const uint64_t P = (uint64_t)__popc(m0) | ((uint64_t)__popc(m1) << 16) | ((uint64_t)__popc(m2) << 32) | ((uint64_t)__popc(m3) << 48);
                uint64_t incl = P;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) { const uint64_t o = __shfl_up_sync(XSI_FULL, incl, dd); if (lane >= (uint32_t)dd) incl += o; }
                uint64_t wbase = 0, total = __shfl_sync(XSI_FULL, incl, 31);
                if (WSL > 32) {
                    if (lane == 31) sc64[warp] = incl;
                    named_bar_sync1(WSL);
                    const uint64_t wv = lane < (WSL >> 5) ? sc64[lane] : 0ull;
                    const uint64_t lo = lane < warp ? wv : 0ull;
                    // fields stay below 2^16 (slices of at most 32768 positions): the two halves can be summed separately
                    wbase = (uint64_t)__reduce_add_sync(XSI_FULL, (uint32_t)lo) | ((uint64_t)__reduce_add_sync(XSI_FULL, (uint32_t)(lo >> 32)) << 32);
                    total = (uint64_t)__reduce_add_sync(XSI_FULL, (uint32_t)wv) | ((uint64_t)__reduce_add_sync(XSI_FULL, (uint32_t)(wv >> 32)) << 32);
                }
                before = wbase + incl - P;
                const uint32_t mm[4] = {m0, m1, m2, m3};
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t ent[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t cnt = (uint32_t)(before >> (16 * c)) & 0xFFFFu;
                        if (h) cnt += __popc(mm[c] & 0xFFFFu);
                        ent[c] = (cnt << 16) | ((mm[c] >> (16 * h)) & 0xFFFFu);
                    }
                    const uint4 E = make_uint4(ent[0], ent[1], ent[2], ent[3]);
                    const uint32_t chunk = 2 * gw + h;
                    *reinterpret_cast<uint4*>(T + 4 * chunk) = E;
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (C > 1 && (uint32_t)c != crank) st_async_v4(mapa_u32(T_sa + 16 * chunk, c), E, mapa_u32(mb_sa + 8, c));
                }
                if (tid == 0) {
                    const uint4 Z4 = make_uint4((uint32_t)total & 0xFFFFu, (uint32_t)(total >> 16) & 0xFFFFu,
                                                (uint32_t)(total >> 32) & 0xFFFFu, (uint32_t)(total >> 48) & 0xFFFFu);
                    *reinterpret_cast<uint4*>(zs + 4 * crank) = Z4;
#pragma unroll
                    for (int c = 0; c < C; ++c)
                        if (C > 1 && (uint32_t)c != crank) st_async_v4(mapa_u32(zs_sa + 16 * crank, c), Z4, mapa_u32(mb_sa + 8, c));
                }
            }
        }
        if (!live) break;  // uniform: the flush round ends here
        if (C > 1 && tid == 0) mbar_expect_tx(&mb[1], (C - 1) * (WSL * 32 + 16));  // counts as thread 0's arrival
        else mbar_arrive(&mb[1]);
        mbar_wait(&mb[1], par);
        // ---- 3: bases.  lane c*C+s: positions of the classes below c, plus class c in the slices below s ----
        uint32_t tot0 = 0, tot1 = 0, tot2 = 0, bs0 = 0, bs1 = 0, bs2 = 0, bs3 = 0, zb0base = 0;
        const uint32_t myc = (lane / C) & 3u, mys = lane % C;
#pragma unroll
        for (int s = 0; s < C; ++s) {
            const uint4 z = *reinterpret_cast<const uint4*>(zs + 4 * s);
            if ((uint32_t)s < mys) { bs0 += z.x; bs1 += z.y; bs2 += z.z; bs3 += z.w; }
            if ((uint32_t)s < crank) zb0base += z.x + z.z;  // zeros of line k in the slices below mine (step 5)
            tot0 += z.x; tot1 += z.y; tot2 += z.z;
        }
        const uint32_t B1 = tot0 - pad, B2 = B1 + tot1, B3 = B2 + tot2;
        const uint32_t cbv = myc == 0 ? bs0 : (myc == 1 ? B1 + bs1 : (myc == 2 ? B2 + bs2 : B3 + bs3));
        // ---- 4: the update, one lookup per haplotype ----
        if (wlive) {
#pragma unroll
            for (int q = 0; q < KH; ++q) {
                const uint32_t j = pk[q];
                const uint32_t c = ((x0 >> q) & 1u) | (((x1 >> q) & 1u) << 1);
                const uint32_t e = lds_u32(T_sa + ((((j >> 4) << 2) | c) << 2));
                const uint32_t r = (e >> 16) + __popc(e & ((1u << (j & 15u)) - 1u));
                pk[q] = r + __shfl_sync(XSI_FULL, cbv, c * C + (j >> SH));
            }
            // step 0 of the next pair: carriers are the minority, so this is a separate, branchy pass
            if (x2 | x3) {
#pragma unroll
                for (int q = 0; q < KH; ++q) {
                    const uint32_t np = pk[q];
                    if (x2 & (1u << q)) atomicOr(&yp[np >> 5], 1u << (np & 31u));
                    if (x3 & (1u << q)) atomicOr(&yp[WT + (np >> 5)], 1u << (np & 31u));
                }
            }
        }
        // ---- 5: row k+1 in the order after line k, from the owner's side ----
        if (tid < WSL && e1 != 0xFFFFFFFFu) {
            const uint32_t gw = sw0 + tid;
            // zeros of line k before this word (all slices): classes 0 and 2
            const uint32_t zb = zb0base + ((uint32_t)before & 0xFFFFu) + ((uint32_t)(before >> 32) & 0xFFFFu);
            const uint32_t Z0 = tot0 + tot2 - pad;
            const uint32_t vz = compress32(y1, ~y0), vo = compress32(y1, y0);
            uint32_t* out = yp + 2 * WT;
            if (vz) {
                const uint32_t d = zb, sh = d & 31u;
                atomicOr(&out[d >> 5], vz << sh);
                const uint32_t hi = sh ? vz >> (32u - sh) : 0u;  // the run may straddle two words
                if (hi) atomicOr(&out[(d >> 5) + 1], hi);
            }
            if (vo) {
                const uint32_t d = Z0 + gw * 32 - zb, sh = d & 31u;
                atomicOr(&out[d >> 5], vo << sh);
                const uint32_t hi = sh ? vo >> (32u - sh) : 0u;
                if (hi) atomicOr(&out[(d >> 5) + 1], hi);
            }
        }
        eo = e1;
        x0 = x2; x1 = x3; x2 = extract_x(r4); x3 = extract_x(r5); r4 = r6; r5 = r7;
        e0 = e2; e1 = e3; e2 = e4; e3 = e5; e4 = e6; e5 = e7;
    }
    if (C > 1) cgx::this_cluster().sync();  // nobody exits while a peer may still push into its shared memory
}

// =============================================================================================
// E3 grid: PBWT permute for MORE than 65,534 haplotypes (biobank scale, uint32 indices).  A million
// positions do not fit one SM (nor a 16-CTA cluster: 4 MB of positions + 125 KB bitmaps), so the WHOLE
// GPU works on a few PBWT blocks at once, one WAH line of each per step, as a cooperative launch with
// ONE grid barrier per line:
//   state   pos[i] (inverse permutation) in REGISTERS, KH haplotypes per thread, TPB threads per PBWT block
//   A  (fused with C of the previous line) every carrier sets its bit in the line's bitmap Y and adds one to
//      the carrier count of its 128-position group, both in global memory (L2 atomics)
//   -- grid barrier --
//   B  every CTA takes the group counts of ITS block (N/128 words, 31 KB at a million haplotypes) into
//      shared memory and scans them: zeros before every group, and Z
//   C  pos[i] <- x[i] ? Z + j - zb(j) : zb(j),  j = pos[i],  zb(j) = scan[j >> 7] + zeros of the group's
//      four bitmap words before j: one 16-byte L2 load per haplotype
// Y and the counts are triple buffered: in the step that reads line k and scatters line k+1, the buffer of line
// k-1 is copied out as that line's permuted row (in place, bitrows) and cleared for line k+2.
// Cost per line and group of blocks (two blocks of a million haplotypes, r01r): 21.7 us, of which the random
// 16-byte L2 loads of C are about 7 (the same L2 sector rate that bounds the wide decode kernel) and the rest is
// the barrier, the scan and the atomics: fixed per step, so it amortises over the blocks of a group (4 at KH = 32).
// (The first version exchanged chunk totals through flag words and built a table in a second phase with a second
// grid barrier: same time per line, more moving parts.)
// =============================================================================================
struct PermGridCfg {
    uint32_t b0, nbg;   // PBWT blocks [b0, b0+nbg) of the batch
    uint32_t TPB;       // threads per PBWT block (a multiple of 1024: CTAs never straddle blocks)
    uint32_t WSP;       // row words padded to a multiple of 32 (groups of 4 words, 8 groups per scan thread)
    uint32_t* Y;        // [nbg][3][WSP], zero on entry
    uint32_t* CNT;      // [nbg][3][WSP/4] carriers per 128-position group, zero on entry
    uint32_t* bar;      // grid barrier counter, zero on entry
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

template <int KH>  // 8, 16 or 32
__global__ void __launch_bounds__(1024, 1) pbwt_permute_grid_kernel(EncDev p, PermGridCfg c) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* zpre = reinterpret_cast<uint32_t*>(smem_raw);  // [NG] zeros before every group of this CTA's block
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_Z;
    const uint32_t N = 2 * p.n_samples, WS = p.WS, WSP = c.WSP, NG = WSP >> 2;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t ctas_per_block = c.TPB >> 10;
    const uint32_t bi = blockIdx.x / ctas_per_block;                 // PBWT block of this CTA (uniform)
    const uint32_t cib = blockIdx.x - bi * ctas_per_block;           // CTA index inside the block
    const bool active = bi < c.nbg;
    const uint32_t hb = (cib * 1024u + tid) * KH;
    const bool owner = active && hb < N;
    const uint32_t bnwah = active ? p.blk_nwah[c.b0 + bi] : 0u;      // lines of this CTA's block
    const uint32_t nwah = owner ? bnwah : 0u;
    const uint32_t* list = p.wah_list + p.blk_line0[c.b0 + (active ? bi : 0)];
    uint32_t* Yb = c.Y + (size_t)(active ? bi : 0) * 3 * WSP;
    uint32_t* Cb = c.CNT + (size_t)(active ? bi : 0) * 3 * NG;
    uint32_t max_nwah = 0;
    for (uint32_t i = 0; i < c.nbg; ++i) max_nwah = max(max_nwah, p.blk_nwah[c.b0 + i]);
    const uint32_t nvalid = !owner ? 0u : (N - hb >= (uint32_t)KH ? (uint32_t)KH : N - hb);
    uint32_t pk[KH];
    // identity at block start (gt_block.hpp:179); slots past N sit at position 0 and never carry
#pragma unroll
    for (int q = 0; q < KH; ++q) pk[q] = (uint32_t)q < nvalid ? hb + q : 0u;
    const uint32_t vmask = nvalid >= 32 ? 0xFFFFFFFFu : ((1u << nvalid) - 1u);
    // the KH bits of this thread in a natural-order bit-row (rows of later lines are untouched until their copy-out)
    auto load_x = [&](uint32_t k) -> uint32_t {
        if (k >= nwah) return 0u;
        const uint32_t v = __ldcg(p.bitrows + (size_t)(list[k] & 0x7FFFFFFFu) * WS + (hb >> 5));
        return (KH >= 32 ? v : (v >> (hb & 31u))) & vmask;
    };
    auto scatter = [&](uint32_t x, uint32_t buf) {
        if (!x) return;
        uint32_t* Yn = Yb + (size_t)buf * WSP;
        uint32_t* Cn = Cb + (size_t)buf * NG;
#pragma unroll
        for (int q = 0; q < KH; ++q)
            if (x & (1u << q)) { atomicOr(Yn + (pk[q] >> 5), 1u << (pk[q] & 31u)); atomicAdd(Cn + (pk[q] >> 7), 1u); }
    };
    uint32_t bar_target = 0;
    auto grid_arrive = [&]() {
        bar_target += gridDim.x;
        __syncthreads();
        if (tid == 0) { __threadfence(); atomicAdd(c.bar, 1u); }
    };
    auto grid_wait = [&]() {
        if (tid == 0) { while (ld_volatile_u32(c.bar) < bar_target) {} __threadfence(); }
        __syncthreads();
    };

    uint32_t x0 = load_x(0), x1 = load_x(1), x2 = load_x(2);
    scatter(x0, 0);  // A of the first line (positions are the identity)

    for (uint32_t k = 0; k < max_nwah; ++k) {
        const uint32_t cur = k % 3u, nxt = (k + 1) % 3u, old = (k + 2) % 3u;  // old = buffer of line k-1
        grid_arrive();
        grid_wait();  // every carrier of line k has landed; everybody is done with line k-1
        // ---- housekeeping of line k-1's buffer (nobody reads it any more): permuted row out (in place), clear ----
        if (active && k >= 1 && k - 1 < bnwah) {
            uint32_t* Yo = Yb + (size_t)old * WSP;
            uint32_t* Co = Cb + (size_t)old * NG;
            uint32_t* grow = p.bitrows + (size_t)(list[k - 1] & 0x7FFFFFFFu) * WS;
            for (uint32_t w4 = cib * 1024u + tid; w4 < (WSP >> 2); w4 += ctas_per_block * 1024u) {
                const uint4 y = __ldcg(reinterpret_cast<const uint4*>(Yo) + w4);
                if (4 * w4 < WS) *reinterpret_cast<uint4*>(grow + 4 * w4) = y;
                __stcg(reinterpret_cast<uint4*>(Yo) + w4, make_uint4(0, 0, 0, 0));
                __stcg(Co + w4, 0u);  // one group = four words
            }
        }
        if (active && k < bnwah) {
            // ---- B: zeros before every 128-position group of this block (block-wide exclusive scan) ----
            const uint32_t* Cc = Cb + (size_t)cur * NG;
            uint32_t carry = 0;
            for (uint32_t g0 = 0; g0 < NG; g0 += 8192u) {  // 8 groups per thread and pass
                const uint32_t g = g0 + tid * 8u;
                uint4 cn = make_uint4(0, 0, 0, 0), cm = make_uint4(0, 0, 0, 0);
                if (g < NG) { cn = __ldcg(reinterpret_cast<const uint4*>(Cc + g)); cm = __ldcg(reinterpret_cast<const uint4*>(Cc + g) + 1); }
                const uint32_t z0 = 128u - cn.x, z1 = 128u - cn.y, z2 = 128u - cn.z, z3 = 128u - cn.w;
                const uint32_t z4 = 128u - cm.x, z5 = 128u - cm.y, z6 = 128u - cm.z, z7 = 128u - cm.w;
                const uint32_t tot = g < NG ? z0 + z1 + z2 + z3 + z4 + z5 + z6 + z7 : 0u;
                uint32_t incl = tot;
#pragma unroll
                for (int dd = 1; dd < 32; dd <<= 1) { const uint32_t o = __shfl_up_sync(XSI_FULL, incl, dd); if (lane >= (uint32_t)dd) incl += o; }
                if (lane == 31) s_warp[warp] = incl;
                __syncthreads();
                const uint32_t wv = s_warp[lane];
                const uint32_t wbase = __reduce_add_sync(XSI_FULL, lane < warp ? wv : 0u);
                const uint32_t wtot = __reduce_add_sync(XSI_FULL, wv);
                const uint32_t ex = carry + wbase + incl - tot;
                if (g < NG) {
                    const uint32_t h = ex + z0 + z1 + z2 + z3;
                    *reinterpret_cast<uint4*>(zpre + g) = make_uint4(ex, ex + z0, ex + z0 + z1, ex + z0 + z1 + z2);
                    *reinterpret_cast<uint4*>(zpre + g + 4) = make_uint4(h, h + z4, h + z4 + z5, h + z4 + z5 + z6);
                }
                carry += wtot;
                __syncthreads();
            }
            if (tid == 0) s_Z = carry - (WSP * 32u - N);  // the padding past N never holds a carrier
            __syncthreads();
            // ---- C (+ A of line k+1) ----
            if (k < nwah) {
                const uint32_t Z = s_Z;
                const uint4* Y4 = reinterpret_cast<const uint4*>(Yb + (size_t)cur * WSP);
#pragma unroll
                for (int q0 = 0; q0 < KH; q0 += 4) {
                    uint4 y[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) y[q] = __ldcg(Y4 + (pk[q0 + q] >> 7));
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t j = pk[q0 + q];
                        const uint32_t wsel = (j >> 5) & 3u;
                        // zeros of the group before j: whole words below wsel, then the low bits of word wsel
                        const uint32_t n0 = ~y[q].x, n1 = ~y[q].y, n2 = ~y[q].z, n3 = ~y[q].w;
                        uint32_t zb = zpre[j >> 7];
                        zb += wsel > 0 ? __popc(n0) : 0u;
                        zb += wsel > 1 ? __popc(n1) : 0u;
                        zb += wsel > 2 ? __popc(n2) : 0u;
                        const uint32_t wj = wsel == 0 ? n0 : wsel == 1 ? n1 : wsel == 2 ? n2 : n3;
                        zb += __popc(wj & ~(0xFFFFFFFFu << (j & 31u)));
                        pk[q0 + q] = (x0 & (1u << (q0 + q))) ? Z + j - zb : zb;
                    }
                }
                scatter(x1, nxt);
                x0 = x1; x1 = x2; x2 = load_x(k + 3);
            }
        }
    }
    // the block(s) with the most lines still hold their last permuted row (the others were served inside the loop)
    grid_arrive();
    grid_wait();
    if (active && bnwah >= 1 && bnwah == max_nwah) {
        const uint32_t old = (bnwah - 1) % 3u;
        const uint32_t* Yo = Yb + (size_t)old * WSP;
        uint32_t* grow = p.bitrows + (size_t)(list[bnwah - 1] & 0x7FFFFFFFu) * WS;
        for (uint32_t w4 = cib * 1024u + tid; 4 * w4 < WS; w4 += ctas_per_block * 1024u)
            *reinterpret_cast<uint4*>(grow + 4 * w4) = __ldcg(reinterpret_cast<const uint4*>(Yo) + w4);
    }
}

// generic fallback for > 65536 haplotypes: a[] ping-pongs in global memory (L2 resident)
__global__ void __launch_bounds__(1024, 1) pbwt_permute_gmem_kernel(EncDev p, uint32_t* a_pool) {
    __shared__ uint32_t zc[64];
    __shared__ uint32_t s_tot[2];
    const uint32_t N = 2 * p.n_samples;
    const uint32_t W = (N + 31) >> 5, WS = p.WS;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5, NW = blockDim.x >> 5;
    const uint32_t b = blockIdx.x;
    const uint32_t l0 = p.blk_line0[b], nwah = p.blk_nwah[b];
    const uint32_t* list = p.wah_list + l0;
    uint32_t* abuf[2] = {a_pool + (size_t)b * 2 * N, a_pool + (size_t)b * 2 * N + N};
    uint32_t* ytmp = a_pool + (size_t)gridDim.x * 2 * N + (size_t)b * 2 * WS;  // y words, then even masks
    for (uint32_t i = tid; i < N; i += blockDim.x) abuf[0][i] = i;
    __syncthreads();
    const uint32_t wpw = (W + NW - 1) / NW;
    const uint32_t w0 = warp * wpw, w1 = min(W, w0 + wpw);
    const uint32_t ltm = lanemask_lt();
    uint32_t cur = 0;
    for (uint32_t k = 0; k < nwah; ++k) {
        const uint32_t line = list[k] & 0x7FFFFFFFu;
        const bool hap = (list[k] >> 31) != 0;
        uint32_t* row = p.bitrows + (size_t)line * WS;
        const uint32_t* a = abuf[cur];
        uint32_t* an = abuf[cur ^ 1];
        uint32_t zeros = 0, evens = 0;
        for (uint32_t widx = w0; widx < w1; ++widx) {
            const uint32_t j = widx * 32 + lane;
            const bool valid = j < N;
            const uint32_t aj = valid ? a[j] : 0u;
            const uint32_t gi = hap ? (aj >> 1) : aj;
            const uint32_t bit = valid ? ((row[gi >> 5] >> (gi & 31)) & 1u) : 0u;
            const uint32_t yk = __ballot_sync(XSI_FULL, bit);
            const uint32_t vm = __ballot_sync(XSI_FULL, valid);
            zeros += __popc(~yk & vm);
            const uint32_t ek = __ballot_sync(XSI_FULL, valid && !(aj & 1u));
            evens += __popc(ek);
            if (lane == 0) { ytmp[widx] = yk; ytmp[WS + widx] = ek; }
        }
        if (lane == 0) { zc[warp] = zeros; zc[32 + warp] = evens; }
        __syncthreads();
        const uint32_t zv = lane < NW ? zc[lane] : 0u;
        const uint32_t Z = __reduce_add_sync(XSI_FULL, zv);
        uint32_t zbase = __reduce_add_sync(XSI_FULL, lane < warp ? zv : 0u);
        const uint32_t ev = lane < NW ? zc[32 + lane] : 0u;
        uint32_t ebase = __reduce_add_sync(XSI_FULL, lane < warp ? ev : 0u);
        uint32_t obase = Z + (min(w0 * 32, N) - zbase);
        // the natural row has been fully consumed: overwrite it with the permuted row
        if (!hap) { for (uint32_t i = tid; i < WS; i += blockDim.x) row[i] = i < W ? ytmp[i] : 0u; }
        else { for (uint32_t i = tid; i < WS; i += blockDim.x) row[i] = 0u; }
        __syncthreads();
        for (uint32_t widx = w0; widx < w1; ++widx) {
            const uint32_t j = widx * 32 + lane;
            const bool valid = j < N;
            const uint32_t yk = ytmp[widx];
            const uint32_t vm = __ballot_sync(XSI_FULL, valid);
            const uint32_t nz = ~yk & vm;
            const uint32_t bit = (yk >> lane) & 1u;
            const uint32_t dest = bit ? obase + __popc(yk & ltm) : zbase + __popc(nz & ltm);
            if (valid) an[dest] = a[j];
            zbase += __popc(nz);
            obase += __popc(yk);
            if (hap) {
                const uint32_t ek = ytmp[WS + widx];
                if (((ek >> lane) & 1u) && bit) { const uint32_t pp = ebase + __popc(ek & ltm); atomicOr(&row[pp >> 5], 1u << (pp & 31)); }
                ebase += __popc(ek);
            }
        }
        __syncthreads();
        cur ^= 1;
    }
    (void)s_tot;
}

// =============================================================================================
// E4: WAH2-16 encode of bit-rows, one warp per row
// =============================================================================================
// Rules (wah.hpp:376-429,568-573): 15-bit groups LSB first; all-zero / all-one groups are
// counted into 0x8000|n / 0xC000|n words, n <= 16383 (a saturated counter is emitted as
// 0xBFFF / 0xFFFF and counting restarts); anything else is a literal; the tail is zero padded.
__device__ __forceinline__ uint32_t wah_encode_row_warp(const uint32_t* __restrict__ row, uint32_t WS, uint32_t nbits,
                                                        uint16_t* __restrict__ out) {
    const uint32_t lane = lane_id();
    const uint32_t G = (nbits + 14) / 15;
    const uint32_t iters = (G + 31) >> 5;
    uint32_t heads_base = 0, carry_type = 3, carry_len = 0;
    const uint32_t le = lanemask_lt() | (1u << lane);
    // the 16 row words of an iteration are loaded two iterations ahead: the loop is a chain of shuffles and ballots, and an
    // exposed global load per iteration (a line is 136 iterations at 64,976 haplotypes) was most of its time
    uint32_t wnext = (lane < 16 && lane < WS) ? row[lane] : 0u;
    uint32_t wnext2 = (lane < 16 && 15 + lane < WS && iters > 1) ? row[15 + lane] : 0u;
    for (uint32_t it = 0; it < iters; ++it) {
        const uint32_t wv = wnext;
        wnext = wnext2;
        const uint32_t wi2 = (it + 2) * 15 + lane;
        wnext2 = (lane < 16 && wi2 < WS && it + 2 < iters) ? row[wi2] : 0u;
        const uint32_t bitpos = 15 * lane;
        const uint32_t lo = __shfl_sync(XSI_FULL, wv, bitpos >> 5);
        const uint32_t hi = __shfl_sync(XSI_FULL, wv, (bitpos >> 5) + 1);
        const uint32_t grp = __funnelshift_r(lo, hi, bitpos & 31) & 0x7FFFu;
        const uint32_t nextgrp = __shfl_sync(XSI_FULL, wv, 15) & 0x7FFFu;
        const uint32_t g = it * 32 + lane;
        const bool valid = g < G;
        const uint32_t type = !valid ? 3u : (grp == 0 ? 0u : (grp == 0x7FFFu ? 1u : 2u));
        const uint32_t nexttype = ((it + 1) * 32 < G) ? (nextgrp == 0 ? 0u : (nextgrp == 0x7FFFu ? 1u : 2u)) : 3u;
        uint32_t prevtype = __shfl_up_sync(XSI_FULL, type, 1);
        if (lane == 0) prevtype = carry_type;
        uint32_t succtype = __shfl_down_sync(XSI_FULL, type, 1);
        if (lane == 31) succtype = nexttype;
        const bool h0 = valid && (type == 2 || type != prevtype);
        const uint32_t H0 = __ballot_sync(XSI_FULL, h0);
        const uint32_t hb = H0 & le;
        const uint32_t offset = hb ? lane - (31 - __clz(hb)) : carry_len + lane;  // groups since the run's natural start
        const bool h1 = valid && type < 2 && offset > 0 && (offset % 16383u) == 0;
        const uint32_t H = H0 | __ballot_sync(XSI_FULL, h1);
        const uint32_t idx = heads_base + __popc(H & le) - 1;
        if (valid) {
            if (type == 2) out[idx] = (uint16_t)grp;
            else if (succtype != type || ((offset + 1) % 16383u) == 0)
                out[idx] = (uint16_t)(0x8000u | (type << 14) | ((offset % 16383u) + 1));
        }
        heads_base += __popc(H);
        const uint32_t t31 = __shfl_sync(XSI_FULL, type, 31);
        const uint32_t o31 = __shfl_sync(XSI_FULL, offset, 31);
        carry_type = t31;
        carry_len = t31 < 2 ? o31 + 1 : 0;
    }
    return heads_base;
}

constexpr int E4_WARPS = 4;
__global__ void __launch_bounds__(E4_WARPS * 32) wah_encode_rows_kernel(EncDev p) {
    const uint32_t job = blockIdx.x * E4_WARPS + (threadIdx.x >> 5);
    if (job < p.L) {
        if (!(p.line_flags[job] & LF_WAH)) return;
        const uint32_t n = wah_encode_row_warp(p.bitrows + (size_t)job * p.WS, p.WS, p.rec_ngt[p.line_rec[job]],
                                               p.wahslots + (size_t)job * p.SLOTW);
        if (lane_id() == 0) p.line_wah_n[job] = n;
    } else if (job < p.L + p.R) {
        const uint32_t r = job - p.L;
        const int32_t slot = p.rec_aux[r * 3 + 2];
        if (slot < 0) return;
        const uint32_t n = wah_encode_row_warp(p.phrows + (size_t)slot * p.WS, p.WS, p.rec_ngt[r],
                                               p.phslots + (size_t)slot * p.SLOTW);
        if (lane_id() == 0) p.rec_phase_n[r] = n;
    } else if (p.wah_missing && job < p.L + 3 * p.R) {
        // gt_block.hpp:340-372 under WS_WAH: a_weirdness is the identity, the predicate rows are encoded as they are
        const uint32_t which = (job - p.L) / p.R - 1;  // 0: missing, 1: end of vector
        const uint32_t r = (job - p.L) - (which + 1) * p.R;
        const int32_t slot = p.rec_aux[r * 3 + which];
        if (slot < 0) return;
        const uint32_t n = wah_encode_row_warp(p.auxrows + (size_t)slot * p.WS, p.WS, p.rec_ngt[r],
                                               p.auxslots + (size_t)slot * p.SLOTW);
        if (lane_id() == 0) (which == 0 ? p.rec_missw_n : p.rec_eovw_n)[r] = n;
    }
}

// =============================================================================================
// exclusive scan of u32 -> u64 over several arrays at once: tiles of 4096 entries, one CTA per (tile, array); pass 1 sums
// the tiles, pass 2 adds up the sums of the tiles before its own (at most a few hundred) and scans its tile.  (One CTA per
// array walked 1.8 M entries in 440 dependent iterations: 1.4 ms per batch on the 1KGP3 shape.)
// =============================================================================================
struct ScanJob { const uint32_t* in; uint64_t* out; uint32_t n; uint32_t pad; };  // out has n+1 entries
constexpr int SCAN_THREADS = 1024;
constexpr uint32_t SCAN_TILE = SCAN_THREADS * 4;
__host__ __device__ inline uint32_t scan_tiles(uint32_t n) { return n ? (n + SCAN_TILE - 1) / SCAN_TILE : 1u; }

__device__ __forceinline__ uint64_t scan_block_sum(uint64_t v, uint64_t* s_warp) {  // sum over the CTA, returned to every thread
    const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(XSI_FULL, v, d);
    __syncthreads();  // s_warp may still be read by the previous use
    if (lane == 0) s_warp[warp] = v;
    __syncthreads();
    uint64_t t = s_warp[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) t += __shfl_xor_sync(XSI_FULL, t, d);
    return t;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_u32_sums_kernel(const ScanJob* jobs, uint64_t* sums, uint32_t max_tiles) {
    __shared__ uint64_t s_warp[32];
    const ScanJob j = jobs[blockIdx.y];
    const uint32_t tile = blockIdx.x;
    if (tile >= scan_tiles(j.n)) return;
    const uint32_t i0 = tile * SCAN_TILE + threadIdx.x * 4;
    uint64_t t = 0;
#pragma unroll
    for (int q = 0; q < 4; ++q) t += (i0 + q < j.n) ? j.in[i0 + q] : 0u;
    t = scan_block_sum(t, s_warp);
    if (threadIdx.x == 0) sums[(size_t)blockIdx.y * max_tiles + tile] = t;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_u32_kernel(const ScanJob* jobs, const uint64_t* sums, uint32_t max_tiles) {
    __shared__ uint64_t s_warp[32];
    const ScanJob j = jobs[blockIdx.y];
    const uint32_t tile = blockIdx.x, ntiles = scan_tiles(j.n);
    if (tile >= ntiles) return;
    const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
    // entries of the tiles before this one
    uint64_t before = 0;
    for (uint32_t t = tid; t < tile; t += SCAN_THREADS) before += sums[(size_t)blockIdx.y * max_tiles + t];
    before = scan_block_sum(before, s_warp);
    const uint32_t i0 = tile * SCAN_TILE + tid * 4;
    uint32_t v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = (i0 + q < j.n) ? j.in[i0 + q] : 0u;
    const uint64_t t = (uint64_t)v[0] + v[1] + v[2] + v[3];
    uint64_t incl = t;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint64_t o = __shfl_up_sync(XSI_FULL, incl, d); if (lane >= (uint32_t)d) incl += o; }
    __syncthreads();
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = s_warp[lane], wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint64_t o = __shfl_up_sync(XSI_FULL, wi, d); if (lane >= (uint32_t)d) wi += o; }
        s_warp[lane] = wi - w;  // exclusive
    }
    __syncthreads();
    uint64_t ex = before + s_warp[warp] + (incl - t);
#pragma unroll
    for (int q = 0; q < 4; ++q) { if (i0 + q < j.n) j.out[i0 + q] = ex; ex += v[q]; }
    if (tile == ntiles - 1 && tid == SCAN_THREADS - 1) j.out[j.n] = ex;  // the last thread's running total covers the whole array
}

// The host lays the blocks out from the section offsets AT BLOCK BOUNDARIES only (first line / first record of every block, and
// the totals): gather those instead of sending every per-line offset back (72 MB per batch at 1.8 M lines).
struct GatherOffs {
    const uint64_t* src[7];   // exclusive scans, n+1 entries each
    uint32_t by_line[7];      // 1: indexed by binary line (blk_line0), 0: by record (b * block_len, last = R)
    uint32_t n_arrays, nb, block_len, R;
    const uint32_t* blk_line0;  // [nb+1]
    uint64_t* dst;              // [n_arrays][nb+1]
};
__global__ void __launch_bounds__(256) gather_block_offsets_kernel(GatherOffs g) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.n_arrays * (g.nb + 1)) return;
    const uint32_t a = i / (g.nb + 1), b = i - a * (g.nb + 1);
    const uint64_t at = g.by_line[a] ? g.blk_line0[b] : min((uint64_t)b * g.block_len, (uint64_t)g.R);
    g.dst[i] = g.src[a][at];
}

// =============================================================================================
// E5: sparse index lists.  job < L: GT sparse line; L..L+R: missing list; L+R..L+2R: EOV list
// =============================================================================================
struct EmitDev {
    const uint64_t* line_sparse_off;  // [L+1] in A_T entries
    const uint64_t* rec_miss_off;     // [R+1]
    const uint64_t* rec_eov_off;      // [R+1]
    const uint64_t* line_wah_off;     // [L+1] in u16 words
    const uint64_t* rec_phase_off;    // [R+1]
    void* out_sparse; void* out_miss; void* out_eov;
    uint16_t* out_wah; uint16_t* out_phase;
    const uint64_t* rec_missw_off;    // [R+1] (WS_WAH only)
    const uint64_t* rec_eovw_off;     // [R+1]
    uint16_t* out_missw; uint16_t* out_eovw;
};

template <typename AT>
__device__ __forceinline__ void sparse_emit_row_warp(const uint32_t* __restrict__ row, uint32_t nbits, AT* __restrict__ out,
                                                     bool negated) {
    const uint32_t lane = lane_id();
    const uint32_t nwords = (nbits + 31) >> 5;
    uint32_t base = 0;
    for (uint32_t w0 = 0; w0 < nwords; w0 += 32) {
        const uint32_t wi = w0 + lane;
        uint32_t w = wi < nwords ? row[wi] : 0u;
        if (wi == nwords - 1 && (nbits & 31)) w &= (1u << (nbits & 31)) - 1u;
        const uint32_t c = __popc(w);
        uint32_t incl = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(XSI_FULL, incl, d); if (lane >= (uint32_t)d) incl += o; }
        uint32_t pos = base + incl - c;
        while (w) { const uint32_t bpos = __ffs(w) - 1; w &= w - 1; out[1 + pos++] = (AT)(wi * 32 + bpos); }
        base += __shfl_sync(XSI_FULL, incl, 31);
    }
    if (lane == 0) out[0] = (AT)(base | (negated ? ((AT)1 << (sizeof(AT) * 8 - 1)) : (AT)0));
}

constexpr int E5_WARPS = 4;
template <typename AT>
__global__ void __launch_bounds__(E5_WARPS * 32) sparse_emit_kernel(EncDev p, EmitDev e) {
    const uint32_t job = blockIdx.x * E5_WARPS + (threadIdx.x >> 5);
    if (job < p.L) {
        const uint8_t lf = p.line_flags[job];
        if (lf & LF_WAH) return;
        sparse_emit_row_warp<AT>(p.bitrows + (size_t)job * p.WS, p.rec_ngt[p.line_rec[job]],
                                 reinterpret_cast<AT*>(e.out_sparse) + e.line_sparse_off[job], (lf & LF_NEGATED) != 0);
    } else if (job < p.L + 2 * p.R) {
        const uint32_t kind = (job - p.L) / p.R;  // 0 missing, 1 eov
        const uint32_t r = (job - p.L) - kind * p.R;
        const int32_t slot = p.rec_aux[r * 3 + kind];
        if (slot < 0) return;
        AT* out = kind == 0 ? reinterpret_cast<AT*>(e.out_miss) + e.rec_miss_off[r]
                            : reinterpret_cast<AT*>(e.out_eov) + e.rec_eov_off[r];
        sparse_emit_row_warp<AT>(p.auxrows + (size_t)slot * p.WS, p.rec_ngt[r], out, false);
    }
}

// =============================================================================================
// E6: pack the WAH words of every line / phase line into the contiguous matrices
// =============================================================================================
constexpr int E6_WARPS = 4;
__global__ void __launch_bounds__(E6_WARPS * 32) pack_wah_kernel(EncDev p, EmitDev e) {
    const uint32_t job = blockIdx.x * E6_WARPS + (threadIdx.x >> 5);
    const uint32_t lane = lane_id();
    const uint16_t* src; uint16_t* dst; uint32_t n;
    if (job < p.L) {
        n = p.line_wah_n[job];
        if (!n) return;
        src = p.wahslots + (size_t)job * p.SLOTW;
        dst = e.out_wah + e.line_wah_off[job];
    } else if (job < p.L + p.R) {
        const uint32_t r = job - p.L;
        n = p.rec_phase_n[r];
        if (!n) return;
        src = p.phslots + (size_t)p.rec_aux[r * 3 + 2] * p.SLOTW;
        dst = e.out_phase + e.rec_phase_off[r];
    } else if (p.wah_missing && job < p.L + 3 * p.R) {
        const uint32_t which = (job - p.L) / p.R - 1;
        const uint32_t r = (job - p.L) - (which + 1) * p.R;
        n = (which == 0 ? p.rec_missw_n : p.rec_eovw_n)[r];
        if (!n) return;
        src = p.auxslots + (size_t)p.rec_aux[r * 3 + which] * p.SLOTW;
        dst = which == 0 ? e.out_missw + e.rec_missw_off[r] : e.out_eovw + e.rec_eovw_off[r];
    } else return;
    for (uint32_t i = lane; i < n; i += 32) dst[i] = src[i];
}

}  // namespace xsi
