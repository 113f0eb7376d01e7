/* oracle/xsi_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT part of the product.
 *
 * A plain-C, single-threaded restatement of the xSqueezeIt (rwk-unil/xSqueezeIt @55ad8c7)
 * genotype encode/decode path, written from the reference's behaviour.  It exists only to
 * check the CUDA path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
 * may load it.  The product (xsqueezeit_b200/) never links, imports or calls this file.
 *
 * Parity pinning: this restatement is checked byte-for-byte against the reference's own
 * output (oracle/_ref, the unmodified reference compiled in place) on every fixture of
 * test/test_files and on the SHA-256 table of SURVEY.md section 8(c); see
 * tests/test_oracle_golden.py and tests/golden/.
 *
 * Each function cites the reference file:line (paths under /root/reference) it follows.
 */
#include "xsi_oracle.h"

#include <stdlib.h>
#include <string.h>

#define BCF_MISSING_I32 ((int32_t)0x80000000)    /* htslib/vcf.h:1324 bcf_int32_missing   */
#define BCF_VECTOR_END_I32 ((int32_t)0x80000001) /* htslib/vcf.h:1329 bcf_int32_vector_end */
static inline int32_t gt_allele(int32_t v) { return (v >> 1) - 1; }           /* vcf.h:897 */
static inline int gt_phased(int32_t v) { return v & 1; }                      /* vcf.h:895 */
static inline int gt_missing(int32_t v) { return ((v >> 1) == 0) || v == BCF_MISSING_I32; } /* vcf.h:894 + gt_block.hpp:79 */
static inline int32_t gt_unphased(int32_t allele) { return (allele + 1) << 1; } /* vcf.h:892 */

/* ------------------------------------------------------------------ growable byte buffer */
typedef struct { uint8_t* p; size_t n, cap; } buf_t;
static int buf_reserve(buf_t* b, size_t extra) {
    if (b->n + extra <= b->cap) return 0;
    size_t nc = b->cap ? b->cap : 4096;
    while (nc < b->n + extra) nc *= 2;
    uint8_t* q = (uint8_t*)realloc(b->p, nc);
    if (!q) return -1;
    b->p = q; b->cap = nc;
    return 0;
}
static void buf_put(buf_t* b, const void* src, size_t len) {
    if (buf_reserve(b, len)) abort();
    if (len) memcpy(b->p + b->n, src, len);
    b->n += len;
}
static void buf_u32(buf_t* b, uint32_t v) { buf_put(b, &v, 4); }
static void buf_u16(buf_t* b, uint16_t v) { buf_put(b, &v, 2); }
static void buf_zero(buf_t* b, size_t len) { if (buf_reserve(b, len)) abort(); memset(b->p + b->n, 0, len); b->n += len; }

/* ------------------------------------------------------------------ WAH2-16 */
/* wah.hpp:376-429 process_wah_word */
typedef struct { uint16_t ones, zeros; } wah_state;
static void wah_push_group(buf_t* out, wah_state* st, uint16_t word) {
    if (word == 0) {
        if (st->ones) { buf_u16(out, (uint16_t)(0xC000u | st->ones)); st->ones = 0; }
        if (st->zeros == 0x3FFF) { buf_u16(out, 0xBFFF); st->zeros = 0; }
        st->zeros++;
    } else if (word == 0x7FFF) {
        if (st->zeros) { buf_u16(out, (uint16_t)(0x8000u | st->zeros)); st->zeros = 0; }
        if (st->ones == 0x3FFF) { buf_u16(out, 0xFFFF); st->ones = 0; }
        st->ones++;
    } else {
        if (st->ones) { buf_u16(out, (uint16_t)(0xC000u | st->ones)); st->ones = 0; }
        if (st->zeros) { buf_u16(out, (uint16_t)(0x8000u | st->zeros)); st->zeros = 0; }
        buf_u16(out, word);
    }
}
/* wah.hpp:568-573 / 330-335 final flush: zeros first, then ones */
static void wah_finish(buf_t* out, wah_state* st) {
    if (st->zeros) buf_u16(out, (uint16_t)(0x8000u | st->zeros));
    if (st->ones) buf_u16(out, (uint16_t)(0xC000u | st->ones));
}
/* wah.hpp:238-342 wah_encode2(vector<bool>): one byte per bit in, zero padded tail */
static void wah_encode_bytes(buf_t* out, const uint8_t* bits, uint64_t n) {
    wah_state st = {0, 0};
    uint64_t groups = (n + 14) / 15;
    for (uint64_t g = 0; g < groups; ++g) {
        uint16_t w = 0;
        for (unsigned j = 0; j < 15; ++j) {
            uint64_t i = g * 15 + j;
            if (i < n && bits[i]) w |= (uint16_t)(1u << j);
        }
        wah_push_group(out, &st, w);
    }
    wah_finish(out, &st);
}
uint64_t xo_wah_encode_bits(const uint8_t* bits, uint64_t n, uint16_t* outw) {
    buf_t b = {0, 0, 0};
    wah_encode_bytes(&b, bits, n);
    memcpy(outw, b.p, b.n);
    uint64_t words = b.n / 2;
    free(b.p);
    return words;
}
/* wah.hpp:177-223 wah2_extract_template: consumes whole words until >= n_bits.
 * `bits` must hold n_bits+15 entries at least (plus run overshoot is clipped here). */
static const uint16_t* wah_extract(const uint16_t* w, const uint16_t* end, uint8_t* bits, uint64_t n_bits,
                                   uint64_t cap, uint64_t* ones) {
    uint64_t pos = 0, cnt = 0;
    while (pos < n_bits) {
        uint16_t word;
        if (w >= end) word = 0x8000u | 0x3FFF; /* ran off the image: behave as zeros (reference would read garbage) */
        else memcpy(&word, w, 2);
        if (word & 0x8000u) {
            uint64_t len = (uint64_t)(word & 0x3FFFu) * 15;
            uint8_t v = (word & 0x4000u) ? 1 : 0;
            for (uint64_t i = pos; i < pos + len && i < cap; ++i) bits[i] = v;
            if (v) cnt += len;
            pos += len;
            if (len == 0 && w >= end) break;
        } else {
            for (unsigned j = 0; j < 15; ++j) {
                uint8_t v = (word >> j) & 1;
                if (pos + j < cap) bits[pos + j] = v;
                cnt += v;
            }
            pos += 15;
        }
        w++;
    }
    if (ones) *ones = cnt;
    return w;
}
uint64_t xo_wah_decode_bits(const uint16_t* wah, uint64_t n_words, uint64_t n_bits, uint8_t* bits, uint64_t* ones) {
    const uint16_t* e = wah_extract(wah, wah + n_words, bits, n_bits, n_bits + 15, ones);
    return (uint64_t)(e - wah);
}
/* wah.hpp:159-174 wah2_advance_pointer */
static const uint16_t* wah_skip(const uint16_t* w, const uint16_t* end, uint64_t n_bits) {
    uint64_t pos = 0;
    while (pos < n_bits && w < end) {
        uint16_t word; memcpy(&word, w, 2);
        pos += (word & 0x8000u) ? (uint64_t)(word & 0x3FFFu) * 15 : 15;
        w++;
    }
    return w;
}

/* ------------------------------------------------------------------ unordered_map order
 * The reference writes both per-block dictionaries by iterating a
 * std::unordered_map<uint32_t,uint32_t> (interfaces.hpp:37-54, gt_block.hpp:461,464-510), so
 * the byte layout follows libstdc++'s _Hashtable: identity hash, _Prime_rehash_policy
 * (first growth to 13 buckets, then the prime >= 2*buckets), nodes of an empty bucket are
 * linked at the list head, nodes of a non-empty bucket right after that bucket's
 * before-node; rehash re-links in iteration order with the same rule.               */
static const uint32_t k_primes[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71,
                                    73, 79, 83, 89, 97, 103, 109, 113, 127, 137, 139, 149, 157, 167, 179, 193,
                                    199, 211, 227, 241, 257, 277, 293, 313, 337, 359, 383, 409, 439, 467, 503};
static uint32_t next_bkt(uint32_t n, uint32_t* next_resize) {
    static const uint8_t fast[] = {2, 2, 2, 3, 5, 5, 7, 7, 11, 11, 11, 11, 13, 13};
    uint32_t r = 0;
    if (n < 14) r = fast[n];
    else for (size_t i = 0; i < sizeof(k_primes) / sizeof(k_primes[0]); ++i) if (k_primes[i] >= n) { r = k_primes[i]; break; }
    if (!r) abort();
    *next_resize = r; /* floor(r * max_load_factor 1.0) */
    return r;
}
#define XO_MAXD 256
typedef struct { uint32_t key[XO_MAXD]; int next[XO_MAXD]; int n; int head; int bkt[512]; uint32_t nb; uint32_t next_resize; } umap_t;
/* bkt[i]: -2 = empty, -1 = before_begin, else index of the node preceding the bucket's first node */
static void umap_link(umap_t* m, int node, uint32_t b) {
    if (m->bkt[b] != -2) {
        int prev = m->bkt[b];
        if (prev == -1) { m->next[node] = m->head; m->head = node; }
        else { m->next[node] = m->next[prev]; m->next[prev] = node; }
    } else {
        m->next[node] = m->head;
        m->head = node;
        if (m->next[node] >= 0) m->bkt[m->key[m->next[node]] % m->nb] = node;
        m->bkt[b] = -1;
    }
}
static void umap_rehash(umap_t* m, uint32_t nb) {
    int order[XO_MAXD], cnt = 0;
    for (int p = m->head; p >= 0; p = m->next[p]) order[cnt++] = p;
    m->nb = nb;
    for (uint32_t i = 0; i < nb; ++i) m->bkt[i] = -2;
    m->head = -1;
    uint32_t bbegin = 0;
    for (int i = 0; i < cnt; ++i) {
        int p = order[i];
        uint32_t b = m->key[p] % nb;
        if (m->bkt[b] == -2) {
            m->next[p] = m->head;
            m->head = p;
            m->bkt[b] = -1;
            if (m->next[p] >= 0) m->bkt[bbegin] = p;
            bbegin = b;
        } else {
            int prev = m->bkt[b];
            if (prev == -1) { m->next[p] = m->head; m->head = p; }
            else { m->next[p] = m->next[prev]; m->next[prev] = p; }
        }
    }
}
static void umap_init(umap_t* m) { m->n = 0; m->head = -1; m->nb = 1; m->bkt[0] = -2; m->next_resize = 0; }
static void umap_insert(umap_t* m, uint32_t key) {
    for (int i = 0; i < m->n; ++i) if (m->key[i] == key) return; /* operator[] on an existing key */
    if ((uint32_t)m->n + 1 > m->next_resize) { /* _Prime_rehash_policy::_M_need_rehash */
        uint32_t want = (uint32_t)m->n + 1;
        if (!m->next_resize && want < 11) want = 11;
        if (want >= m->nb) {
            uint32_t target = want + 1 > m->nb * 2 ? want + 1 : m->nb * 2;
            umap_rehash(m, next_bkt(target, &m->next_resize));
        } else {
            m->next_resize = m->nb;
        }
    }
    int node = m->n++;
    if (node >= XO_MAXD) abort();
    m->key[node] = key;
    umap_link(m, node, key % m->nb);
}
void xo_unordered_order(const uint32_t* keys, uint32_t n, uint32_t* order_out) {
    umap_t m; umap_init(&m);
    for (uint32_t i = 0; i < n; ++i) umap_insert(&m, keys[i]);
    uint32_t k = 0;
    for (int p = m.head; p >= 0; p = m.next[p]) order_out[k++] = m.key[p];
}

/* ------------------------------------------------------------------ file-level parameters */
/* xcf.cpp:811-836 seek_default_phased (limit = 3, xcf.hpp:301) */
int xo_default_phased(const int32_t* gt, const uint64_t* rec_off, const int32_t* ngt, uint64_t n_records,
                      uint64_t n_samples) {
    uint64_t counts[2] = {0, 0};
    for (uint64_t r = 0; r < n_records && r < 3; ++r) {
        uint64_t p = n_samples ? (uint64_t)ngt[r] / n_samples : 0;
        if (p == 1) return 0;
        for (uint64_t i = 0; i < n_samples; ++i) counts[gt_phased(gt[rec_off[r] + i * p + 1])]++;
    }
    return counts[0] > counts[1] ? 0 : 1;
}
/* gt_compressor_new.hpp:98-99 */
uint64_t xo_mac_threshold(uint64_t n_samples, uint64_t first_record_ploidy, double maf) {
    return (uint64_t)((double)(n_samples * first_record_ploidy) * maf);
}

/* ------------------------------------------------------------------ GT block encoder */
enum { K_BCF_LINES = 0, K_BIN_LINES = 1, K_MAX_PLOIDY = 2, K_DEF_PHASING = 3, K_WEIRD_STRAT = 4,
       K_LINE_SORT = 0x10, K_LINE_SELECT = 0x11, K_LINE_HAPLOID = 0x12, K_LINE_MISSING = 0x16,
       K_LINE_PHASE = 0x17, K_LINE_EOV = 0x18, K_MAT_WAH = 0x20, K_MAT_SPARSE = 0x21, K_MAT_MISSING = 0x26,
       K_MAT_PHASE = 0x27, K_MAT_EOV = 0x28, K_MAT_MISSING_SPARSE = 0x36, K_MAT_EOV_SPARSE = 0x38 }; /* gt_block.hpp:36-60 */

typedef struct {
    uint64_t n_samples, mac_thr;
    int default_phasing, aet; /* aet: bytes of A_T (xsi_factory.hpp:425) */
    int wah_missing;          /* --wah-encode-missing: weirdness_strat = WS_WAH (gt_block.hpp:174-176) */
    buf_t miss_wah, eov_wah;  /* WAH lines of the missing / end-of-vector predicates (gt_block.hpp:340-372) */
    uint32_t *a, *b;          /* PBWT order, 2*n_samples entries (gt_block.hpp:171,179) */
    uint32_t bcf_lines, bin_lines, max_ploidy;
    int missing_found, eov_found, phase_found, haploid_found;
    buf_t is_wah, has_missing, has_eov, has_phase, haploid, n_alt; /* one byte per flag / u32 per record for n_alt */
    buf_t wah, sparse, miss, eov, phase;
} gtblock_t;

static void gtblock_init(gtblock_t* g, uint64_t n_samples, uint64_t mac_thr, int default_phasing, int aet) {
    memset(g, 0, sizeof(*g));
    g->n_samples = n_samples; g->mac_thr = mac_thr; g->default_phasing = default_phasing; g->aet = aet;
    g->a = (uint32_t*)malloc(sizeof(uint32_t) * (2 * n_samples + 1));
    g->b = (uint32_t*)malloc(sizeof(uint32_t) * (2 * n_samples + 1));
    for (uint64_t i = 0; i < 2 * n_samples; ++i) g->a[i] = (uint32_t)i;
    g->max_ploidy = 1; /* gt_block.hpp:168 */
}
static void gtblock_free(gtblock_t* g) {
    free(g->a); free(g->b);
    free(g->is_wah.p); free(g->has_missing.p); free(g->has_eov.p); free(g->has_phase.p); free(g->haploid.p); free(g->n_alt.p);
    free(g->wah.p); free(g->sparse.p); free(g->miss.p); free(g->eov.p); free(g->phase.p);
    free(g->miss_wah.p); free(g->eov_wah.p);
}
static void put_index(buf_t* b, int aet, uint32_t v) { if (aet == 2) buf_u16(b, (uint16_t)v); else buf_u32(b, v); }

/* block.hpp:54-99 Sparse / SparseGtLine.  kind 0: allele == key, 1: missing, 2: end-of-vector */
static void sparse_line(buf_t* out, int aet, const int32_t* gt, uint32_t ngt, int kind, int32_t key, int set_msb) {
    size_t at = out->n;
    put_index(out, aet, 0);
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < ngt; ++i) {
        int hit = kind == 0 ? (gt_allele(gt[i]) == key) : kind == 1 ? gt_missing(gt[i]) : (gt[i] == BCF_VECTOR_END_I32);
        if (hit) { put_index(out, aet, i); cnt++; }
    }
    if (aet == 2) { uint16_t c = (uint16_t)cnt; if (set_msb) c |= 0x8000u; memcpy(out->p + at, &c, 2); }
    else { uint32_t c = cnt; if (set_msb) c |= 0x80000000u; memcpy(out->p + at, &c, 4); }
}

/* gt_block.hpp:279-406 encode_line (+ scan_genotypes :207-269). Returns 0 or <0. */
static int gtblock_encode_line(gtblock_t* g, const int32_t* gt, uint32_t ngt, uint32_t n_allele) {
    const uint64_t S = g->n_samples;
    const uint32_t P = S ? (uint32_t)(ngt / S) : 0;
    if (P > g->max_ploidy) g->max_ploidy = P;
    uint8_t hap = (P == 1);
    if (hap) g->haploid_found = 1;
    buf_put(&g->haploid, &hap, 1);
    uint64_t* cnt = (uint64_t*)calloc(n_allele ? n_allele : 1, sizeof(uint64_t));
    uint8_t has_missing = 0, has_eov = 0, has_phase = 0;
    for (uint64_t i = 0; i < S; ++i) {
        for (uint32_t j = 0; j < P; ++j) {
            int32_t v = gt[i * P + j];
            if (j && gt_phased(v) != g->default_phasing) has_phase = 1;
            if (gt_missing(v)) has_missing = 1;
            else if (v == BCF_VECTOR_END_I32) has_eov = 1;
            else {
                int32_t al = gt_allele(v);
                if (al < 0 || (uint32_t)al >= n_allele) { free(cnt); return -3; } /* "Unknown allele error !" */
                cnt[al]++;
            }
        }
    }
    if (has_missing) g->missing_found = 1;
    if (has_eov) g->eov_found = 1;
    if (has_phase) g->phase_found = 1;
    buf_put(&g->has_missing, &has_missing, 1);
    buf_put(&g->has_eov, &has_eov, 1);
    buf_put(&g->has_phase, &has_phase, 1);
    uint32_t nalt = n_allele ? n_allele - 1 : 0;
    buf_put(&g->n_alt, &nalt, 4);

    for (uint32_t alt = 1; alt < n_allele; ++alt) {
        uint64_t c = cnt[alt];
        uint64_t mac = c < (uint64_t)ngt - c ? c : (uint64_t)ngt - c;
        if (mac > g->mac_thr) {
            uint8_t one = 1; buf_put(&g->is_wah, &one, 1);
            if (P != 1 && P != 2) { free(cnt); return -4; } /* "PLOIDY ERROR" */
            wah_state st = {0, 0};
            /* wah.hpp:506-578: groups of 15 over the permuted order, zero padded tail */
            if (P == 2) {
                uint64_t groups = ((uint64_t)ngt + 14) / 15;
                for (uint64_t gi = 0; gi < groups; ++gi) {
                    uint16_t w = 0;
                    for (unsigned j = 0; j < 15; ++j) {
                        uint64_t k = gi * 15 + j;
                        if (k < ngt && gt_allele(gt[g->a[k]]) == (int32_t)alt) w |= (uint16_t)(1u << j);
                    }
                    wah_push_group(&g->wah, &st, w);
                }
                wah_finish(&g->wah, &st);
                /* internal_gt_record.hpp:32-53 pbwt_sort (loops over a.size()) */
                uint64_t u = 0, v = 0;
                for (uint64_t i = 0; i < 2 * S; ++i) {
                    uint32_t h = g->a[i];
                    if (gt_allele(gt[h]) != (int32_t)alt) g->a[u++] = h; else g->b[v++] = h;
                }
                memcpy(g->a + u, g->b, v * sizeof(uint32_t));
            } else {
                /* interfaces.hpp:318-333 haploid_rearrangement_from_diploid, then WAH over a1 */
                uint32_t* a1 = (uint32_t*)malloc(sizeof(uint32_t) * (S + 1));
                uint64_t n1 = 0;
                for (uint64_t i = 0; i < 2 * S; ++i) if ((g->a[i] & 1) == 0) a1[n1++] = g->a[i] / 2;
                uint64_t groups = ((uint64_t)ngt + 14) / 15;
                for (uint64_t gi = 0; gi < groups; ++gi) {
                    uint16_t w = 0;
                    for (unsigned j = 0; j < 15; ++j) {
                        uint64_t k = gi * 15 + j;
                        if (k < ngt && gt_allele(gt[a1[k]]) == (int32_t)alt) w |= (uint16_t)(1u << j);
                    }
                    wah_push_group(&g->wah, &st, w);
                }
                wah_finish(&g->wah, &st);
                free(a1);
                /* internal_gt_record.hpp:55-59 pbwt_sort1: V_LEN_RATIO = 2 */
                uint64_t u = 0, v = 0;
                for (uint64_t i = 0; i < 2 * S; ++i) {
                    uint32_t h = g->a[i];
                    if (gt_allele(gt[h / 2]) != (int32_t)alt) g->a[u++] = h; else g->b[v++] = h;
                }
                memcpy(g->a + u, g->b, v * sizeof(uint32_t));
            }
        } else {
            uint8_t zero = 0; buf_put(&g->is_wah, &zero, 1);
            int32_t sparse_allele = (c == mac) ? (int32_t)alt : 0; /* gt_block.hpp:318-321 */
            sparse_line(&g->sparse, g->aet, gt, ngt, 0, sparse_allele, sparse_allele == 0);
        }
        g->bin_lines++;
    }
    if (has_missing) sparse_line(&g->miss, g->aet, gt, ngt, 1, 0, 0); /* gt_block.hpp:330-333 */
    if (has_eov) sparse_line(&g->eov, g->aet, gt, ngt, 2, 0, 0);      /* gt_block.hpp:335-338 */
    if (g->wah_missing) {
        /* gt_block.hpp:340-372 with WS_WAH: a_weirdness stays the identity (it is only sorted under WS_PBWT_WAH, :377),
         * and for an all-haploid line a1 = haploid_rearrangement_from_diploid(identity) is the identity over
         * ngt = n_samples entries, so both predicates are WAH encoded in natural order over ngt bits. */
        for (int kind = 1; kind <= 2; ++kind) {
            if (!(kind == 1 ? has_missing : has_eov)) continue;
            buf_t* dst = kind == 1 ? &g->miss_wah : &g->eov_wah;
            wah_state st = {0, 0};
            uint64_t groups = ((uint64_t)ngt + 14) / 15;
            for (uint64_t gi = 0; gi < groups; ++gi) {
                uint16_t w = 0;
                for (unsigned j = 0; j < 15; ++j) {
                    uint64_t k = gi * 15 + j;
                    if (k < ngt && (kind == 1 ? gt_missing(gt[k]) : gt[k] == BCF_VECTOR_END_I32)) w |= (uint16_t)(1u << j);
                }
                wah_push_group(dst, &st, w);
            }
            wah_finish(dst, &st);
        }
    }
    if (has_phase) {                                                    /* gt_block.hpp:398-401, wah.hpp:441-501 */
        wah_state st = {0, 0};
        uint64_t groups = ((uint64_t)ngt + 14) / 15;
        for (uint64_t gi = 0; gi < groups; ++gi) {
            uint16_t w = 0;
            for (unsigned j = 0; j < 15; ++j) {
                uint64_t k = gi * 15 + j;
                if (k < ngt && (k & 1) && gt_phased(gt[k]) != g->default_phasing) w |= (uint16_t)(1u << j);
            }
            wah_push_group(&g->phase, &st, w);
        }
        wah_finish(&g->phase, &st);
    }
    g->bcf_lines++;
    free(cnt);
    return 0;
}

/* gt_block.hpp:650-666 reindex_binary_vector_from_bcf_to_binary_lines, then WAH (:676-679) */
static void put_reindexed(buf_t* out, const gtblock_t* g, const buf_t* flags) {
    uint8_t* v = (uint8_t*)calloc(g->bin_lines + 1, 1);
    uint32_t o = 0;
    for (uint32_t i = 0; i < g->bcf_lines; ++i) {
        uint32_t nalt; memcpy(&nalt, g->n_alt.p + 4 * (size_t)i, 4);
        if (o < g->bin_lines) v[o] = flags->p[i];
        o++; /* result[binary_offset++] = v[i] happens even for n_alt == 0 (reference quirk) */
        if (nalt > 1) o += nalt - 1;
    }
    wah_encode_bytes(out, v, g->bin_lines);
    free(v);
}

/* gt_block.hpp:185-204 write_to_stream (+ fill_dictionary :464-510, write_writables :512-647) */
static void gtblock_write(const gtblock_t* g, buf_t* out) {
    const size_t start = out->n;
    uint32_t keys[32]; uint32_t nk = 0;
    keys[nk++] = K_BCF_LINES; keys[nk++] = K_BIN_LINES; keys[nk++] = K_MAX_PLOIDY; keys[nk++] = K_DEF_PHASING;
    keys[nk++] = K_WEIRD_STRAT; keys[nk++] = K_LINE_SORT; keys[nk++] = K_LINE_SELECT; keys[nk++] = K_MAT_WAH;
    keys[nk++] = K_MAT_SPARSE;
    if (g->missing_found) { keys[nk++] = K_LINE_MISSING; keys[nk++] = K_MAT_MISSING; keys[nk++] = K_MAT_MISSING_SPARSE; }
    if (g->eov_found) { keys[nk++] = K_LINE_EOV; keys[nk++] = K_MAT_EOV; keys[nk++] = K_MAT_EOV_SPARSE; }
    if (g->phase_found) { keys[nk++] = K_LINE_PHASE; keys[nk++] = K_MAT_PHASE; }
    if (g->haploid_found) keys[nk++] = K_LINE_HAPLOID;
    uint32_t order[32];
    xo_unordered_order(keys, nk, order);
    uint32_t val[0x40];
    memset(val, 0xFF, sizeof(val));
    val[K_BCF_LINES] = g->bcf_lines; val[K_BIN_LINES] = g->bin_lines; val[K_MAX_PLOIDY] = g->max_ploidy;
    val[K_DEF_PHASING] = (uint32_t)g->default_phasing;
    val[K_WEIRD_STRAT] = g->wah_missing ? 1 : 2; /* WS_WAH (gt_block.hpp:174-176) / WS_SPARSE (:417) */
    buf_u32(out, 0xFFFFFFFFu);
    buf_u32(out, nk);
    const size_t dict_at = out->n;
    buf_zero(out, (size_t)nk * 8);
    val[K_LINE_SORT] = val[K_LINE_SELECT] = (uint32_t)(out->n - start);
    wah_encode_bytes(out, g->is_wah.p, g->bin_lines);
    val[K_MAT_WAH] = (uint32_t)(out->n - start);
    buf_put(out, g->wah.p, g->wah.n);
    val[K_MAT_SPARSE] = (uint32_t)(out->n - start);
    buf_put(out, g->sparse.p, g->sparse.n);
    if (g->missing_found) {
        val[K_LINE_MISSING] = (uint32_t)(out->n - start);
        put_reindexed(out, g, &g->has_missing);
        if (g->wah_missing) { val[K_MAT_MISSING] = (uint32_t)(out->n - start); buf_put(out, g->miss_wah.p, g->miss_wah.n); } /* :574-576 */
        else { val[K_MAT_MISSING_SPARSE] = (uint32_t)(out->n - start); buf_put(out, g->miss.p, g->miss.n); }
    }
    if (g->eov_found) {
        val[K_LINE_EOV] = (uint32_t)(out->n - start);
        put_reindexed(out, g, &g->has_eov);
        if (g->wah_missing) { val[K_MAT_EOV] = (uint32_t)(out->n - start); buf_put(out, g->eov_wah.p, g->eov_wah.n); } /* :596-598 */
        else { val[K_MAT_EOV_SPARSE] = (uint32_t)(out->n - start); buf_put(out, g->eov.p, g->eov.n); }
    }
    if (g->phase_found) {
        val[K_LINE_PHASE] = (uint32_t)(out->n - start);
        put_reindexed(out, g, &g->has_phase);
        val[K_MAT_PHASE] = (uint32_t)(out->n - start);
        buf_put(out, g->phase.p, g->phase.n);
    }
    if (g->haploid_found) {
        val[K_LINE_HAPLOID] = (uint32_t)(out->n - start);
        wah_encode_bytes(out, g->haploid.p, g->bcf_lines); /* one bit per BCF line, gt_block.hpp:219-224,639-642 */
    }
    for (uint32_t i = 0; i < nk; ++i) {
        memcpy(out->p + dict_at + 8 * (size_t)i, &order[i], 4);
        memcpy(out->p + dict_at + 8 * (size_t)i + 4, &val[order[i]], 4);
    }
}

/* ------------------------------------------------------------------ .xsi writer */
/* compression.hpp:40-104 header_t, packed, 256 bytes */
static void put_header(uint8_t* h, uint8_t ploidy, uint8_t aet, int default_phased, uint64_t hap_samples,
                       uint64_t num_variants, uint32_t ss_rate, uint32_t n_ssas, uint64_t indices_off,
                       uint64_t samples_off, uint32_t rare_thr, uint64_t entries, uint64_t num_samples) {
    memset(h, 0, 256);
    uint32_t u32; uint64_t u64;
#define P32(off, v) do { u32 = (uint32_t)(v); memcpy(h + (off), &u32, 4); } while (0)
#define P64(off, v) do { u64 = (uint64_t)(v); memcpy(h + (off), &u64, 8); } while (0)
    P32(0, 0xaabbccddu); P32(4, 0xfeed1767u); P32(8, 5); /* xsi_factory.hpp:469 */
    h[12] = ploidy; h[13] = 4; h[14] = aet; h[15] = 2;
    h[16] = (uint8_t)((default_phased ? 1 : 0) << 2); /* special_bitset.default_phased */
    h[17] = 0x01;                                      /* specific_bitset.iota_ppa; zstd (bit 2) not produced here */
    P64(32, hap_samples); P64(40, num_variants); P32(48, 0); P32(52, 1); P32(56, ss_rate); P32(60, n_ssas);
    P64(64, 256); P64(72, indices_off); P64(80, samples_off); P32(88, 0xFFFFFFFFu); P32(92, 0xFFFFFFFFu);
    P32(96, rare_thr); P64(100, entries); P32(108, 0); P64(112, num_samples);
    P32(252, 0xfeed1767u);
#undef P32
#undef P64
}

/* xsi_factory.hpp:436-606 XsiFactoryExt as driven by gt_compressor_new.hpp:84-142 */
int xo_encode(const int32_t* gt, const uint64_t* rec_off, const int32_t* ngt, const int32_t* n_allele,
              uint64_t n_records, uint64_t n_samples, uint64_t block_len, uint64_t mac_threshold,
              int default_phased, const char* sample_names, uint8_t** out, uint64_t* out_len) {
    return xo_encode_opt(gt, rec_off, ngt, n_allele, n_records, n_samples, block_len, mac_threshold, default_phased, sample_names,
                         0, out, out_len);
}
/* wah_encode_missing: the reference's --wah-encode-missing (xsqueezeit.hpp:58, gt_block.hpp:174-176) */
int xo_encode_opt(const int32_t* gt, const uint64_t* rec_off, const int32_t* ngt, const int32_t* n_allele,
                  uint64_t n_records, uint64_t n_samples, uint64_t block_len, uint64_t mac_threshold,
                  int default_phased, const char* sample_names, int wah_encode_missing, uint8_t** out, uint64_t* out_len) {
    if (n_samples > 32767 && n_samples <= 65535) return -10; /* reference mixes uint16 a[] with >65535 haplotypes: undefined */
    const int aet_block = n_samples <= 65535 ? 2 : 4;     /* xsi_factory.hpp:425 */
    const int aet_header = n_samples * 2 <= 65535 ? 2 : 4; /* gt_compressor_new.hpp:182 */
    buf_t f = {0, 0, 0};
    buf_zero(&f, 256);
    uint64_t* indices = (uint64_t*)malloc(sizeof(uint64_t) * (n_records / (block_len ? block_len : 1) + 2));
    uint64_t nblocks = 0, num_variants = 0, max_ploidy = 0;
    gtblock_t g; int have = 0, rc = 0;
    for (uint64_t r = 0; r < n_records; ++r) {
        if (r % block_len == 0) { /* check_flush_block, xsi_factory.hpp:527-539 */
            if (have) {
                indices[nblocks++] = f.n;
                buf_u32(&f, 0xFFFFFFFFu); buf_u32(&f, 1); buf_u32(&f, 256); buf_u32(&f, 16); /* interfaces.hpp:176-238 */
                gtblock_write(&g, &f);
                while (f.n % 4) buf_zero(&f, 1); /* interfaces.hpp:254-263 */
                gtblock_free(&g);
            }
            gtblock_init(&g, n_samples, mac_threshold, default_phased, aet_block);
            g.wah_missing = wah_encode_missing ? 1 : 0;
            have = 1;
        }
        uint64_t lp = n_samples ? (uint64_t)ngt[r] / n_samples : 0;
        if (lp > max_ploidy) { if (lp > 2) { rc = -5; break; } max_ploidy = lp; }
        rc = gtblock_encode_line(&g, gt + rec_off[r], (uint32_t)ngt[r], (uint32_t)n_allele[r]);
        if (rc) break;
        num_variants += (uint64_t)n_allele[r] - 1;
    }
    if (rc) { if (have) gtblock_free(&g); free(indices); free(f.p); return rc; }
    if (have && g.bcf_lines) { /* finalize_file, xsi_factory.hpp:543-606 */
        indices[nblocks++] = f.n;
        buf_u32(&f, 0xFFFFFFFFu); buf_u32(&f, 1); buf_u32(&f, 256); buf_u32(&f, 16);
        gtblock_write(&g, &f);
        while (f.n % 4) buf_zero(&f, 1);
    }
    if (have) gtblock_free(&g);
    while (f.n % 8) buf_zero(&f, 1);
    uint64_t indices_off = f.n;
    buf_put(&f, indices, nblocks * 8);
    uint64_t samples_off = f.n;
    const char* sn = sample_names;
    for (uint64_t i = 0; i < n_samples; ++i) {
        if (sn) { size_t l = strlen(sn) + 1; buf_put(&f, sn, l); sn += l; }
        else { char tmp[32]; int l = 0; uint64_t v = i; char d[24]; int nd = 0; do { d[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
               tmp[l++] = 'S'; while (nd) tmp[l++] = d[--nd]; tmp[l++] = 0; buf_put(&f, tmp, (size_t)l); }
    }
    uint32_t bl32 = (uint32_t)block_len;
    put_header(f.p, (uint8_t)max_ploidy, (uint8_t)aet_header, default_phased, n_samples * max_ploidy, num_variants,
               bl32, bl32 ? (uint32_t)((n_records + bl32 - 1) / bl32) : 0, indices_off, samples_off,
               (uint32_t)mac_threshold, n_records, n_samples);
    free(indices);
    *out = f.p; *out_len = f.n;
    return 0;
}
void xo_free(void* p) { free(p); }

/* ------------------------------------------------------------------ reader (Accessor side) */
typedef struct {
    const uint8_t* base; /* GT block start */
    uint32_t bcf_lines, bin_lines;
    int default_phasing;
    uint8_t *is_wah, *is_sort, *has_missing, *has_eov, *has_phase, *haploid; /* one byte per binary line (+15) */
    int weird, phase;
    int ws; /* KEY_WEIRDNESS_STRATEGY: 2 = WS_SPARSE, 1 = WS_WAH (missing / end-of-vector lines as natural-order WAH) */
    const uint8_t *wah0, *sparse0, *miss0, *eov0, *phase0;
    const uint8_t *wah_p, *sparse_p, *miss_p, *eov_p, *phase_p;
    uint64_t pos, weird_pos, phase_pos;
    uint32_t *a, *b;
} cursor_t;

struct xo_reader {
    const uint8_t* file; uint64_t len;
    uint32_t version; uint8_t ploidy, aet; int zstd;
    uint64_t hap_samples, num_samples, indices_off, n_blocks_hint;
    uint64_t N_SAMPLES, N_HAPS;
    int64_t cur_block;
    cursor_t c;
    uint8_t *y, *x;
    uint64_t allele_counts[256]; uint64_t n_counts;
    uint64_t ones; int sparse_negated;
};

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static uint32_t rd_idx(const uint8_t* p, int aet) { if (aet == 2) { uint16_t v; memcpy(&v, p, 2); return v; } return rd32(p); }

/* interfaces.hpp:77-90 read_dictionary into std::map: later duplicates overwrite */
static uint32_t dict_get(const uint8_t* blk, uint32_t key, int* found) {
    uint32_t n = rd32(blk + 4), val = 0xFFFFFFFFu; int f = 0;
    for (uint32_t i = 0; i < n; ++i) if (rd32(blk + 8 + 8 * (size_t)i) == key) { val = rd32(blk + 12 + 8 * (size_t)i); f = 1; }
    if (found) *found = f;
    return val;
}

xo_reader* xo_open(const uint8_t* file, uint64_t len) {
    if (len < 256) return NULL;
    if (rd32(file + 4) != 0xfeed1767u || rd32(file + 252) != 0xfeed1767u) return NULL; /* accessor.cpp:37-41 */
    if (rd32(file) != 0xaabbccddu) return NULL;
    xo_reader* r = (xo_reader*)calloc(1, sizeof(xo_reader));
    r->file = file; r->len = len;
    r->version = rd32(file + 8);
    if (r->version != 4 && r->version != 5) { free(r); return NULL; } /* accessor_internals_new.hpp:782-785 */
    r->ploidy = file[12]; r->aet = file[14]; r->zstd = (file[17] >> 2) & 1;
    if (r->zstd || (r->aet != 2 && r->aet != 4) || r->ploidy == 0) { free(r); return NULL; }
    r->hap_samples = rd64(file + 32); r->num_samples = rd64(file + 112);
    r->indices_off = rd64(file + 72);
    r->N_SAMPLES = r->num_samples;
    r->N_HAPS = r->N_SAMPLES ? r->N_SAMPLES * 2 : r->hap_samples; /* accessor_internals_new.hpp:53 */
    r->cur_block = -1;
    r->y = (uint8_t*)calloc(r->N_HAPS + 64, 1);
    r->x = (uint8_t*)calloc(r->N_HAPS + 64, 1);
    return r;
}
static void cursor_free(cursor_t* c) {
    free(c->is_wah); free(c->is_sort); free(c->has_missing); free(c->has_eov); free(c->has_phase); free(c->haploid);
    free(c->a); free(c->b);
    memset(c, 0, sizeof(*c));
}
void xo_close(xo_reader* r) { if (!r) return; cursor_free(&r->c); free(r->y); free(r->x); free(r); }
uint64_t xo_hap_samples(const xo_reader* r) { return r->hap_samples; }
uint64_t xo_num_blocks(const xo_reader* r) { return (rd64(r->file + 80) - r->indices_off) / (r->version >= 5 ? 8 : 4); }

/* accessor_internals_new.hpp:591-604 fill_bool_vector_from_1d_dict_key */
static uint8_t* load_flags(const xo_reader* r, const uint8_t* blk, uint32_t key, uint64_t size, int* present) {
    int f; uint32_t off = dict_get(blk, key, &f);
    *present = 0;
    if (!f || off == 0xFFFFFFFFu) return NULL;
    uint8_t* v = (uint8_t*)calloc(size + 32, 1);
    wah_extract((const uint16_t*)(blk + off), (const uint16_t*)(r->file + r->len), v, size, size + 15, NULL);
    *present = 1;
    return v;
}
static const uint8_t* dict_ptr(const uint8_t* blk, uint32_t key) {
    int f; uint32_t off = dict_get(blk, key, &f);
    return (!f || off == 0xFFFFFFFFu) ? NULL : blk + off;
}
static void cursor_reset(xo_reader* r) { /* accessor_internals_new.hpp:386-405 */
    cursor_t* c = &r->c;
    for (uint64_t i = 0; i < r->N_HAPS; ++i) c->a[i] = (uint32_t)i;
    c->pos = 0; c->wah_p = c->wah0; c->sparse_p = c->sparse0;
    c->weird_pos = 0; c->miss_p = c->miss0; c->eov_p = c->eov0;
    c->phase_pos = 0; c->phase_p = c->phase0;
}
/* accessor_internals_new.hpp:845-893 set_block_ptr + :52-148 DecompressPointerGTBlock ctor */
static int open_block(xo_reader* r, uint64_t block_id) {
    cursor_free(&r->c);
    uint64_t off = r->version >= 5 ? rd64(r->file + r->indices_off + 8 * block_id) : rd32(r->file + r->indices_off + 4 * block_id);
    const uint8_t* outer = r->file + off;
    int f; uint32_t gt_off = dict_get(outer, 256, &f);
    if (!f) return -1;
    const uint8_t* blk = outer + gt_off;
    cursor_t* c = &r->c;
    c->base = blk;
    c->bcf_lines = dict_get(blk, K_BCF_LINES, &f); if (!f) return -1;
    c->bin_lines = dict_get(blk, K_BIN_LINES, &f); if (!f) return -1;
    uint32_t dp = dict_get(blk, K_DEF_PHASING, &f); if (!f) return -1;
    c->default_phasing = (dp == 1) ? 1 : 0; /* accessor_internals_new.hpp:77-81 */
    uint32_t ws = dict_get(blk, K_WEIRD_STRAT, &f);
    if (!f || (ws != 2 && ws != 1)) return -2; /* WS_SPARSE (default) and WS_WAH (--wah-encode-missing); WS_PBWT_WAH is not reachable from the CLI */
    c->ws = (int)ws;
    int p;
    c->is_wah = load_flags(r, blk, K_LINE_SELECT, c->bin_lines, &p); if (!p) return -1;
    c->is_sort = load_flags(r, blk, K_LINE_SORT, c->bin_lines, &p);
    if (!p) { c->is_sort = (uint8_t*)malloc(c->bin_lines + 32); memcpy(c->is_sort, c->is_wah, c->bin_lines); }
    int pm, pe;
    c->has_missing = load_flags(r, blk, K_LINE_MISSING, c->bin_lines, &pm);
    c->has_eov = load_flags(r, blk, K_LINE_EOV, c->bin_lines, &pe);
    c->weird = pm || pe;
    c->has_phase = load_flags(r, blk, K_LINE_PHASE, c->bin_lines, &c->phase);
    c->haploid = load_flags(r, blk, K_LINE_HAPLOID, c->bin_lines, &p);
    if (!p) c->haploid = (uint8_t*)calloc(c->bin_lines + 32, 1);
    c->wah0 = dict_ptr(blk, K_MAT_WAH); c->sparse0 = dict_ptr(blk, K_MAT_SPARSE);
    if (ws == 2) { c->miss0 = dict_ptr(blk, K_MAT_MISSING_SPARSE); c->eov0 = dict_ptr(blk, K_MAT_EOV_SPARSE); }
    else { c->miss0 = dict_ptr(blk, K_MAT_MISSING); c->eov0 = dict_ptr(blk, K_MAT_EOV); } /* accessor_internals_new.hpp:128-137 */
    c->phase0 = dict_ptr(blk, K_MAT_PHASE);
    c->a = (uint32_t*)malloc(sizeof(uint32_t) * (r->N_HAPS + 1));
    c->b = (uint32_t*)malloc(sizeof(uint32_t) * (r->N_HAPS + 1));
    cursor_reset(r);
    r->cur_block = (int64_t)block_id;
    return 0;
}
static uint64_t cur_n(const xo_reader* r) { return r->c.haploid[r->c.pos] ? r->N_SAMPLES : r->N_HAPS; }
/* accessor_internals_new.hpp:639-653 sparse_advance_pointer / :619-637 sparse_extract (header part) */
static const uint8_t* sparse_header(xo_reader* r, const uint8_t* p, uint32_t* num) {
    uint32_t n = rd_idx(p, r->aet);
    uint32_t msb = r->aet == 2 ? 0x8000u : 0x80000000u;
    r->sparse_negated = (n & msb) != 0;
    n &= ~msb;
    *num = n;
    uint64_t N = cur_n(r);
    r->ones = r->sparse_negated ? N - n : n;
    return p + r->aet;
}
/* accessor_internals_new.hpp:548-589 update_a_if_needed / private_pbwt_sort / gt_block.hpp:124-136 */
static void update_a(xo_reader* r) {
    cursor_t* c = &r->c;
    if (!c->is_sort[c->pos]) return;
    uint64_t u = 0, v = 0;
    if (!c->haploid[c->pos]) {
        for (uint64_t i = 0; i < r->N_HAPS; ++i) { if (!r->y[i]) c->a[u++] = c->a[i]; else c->b[v++] = c->a[i]; }
    } else {
        uint64_t k = 0;
        memset(r->x, 0, r->N_SAMPLES + 1);
        for (uint64_t i = 0; i < r->N_HAPS; ++i) if ((c->a[i] & 1) == 0) { if (k < r->N_SAMPLES) r->x[c->a[i] / 2] = r->y[k]; k++; }
        for (uint64_t i = 0; i < r->N_SAMPLES * 2; ++i) { if (!r->x[c->a[i] / 2]) c->a[u++] = c->a[i]; else c->b[v++] = c->a[i]; }
    }
    memcpy(c->a + u, c->b, v * sizeof(uint32_t));
}
static const uint16_t* wah_skip(const uint16_t* w, const uint16_t* end, uint64_t n_bits);
/* accessor_internals_new.hpp:478-537: WS_SPARSE branch, and the WS_WAH branch (:492-501, no PBWT on a_weird) */
static void weird_advance(xo_reader* r, uint64_t steps, uint64_t N) {
    cursor_t* c = &r->c; uint32_t n;
    for (uint64_t i = 0; i < steps; ++i) {
        if (c->ws == 1) {
            const uint16_t* end = (const uint16_t*)(r->file + r->len);
            if (c->has_missing && c->has_missing[c->weird_pos]) c->miss_p = (const uint8_t*)wah_skip((const uint16_t*)c->miss_p, end, N);
            if (c->has_eov && c->has_eov[c->weird_pos]) c->eov_p = (const uint8_t*)wah_skip((const uint16_t*)c->eov_p, end, N);
            c->weird_pos++;
            continue;
        }
        if (c->has_missing && c->has_missing[c->weird_pos]) { const uint8_t* p = c->miss_p; n = rd_idx(p, r->aet) & (r->aet == 2 ? 0x7FFFu : 0x7FFFFFFFu); c->miss_p = p + (size_t)r->aet * (1 + n); }
        if (c->has_eov && c->has_eov[c->weird_pos]) { const uint8_t* p = c->eov_p; n = rd_idx(p, r->aet) & (r->aet == 2 ? 0x7FFFu : 0x7FFFFFFFu); c->eov_p = p + (size_t)r->aet * (1 + n); }
        c->weird_pos++;
    }
}
/* accessor_internals_new.hpp:539-546 */
static void phase_advance(xo_reader* r, uint64_t steps, uint64_t N) {
    cursor_t* c = &r->c;
    for (uint64_t i = 0; i < steps; ++i) {
        if (c->has_phase && c->has_phase[c->phase_pos]) c->phase_p = (const uint8_t*)wah_skip((const uint16_t*)c->phase_p, (const uint16_t*)(r->file + r->len), N);
        c->phase_pos++;
    }
}
/* accessor_internals_new.hpp:154-196 seek */
static void cursor_seek(xo_reader* r, uint64_t position) {
    cursor_t* c = &r->c;
    if (c->pos == position) return;
    if (c->pos > position) cursor_reset(r);
    const uint16_t* end = (const uint16_t*)(r->file + r->len);
    while (c->pos < position) {
        uint64_t N = cur_n(r);
        if (c->is_wah[c->pos]) {
            if (c->is_sort[c->pos]) c->wah_p = (const uint8_t*)wah_extract((const uint16_t*)c->wah_p, end, r->y, N, r->N_HAPS + 15, NULL);
            else c->wah_p = (const uint8_t*)wah_skip((const uint16_t*)c->wah_p, end, N);
        } else {
            uint32_t n; const uint8_t* p = sparse_header(r, c->sparse_p, &n);
            c->sparse_p = p + (size_t)n * r->aet; /* a sparse line never sorts in v5 files; if flagged, y is stale as in the reference */
        }
        update_a(r);
        if (c->weird) weird_advance(r, 1, N);
        if (c->phase) phase_advance(r, 1, N);
        c->pos++;
    }
}

/* accessor_internals_new.hpp:198-384 fill_genotype_array_advance */
static int64_t fill_advance(xo_reader* r, int32_t* gt, uint64_t gt_size, uint64_t n_alleles) {
    cursor_t* c = &r->c;
    const uint16_t* end = (const uint16_t*)(r->file + r->len);
    const int DP = c->default_phasing;
    const uint64_t N = cur_n(r);
    const uint64_t START = c->pos;
    if (N > gt_size || n_alleles < 2 || n_alleles > 255) return -1;
    uint64_t total_alt = 0, n_missing = 0, n_eovs = 0;
    r->n_counts = n_alleles;
    memset(r->allele_counts, 0, sizeof(r->allele_counts));
    const int hap_line = c->haploid[c->pos];
    uint32_t* a1 = NULL;
    /* first ALT */
    if (!c->is_wah[c->pos]) {
        uint32_t n; const uint8_t* p = sparse_header(r, c->sparse_p, &n);
        int32_t dflt = r->sparse_negated ? 1 : 0, sp = r->sparse_negated ? 0 : 1;
        for (uint64_t i = 0; i < N; ++i) gt[i] = gt_unphased(dflt) | (int32_t)((i & 1) & (uint64_t)DP);
        for (uint32_t k = 0; k < n; ++k) { uint32_t i = rd_idx(p + (size_t)k * r->aet, r->aet); if (i < gt_size) gt[i] = gt_unphased(sp) | (int32_t)((i & 1) & (uint32_t)DP); }
        c->sparse_p = p + (size_t)n * r->aet;
    } else {
        c->wah_p = (const uint8_t*)wah_extract((const uint16_t*)c->wah_p, end, r->y, N, r->N_HAPS + 15, &r->ones);
        if (hap_line) {
            a1 = (uint32_t*)malloc(sizeof(uint32_t) * (r->N_SAMPLES + 1)); uint64_t k = 0;
            for (uint64_t i = 0; i < r->N_HAPS; ++i) if ((c->a[i] & 1) == 0) a1[k++] = c->a[i] / 2;
            for (uint64_t i = 0; i < N; ++i) gt[a1[i]] = gt_unphased(r->y[i]);
            free(a1); a1 = NULL;
        } else {
            for (uint64_t i = 0; i < N; ++i) gt[c->a[i]] = gt_unphased(r->y[i]) | (int32_t)((c->a[i] & 1) & (uint32_t)DP);
        }
    }
    r->allele_counts[1] = r->ones; total_alt = r->ones;
    update_a(r);
    c->pos++;
    for (uint64_t alt = 2; alt < n_alleles; ++alt) {
        if (!c->is_wah[c->pos]) {
            uint32_t n; const uint8_t* p = sparse_header(r, c->sparse_p, &n);
            if (r->sparse_negated) {
                for (uint64_t i = 0; i < N; ++i) if (gt_allele(gt[i]) == 0) gt[i] = gt_unphased((int32_t)alt) | (int32_t)((i & 1) & (uint64_t)DP);
                for (uint32_t k = 0; k < n; ++k) { uint32_t i = rd_idx(p + (size_t)k * r->aet, r->aet); if (i < gt_size && gt_allele(gt[i]) == (int32_t)alt) gt[i] = gt_unphased(0) | (int32_t)((i & 1) & (uint32_t)DP); }
            } else {
                for (uint32_t k = 0; k < n; ++k) { uint32_t i = rd_idx(p + (size_t)k * r->aet, r->aet); if (i < gt_size) gt[i] = gt_unphased((int32_t)alt) | (int32_t)((i & 1) & (uint32_t)DP); }
            }
            c->sparse_p = p + (size_t)n * r->aet;
        } else {
            c->wah_p = (const uint8_t*)wah_extract((const uint16_t*)c->wah_p, end, r->y, N, r->N_HAPS + 15, &r->ones);
            if (c->haploid[c->pos]) {
                a1 = (uint32_t*)malloc(sizeof(uint32_t) * (r->N_SAMPLES + 1)); uint64_t k = 0;
                for (uint64_t i = 0; i < r->N_HAPS; ++i) if ((c->a[i] & 1) == 0) a1[k++] = c->a[i] / 2;
                for (uint64_t i = 0; i < N; ++i) if (r->y[i]) gt[a1[i]] = gt_unphased(r->y[i]); /* sic: allele 1, :269 */
                free(a1); a1 = NULL;
            } else {
                for (uint64_t i = 0; i < N; ++i) if (r->y[i]) gt[c->a[i]] = gt_unphased((int32_t)alt) | (int32_t)((c->a[i] & 1) & (uint32_t)DP);
            }
        }
        r->allele_counts[alt] = r->ones; total_alt += r->ones;
        update_a(r);
        c->pos++;
    }
    if (c->weird && c->ws == 1) {
        /* accessor_internals_new.hpp:307-321, 330-341: extract the WAH line, a_weird is the identity under WS_WAH */
        if (c->has_missing && c->has_missing[START]) {
            uint64_t ones = 0;
            wah_extract((const uint16_t*)c->miss_p, end, r->x, N, r->N_HAPS + 15, &ones);
            n_missing = ones;
            for (uint64_t i = 0; i < N; ++i) if (r->x[i]) gt[i] = 0 | (int32_t)((i & 1) & (uint32_t)DP);
        }
        if (c->has_eov && c->has_eov[START]) {
            uint64_t ones = 0;
            wah_extract((const uint16_t*)c->eov_p, end, r->x, N, r->N_HAPS + 15, &ones);
            n_eovs = ones;
            for (uint64_t i = 0; i < N; ++i) if (r->x[i]) gt[i] = BCF_VECTOR_END_I32;
        }
        weird_advance(r, n_alleles - 1, N);
    } else if (c->weird) {
        if (c->has_missing && c->has_missing[START]) {
            uint32_t n = rd_idx(c->miss_p, r->aet) & (r->aet == 2 ? 0x7FFFu : 0x7FFFFFFFu);
            n_missing = n;
            for (uint32_t k = 0; k < n; ++k) { uint32_t i = rd_idx(c->miss_p + (size_t)(k + 1) * r->aet, r->aet); if (i < gt_size) gt[i] = 0 | (int32_t)((i & 1) & (uint32_t)DP); }
        }
        if (c->has_eov && c->has_eov[START]) {
            uint32_t n = rd_idx(c->eov_p, r->aet) & (r->aet == 2 ? 0x7FFFu : 0x7FFFFFFFu);
            n_eovs = n;
            for (uint32_t k = 0; k < n; ++k) { uint32_t i = rd_idx(c->eov_p + (size_t)(k + 1) * r->aet, r->aet); if (i < gt_size) gt[i] = BCF_VECTOR_END_I32; }
        }
        weird_advance(r, n_alleles - 1, N);
    }
    if (c->phase) {
        if (c->has_phase && c->has_phase[START]) {
            wah_extract((const uint16_t*)c->phase_p, end, r->x, N, r->N_HAPS + 15, NULL);
            for (uint64_t i = 0; i < N; ++i) if (r->x[i] && gt[i] != BCF_VECTOR_END_I32) gt[i] ^= (int32_t)(i & 1);
        }
        phase_advance(r, n_alleles - 1, N);
    }
    r->allele_counts[0] = N - (total_alt + n_missing + n_eovs);
    return (int64_t)N;
}

/* accessor_internals_new.hpp:722-745 */
int64_t xo_fill_genotype_array(xo_reader* r, int32_t* gt_arr, uint64_t gt_arr_size, uint64_t n_alleles, uint64_t position) {
    uint64_t block_id = (position & 0xFFFFFFFFull) >> 15;
    uint64_t offset = position & 0x7FFF;
    if (r->cur_block != (int64_t)block_id) { int rc = open_block(r, block_id); if (rc) { r->cur_block = -1; return rc; } }
    cursor_seek(r, offset);
    return fill_advance(r, gt_arr, gt_arr_size, n_alleles);
}
/* accessor_internals_new.hpp:407-440 fill_allele_counts_advance behind :747-752 fill_allele_counts: counts only.
 * As in the reference, allele_counts[0] = CURRENT_N_HAPS - total_alt (missing / end-of-vector entries are NOT
 * subtracted here, unlike fill_genotype_array) and the missing / phase cursors are not advanced. */
int64_t xo_fill_allele_counts(xo_reader* r, uint64_t n_alleles, uint64_t position) {
    uint64_t block_id = (position & 0xFFFFFFFFull) >> 15;
    uint64_t offset = position & 0x7FFF;
    if (r->cur_block != (int64_t)block_id) { int rc = open_block(r, block_id); if (rc) { r->cur_block = -1; return rc; } }
    cursor_seek(r, offset);
    cursor_t* c = &r->c;
    const uint16_t* end = (const uint16_t*)(r->file + r->len);
    if (n_alleles < 1 || n_alleles > 255) return -1;
    const uint64_t N = cur_n(r);
    uint64_t total_alt = 0;
    r->n_counts = n_alleles;
    memset(r->allele_counts, 0, sizeof(r->allele_counts));
    for (uint64_t alt = 1; alt < n_alleles; ++alt) {
        if (c->is_wah[c->pos]) {
            c->wah_p = (const uint8_t*)wah_extract((const uint16_t*)c->wah_p, end, r->y, N, r->N_HAPS + 15, &r->ones);
        } else {
            uint32_t n; const uint8_t* p = sparse_header(r, c->sparse_p, &n);
            c->sparse_p = p + (size_t)n * r->aet;
        }
        update_a(r);
        c->pos++;
        r->allele_counts[alt] = r->ones; total_alt += r->ones;
    }
    r->allele_counts[0] = N - total_alt;
    return (int64_t)n_alleles;
}
uint64_t xo_allele_counts(const xo_reader* r, uint64_t* out, uint64_t cap) {
    for (uint64_t i = 0; i < r->n_counts && i < cap; ++i) out[i] = r->allele_counts[i];
    return r->n_counts;
}
