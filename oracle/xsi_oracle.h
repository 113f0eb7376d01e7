/* oracle/xsi_oracle.h -- TEST INFRASTRUCTURE ONLY.  See xsi_oracle.c. */
#ifndef XSI_ORACLE_H
#define XSI_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* WAH2-16 primitives (reference include/wah.hpp) */
uint64_t xo_wah_encode_bits(const uint8_t* bits, uint64_t n, uint16_t* out /* >= n/15+2 */);
uint64_t xo_wah_decode_bits(const uint16_t* wah, uint64_t n_words, uint64_t n_bits, uint8_t* bits /* >= n_bits+15 */,
                            uint64_t* ones /* may be NULL */);

/* libstdc++ unordered_map<uint32,uint32> iteration order after inserting keys[0..n) (distinct) */
void xo_unordered_order(const uint32_t* keys, uint32_t n, uint32_t* order_out);

/* file-level parameters (reference xcf.cpp:811-862, gt_compressor_new.hpp:98-99) */
int      xo_default_phased(const int32_t* gt, const uint64_t* rec_off, const int32_t* ngt,
                           uint64_t n_records, uint64_t n_samples);
uint64_t xo_mac_threshold(uint64_t n_samples, uint64_t first_record_ploidy, double maf);

/* whole-file encoder: returns 0 and a malloc'ed .xsi image, or <0 */
int xo_encode(const int32_t* gt, const uint64_t* rec_off, const int32_t* ngt, const int32_t* n_allele,
              uint64_t n_records, uint64_t n_samples, uint64_t block_len, uint64_t mac_threshold,
              int default_phased, const char* sample_names, uint8_t** out, uint64_t* out_len);
/* same with the reference's --wah-encode-missing (missing / end-of-vector lines as natural-order WAH, WS_WAH) */
int xo_encode_opt(const int32_t* gt, const uint64_t* rec_off, const int32_t* ngt, const int32_t* n_allele,
                  uint64_t n_records, uint64_t n_samples, uint64_t block_len, uint64_t mac_threshold,
                  int default_phased, const char* sample_names, int wah_encode_missing, uint8_t** out, uint64_t* out_len);
void xo_free(void* p);

/* reader / cursor decoder (reference Accessor + DecompressPointerGTBlock) */
typedef struct xo_reader xo_reader;
xo_reader* xo_open(const uint8_t* file, uint64_t len); /* borrows `file` (uncompressed blocks only) */
void       xo_close(xo_reader* r);
uint64_t   xo_hap_samples(const xo_reader* r);
uint64_t   xo_num_blocks(const xo_reader* r);
/* returns number of filled entries, <0 on error */
int64_t xo_fill_genotype_array(xo_reader* r, int32_t* gt_arr, uint64_t gt_arr_size, uint64_t n_alleles,
                               uint64_t position);
/* counts only (Accessor::fill_allele_counts); read them with xo_allele_counts. returns n_alleles, <0 on error */
int64_t xo_fill_allele_counts(xo_reader* r, uint64_t n_alleles, uint64_t position);
uint64_t xo_allele_counts(const xo_reader* r, uint64_t* out, uint64_t cap);

#ifdef __cplusplus
}
#endif
#endif
