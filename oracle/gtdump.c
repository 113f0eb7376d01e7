/* oracle/gtdump.c -- TEST INFRASTRUCTURE ONLY (fixture generation, runs only where
 * /root/reference and its vendored htslib exist).
 *
 * Dumps what the reference compressor sees for every record of a VCF/BCF:
 *   bcf_unpack + bcf_get_genotypes  (reference: bcf_traversal.cpp:9-11)
 * into two flat little-endian files:
 *   <out>.meta : u32 magic 'GTD1', u32 n_samples, u64 n_records, then per record
 *                u32 n_allele, u32 ngt
 *   <out>.gt   : the int32 genotype rows back to back (ngt each)
 * Sample names go to <out>.samples, one per line.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "htslib/vcf.h"
#include "htslib/hts.h"

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: gtdump in.{vcf,bcf} out_prefix\n"); return 2; }
    htsFile* fp = hts_open(argv[1], "r");
    if (!fp) { fprintf(stderr, "cannot open %s\n", argv[1]); return 1; }
    bcf_hdr_t* hdr = bcf_hdr_read(fp);
    if (!hdr) { fprintf(stderr, "cannot read header\n"); return 1; }
    char path[4096];
    snprintf(path, sizeof path, "%s.meta", argv[2]);
    FILE* fm = fopen(path, "wb");
    snprintf(path, sizeof path, "%s.gt", argv[2]);
    FILE* fg = fopen(path, "wb");
    snprintf(path, sizeof path, "%s.samples", argv[2]);
    FILE* fs = fopen(path, "w");
    if (!fm || !fg || !fs) { fprintf(stderr, "cannot open outputs\n"); return 1; }
    uint32_t n_samples = (uint32_t)bcf_hdr_nsamples(hdr);
    for (uint32_t i = 0; i < n_samples; ++i) fprintf(fs, "%s\n", hdr->samples[i]);
    fclose(fs);
    uint32_t magic = 0x31445447u; /* 'GTD1' */
    uint64_t n_records = 0;
    fwrite(&magic, 4, 1, fm);
    fwrite(&n_samples, 4, 1, fm);
    fwrite(&n_records, 8, 1, fm);
    bcf1_t* rec = bcf_init();
    int32_t* gt = NULL;
    int ngt_cap = 0;
    while (bcf_read(fp, hdr, rec) == 0) {
        bcf_unpack(rec, BCF_UN_STR);
        int ngt = bcf_get_genotypes(hdr, rec, &gt, &ngt_cap);
        if (ngt < 0) ngt = 0;
        uint32_t m[2] = {(uint32_t)rec->n_allele, (uint32_t)ngt};
        fwrite(m, 4, 2, fm);
        fwrite(gt, 4, (size_t)ngt, fg);
        n_records++;
    }
    fseek(fm, 8, SEEK_SET);
    fwrite(&n_records, 8, 1, fm);
    fclose(fm);
    fclose(fg);
    free(gt);
    bcf_destroy(rec);
    bcf_hdr_destroy(hdr);
    hts_close(fp);
    fprintf(stderr, "gtdump: %u samples, %llu records\n", n_samples, (unsigned long long)n_records);
    return 0;
}
