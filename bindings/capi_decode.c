/* bindings/capi_decode.c -- timing loop over the reference's C API (include/c_api.h:38-93), the way a consumer such
 * as SHAPEIT4 or the reference's own c_api_test/main.c uses it: one c_xcf_get_genotypes call per record of the
 * `_var.bcf` companion.  Linked twice by bindings/Makefile: with the reference's accessor.o (CPU) and with
 * accessor_b200.o (the adapter, GPU).  Prints one line:
 *     records <R> genotypes <G> seconds <T> checksum <C> setup <seconds before the loop> teardown <seconds in c_xcf_delete>
 * where the checksum (four multiply-add lanes over every returned int32; XSI_CAPI_NO_CHECKSUM=1 skips it) lets the two builds be compared without storing rows.
 * usage: capi_decode <file.xsi_var.bcf | file.bcf> [max_records]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include "c_api.h"
#include "synced_bcf_reader.h"
#include "vcf.h"

static double now(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file [max_records]\n", argv[0]); return 2; }
    const long max_records = argc > 2 ? atol(argv[2]) : -1;
    const double ts = now();
    c_xcf* x = c_xcf_new();
    bcf_srs_t* sr = bcf_sr_init();
    if (!bcf_sr_add_reader(sr, argv[1])) { fprintf(stderr, "could not load %s\n", argv[1]); return 1; }
    c_xcf_add_readers(x, sr);
    int* gt = NULL;
    int ngt_arr = 0;
    long records = 0;
    uint64_t genotypes = 0, h0 = 1, h1 = 2, h2 = 3, h3 = 4;
    const int do_sum = !(getenv("XSI_CAPI_NO_CHECKSUM") && atoi(getenv("XSI_CAPI_NO_CHECKSUM")));
    const double t0 = now();
    while (bcf_sr_next_line(sr)) {
        bcf1_t* line = bcf_sr_get_line(sr, 0);
        const int ngt = c_xcf_get_genotypes(x, 0, sr->readers[0].header, line, (void**)&gt, &ngt_arr);
        if (ngt < 0) { fprintf(stderr, "get_genotypes failed at record %ld\n", records); return 1; }
        if (do_sum) { /* four independent multiply-add lanes over 64-bit pairs: cheap next to the decode being timed */
            const uint64_t* w = (const uint64_t*)gt;
            const int nw = ngt / 2;
            int i = 0;
            for (; i + 4 <= nw; i += 4) {
                h0 = h0 * 0x9E3779B97F4A7C15ull + w[i];
                h1 = h1 * 0xC2B2AE3D27D4EB4Full + w[i + 1];
                h2 = h2 * 0x165667B19E3779F9ull + w[i + 2];
                h3 = h3 * 0x27D4EB2F165667C5ull + w[i + 3];
            }
            for (; i < nw; ++i) h0 = h0 * 0x9E3779B97F4A7C15ull + w[i];
            if (ngt & 1) h1 = h1 * 0xC2B2AE3D27D4EB4Full + (uint32_t)gt[ngt - 1];
            h0 += (uint64_t)ngt;
        }
        genotypes += (uint64_t)ngt;
        ++records;
        if (max_records >= 0 && records >= max_records) break;
    }
    const double t1 = now();
    c_xcf_delete(x);
    const double t2 = now();
    printf("records %ld genotypes %llu seconds %.6f checksum %016llx setup %.6f teardown %.6f\n", records, (unsigned long long)genotypes, t1 - t0,
           (unsigned long long)(h0 ^ (h1 * 3) ^ (h2 * 5) ^ (h3 * 7)), t0 - ts, t2 - t1);
    return 0;
}
