/* bindings/capi_decode.c -- timing loop over the reference's C API (include/c_api.h:38-93), the way a consumer such
 * as SHAPEIT4 or the reference's own c_api_test/main.c uses it: one c_xcf_get_genotypes call per record of the
 * `_var.bcf` companion.  Linked twice by bindings/Makefile: with the reference's accessor.o (CPU) and with
 * accessor_b200.o (the adapter, GPU).  Prints one line:
 *     records <R> genotypes <G> seconds <T> checksum <C>
 * where the checksum (FNV-1a over every returned int32) lets the two builds be compared without storing rows.
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
    c_xcf* x = c_xcf_new();
    bcf_srs_t* sr = bcf_sr_init();
    if (!bcf_sr_add_reader(sr, argv[1])) { fprintf(stderr, "could not load %s\n", argv[1]); return 1; }
    c_xcf_add_readers(x, sr);
    int* gt = NULL;
    int ngt_arr = 0;
    long records = 0;
    uint64_t genotypes = 0, h = 1469598103934665603ull;
    const double t0 = now();
    while (bcf_sr_next_line(sr)) {
        bcf1_t* line = bcf_sr_get_line(sr, 0);
        const int ngt = c_xcf_get_genotypes(x, 0, sr->readers[0].header, line, (void**)&gt, &ngt_arr);
        if (ngt < 0) { fprintf(stderr, "get_genotypes failed at record %ld\n", records); return 1; }
        /* 8 values per multiply keeps the checksum cheap next to the decode being timed */
        int i = 0;
        for (; i + 8 <= ngt; i += 8) {
            uint64_t v = 0;
            for (int k = 0; k < 8; ++k) v = v * 31 + (uint32_t)gt[i + k];
            h = (h ^ v) * 1099511628211ull;
        }
        for (; i < ngt; ++i) h = (h ^ (uint32_t)gt[i]) * 1099511628211ull;
        genotypes += (uint64_t)ngt;
        ++records;
        if (max_records >= 0 && records >= max_records) break;
    }
    const double t1 = now();
    c_xcf_delete(x);
    printf("records %ld genotypes %llu seconds %.6f checksum %016llx\n", records, (unsigned long long)genotypes, t1 - t0,
           (unsigned long long)h);
    return 0;
}
