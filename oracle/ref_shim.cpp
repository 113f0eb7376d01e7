// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" door onto the UNMODIFIED reference classes, compiled against the
// reference headers where they lie (/root/reference/include, see oracle/Makefile).
// Nothing of the reference is copied here: this file only drives
//   * XsiFactoryExt<A_T>            (include/xsi_factory.hpp:435-639)   -- the .xsi writer
//   * Accessor                      (include/accessor.hpp:31-124)       -- the reader
// the way GtCompressorStream (include/gt_compressor_new.hpp:84-142) and
// Accessor::get_genotypes (include/accessor.hpp:58-67) drive them, but fed from
// in-memory int32 genotype arrays instead of htslib records, so that tests and
// bench.py's reference arm can run the real reference on synthetic data.
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "xsqueezeit.hpp"
GlobalAppOptions global_app_options;  // the reference expects its user to define this

#include "gt_compressor_new.hpp"
#include "accessor.hpp"

extern "C" {

// The reference's --wah-encode-missing switch (xsqueezeit.hpp:58; read by GtBlock's constructor, gt_block.hpp:174-176).
void xsi_ref_set_wah_encode_missing(int on) { global_app_options.wah_encode_missing = on != 0; }

// Returns 0 on success, <0 on a reference `throw`.
// gt: concatenated rows, row r starts at rec_off[r] and holds ngt[r] int32 (htslib GT encoding).
int xsi_ref_encode_file(const char* out_path, const int32_t* gt, const uint64_t* rec_off,
                        const int32_t* ngt, const int32_t* n_allele, uint64_t n_records,
                        uint64_t n_samples, uint64_t block_len, uint64_t mac_threshold,
                        int default_phased, int zstd_on, int zstd_level, int quiet,
                        const char* sample_names /* n_samples NUL-terminated strings back to back, or NULL */) {
    FILE* saved = nullptr;
    std::streambuf* old_cout = nullptr;
    std::ostringstream sink;
    if (quiet) old_cout = std::cout.rdbuf(sink.rdbuf());
    (void)saved;
    int rc = 0;
    try {
        std::vector<std::string> samples(n_samples);
        const char* sn = sample_names;
        for (uint64_t i = 0; i < n_samples; ++i) {
            if (sn) { samples[i] = sn; sn += samples[i].size() + 1; }
            else samples[i] = "S" + std::to_string(i);
        }
        std::unique_ptr<XsiFactoryInterface> factory;
        // NewCompressor::init_compression, gt_compressor_new.hpp:174-191
        if (n_samples * 2 <= std::numeric_limits<uint16_t>::max())
            factory = make_unique<XsiFactoryExt<uint16_t> >(std::string(out_path), block_len, mac_threshold,
                                                            default_phased, samples, zstd_on != 0, zstd_level);
        else
            factory = make_unique<XsiFactoryExt<uint32_t> >(std::string(out_path), block_len, mac_threshold,
                                                            default_phased, samples, zstd_on != 0, zstd_level);
        bcf1_t line;
        memset(&line, 0, sizeof(line));
        bcf_file_reader_info_t fri;
        fri.n_samples = n_samples;
        fri.line = &line;
        size_t PLOIDY = 0;
        for (uint64_t r = 0; r < n_records; ++r) {
            fri.gt_arr = const_cast<int*>(reinterpret_cast<const int*>(gt + rec_off[r]));
            fri.ngt = ngt[r];
            fri.size_gt_arr = ngt[r];
            line.n_allele = n_allele[r];
            size_t line_max_ploidy = n_samples ? (size_t)ngt[r] / n_samples : 0;
            // GtCompressorStream::handle_bcf_line, gt_compressor_new.hpp:111-124
            if (line_max_ploidy > PLOIDY) {
                if (line_max_ploidy > 2) throw "Ploidy higher than 2 is not yet supported";
                PLOIDY = line_max_ploidy;
            }
            factory->append(fri);
        }
        factory->finalize_file(PLOIDY);
    } catch (const char* e) {
        fprintf(stderr, "xsi_ref_encode_file: reference threw: %s\n", e);
        rc = -1;
    } catch (std::exception& e) {
        fprintf(stderr, "xsi_ref_encode_file: exception: %s\n", e.what());
        rc = -2;
    }
    if (quiet) std::cout.rdbuf(old_cout);
    return rc;
}

void* xsi_ref_accessor_open(const char* path) {
    try {
        std::string p(path);
        return new Accessor(p);
    } catch (const char* e) {
        fprintf(stderr, "xsi_ref_accessor_open: reference threw: %s\n", e);
        return nullptr;
    }
}

uint64_t xsi_ref_hap_samples(void* h) { return static_cast<Accessor*>(h)->get_header_ref().hap_samples; }

// Accessor::fill_genotype_array (accessor.hpp:48-50). Returns number of filled entries, or (uint64)-1.
uint64_t xsi_ref_fill_genotype_array(void* h, int32_t* gt_arr, uint64_t gt_arr_size, uint64_t n_alleles,
                                     uint64_t position) {
    try {
        return static_cast<Accessor*>(h)->fill_genotype_array(gt_arr, gt_arr_size, n_alleles, position);
    } catch (const char* e) {
        fprintf(stderr, "xsi_ref_fill_genotype_array: reference threw: %s\n", e);
        return (uint64_t)-1;
    }
}

// Accessor::fill_allele_counts (accessor.hpp:52-54): counts without materialising the row
int xsi_ref_fill_allele_counts(void* h, uint64_t n_alleles, uint64_t position) {
    try {
        static_cast<Accessor*>(h)->fill_allele_counts(n_alleles, position);
        return 0;
    } catch (const char* e) {
        fprintf(stderr, "xsi_ref_fill_allele_counts: reference threw: %s\n", e);
        return -1;
    }
}

// allele counts as left by the last fill_genotype_array / fill_allele_counts (accessor.hpp:56)
uint64_t xsi_ref_allele_counts(void* h, uint64_t* out, uint64_t cap) {
    const std::vector<size_t>& ac = static_cast<Accessor*>(h)->get_allele_counts();
    for (size_t i = 0; i < ac.size() && i < cap; ++i) out[i] = ac[i];
    return ac.size();
}

// InternalGtAccess of a record (Accessor::get_internal_access, accessor.hpp:69-72 -> accessor_internals_new.hpp:444-471): per
// line the sparse flag and the first `nbytes` bytes at its pointer, the default allele, and the arrangement widened to uint32.
namespace {
struct Peek : public Accessor {
    static AccessorInternals* internals_of(Accessor* a) { return (a->*(&Peek::internals)).get(); }
};
}
int xsi_ref_internal_access(void* h, uint64_t n_alleles, uint64_t position, uint32_t* a_out, uint64_t a_n, int a_bytes,
                            uint8_t* sparse_out, uint8_t* bytes_out, uint32_t nbytes, int32_t* default_allele) {
    try {
        InternalGtAccess ia = Peek::internals_of(static_cast<Accessor*>(h))->get_internal_access(n_alleles, position);
        if ((int)ia.a_bytes != a_bytes) return -2;
        for (uint64_t i = 0; i < a_n; ++i)
            a_out[i] = a_bytes == 2 ? static_cast<const uint16_t*>(ia.a)[i] : static_cast<const uint32_t*>(ia.a)[i];
        for (size_t k = 0; k < ia.pointers.size(); ++k) {
            sparse_out[k] = ia.sparse[k] ? 1 : 0;
            memcpy(bytes_out + k * nbytes, ia.pointers[k], nbytes);
        }
        if (default_allele) *default_allele = ia.default_allele;
        return (int)ia.pointers.size();
    } catch (const char* e) {
        fprintf(stderr, "xsi_ref_internal_access: threw: %s\n", e);
        return -1;
    }
}

void xsi_ref_accessor_close(void* h) { delete static_cast<Accessor*>(h); }

}  // extern "C"
