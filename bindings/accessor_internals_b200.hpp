// bindings/accessor_internals_b200.hpp -- reference-side DECODE adapter: the B200 path behind the reference's abstract
// AccessorInternals (include/accessor_internals.hpp:399-413), a drop-in for AccessorInternalsNewTemplate<A_T,WAH_T>
// (include/accessor_internals_new.hpp:719-906) where Accessor::Accessor picks its internals (accessor.cpp:62-79).
//
// Wiring (bindings/Makefile): the reference's accessor.cpp is compiled from where it lies with
//     g++ -include bindings/accessor_internals_b200.hpp ... /root/reference/accessor.cpp
// The header includes the reference's accessor_internals_new.hpp first, derives the adapter from the reference class
// (so the mmap, header checks, block index lookup and ZSTD_decompress stay the reference's own code,
// accessor_internals_new.hpp:763-893) and then renames AccessorInternalsNewTemplate for the rest of the translation
// unit, so that the two make_unique<> at accessor.cpp:64,71 build the adapter.  Accessor, Xcf and the C API
// (c_api.h: c_xcf_get_genotypes) reach it unchanged.
//
// What is replaced is the cursor decoder DecompressPointerGTBlock (accessor_internals_new.hpp:49-717): a block is
// expanded and un-permuted on the device when first touched, after which any of its records is addressable without
// replay (seek, :154-196, disappears -- forward, backward and repeated positions cost the same).
//
// One record per call is the contract of fill_genotype_array (c_api.cpp:78-80 -> accessor.hpp:58-67); a device round
// trip per record would waste the GPU, so the adapter DECODES AHEAD: on a miss it materialises the next window of
// binary lines of the block as bi-allelic records (raw BCF int8 rows, a quarter of the int32 bytes over PCIe) into a
// pinned window and serves calls from there, widening one row into the caller's int32 array.  A line that turns out
// to belong to a multi-allelic record is simply never asked for under its own position; records with more than two
// alleles are decoded on demand.
#ifndef ACCESSOR_INTERNALS_B200_HPP
#define ACCESSOR_INTERNALS_B200_HPP

// the reference's own headers, in the order accessor.hpp:28-29 includes them
#include "accessor_internals.hpp"
#include "accessor_internals_new.hpp"

#include <algorithm>
#include <vector>

#include "xsi_b200_runtime.hpp"

template <typename A_T = uint32_t, typename WAH_T = uint16_t>
class AccessorInternalsB200 : public AccessorInternalsNewTemplate<A_T, WAH_T> {
    using Base = AccessorInternalsNewTemplate<A_T, WAH_T>;

public:
    AccessorInternalsB200(std::string filename) : Base(filename) {
        const int rc = xsi_create(xsi_b200::device_from_env(), &ctx);
        if (rc != XSI_OK) { ctx = nullptr; xsi_b200::raise(nullptr, rc, "xsi_create (no CUDA device? there is no CPU fallback)"); }
        N = (size_t)this->header.num_samples * 2;
        stride8 = (N + 15) / 16 * 16;
        const char* e = getenv("XSI_B200_WINDOW_BYTES");
        window_bytes = e ? (size_t)atoll(e) : (size_t)32 << 20;
    }

    virtual ~AccessorInternalsB200() {
        if (ctx) xsi_destroy(ctx);
    }

    // accessor_internals_new.hpp:741-745
    size_t fill_genotype_array(int32_t* gt_arr, size_t gt_arr_size, size_t n_alleles, size_t new_position) override {
        uint32_t line;
        locate(new_position, line);
        if (n_alleles == 2) {
            if (!(line >= win_line0 && line < win_line0 + win_n)) fill_window(line);
            const size_t i = line - win_line0;
            const uint32_t len = win_filled[i];
            if (gt_arr_size < len) throw "fill_genotype_array: array too small";
            xsi_host_widen_i8_i32(win.as<int8_t>() + i * stride8, stride8, gt_arr, len, &len, 1);
            this->allele_counts.assign(win_counts.begin() + 2 * i, win_counts.begin() + 2 * i + 2);
            return len;
        }
        // multi-allelic records: on demand
        if (n_alleles < 2) throw "fill_genotype_array: fewer than 2 alleles";
        if (gt_arr_size < N) throw "fill_genotype_array: array too small";
        const uint32_t b0 = 0, na = (uint32_t)n_alleles;
        uint32_t filled = 0;
        counts_tmp.assign(n_alleles, 0);
        const int rc = xsi_decode_records(ctx, 1, &b0, &line, &na, gt_arr, gt_arr_size, 0, &filled, counts_tmp.data(), na);
        if (rc != XSI_OK) xsi_b200::raise(ctx, rc, "xsi_decode_records");
        this->allele_counts.assign(counts_tmp.begin(), counts_tmp.end());
        return filled;
    }

    // accessor_internals_new.hpp:747-752: counts without materialising the row
    void fill_allele_counts(size_t n_alleles, size_t new_position) override {
        uint32_t line;
        locate(new_position, line);
        if (n_alleles < 2) throw "fill_allele_counts: fewer than 2 alleles";
        const uint32_t b0 = 0, na = (uint32_t)n_alleles;
        counts_tmp.assign(n_alleles, 0);
        const int rc = xsi_decode_allele_counts(ctx, 1, &b0, &line, &na, counts_tmp.data(), na);
        if (rc != XSI_OK) xsi_b200::raise(ctx, rc, "xsi_decode_allele_counts");
        this->allele_counts.assign(counts_tmp.begin(), counts_tmp.end());
    }

    // the base returns its cursor's vector (accessor_internals_new.hpp:755-757); there is no cursor here
    inline const std::vector<size_t>& get_allele_counts() const override { return this->allele_counts; }

    // accessor_internals_new.hpp:444-471: pointers to the record's encoded lines inside the mapped block, and the PBWT
    // arrangement `a` in force at the record.  The device has located every line when the block was loaded; the arrangement is
    // what the lazy chain has parked after the WAH lines before the record (the chain is continued up to there; a request behind
    // the chain reloads the block, like the reference's backward seek resets its cursor, :178-186).
    inline InternalGtAccess get_internal_access(size_t n_alleles, size_t new_position) override {
        uint32_t line;
        locate(new_position, line);
        InternalGtAccess ia;
        ia.position = line;
        ia.n_alleles = n_alleles;
        ia.sparse_bytes = sizeof(A_T);
        ia.wah_bytes = sizeof(WAH_T);
        ia.a_bytes = sizeof(A_T);
        if (n_alleles == 0) return ia;
        if (n_alleles < 2) throw "get_internal_access: fewer than 2 alleles";
        std::vector<xsi_line_access> la(n_alleles - 1);
        a_host.resize(N);
        int32_t dflt = 0;
        int rc = xsi_decode_internal_access(ctx, 0, line, (uint32_t)n_alleles, la.data(), &dflt, a_host.data());
        if (rc == XSI_E_UNSUPPORTED) {  // the chain is already past this line: start the block again
            loaded = false;
            locate(new_position, line);
            rc = xsi_decode_internal_access(ctx, 0, line, (uint32_t)n_alleles, la.data(), &dflt, a_host.data());
        }
        if (rc != XSI_OK) xsi_b200::raise(ctx, rc, "xsi_decode_internal_access");
        ia.default_allele = dflt;
        ia.a = a_host.data();
        char* base = static_cast<char*>(this->gt_block_p);
        for (const auto& l : la) {
            ia.sparse.push_back(l.is_sparse != 0);
            ia.pointers.push_back(base + l.byte_offset);
        }
        return ia;
    }

private:
    // BM split (accessor_internals_new.hpp:722-726) + block residency
    void locate(size_t position, uint32_t& line) {
        const size_t block_id = (position & 0xFFFFFFFF) >> this->BM_BLOCK_BITS;
        line = (uint32_t)(position & (((size_t)1 << this->BM_BLOCK_BITS) - 1));
        if (loaded && this->current_block == block_id) return;
        loaded = false;
        win_n = 0;
        this->set_gt_block_ptr(block_id);  // reference code: index lookup, optional ZSTD_decompress, outer dictionary
        // the GT block runs to the end of the (inflated) outer block; the next block offset bounds it in the mmap
        const uint8_t* p = static_cast<const uint8_t*>(this->gt_block_p);
        uint64_t size = 0;
        if (this->header.zstd) {
            const uint8_t* m = static_cast<const uint8_t*>(this->file_mmap_p) + block_offset(block_id);
            const uint64_t usize = *reinterpret_cast<const uint64_t*>(m + sizeof(uint64_t));
            size = usize - (uint64_t)(p - static_cast<const uint8_t*>(this->block_p));
        } else {
            const uint64_t end = block_id + 1 < this->header.number_of_ssas ? block_offset(block_id + 1) : (uint64_t)this->header.indices_offset;
            size = end - (uint64_t)(p - static_cast<const uint8_t*>(this->file_mmap_p));
        }
        // lazy: every WAH line is expanded now, the sequential inverse-PBWT chain only runs as far as the records that are
        // actually asked for (a region query near the start of a block does not pay for the whole block)
        const int rc = xsi_decode_load_blocks_lazy(ctx, 1, &p, &size, this->header.num_samples, this->header.aet_bytes, 0);
        if (rc != XSI_OK) xsi_b200::raise(ctx, rc, "xsi_decode_load_blocks_lazy");
        uint32_t bcf_lines = 0;
        if (xsi_decode_block_info(ctx, 0, &bcf_lines, &bin_lines) != XSI_OK) throw "xsi_decode_block_info";
        loaded = true;
    }

    uint64_t block_offset(size_t block_id) const {  // accessor_internals_new.hpp:849-855 (version 5: u64 index)
        const uint8_t* base = static_cast<const uint8_t*>(this->file_mmap_p) + this->header.indices_offset;
        if (this->header.version <= 4) return reinterpret_cast<const uint32_t*>(base)[block_id];
        return reinterpret_cast<const uint64_t*>(base)[block_id];
    }

    void fill_window(uint32_t line) {
        if (line >= bin_lines) throw "fill_genotype_array: position past the end of its block";
        const size_t rows = std::max<size_t>(1, std::min<size_t>(window_bytes / stride8, bin_lines - line));
        win.reserve(rows * stride8);
        win_filled.resize(rows);
        win_counts.resize(rows * 2);
        req_blk.assign(rows, 0);
        req_na.assign(rows, 2);
        req_line.resize(rows);
        for (size_t i = 0; i < rows; ++i) req_line[i] = line + (uint32_t)i;
        const int rc = xsi_decode_records_i8(ctx, rows, req_blk.data(), req_line.data(), req_na.data(), win.as<int8_t>(), stride8, 0,
                                             win_filled.data(), win_counts.data(), 2);
        if (rc != XSI_OK) { win_n = 0; xsi_b200::raise(ctx, rc, "xsi_decode_records_i8"); }
        win_line0 = line;
        win_n = (uint32_t)rows;
    }

    xsi_ctx* ctx = nullptr;
    size_t N = 0, stride8 = 0, window_bytes = 0;
    bool loaded = false;
    uint32_t bin_lines = 0;
    // decoded-ahead window: rows of binary lines [win_line0, win_line0 + win_n) of the resident block, as int8
    xsi_b200::Pinned win;
    uint32_t win_line0 = 0, win_n = 0;
    std::vector<uint32_t> win_filled, req_blk, req_line, req_na;
    std::vector<uint64_t> win_counts, counts_tmp;
    std::vector<A_T> a_host;  // the arrangement handed out by get_internal_access
};

// from here on, every mention of AccessorInternalsNewTemplate in this translation unit (accessor.cpp:64,71) is the adapter
#define AccessorInternalsNewTemplate AccessorInternalsB200

#endif
