// bindings/gt_block_b200.hpp -- reference-side ENCODE adapter: the B200 path behind the reference's own plugin
// interface IWritableBCFLineEncoder (include/interfaces.hpp:99-152), i.e. a drop-in for GtBlock<A_T,WAH_T>
// (include/gt_block.hpp:154-687) where EncodingBinaryBlockWithGT registers it under KEY_GT_ENTRY
// (include/xsi_factory.hpp:419-433).
//
// How it is wired in (bindings/Makefile): the reference's translation units are compiled from where they lie with
//     g++ -include bindings/gt_block_b200.hpp ... /root/reference/xsqueezeit.cpp
// This header pulls in the reference's gt_block.hpp first (so its include guard is set), defines GtBlockB200 with
// the same constructor signature, and then renames `GtBlock` for the rest of the translation unit, so that the
// two `std::make_shared<GtBlock<...>>` at xsi_factory.hpp:427-428 build the adapter.  The equivalent source edit a
// maintainer would make is those two lines.  Nothing else of the reference changes: the file
// header, outer dictionary, zstd framing, padding, index and sample names are still written by XsiFactoryExt.
//
// Contract kept: one instance per block; encode_line once per BCF record in file order on one
// thread with a borrowed bcf_fri; write_to_stream once at flush, writing [u32 -1][u32 n][dictionary][sections] at
// the current stream position with offsets relative to it; errors are `throw const char*`.
#ifndef GT_BLOCK_B200_HPP
#define GT_BLOCK_B200_HPP

// the reference's own headers, in the order its translation units see them (xsqueezeit.cpp:26-30,
// gt_compressor_new.hpp:28-49, xsi_factory.hpp:28-34): they lean on each other's includes
#include <iostream>
#include "fs.hpp"
#include <thread>
#include "bcf_traversal.hpp"
#include "xcf.hpp"
#include "wah.hpp"
#include "compression.hpp"
#include "make_unique.hpp"
#include "internal_gt_record.hpp"
#include <algorithm>
#include <numeric>
#include <sstream>
#include <string>
#include <memory>
#include "block.hpp"
#include "xsqueezeit.hpp"
extern GlobalAppOptions global_app_options;  // gt_compressor_new.hpp:51
#include "gt_block.hpp"                      // IWritableBCFLineEncoder, BCFBlock, IBinaryBlock keys

#include <vector>

#include "xsi_b200_runtime.hpp"

template <typename A_T = uint32_t, typename WAH_T = uint16_t>
class GtBlockB200 : public IWritableBCFLineEncoder, public BCFBlock {
public:
    GtBlockB200(const size_t NUM_SAMPLES, const size_t BLOCK_BCF_LINES, const size_t MAC_THRESHOLD, const int32_t default_phasing = 0)
        : BCFBlock(BLOCK_BCF_LINES), num_samples(NUM_SAMPLES), mac_threshold(MAC_THRESHOLD), default_phasing(default_phasing),
          wah_encode_missing(global_app_options.wah_encode_missing),  // gt_block.hpp:174-176
          ts(xsi_b200::thread_state()) {
        n_allele.reserve(BLOCK_BCF_LINES);
        ploidy.reserve(BLOCK_BCF_LINES);
    }

    inline uint32_t get_id() const override { return IBinaryBlock<uint32_t, uint32_t>::KEY_GT_ENTRY; }

    // gt_block.hpp:279-406 buffers nothing and encodes at once; here the row is staged (pinned memory) and the whole
    // block is encoded on the device at flush.  The row is taken as the record's raw FORMAT/GT payload (int8, one
    // byte per genotype, htslib vcf.h:152-158) when the record has one: bcf_get_genotypes only widened those bytes
    // (htslib vcf.c:4728-4795) and the kernels read them as they are (gt_elem_bytes = 1); a quarter of the bytes to
    // stage and to move over PCIe.  Records whose GT is int16/int32 (more than 63 alleles) switch the block to the
    // int32 rows of bcf_fri.gt_arr.
    void encode_line(const bcf_file_reader_info_t& bcf_fri) override {
        if (bcf_fri.ngt < 0 || bcf_fri.n_samples == 0) throw "Unknown allele error !";
        const size_t ngt = (size_t)bcf_fri.ngt;
        const size_t pl = ngt / bcf_fri.n_samples;
        if (pl > 2) throw "Ploidy higher than 2 is not yet supported";  // gt_compressor_new.hpp:118-120
        if (pl == 0 || pl * bcf_fri.n_samples != ngt || bcf_fri.n_samples != num_samples) throw "Unknown allele error !";

        const int8_t* raw = nullptr;
        if (elem_bytes == 1 && bcf_fri.sr && bcf_fri.line) {
            bcf_fmt_t* fmt = bcf_get_fmt(bcf_fri.sr->readers[0].header, bcf_fri.line, "GT");
            if (fmt && fmt->type == BCF_BT_INT8 && (size_t)fmt->n == pl && fmt->p) raw = reinterpret_cast<const int8_t*>(fmt->p);
        }
        if (elem_bytes == 1 && !raw) widen_staged_rows();  // from here on this block moves int32
        ts.rows.reserve((n_elems + ngt) * elem_bytes, n_elems * elem_bytes);
        if (elem_bytes == 1) memcpy(ts.rows.as<int8_t>() + n_elems, raw, ngt);
        else memcpy(ts.rows.as<int32_t>() + n_elems, bcf_fri.gt_arr, ngt * sizeof(int32_t));
        n_elems += ngt;
        n_allele.push_back((uint32_t)bcf_fri.line->n_allele);
        ploidy.push_back((uint8_t)pl);
        effective_bcf_lines_in_block++;
    }

    // gt_block.hpp:185-204: one launch for the block, then the bytes exactly as GtBlock would have written them
    void write_to_stream(std::fstream& ofs) override {
        xsi_ctx* ctx = ts.context();
        xsi_encode_desc d;
        memset(&d, 0, sizeof(d));
        d.n_records = n_allele.size();
        d.n_samples = (uint32_t)num_samples;
        d.block_len = (uint32_t)(n_allele.size() > BLOCK_BCF_LINES ? n_allele.size() : BLOCK_BCF_LINES);
        d.mac_threshold = mac_threshold;
        d.default_phasing = default_phasing;
        d.gt_elem_bytes = elem_bytes;
        d.gt_on_device = 0;
        d.wah_encode_missing = wah_encode_missing ? 1 : 0;
        d.gt = ts.rows.p;
        d.n_allele = n_allele.data();
        d.ploidy = ploidy.data();
        int rc = xsi_encode_launch(ctx, &d);
        if (rc != XSI_OK) xsi_b200::raise(ctx, rc, "xsi_encode_launch");
        uint32_t nb = 0;
        const uint8_t* const* blk = nullptr;
        const uint64_t* sz = nullptr;
        rc = xsi_encode_collect(ctx, &nb, &blk, &sz);
        if (rc != XSI_OK) xsi_b200::raise(ctx, rc, "xsi_encode_collect");
        if (nb != 1) throw "GtBlockB200: one block expected";
        ofs.write(reinterpret_cast<const char*>(blk[0]), sz[0]);
    }

private:
    // rare: a record without an int8 GT payload arrived after int8 rows were staged
    void widen_staged_rows() {
        if (n_elems) {
            std::vector<int8_t> tmp(ts.rows.as<int8_t>(), ts.rows.as<int8_t>() + n_elems);
            ts.rows.reserve(n_elems * sizeof(int32_t));
            const uint32_t len = (uint32_t)n_elems;
            // one "row" of n_elems values; n_elems of a block stays far below 2^32 only for small shapes, so go by pieces
            size_t done = 0;
            while (done < n_elems) {
                const uint32_t piece = (uint32_t)std::min<size_t>(n_elems - done, (size_t)1 << 30);
                xsi_host_widen_i8_i32(tmp.data() + done, piece, ts.rows.as<int32_t>() + done, piece, &piece, 1);
                done += piece;
            }
            (void)len;
        }
        elem_bytes = 4;
    }

    const size_t num_samples, mac_threshold;
    const int32_t default_phasing;
    const bool wah_encode_missing;
    xsi_b200::ThreadState& ts;
    int32_t elem_bytes = 1;
    size_t n_elems = 0;  // genotypes staged so far
    std::vector<uint32_t> n_allele;
    std::vector<uint8_t> ploidy;
};

// from here on, every mention of GtBlock in this translation unit (xsi_factory.hpp:427-428) is the adapter
#define GtBlock GtBlockB200

#endif
