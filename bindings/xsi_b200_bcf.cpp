// bindings/xsi_b200_bcf.cpp -- BCF ingest for the B200 path: `xsqueezeit -c` rebuilt around the C ABI
// so that the GPU is fed at the speed htslib can read, not at the speed of the reference's record-at-a-time loop.
// Host C++ keeps htslib BCF parsing, the variant-info BCF writer and the container layout (north_star); genotype work
// happens only in libxsi_b200.so.  Output is the same file pair the reference writes, byte for byte.
//
//   xsi_b200_bcf compress <in.bcf|vcf.gz> <out.xsi> [--maf 0.001] [--variant-block-length 8192] [--zstd] [--zstd-level 7]
//                [--wah-encode-missing] [--threads T] [--batch-blocks K] [--device D]
//
// What differs from the reference's compress loop (bcf_traversal.cpp:3-16, gt_compressor_new.hpp:84-142):
//  * the reader thread never calls bcf_get_genotypes: the record's FORMAT/GT payload (bcf_fmt_t.p, int8 for up to 63
//    alleles, htslib/vcf.h:152-158) is copied as is into a pinned batch and read by the kernels as gt_elem_bytes = 1
//    (no int32 widening, htslib/vcf.c:4728-4795; a quarter of the bytes to stage and to move over PCIe);
//  * BGZF inflate runs on a thread pool (hts_set_threads);
//  * K whole blocks go to the device per launch (their PBWT chains run side by side), and the encode of batch i runs
//    on its own thread while the reader fills batch i+1.
//  * ONE pass over the input: the reference opens and parses it eight times (two checks, a sample count, the default-phase
//    and ploidy pre-scans twice, the traversal, and the `_var.bcf` thread; each open parses a header with one dictionary
//    entry per sample, 0.45 s at 32,488 samples).  Here the default phase (xcf.cpp:811-836), the first record's ploidy
//    (xcf.cpp:838-862) and the `_var.bcf` companion (xcf.cpp:641-714: samples replaced by the pseudo-sample
//    BIN_MATRIX_POS with FORMAT/BM = block<<15 | binary line offset) all come from the records the reader holds anyway;
//    the companion is byte-identical to the reference's (tests/test_bindings.py).  --reference-var runs the reference's
//    own function on a second thread instead (xsqueezeit.cpp:119-128).
// The CSI index of the companion is built by the reference's create_index_file (xcf.cpp:39-57).
//
//   xsi_b200_bcf extract <in.xsi> <out.bcf> [-O b|u] [--threads T] [--window-bytes N] [--device D]
// is the matching egress (`xsqueezeit -x`, gt_decompressor_new.hpp:109-206,275-320): records of the companion are read,
// their rows decoded on the device as raw BCF int8 FORMAT/GT payload (xsi_decode_records_i8) in windows, and spliced into
// the records as the typed vector bcf_update_genotypes would have built, so the CPU neither widens to int32 nor narrows
// back (bcf_enc_vint); BGZF deflate of the output runs on the thread pool.  The output equals the reference's byte for byte.
//
//   xsi_b200_bcf subset <in.xsi> <out.xsi> [--samples a,b,c | --samples ^a,b | --samples-file f] [--maf 0.001] [--zstd] [--zstd-level 7]
//                [--batch-blocks K] [--threads T] [--device D]
// is `xsqueezeit -x -O x [-s/-S]` (gt_decompressor_new.hpp:130-143,241-273: decode every record, keep the selected samples,
// append the row to a new XsiFactoryExt, write the companion with the new BM and, under -s/-S, AC / AN recomputed) with the
// rows never leaving the device: whole blocks are decoded and gathered by xsi_decode_records_subset into device rows, and
// xsi_encode_launch_strided encodes them from there; only the selected carriers' counts (ac_s) and the encoded blocks come back.
// Whole files only (the reference's -r/-t with -Ox goes through bindings/_out/xsqueezeit_b200).  Both output files equal the
// reference's byte for byte.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <libgen.h>

#include "xcf.hpp"  // reference host helpers (declarations only; objects come from oracle/_ref/obj)

#include "xsi_b200_runtime.hpp"

namespace {

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// Row staging of a batch.  Pageable on purpose: page-locking is slow on some hosts (2.5 s per GB measured on the B200 box,
// i.e. longer than reading the records), while the H2D copy of a pageable batch through the driver's bounce buffers costs
// tens of milliseconds and runs on the encoder thread beside the reader.
struct Rows {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes, size_t keep = 0) {
        if (bytes <= cap) return;
        size_t want = cap ? cap : (size_t)1 << 20;
        while (want < bytes) want *= 2;
        void* q = malloc(want);
        if (!q) throw "out of memory";
        if (keep) memcpy(q, p, keep);
        free(p);
        p = q; cap = want;
    }
    ~Rows() { free(p); }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct Batch {
    Rows rows;
    std::vector<uint32_t> n_allele;
    std::vector<uint8_t> ploidy;
    size_t n_elems = 0;
    int32_t elem_bytes = 1;
    uint64_t variants = 0;
    bool last = false;
    void clear() { n_allele.clear(); ploidy.clear(); n_elems = 0; elem_bytes = 1; variants = 0; last = false; }
};

// two batches ping-pong between the reader (fills) and the encoder (drains)
struct Exchange {
    std::mutex m;
    std::condition_variable cv;
    Batch b[2];
    int state[2] = {0, 2};  // 0 = free (reader owns), 1 = full (encoder owns), 2 = being allocated by the encoder thread
    std::string error;
};

struct Options {
    std::string in, out;
    double maf = 0.001;
    size_t block_len = 8192;
    bool zstd = false, wah_missing = false, reference_var = false;
    int zstd_level = 7, threads = 8, batch_blocks = 2, device = 0;
    std::string output_type = "b";
    size_t window_bytes = (size_t)64 << 20;
    std::string samples, samples_file;
};

void widen_batch(Batch& b) {  // a record without an int8 GT payload arrived: the batch moves int32 from here on
    if (b.n_elems) {
        std::vector<int8_t> tmp(b.rows.as<int8_t>(), b.rows.as<int8_t>() + b.n_elems);
        b.rows.reserve(b.n_elems * 4);
        size_t done = 0;
        while (done < b.n_elems) {
            const uint32_t piece = (uint32_t)std::min<size_t>(b.n_elems - done, (size_t)1 << 30);
            xsi_host_widen_i8_i32(tmp.data() + done, piece, b.rows.as<int32_t>() + done, piece, &piece, 1);
            done += piece;
        }
    }
    b.elem_bytes = 4;
}

// the companion `_var.bcf`, written from the records of the one reader (what xcf.cpp:641-714 does in its own pass)
struct VarWriter {
    htsFile* fp = nullptr;
    bcf_hdr_t* hdr = nullptr;
    bcf1_t* v = nullptr;
    size_t line = 0, block_len = 0;
    int32_t offset = 0, block = 0;
    bool open(const bcf_hdr_t* in_hdr, const std::string& path, const std::string& xsi_name, size_t bl, int threads) {
        block_len = bl;
        fp = hts_open(path.c_str(), "wz");  // bgzipped VCF text, xcf.cpp:647
        if (!fp) return false;
        if (threads > 1) hts_set_threads(fp, threads);
        bcf_hdr_t* h0 = bcf_hdr_dup(in_hdr);
        if (bcf_hdr_set_samples(h0, NULL, 0) < 0) return false;  // xcf.cpp:650
        hdr = bcf_hdr_dup(h0);
        bcf_hdr_destroy(h0);
        bcf_hdr_add_sample(hdr, "BIN_MATRIX_POS");
        bcf_hdr_append(hdr, "##FORMAT=<ID=BM,Number=1,Type=Integer,Description=\"Position in GT Binary Matrix\">");
        std::string tmp(xsi_name);
        bcf_hdr_append(hdr, std::string("##XSI=").append(std::string(basename((char*)tmp.c_str()))).c_str());
        if (bcf_hdr_sync(hdr) < 0) fprintf(stderr, "bcf_hdr_sync() failed ... oh well\n");
        v = bcf_init();
        return bcf_hdr_write(fp, hdr) >= 0;
    }
    // rec: the full record as read; only its shared part (CHROM..INFO) is carried over
    bool add(const bcf1_t* rec) {
        bcf_clear(v);
        v->rid = rec->rid; v->pos = rec->pos; v->rlen = rec->rlen; v->qual = rec->qual;
        v->n_info = rec->n_info; v->n_allele = rec->n_allele; v->n_fmt = 0; v->n_sample = 0;
        if (ks_resize(&v->shared, rec->shared.l ? rec->shared.l : 1) != 0) return false;
        v->shared.l = rec->shared.l;
        memcpy(v->shared.s, rec->shared.s, rec->shared.l);
        v->indiv.l = 0;
        bcf_unpack(v, BCF_UN_STR);
        v->n_sample = 1;
        if (line && (line % block_len) == 0) { block++; offset = 0; }
        if (offset >> 15) throw "Variant BCF generation error, BM bits";  // xcf.cpp:691-694
        int32_t bm = block << 15 | offset;
        bcf_update_format_int32(hdr, v, "BM", &bm, 1);
        if (bcf_write1(fp, hdr, v) < 0) return false;
        if (rec->n_allele) offset += rec->n_allele - 1;
        line++;
        return true;
    }
    bool close() {
        bool ok = true;
        if (fp) ok = hts_close(fp) >= 0;
        if (hdr) bcf_hdr_destroy(hdr);
        if (v) bcf_destroy(v);
        fp = nullptr; hdr = nullptr; v = nullptr;
        return ok;
    }
};

int compress(const Options& o) {
    const double t0 = now();
    bool fail = false;
    std::thread variant_thread;
    if (o.reference_var)
        variant_thread = std::thread([&] {  // xsqueezeit.cpp:119-128
            try {
                replace_samples_by_pos_in_binary_matrix(o.in, o.out + "_var.bcf", o.out, true, o.block_len);
            } catch (const char* e) {
                fprintf(stderr, "%s\n", e);
                fail = true;
            }
            create_index_file(o.out + "_var.bcf");
        });
    auto join_var = [&] { if (variant_thread.joinable()) variant_thread.join(); };

    htsFile* fp = hts_open(o.in.c_str(), "r");
    if (!fp) { fprintf(stderr, "Failed to open file %s\n", o.in.c_str()); join_var(); return 1; }
    if (o.threads > 1) hts_set_threads(fp, o.threads);
    bcf_hdr_t* hdr = bcf_hdr_read(fp);
    if (!hdr) { fprintf(stderr, "Failed to read the header of %s\n", o.in.c_str()); join_var(); return 1; }
    const size_t S = (size_t)bcf_hdr_nsamples(hdr);
    if (S == 0) { fprintf(stderr, "The file %s has no samples\n", o.in.c_str()); join_var(); return 1; }  // xsqueezeit.cpp:111
    std::string names;
    for (size_t i = 0; i < S; ++i) { names += hdr->samples[i]; names.push_back('\0'); }
    VarWriter var;
    if (!o.reference_var && !var.open(hdr, o.out + "_var.bcf", o.out, o.block_len, std::max(1, o.threads / 4))) {
        fprintf(stderr, "Failed to write header to file %s_var.bcf\n", o.out.c_str());
        return 1;
    }
    // file-level parameters: found from the first records (below), needed when the first batch is handed over
    int32_t default_phased = 1;
    size_t first_ploidy = 0;
    uint64_t mac = 0;
    xsi_writer* w = nullptr;
    int rc = XSI_OK;

    Exchange ex;
    int max_ploidy = 0;
    uint64_t records = 0, genotypes = 0;
    double t_encode = 0;
    std::thread encoder([&] {
        xsi_ctx* ctx = nullptr;
        int r = xsi_create(o.device, &ctx);
        if (r != XSI_OK) {
            std::lock_guard<std::mutex> l(ex.m);
            ex.error = "xsi_create failed (no CUDA device? there is no CPU fallback)";
            ex.cv.notify_all();
            return;
        }
        try {
            ex.b[1].rows.reserve(o.block_len * (size_t)o.batch_blocks * S * 2);
        } catch (const char*) {
            std::lock_guard<std::mutex> l(ex.m);
            ex.error = "pinned allocation failed";
        }
        {
            std::lock_guard<std::mutex> l(ex.m);
            ex.state[1] = 0;
        }
        ex.cv.notify_all();
        for (int k = 0;; k ^= 1) {
            {
                std::unique_lock<std::mutex> l(ex.m);
                ex.cv.wait(l, [&] { return ex.state[k] == 1; });
            }
            Batch& b = ex.b[k];
            const double te = now();
            if (!b.n_allele.empty()) {
                xsi_encode_desc d;
                memset(&d, 0, sizeof d);
                d.n_records = b.n_allele.size();
                d.n_samples = (uint32_t)S;
                d.block_len = (uint32_t)o.block_len;
                d.mac_threshold = mac;
                d.default_phasing = default_phased;
                d.gt_elem_bytes = b.elem_bytes;
                d.wah_encode_missing = o.wah_missing ? 1 : 0;
                d.gt = b.rows.p;
                d.n_allele = b.n_allele.data();
                d.ploidy = b.ploidy.data();
                uint32_t nb = 0;
                const uint8_t* const* blk = nullptr;
                const uint64_t* sz = nullptr;
                r = xsi_encode_launch(ctx, &d);
                if (r == XSI_OK) r = xsi_encode_collect(ctx, &nb, &blk, &sz);
                if (r == XSI_OK) r = xsi_writer_add_blocks(w, nb, blk, sz, b.n_allele.size(), b.variants);
                if (r != XSI_OK) {
                    std::lock_guard<std::mutex> l(ex.m);
                    ex.error = std::string("encode failed: ") + xsi_last_error(ctx) + " (rc " + std::to_string(r) + ")";
                    ex.state[k] = 0;
                    ex.cv.notify_all();
                    break;
                }
                max_ploidy = std::max(max_ploidy, xsi_encode_max_ploidy(ctx));
            }
            t_encode += now() - te;
            const bool last = b.last;
            {
                std::lock_guard<std::mutex> l(ex.m);
                ex.state[k] = 0;
            }
            ex.cv.notify_all();
            if (last) break;
        }
        xsi_destroy(ctx);
    });

    // ---- reader: the BcfTraversal loop (bcf_traversal.cpp:3-16) without bcf_get_genotypes ----
    bcf1_t* rec = bcf_init();
    int32_t* gt32 = nullptr;
    int n_gt32 = 0;
    const size_t batch_records = o.block_len * (size_t)o.batch_blocks;
    int k = 0;
    bool eof = false, err = false;
    const int gt_id = bcf_hdr_id2int(hdr, BCF_DT_ID, "GT");
    size_t phase_counts[2] = {0, 0};
    bool phase_known = false;
    ex.b[0].rows.reserve(batch_records * S * 2);  // one allocation per batch buffer (int8, diploid); the encoder thread prepares the other
    while (!eof && !err) {
        {
            std::unique_lock<std::mutex> l(ex.m);
            ex.cv.wait(l, [&] { return ex.state[k] == 0 || !ex.error.empty(); });
            if (!ex.error.empty()) { err = true; break; }
        }
        Batch& b = ex.b[k];
        b.clear();
        while (b.n_allele.size() < batch_records) {
            const int rr = bcf_read(fp, hdr, rec);
            if (rr < -1) { fprintf(stderr, "read error in %s\n", o.in.c_str()); err = true; break; }
            if (rr < 0) { eof = true; break; }
            bcf_unpack(rec, BCF_UN_FMT);
            bcf_fmt_t* fmt = nullptr;
            for (int i = 0; i < (int)rec->n_fmt; ++i)
                if (rec->d.fmt[i].id == gt_id) { fmt = &rec->d.fmt[i]; break; }
            if (!fmt || S == 0) { fprintf(stderr, "record %llu has no GT\n", (unsigned long long)records); err = true; break; }
            const size_t pl = (size_t)fmt->n, ngt = pl * S;
            if (pl > 2) { fprintf(stderr, "Ploidy higher than 2 is not yet supported\n"); err = true; break; }  // gt_compressor_new.hpp:118-120
            if (b.elem_bytes == 1 && fmt->type != BCF_BT_INT8) widen_batch(b);
            b.rows.reserve((b.n_elems + ngt) * b.elem_bytes, b.n_elems * b.elem_bytes);
            if (b.elem_bytes == 1) {
                memcpy(b.rows.as<int8_t>() + b.n_elems, fmt->p, ngt);
            } else {
                const int n = bcf_get_genotypes(hdr, rec, &gt32, &n_gt32);
                if (n != (int)ngt) { fprintf(stderr, "bcf_get_genotypes failed\n"); err = true; break; }
                memcpy(b.rows.as<int32_t>() + b.n_elems, gt32, ngt * 4);
            }
            if (records < 3 && !phase_known) {  // seek_default_phased, xcf.cpp:811-836: phase bit of every sample's 2nd allele
                if (pl == 1) { default_phased = 0; phase_known = true; }
                else for (size_t i = 0; i < S; ++i)
                    phase_counts[(b.elem_bytes == 1 ? (int)b.rows.as<int8_t>()[b.n_elems + i * pl + 1] : b.rows.as<int32_t>()[b.n_elems + i * pl + 1]) & 1]++;
            }
            if (records == 0) first_ploidy = pl;  // seek_max_ploidy_from_first_entry, xcf.cpp:838-862
            if (!o.reference_var && !var.add(rec)) { fprintf(stderr, "Failed to write the variant file\n"); err = true; break; }
            b.n_elems += ngt;
            b.n_allele.push_back((uint32_t)rec->n_allele);
            b.ploidy.push_back((uint8_t)pl);
            b.variants += rec->n_allele ? rec->n_allele - 1 : 0;
            ++records;
            genotypes += ngt;
        }
        b.last = eof || err;
        if (!w && !err) {  // first batch complete (at least 3 records unless the file is shorter): the file-level parameters are known
            if (records == 0) { fprintf(stderr, "The file %s has no entries\n", o.in.c_str()); err = true; b.last = true; }  // xsqueezeit.cpp:114
            if (!phase_known) default_phased = phase_counts[0] > phase_counts[1] ? 0 : 1;
            mac = (uint64_t)((double)(S * first_ploidy) * o.maf);  // gt_compressor_new.hpp:98-99
            max_ploidy = (int)first_ploidy;
            rc = xsi_writer_open(o.out.c_str(), (uint32_t)S, names.data(), (uint32_t)o.block_len, mac, default_phased, o.zstd ? 1 : 0,
                                 o.zstd_level, &w);
            if (rc != XSI_OK) { fprintf(stderr, "Failed to open file %s (rc %d)\n", o.out.c_str(), rc); err = true; b.last = true; b.n_allele.clear(); }
        }
        {
            std::lock_guard<std::mutex> l(ex.m);
            ex.state[k] = 1;
        }
        ex.cv.notify_all();
        k ^= 1;
    }
    encoder.join();
    const double t_read_done = now();
    free(gt32);
    bcf_destroy(rec);
    bcf_hdr_destroy(hdr);
    hts_close(fp);
    if (!ex.error.empty()) { fprintf(stderr, "%s\n", ex.error.c_str()); err = true; }
    if (w) {
        rc = xsi_writer_close(w, max_ploidy);
        if (rc != XSI_OK) { fprintf(stderr, "finalize failed (rc %d)\n", rc); err = true; }
    }
    const double t_gt = now();
    if (!o.reference_var) {
        if (!var.close()) { fprintf(stderr, "Failed to close the variant file\n"); err = true; }
        if (!err) create_index_file(o.out + "_var.bcf");
    }
    join_var();
    const double t1 = now();
    if (fail || err) { fprintf(stderr, "Failure occurred, exiting...\n"); return 1; }
    printf("xsi_b200_bcf compress: records %llu genotypes %llu seconds %.6f gt_path_seconds %.6f encode_thread_seconds %.6f "
           "reader_seconds %.6f threads %d batch_blocks %d\n",
           (unsigned long long)records, (unsigned long long)genotypes, t1 - t0, t_gt - t0, t_encode, t_read_done - t0, o.threads, o.batch_blocks);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// extract: `xsqueezeit -x` (gt_decompressor_new.hpp:109-206) with the rows spliced in as raw int8 FORMAT/GT payload
// ------------------------------------------------------------------------------------------------------------------
int extract(const Options& o) {
    const double t0 = now();
    xsi_reader* rd = nullptr;
    int rc = xsi_reader_open(o.in.c_str(), &rd);
    if (rc != XSI_OK) { fprintf(stderr, "Failed to open file %s (rc %d)\n", o.in.c_str(), rc); return 1; }
    uint64_t S = 0, hap = 0, entries = 0, nvar = 0, rare = 0;
    uint32_t ploidy = 0, aet = 0, nblocks = 0, block_len = 0;
    int32_t zstd = 0, dph = 0;
    xsi_reader_info(rd, &S, &hap, &ploidy, &aet, &nblocks, &block_len, &entries, &nvar, &zstd, &rare, &dph);
    const std::string var_name = o.in + "_var.bcf";
    htsFile* vin = hts_open(var_name.c_str(), "r");
    if (!vin) { fprintf(stderr, "Failed to open file %s\n", var_name.c_str()); return 1; }
    if (o.threads > 1) hts_set_threads(vin, 2);
    bcf_hdr_t* hin = bcf_hdr_read(vin);
    if (!hin) { fprintf(stderr, "Failed to read the header of %s\n", var_name.c_str()); return 1; }
    // output header: create_output_file, gt_decompressor_new.hpp:471-530
    bcf_hdr_t* hout = bcf_hdr_dup(hin);
    bcf_hdr_remove(hout, BCF_HL_GEN, "XSI");
    bcf_hdr_remove(hout, BCF_HL_FMT, "BM");
    if (bcf_hdr_set_samples(hout, NULL, 0) < 0) { fprintf(stderr, "Failed to remove samples\n"); return 1; }
    for (uint64_t i = 0; i < S; ++i) bcf_hdr_add_sample(hout, xsi_reader_sample_name(rd, i));
    bcf_hdr_add_sample(hout, NULL);
    if (bcf_hdr_sync(hout) < 0) fprintf(stderr, "bcf_hdr_sync() failed ...\n");
    const char* flags = o.output_type == "u" ? "wbu" : "wb";  // gt_decompressor_new.hpp:439-447
    htsFile* fout = hts_open(o.out.c_str(), flags);
    if (!fout) { fprintf(stderr, "Could not open %s\n", o.out.c_str()); return 1; }
    if (o.threads > 1) hts_set_threads(fout, o.threads);
    if (bcf_hdr_write(fout, hout) < 0) { fprintf(stderr, "Could not write header to file %s\n", o.out.c_str()); return 1; }
    const int gt_id = bcf_hdr_id2int(hout, BCF_DT_ID, "GT");
    if (gt_id < 0) { fprintf(stderr, "no GT FORMAT in the header\n"); return 1; }

    xsi_ctx* ctx = nullptr;
    rc = xsi_create(o.device, &ctx);
    if (rc != XSI_OK) { fprintf(stderr, "xsi_create failed (no CUDA device? there is no CPU fallback)\n"); return 1; }
    const size_t N = (size_t)S * 2, stride8 = (N + 15) / 16 * 16;
    const size_t win_rows = std::max<size_t>(1, o.window_bytes / stride8);
    xsi_b200::Pinned win;
    win.reserve(win_rows * stride8);
    std::vector<bcf1_t*> recs(win_rows, nullptr);
    for (auto& r : recs) r = bcf_init();
    std::vector<uint32_t> blk(win_rows, 0), line(win_rows), nall(win_rows), filled(win_rows);
    std::vector<int32_t> wide;  // only for records with more than 63 alleles
    int32_t* bm = nullptr;
    int n_bm = 0;
    int64_t loaded = -1;
    uint64_t records = 0, genotypes = 0;
    bool eof = false, err = false;
    bcf1_t* pending = bcf_init();  // a record of the NEXT block, read while filling a window, opens the next window
    bool have_pending = false;
    while (!err && (!eof || have_pending)) {
        size_t n = 0;
        int64_t want = -1;
        while (n < win_rows) {
            bcf1_t* rec = recs[n];
            if (have_pending) { std::swap(recs[n], pending); rec = recs[n]; have_pending = false; }
            else {
                if (eof) break;
                const int rr = bcf_read(vin, hin, rec);
                if (rr < -1) { fprintf(stderr, "read error in %s\n", var_name.c_str()); err = true; break; }
                if (rr < 0) { eof = true; break; }
            }
            // Accessor::position_from_bm_entry, accessor.hpp:37-46
            if (bcf_unpack(rec, BCF_UN_ALL)) fprintf(stderr, "bcf_unpack error\n");
            if (bcf_get_format_int32(hin, rec, "BM", &bm, &n_bm) < 1) { fprintf(stderr, "BM key value not found\n"); err = true; break; }
            const uint32_t pos = (uint32_t)bm[0];
            const int64_t b = pos >> 15;
            if (want < 0) want = b;
            if (b != want) { std::swap(recs[n], pending); have_pending = true; break; }
            line[n] = pos & 0x7FFF;
            nall[n] = rec->n_allele;
            ++n;
        }
        if (err || n == 0) break;
        if (want != loaded) {
            const uint8_t* p = nullptr;
            uint64_t sz = 0;
            rc = xsi_reader_gt_block(rd, (uint32_t)want, &p, &sz);
            if (rc == XSI_OK) rc = xsi_decode_load_blocks(ctx, 1, &p, &sz, S, (int32_t)aet);
            if (rc != XSI_OK) { fprintf(stderr, "block %lld: %s (rc %d)\n", (long long)want, xsi_last_error(ctx), rc); err = true; break; }
            loaded = want;
        }
        bool all_i8 = true;
        for (size_t i = 0; i < n; ++i) all_i8 &= nall[i] >= 2 && nall[i] <= 63;
        if (all_i8) {
            rc = xsi_decode_records_i8(ctx, n, blk.data(), line.data(), nall.data(), win.as<int8_t>(), stride8, 0, filled.data(), nullptr, 0);
            if (rc != XSI_OK) { fprintf(stderr, "decode: %s (rc %d)\n", xsi_last_error(ctx), rc); err = true; break; }
        }
        for (size_t i = 0; i < n && !err; ++i) {
            bcf1_t* rec = recs[i];
            if (all_i8) {
                // the record leaves with ONE FORMAT field: [typed int key = GT][type: ploidy x int8][S * ploidy bytes], i.e. what
                // bcf_update_format(BM, NULL) + bcf_update_genotypes + bcf1_sync build (gt_decompressor_new.hpp:275-320)
                const uint32_t len = filled[i], pl = (uint32_t)(len / S);
                if (pl == 0 || pl > 2) { fprintf(stderr, "PLOIDY ERROR\n"); err = true; break; }
                rec->indiv.l = 0;
                bcf_enc_int1(&rec->indiv, gt_id);
                bcf_enc_size(&rec->indiv, (int)pl, BCF_BT_INT8);
                if (ks_resize(&rec->indiv, rec->indiv.l + len) != 0) { err = true; break; }
                memcpy(rec->indiv.s + rec->indiv.l, win.as<int8_t>() + i * stride8, len);
                rec->indiv.l += len;
                rec->n_fmt = 1;
                rec->n_sample = (uint32_t)S;
                rec->d.indiv_dirty = 0;
                rec->unpacked &= ~BCF_UN_FMT;  // d.fmt[] described the BM field that is gone
                genotypes += len;
            } else {  // wide alleles: the reference's own way
                wide.resize(N);
                uint32_t f = 0;
                rc = xsi_decode_records(ctx, 1, &blk[i], &line[i], &nall[i], wide.data(), N, 0, &f, nullptr, 0);
                if (rc != XSI_OK) { fprintf(stderr, "decode: %s (rc %d)\n", xsi_last_error(ctx), rc); err = true; break; }
                bcf_update_format(hin, rec, "BM", NULL, 0, BCF_HT_INT);
                if (bcf_update_genotypes(hout, rec, wide.data(), (int)f)) { fprintf(stderr, "Failed to update genotypes\n"); err = true; break; }
                genotypes += f;
            }
            if (bcf_write1(fout, hout, rec)) { fprintf(stderr, "Failed to write record\n"); err = true; break; }
            ++records;
        }
    }
    free(bm);
    bcf_destroy(pending);
    for (auto& r : recs) bcf_destroy(r);
    xsi_destroy(ctx);
    if (hts_close(fout) < 0) err = true;
    hts_close(vin);
    bcf_hdr_destroy(hin);
    bcf_hdr_destroy(hout);
    xsi_reader_close(rd);
    if (err) { fprintf(stderr, "Failure occurred, exiting...\n"); return 1; }
    printf("xsi_b200_bcf extract: records %llu genotypes %llu seconds %.6f threads %d\n", (unsigned long long)records,
           (unsigned long long)genotypes, now() - t0, o.threads);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// subset: `xsqueezeit -x -O x [-s list | -S file]` (gt_decompressor_new.hpp:130-143,241-273) with device-resident rows
// ------------------------------------------------------------------------------------------------------------------
int subset(const Options& o) {
    const double t0 = now();
    xsi_reader* rd = nullptr;
    int rc = xsi_reader_open(o.in.c_str(), &rd);
    if (rc != XSI_OK) { fprintf(stderr, "Failed to open file %s (rc %d)\n", o.in.c_str(), rc); return 1; }
    uint64_t S = 0, hap = 0, entries = 0, nvar = 0, rare = 0;
    uint32_t ploidy = 0, aet = 0, nblocks = 0, block_len = 0;
    int32_t zstd = 0, dph = 0;
    xsi_reader_info(rd, &S, &hap, &ploidy, &aet, &nblocks, &block_len, &entries, &nvar, &zstd, &rare, &dph);
    std::vector<std::string> sample_list;
    for (uint64_t i = 0; i < S; ++i) sample_list.push_back(xsi_reader_sample_name(rd, i));
    // enable_select_samples, gt_decompressor_new.hpp:324-365 (and the -S file form, :391-421)
    std::vector<uint32_t> use;
    bool select = false;
    std::string opt = o.samples;
    if (opt.empty() && !o.samples_file.empty()) {
        std::string file = o.samples_file;
        bool exclude = false;
        if (file[0] == '^') { exclude = true; file.erase(0, 1); }
        FILE* f = fopen(file.c_str(), "r");
        if (!f) { fprintf(stderr, "Could not open file %s\n", file.c_str()); return 1; }
        if (exclude) opt = "^";
        char buf[4096];
        while (fgets(buf, sizeof buf, f)) {
            std::string l(buf);
            while (!l.empty() && (l.back() == '\n' || l.back() == '\r')) l.pop_back();
            opt += l.substr(0, l.find('\t')) + ",";
        }
        fclose(f);
    }
    if (!opt.empty()) {
        select = true;
        bool inverse = opt[0] == '^';
        std::vector<std::string> named;
        size_t p = inverse ? 1 : 0;
        while (p <= opt.size()) {
            const size_t q = std::min(opt.find(',', p), opt.size());
            if (q > p) named.push_back(opt.substr(p, q - p));
            p = q + 1;
        }
        if (inverse) {
            for (uint32_t i = 0; i < S; ++i)
                if (std::find(named.begin(), named.end(), sample_list[i]) == named.end()) use.push_back(i);
        } else {  // bcftools has samples in order of option
            for (const auto& n : named) {
                auto it = std::find(sample_list.begin(), sample_list.end(), n);
                if (it != sample_list.end()) use.push_back((uint32_t)(it - sample_list.begin()));
            }
        }
    } else {
        for (uint32_t i = 0; i < S; ++i) use.push_back(i);
    }
    if (use.empty()) { fprintf(stderr, "No samples to extract\nNo samples found\n"); return 1; }
    const uint32_t n_sel = (uint32_t)use.size();
    // the new file's sample list: the selected names only when FEWER samples are kept (create_output_file, :483-490)
    std::string names;
    if (use.size() < sample_list.size()) for (uint32_t i : use) { names += sample_list[i]; names.push_back('\0'); }
    else for (const auto& n : sample_list) { names += n; names.push_back('\0'); }
    const uint64_t n_haps = (uint64_t)n_sel * ploidy;               // :491
    const uint64_t mac = (uint64_t)((double)n_haps * o.maf);        // :492

    const std::string var_name = o.in + "_var.bcf";
    htsFile* vin = hts_open(var_name.c_str(), "r");
    if (!vin) { fprintf(stderr, "Failed to open file %s\n", var_name.c_str()); return 1; }
    if (o.threads > 1) hts_set_threads(vin, 2);
    bcf_hdr_t* hin = bcf_hdr_read(vin);
    if (!hin) { fprintf(stderr, "Failed to read the header of %s\n", var_name.c_str()); return 1; }
    bcf_hdr_t* hout = bcf_hdr_dup(hin);
    bcf_hdr_remove(hout, BCF_HL_GEN, "XSI");
    {
        std::string tmp(o.out);
        bcf_hdr_append(hout, std::string("##XSI=").append(std::string(basename((char*)tmp.c_str()))).c_str());
    }
    bcf_hdr_add_sample(hout, NULL);
    if (bcf_hdr_sync(hout) < 0) fprintf(stderr, "bcf_hdr_sync() failed ...\n");
    const std::string var_out = o.out + "_var.bcf";
    htsFile* fout = hts_open(var_out.c_str(), "wb");
    if (!fout) { fprintf(stderr, "Could not open %s\n", var_out.c_str()); return 1; }
    if (o.threads > 1) hts_set_threads(fout, o.threads);
    if (bcf_hdr_write(fout, hout) < 0) { fprintf(stderr, "Could not write header to file %s\n", var_out.c_str()); return 1; }

    xsi_ctx* ctx = nullptr;
    rc = xsi_create(o.device, &ctx);
    if (rc != XSI_OK) { fprintf(stderr, "xsi_create failed (no CUDA device? there is no CPU fallback)\n"); return 1; }
    xsi_writer* w = nullptr;
    rc = xsi_writer_open(o.out.c_str(), n_sel, names.data(), block_len, mac, dph ? 1 : 0, (o.zstd || zstd) ? 1 : 0, o.zstd_level, &w);
    if (rc != XSI_OK) { fprintf(stderr, "Failed to open file %s (rc %d)\n", o.out.c_str(), rc); return 1; }

    const size_t batch_records = (size_t)block_len * (size_t)o.batch_blocks;
    const uint64_t stride = 2ull * n_sel;
    void* d_rows = nullptr;
    rc = xsi_device_alloc(ctx, &d_rows, batch_records * stride * 4);
    if (rc != XSI_OK) { fprintf(stderr, "device allocation failed\n"); return 1; }
    std::vector<bcf1_t*> recs(batch_records, nullptr);
    for (auto& r : recs) r = bcf_init();
    std::vector<uint32_t> blk(batch_records), line(batch_records), nall(batch_records), filled(batch_records);
    std::vector<uint8_t> pl(batch_records);
    std::vector<uint32_t> ac;
    std::vector<int32_t> ac_s;
    std::vector<std::vector<uint8_t>> held;  // copies of the GT blocks of a batch (a zstd reader reuses its buffer)
    int32_t* bm = nullptr;
    int n_bm = 0;
    uint64_t records = 0, genotypes = 0, out_block = 0;
    int max_ploidy = 0;
    bool eof = false, err = false;
    while (!eof && !err) {
        size_t n = 0;
        int64_t first_block = -1, last_block = -1;
        uint32_t max_all = 2;
        uint64_t variants = 0;
        while (n < batch_records) {
            bcf1_t* rec = recs[n];
            const int rr = bcf_read(vin, hin, rec);
            if (rr < -1) { fprintf(stderr, "read error in %s\n", var_name.c_str()); err = true; break; }
            if (rr < 0) { eof = true; break; }
            if (bcf_unpack(rec, BCF_UN_ALL)) fprintf(stderr, "bcf_unpack error\n");
            if (bcf_get_format_int32(hin, rec, "BM", &bm, &n_bm) < 1) { fprintf(stderr, "BM key value not found\n"); err = true; break; }
            const uint32_t pos = (uint32_t)bm[0];  // Accessor::position_from_bm_entry, accessor.hpp:37-46
            const int64_t b = pos >> 15;
            if (first_block < 0) first_block = b;
            if (b < last_block || b >= first_block + o.batch_blocks) { fprintf(stderr, "records of %s are not in block order\n", var_name.c_str()); err = true; break; }
            last_block = b;
            blk[n] = (uint32_t)(b - first_block);
            line[n] = pos & 0x7FFF;
            nall[n] = rec->n_allele;
            max_all = std::max(max_all, nall[n]);
            variants += rec->n_allele ? rec->n_allele - 1 : 0;
            ++n;
        }
        if (err || n == 0) break;
        // the blocks of this batch -> device, decoded and gathered there
        const uint32_t nb = (uint32_t)(last_block - first_block + 1);
        held.resize(nb);
        std::vector<const uint8_t*> ptrs(nb);
        std::vector<uint64_t> sizes(nb);
        for (uint32_t b = 0; b < nb && rc == XSI_OK; ++b) {
            const uint8_t* p = nullptr;
            uint64_t sz = 0;
            rc = xsi_reader_gt_block(rd, (uint32_t)first_block + b, &p, &sz);
            if (rc == XSI_OK) { held[b].assign(p, p + sz); ptrs[b] = held[b].data(); sizes[b] = sz; }
        }
        if (rc == XSI_OK) rc = xsi_decode_load_blocks(ctx, nb, ptrs.data(), sizes.data(), S, (int32_t)aet);
        const uint32_t ac_stride = max_all - 1;
        ac.assign(n * (size_t)ac_stride, 0);
        if (rc == XSI_OK)
            rc = xsi_decode_records_subset(ctx, n, blk.data(), line.data(), nall.data(), use.data(), n_sel, static_cast<int32_t*>(d_rows),
                                           stride, 1, filled.data(), ac.data(), ac_stride);
        if (rc != XSI_OK) { fprintf(stderr, "decode: %s (rc %d)\n", xsi_last_error(ctx), rc); err = true; break; }
        for (size_t i = 0; i < n; ++i) {
            const uint32_t p = filled[i] / n_sel;  // CURRENT_LINE_PLOIDY of the selected row, :209-220
            if (p == 0 || p > 2) { fprintf(stderr, "PLOIDY ERROR\n"); err = true; break; }
            pl[i] = (uint8_t)p;
            genotypes += filled[i];
        }
        if (err) break;
        // ... and encoded from there as the next blocks of the new file
        xsi_encode_desc d;
        memset(&d, 0, sizeof d);
        d.n_records = n; d.n_samples = n_sel; d.block_len = block_len; d.mac_threshold = mac; d.default_phasing = dph ? 1 : 0;
        d.gt_elem_bytes = 4; d.gt_on_device = 1; d.gt = d_rows; d.n_allele = nall.data(); d.ploidy = pl.data();
        uint32_t nbo = 0;
        const uint8_t* const* bo = nullptr;
        const uint64_t* so = nullptr;
        rc = xsi_encode_launch_strided(ctx, &d, stride);
        if (rc == XSI_OK) rc = xsi_encode_collect(ctx, &nbo, &bo, &so);
        if (rc == XSI_OK) rc = xsi_writer_add_blocks(w, nbo, bo, so, n, variants);
        if (rc != XSI_OK) { fprintf(stderr, "encode: %s (rc %d)\n", xsi_last_error(ctx), rc); err = true; break; }
        max_ploidy = std::max(max_ploidy, xsi_encode_max_ploidy(ctx));
        // the companion: new BM, AC / AN as bcftools view -s recomputes them (update_and_write_xsi, :241-273)
        for (size_t i = 0; i < n && !err; ++i) {
            bcf1_t* rec = recs[i];
            const uint64_t r = records + i;
            const uint64_t nblk = r / block_len;
            if (nblk != out_block) out_block = nblk;
            // block_id << 15 | offset of the NEW file (:169-180); records are not filtered here, so the offset inside the block is the old one
            int32_t v = (int32_t)((uint32_t)nblk << 15 | line[i]);
            bcf_update_format(hin, rec, "BM", &v, 1, BCF_HT_INT);
            if (select) {
                ac_s.assign(nall[i] - 1, 0);
                for (uint32_t a = 0; a + 1 < nall[i]; ++a) ac_s[a] = (int32_t)ac[i * (size_t)ac_stride + a];
                int32_t an = (int32_t)filled[i];
                bcf_update_info_int32(hout, rec, "AC", ac_s.data(), (int)(nall[i] - 1));
                bcf_update_info_int32(hout, rec, "AN", &an, 1);
            }
            if (bcf_write1(fout, hout, rec)) { fprintf(stderr, "Failed to write record\n"); err = true; break; }
        }
        records += n;
    }
    free(bm);
    for (auto& r : recs) bcf_destroy(r);
    xsi_device_free(ctx, d_rows);
    xsi_destroy(ctx);
    if (w) {
        // sic: the extractor finalises its factory WITHOUT the ploidy it has seen (gt_decompressor_new.hpp:138), so the interface's
        // default applies (finalize_file(max_ploidy = 2), xsi_factory.hpp:43): header.ploidy = 2 and hap_samples = 2 * samples even
        // when every record of the new file is haploid
        (void)max_ploidy;
        rc = xsi_writer_close(w, 2);
        if (rc != XSI_OK) { fprintf(stderr, "finalize failed (rc %d)\n", rc); err = true; }
    }
    if (hts_close(fout) < 0) err = true;
    hts_close(vin);
    bcf_hdr_destroy(hin);
    bcf_hdr_destroy(hout);
    xsi_reader_close(rd);
    if (err) { fprintf(stderr, "Failure occurred, exiting...\n"); return 1; }
    printf("xsi_b200_bcf subset: records %llu selected_samples %u genotypes %llu seconds %.6f\n", (unsigned long long)records, n_sel,
           (unsigned long long)genotypes, now() - t0);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    const std::string cmd = argc > 1 ? argv[1] : "";
    if (argc < 4 || (cmd != "compress" && cmd != "extract" && cmd != "subset")) {
        fprintf(stderr, "usage: %s subset in.xsi out.xsi [--samples a,b|^a,b] [--samples-file f] [--maf f] [--zstd] [--batch-blocks k]\n", argv[0]);
        fprintf(stderr, "usage: %s extract in.xsi out.bcf [-O b|u] [--threads t] [--window-bytes n] [--device d]\n", argv[0]);
        fprintf(stderr, "usage: %s compress in.bcf out.xsi [--maf f] [--variant-block-length n] [--zstd] [--zstd-level l]\n"
                        "          [--wah-encode-missing] [--threads t] [--batch-blocks k] [--device d]\n", argv[0]);
        return 2;
    }
    Options o;
    o.in = argv[2];
    o.out = argv[3];
    for (int i = 4; i < argc; ++i) {
        const std::string a = argv[i];
        auto val = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "%s needs a value\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--maf") o.maf = atof(val());
        else if (a == "--variant-block-length") o.block_len = (size_t)atoll(val());
        else if (a == "--zstd") o.zstd = true;
        else if (a == "--zstd-level") o.zstd_level = atoi(val());
        else if (a == "--wah-encode-missing") o.wah_missing = true;
        else if (a == "--threads") o.threads = atoi(val());
        else if (a == "--batch-blocks") o.batch_blocks = atoi(val());
        else if (a == "--device") o.device = atoi(val());
        else if (a == "--reference-var") o.reference_var = true;
        else if (a == "-O" || a == "--output-type") o.output_type = val();
        else if (a == "--window-bytes") o.window_bytes = (size_t)atoll(val());
        else if (a == "--samples" || a == "-s") o.samples = val();
        else if (a == "--samples-file" || a == "-S") o.samples_file = val();
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (o.block_len == 0 || o.batch_blocks < 1) return 2;
    try {
        return cmd == "compress" ? compress(o) : cmd == "extract" ? extract(o) : subset(o);
    } catch (const char* e) {
        fprintf(stderr, "%s\nFailure occurred, exiting...\n", e);
        return 1;
    }
}
