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
// What is kept from the reference's host side, called from its own objects (nothing copied): seek_default_phased and
// seek_max_ploidy_from_first_entry (xcf.cpp:811-862), and the `_var.bcf` companion writer + CSI index
// (replace_samples_by_pos_in_binary_matrix xcf.cpp:641-714, create_index_file xcf.cpp:39-57) on a second thread exactly as
// xsqueezeit.cpp:119-128 runs it.
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

#include "xcf.hpp"  // reference host helpers (declarations only; objects come from oracle/_ref/obj)

#include "xsi_b200_runtime.hpp"

namespace {

double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Batch {
    xsi_b200::Pinned rows;
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
    int state[2] = {0, 0};  // 0 = free (reader owns), 1 = full (encoder owns)
    std::string error;
};

struct Options {
    std::string in, out;
    double maf = 0.001;
    size_t block_len = 8192;
    bool zstd = false, wah_missing = false;
    int zstd_level = 7, threads = 8, batch_blocks = 4, device = 0;
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

int compress(const Options& o) {
    const double t0 = now();
    // file-level parameters exactly as GtCompressorStream::compress_to_file finds them (gt_compressor_new.hpp:84-109)
    const int32_t default_phased = seek_default_phased(o.in);
    const size_t first_ploidy = seek_max_ploidy_from_first_entry(o.in);
    bool fail = false;
    std::thread variant_thread([&] {  // xsqueezeit.cpp:119-128
        try {
            replace_samples_by_pos_in_binary_matrix(o.in, o.out + "_var.bcf", o.out, true, o.block_len);
        } catch (const char* e) {
            fprintf(stderr, "%s\n", e);
            fail = true;
        }
        create_index_file(o.out + "_var.bcf");
    });

    htsFile* fp = hts_open(o.in.c_str(), "r");
    if (!fp) { fprintf(stderr, "Failed to open file %s\n", o.in.c_str()); variant_thread.join(); return 1; }
    if (o.threads > 1) hts_set_threads(fp, o.threads);
    bcf_hdr_t* hdr = bcf_hdr_read(fp);
    if (!hdr) { fprintf(stderr, "Failed to read the header of %s\n", o.in.c_str()); variant_thread.join(); return 1; }
    const size_t S = (size_t)bcf_hdr_nsamples(hdr);
    std::string names;
    for (size_t i = 0; i < S; ++i) { names += hdr->samples[i]; names.push_back('\0'); }
    const uint64_t mac = (uint64_t)((double)(S * first_ploidy) * o.maf);  // gt_compressor_new.hpp:98-99

    xsi_writer* w = nullptr;
    int rc = xsi_writer_open(o.out.c_str(), (uint32_t)S, names.data(), (uint32_t)o.block_len, mac, default_phased, o.zstd ? 1 : 0,
                             o.zstd_level, &w);
    if (rc != XSI_OK) { fprintf(stderr, "Failed to open file %s (rc %d)\n", o.out.c_str(), rc); variant_thread.join(); return 1; }

    Exchange ex;
    int max_ploidy = (int)first_ploidy;
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
            b.n_elems += ngt;
            b.n_allele.push_back((uint32_t)rec->n_allele);
            b.ploidy.push_back((uint8_t)pl);
            b.variants += rec->n_allele ? rec->n_allele - 1 : 0;
            ++records;
            genotypes += ngt;
        }
        b.last = eof || err;
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
    rc = xsi_writer_close(w, max_ploidy);
    if (rc != XSI_OK) { fprintf(stderr, "finalize failed (rc %d)\n", rc); err = true; }
    const double t_gt = now();
    variant_thread.join();
    const double t1 = now();
    if (fail || err) { fprintf(stderr, "Failure occurred, exiting...\n"); return 1; }
    printf("xsi_b200_bcf compress: records %llu genotypes %llu seconds %.6f gt_path_seconds %.6f encode_thread_seconds %.6f "
           "reader_seconds %.6f threads %d batch_blocks %d\n",
           (unsigned long long)records, (unsigned long long)genotypes, t1 - t0, t_gt - t0, t_encode, t_read_done - t0, o.threads, o.batch_blocks);
    return 0;
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4 || std::string(argv[1]) != "compress") {
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
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
    }
    if (o.block_len == 0 || o.batch_blocks < 1) return 2;
    try {
        return compress(o);
    } catch (const char* e) {
        fprintf(stderr, "%s\nFailure occurred, exiting...\n", e);
        return 1;
    }
}
