// xsi_container.cpp -- host container layer of include/xsi_b200.h: the .xsi file itself.
// Byte-compatible with the reference writer XsiFactoryExt (include/xsi_factory.hpp:435-639,
// IBinaryBlock::write_to_file include/interfaces.hpp:176-268, header_t include/compression.hpp:40-104)
// and readable like Accessor / AccessorInternalsNewTemplate (accessor.cpp:26-82,
// include/accessor_internals_new.hpp:763-893).  No genotype arithmetic happens here.
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/xsi_b200.h"
#include "host_util.hpp"

using namespace xsi;

namespace {

// ---- optional host libzstd (the reference links it; here it is loaded on demand) ----
struct Zstd {
    void* h = nullptr;
    size_t (*compress)(void*, size_t, const void*, size_t, int) = nullptr;
    size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    bool ok() const { return compress && decompress && isError; }
    static Zstd& get() {
        static Zstd z = load();  // function-local static: initialised once, thread-safe
        return z;
    }
    static Zstd load() {
        Zstd z;
        for (const char* n : {"libzstd.so.1", "libzstd.so"}) { z.h = dlopen(n, RTLD_NOW); if (z.h) break; }
        if (z.h) {
            z.compress = (decltype(z.compress))dlsym(z.h, "ZSTD_compress");
            z.decompress = (decltype(z.decompress))dlsym(z.h, "ZSTD_decompress");
            z.isError = (decltype(z.isError))dlsym(z.h, "ZSTD_isError");
        }
        return z;
    }
};

#pragma pack(push, 1)
struct Header {  // compression.hpp:40-104
    uint32_t endianness, first_magic, version;
    uint8_t ploidy, ind_bytes, aet_bytes, wah_bytes, special_bitset, specific_bitset, rsvd_bs[2];
    uint32_t rsvd_1[3];
    uint64_t hap_samples, num_variants;
    uint32_t block_size, number_of_blocks, ss_rate, number_of_ssas;
    uint64_t wahs_offset, indices_offset, samples_offset;
    uint32_t rearrangement_track_offset, sparse_offset;
    uint32_t rare_threshold;
    uint64_t xcf_entries;
    uint32_t phase_info_offset;
    uint64_t num_samples;
    uint8_t rsvd_3[104];
    uint32_t rsvd_4[3], sample_name_chksum, bcf_file_chksum, data_chksum, header_chksum, last_magic;
};
#pragma pack(pop)
static_assert(sizeof(Header) == 256, "header is 256 bytes");
constexpr uint32_t MAGIC = 0xfeed1767u, ENDIANNESS = 0xaabbccddu;

}  // namespace

struct xsi_writer {
    FILE* f = nullptr;
    std::string path;
    uint32_t n_samples = 0, block_len = 0;
    uint64_t mac_threshold = 0;
    int default_phasing = 0, zstd_on = 0, zstd_level = 7;
    std::vector<std::string> samples;
    std::vector<uint64_t> indices;
    uint64_t entries = 0, variants = 0;
    bool failed = false;  // a block write failed: the file is not finalised (no index over a half-written block)
};

extern "C" int xsi_writer_open(const char* path, uint32_t n_samples, const char* sample_names, uint32_t block_len,
                               uint64_t mac_threshold, int32_t default_phasing, int32_t zstd_on, int32_t zstd_level,
                               xsi_writer** out) {
    if (!path || !out || block_len == 0) return XSI_E_ARG;
    if (zstd_on && !Zstd::get().ok()) return XSI_E_ZSTD;
    xsi_writer* w = new xsi_writer();
    w->f = fopen(path, "wb");
    if (!w->f) { delete w; return XSI_E_IO; }
    w->path = path; w->n_samples = n_samples; w->block_len = block_len; w->mac_threshold = mac_threshold;
    w->default_phasing = default_phasing ? 1 : 0; w->zstd_on = zstd_on ? 1 : 0; w->zstd_level = zstd_level;
    const char* sn = sample_names;
    for (uint32_t i = 0; i < n_samples; ++i) {
        if (sn) { w->samples.emplace_back(sn); sn += w->samples.back().size() + 1; }
        else w->samples.push_back("S" + std::to_string(i));
    }
    Header h;
    memset(&h, 0, sizeof(h));
    if (fwrite(&h, 1, sizeof(h), w->f) != sizeof(h)) { fclose(w->f); delete w; return XSI_E_IO; }  // placeholder, xsi_factory.hpp:500
    *out = w;
    return XSI_OK;
}

extern "C" int xsi_writer_add_blocks(xsi_writer* w, uint32_t n_blocks, const uint8_t* const* blocks, const uint64_t* sizes,
                                     uint64_t n_records, uint64_t n_variants) {
    if (!w || !w->f || (n_blocks && (!blocks || !sizes))) return XSI_E_ARG;
    if (w->failed) return XSI_E_IO;
    struct Fail { xsi_writer* w; bool ok = false; ~Fail() { if (!ok) w->failed = true; } } guard{w};
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const off_t at = ftello(w->f);
        if (at < 0) return XSI_E_IO;
        w->indices.push_back((uint64_t)at);  // xsi_factory.hpp:533
        // outer dictionary: {KEY_GT_ENTRY: 16} (interfaces.hpp:187-221), then the GT block
        const uint32_t outer[4] = {0xFFFFFFFFu, 1u, KEY_GT_ENTRY, 16u};
        if (!w->zstd_on) {
            if (fwrite(outer, 1, 16, w->f) != 16 || fwrite(blocks[b], 1, sizes[b], w->f) != sizes[b]) return XSI_E_IO;
        } else {
            // interfaces.hpp:241-252,291-314: u64 compressed size, u64 original size, zstd frame
            std::vector<uint8_t> raw(16 + sizes[b]);
            memcpy(raw.data(), outer, 16);
            memcpy(raw.data() + 16, blocks[b], sizes[b]);
            std::vector<uint8_t> comp(raw.size() * 2);
            const size_t r = Zstd::get().compress(comp.data(), comp.size(), raw.data(), raw.size(), w->zstd_level);
            if (Zstd::get().isError(r)) return XSI_E_ZSTD;
            const uint64_t cs = r, os = raw.size();
            if (fwrite(&cs, 8, 1, w->f) != 1 || fwrite(&os, 8, 1, w->f) != 1 || fwrite(comp.data(), 1, r, w->f) != r) return XSI_E_IO;
        }
        const off_t end = ftello(w->f);  // interfaces.hpp:254-263
        if (end < 0) return XSI_E_IO;
        const uint64_t pos = (uint64_t)end;
        if (pos % 4) { const char z[4] = {0, 0, 0, 0}; if (fwrite(z, 1, 4 - pos % 4, w->f) != 4 - pos % 4) return XSI_E_IO; }
    }
    w->entries += n_records;
    w->variants += n_variants;
    guard.ok = true;
    return XSI_OK;
}

// finalize_file (xsi_factory.hpp:543-606) at the current position of `f` (just past the last block): padding to 8, the
// u64 block index, the sample names, then the header at offset 0
static int finalize_file(FILE* f, const xsi_writer& w, int32_t max_ploidy) {
    int rc = XSI_OK;
    const off_t at = ftello(f);  // xsi_factory.hpp:558-565
    if (at < 0) return XSI_E_IO;
    uint64_t pos = (uint64_t)at;
    if (pos % 8) { const char z[8] = {0}; if (fwrite(z, 1, 8 - pos % 8, f) != 8 - pos % 8) rc = XSI_E_IO; pos += 8 - pos % 8; }
    Header h;
    memset(&h, 0, sizeof(h));
    h.endianness = ENDIANNESS; h.first_magic = MAGIC; h.last_magic = MAGIC;
    h.version = 5;  // xsi_factory.hpp:469
    h.ploidy = (uint8_t)max_ploidy; h.ind_bytes = 4;
    h.aet_bytes = ((uint64_t)w.n_samples * 2 <= 65535) ? 2 : 4;  // gt_compressor_new.hpp:182
    h.wah_bytes = 2;
    h.special_bitset = (uint8_t)(w.default_phasing << 2);
    h.specific_bitset = (uint8_t)(0x01 | (w.zstd_on << 2));
    h.hap_samples = (uint64_t)w.n_samples * (uint64_t)max_ploidy;
    h.num_variants = w.variants;
    h.block_size = 0; h.number_of_blocks = 1;
    h.ss_rate = w.block_len;
    h.number_of_ssas = (uint32_t)((w.entries + (uint32_t)w.block_len - 1) / (uint32_t)w.block_len);
    h.wahs_offset = 256;
    h.indices_offset = pos;
    if (!w.indices.empty() && fwrite(w.indices.data(), 8, w.indices.size(), f) != w.indices.size()) rc = XSI_E_IO;
    h.samples_offset = pos + w.indices.size() * 8;
    for (const std::string& s : w.samples) if (fwrite(s.c_str(), 1, s.size() + 1, f) != s.size() + 1) rc = XSI_E_IO;
    h.rearrangement_track_offset = 0xFFFFFFFFu; h.sparse_offset = 0xFFFFFFFFu;
    h.rare_threshold = (uint32_t)w.mac_threshold;
    h.xcf_entries = w.entries;
    h.num_samples = w.n_samples;
    if (fflush(f) != 0 || fseeko(f, 0, SEEK_SET) != 0) rc = XSI_E_IO;
    if (rc == XSI_OK && fwrite(&h, 1, sizeof(h), f) != sizeof(h)) rc = XSI_E_IO;  // a failed finalise keeps the magic-less placeholder
    return rc;
}

extern "C" int xsi_writer_close(xsi_writer* w, int32_t max_ploidy) {
    if (!w) return XSI_E_ARG;
    int rc = XSI_OK;
    if (w->f && w->failed) {  // leave the placeholder header (no magic): readers refuse the file
        fclose(w->f);
        delete w;
        return XSI_E_IO;
    }
    if (w->f) {
        rc = finalize_file(w->f, *w, max_ploidy);
        if (fclose(w->f) != 0) rc = XSI_E_IO;
    }
    delete w;
    return rc;
}

// Several writers, one file (one rank per GPU, xsqueezeit_b200/sharded.py): every rank has written its blocks at the offsets
// of the global table (all-gather of the per-block byte counts); one rank then calls this to add what the single writer's
// close adds -- same code, so the file is the single writer's byte for byte.  indices: absolute offset of every block;
// end_of_blocks: first byte after the last block.
extern "C" int xsi_writer_finalize_sharded(const char* path, uint32_t n_samples, const char* sample_names, uint32_t block_len,
                                           uint64_t mac_threshold, int32_t default_phasing, int32_t max_ploidy, uint32_t n_blocks,
                                           const uint64_t* indices, uint64_t end_of_blocks, uint64_t n_records, uint64_t n_variants) {
    if (!path || block_len == 0 || (n_blocks && !indices)) return XSI_E_ARG;
    xsi_writer w;
    w.n_samples = n_samples; w.block_len = block_len; w.mac_threshold = mac_threshold;
    w.default_phasing = default_phasing ? 1 : 0; w.zstd_on = 0;
    const char* sn = sample_names;
    for (uint32_t i = 0; i < n_samples; ++i) {
        if (sn) { w.samples.emplace_back(sn); sn += w.samples.back().size() + 1; }
        else w.samples.push_back("S" + std::to_string(i));
    }
    w.indices.assign(indices, indices + n_blocks);
    w.entries = n_records; w.variants = n_variants;
    FILE* f = fopen(path, "r+b");
    if (!f) return XSI_E_IO;
    int rc = fseeko(f, (off_t)end_of_blocks, SEEK_SET) == 0 ? finalize_file(f, w, max_ploidy) : XSI_E_IO;
    if (fclose(f) != 0) rc = XSI_E_IO;
    return rc;
}

// ---------------------------------------------------------------------------------------------
struct xsi_reader {
    int fd = -1;
    const uint8_t* map = nullptr;
    size_t size = 0;
    Header h;
    std::vector<std::string> samples;
    uint32_t n_blocks = 0;
    std::vector<std::vector<uint8_t>> inflated;  // per block, filled lazily for zstd files
};

extern "C" int xsi_reader_open(const char* path, xsi_reader** out) {
    if (!path || !out) return XSI_E_ARG;
    *out = nullptr;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return XSI_E_IO;
    struct stat st;
    if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(Header)) { close(fd); return XSI_E_FORMAT; }
    void* m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_SHARED, fd, 0);
    if (m == MAP_FAILED) { close(fd); return XSI_E_IO; }
    xsi_reader* r = new xsi_reader();
    r->fd = fd; r->map = (const uint8_t*)m; r->size = (size_t)st.st_size;
    memcpy(&r->h, r->map, sizeof(Header));
    const Header& h = r->h;
    bool ok = h.first_magic == MAGIC && h.last_magic == MAGIC && h.endianness == ENDIANNESS;  // accessor.cpp:37-41
    ok = ok && (h.version == 4 || h.version == 5);                                             // accessor_internals_new.hpp:782
    ok = ok && (h.aet_bytes == 2 || h.aet_bytes == 4) && h.ploidy != 0;
    ok = ok && h.indices_offset <= r->size && h.samples_offset <= r->size && h.indices_offset <= h.samples_offset;
    if (!ok) { xsi_reader_close(r); return XSI_E_FORMAT; }
    const size_t isz = h.version >= 5 ? 8 : 4;  // accessor_internals_new.hpp:849-855
    r->n_blocks = (uint32_t)((h.samples_offset - h.indices_offset) / isz);
    const uint64_t n_names = h.hap_samples / h.ploidy;  // accessor.cpp:55-58
    size_t p = h.samples_offset;
    while (r->samples.size() < n_names && p < r->size) {
        const void* z = memchr(r->map + p, 0, r->size - p);
        if (!z) break;
        r->samples.emplace_back((const char*)(r->map + p));
        p = (const uint8_t*)z - r->map + 1;
    }
    r->inflated.resize(r->n_blocks);
    *out = r;
    return XSI_OK;
}

extern "C" void xsi_reader_close(xsi_reader* r) {
    if (!r) return;
    if (r->map) munmap((void*)r->map, r->size);
    if (r->fd >= 0) close(r->fd);
    delete r;
}

extern "C" int xsi_reader_info(const xsi_reader* r, uint64_t* num_samples, uint64_t* hap_samples, uint32_t* ploidy,
                               uint32_t* aet_bytes, uint32_t* n_blocks, uint32_t* block_len, uint64_t* xcf_entries,
                               uint64_t* num_variants, int32_t* zstd, uint64_t* rare_threshold, int32_t* default_phased) {
    if (!r) return XSI_E_ARG;
    const Header& h = r->h;
    if (num_samples) *num_samples = h.num_samples ? h.num_samples : h.hap_samples / 2;  // accessor_internals_new.hpp:53
    if (hap_samples) *hap_samples = h.hap_samples;
    if (ploidy) *ploidy = h.ploidy;
    if (aet_bytes) *aet_bytes = h.aet_bytes;
    if (n_blocks) *n_blocks = r->n_blocks;
    if (block_len) *block_len = h.ss_rate;
    if (xcf_entries) *xcf_entries = h.xcf_entries;
    if (num_variants) *num_variants = h.num_variants;
    if (zstd) *zstd = (h.specific_bitset >> 2) & 1;
    if (rare_threshold) *rare_threshold = h.rare_threshold;
    if (default_phased) *default_phased = (h.special_bitset >> 2) & 1;
    return XSI_OK;
}

extern "C" const char* xsi_reader_sample_name(const xsi_reader* r, uint64_t i) {
    return (r && i < r->samples.size()) ? r->samples[i].c_str() : nullptr;
}

extern "C" int xsi_reader_gt_block(xsi_reader* r, uint32_t b, const uint8_t** ptr, uint64_t* size) {
    if (!r || b >= r->n_blocks || !ptr || !size) return XSI_E_ARG;
    const Header& h = r->h;
    const uint64_t off = h.version >= 5 ? rd_u64(r->map + h.indices_offset + 8 * (size_t)b)
                                        : rd_u32(r->map + h.indices_offset + 4 * (size_t)b);
    const uint64_t next = (b + 1 < r->n_blocks)
                              ? (h.version >= 5 ? rd_u64(r->map + h.indices_offset + 8 * (size_t)(b + 1))
                                                : rd_u32(r->map + h.indices_offset + 4 * (size_t)(b + 1)))
                              : h.indices_offset;
    if (off >= r->size || next > r->size || next < off + 16) return XSI_E_FORMAT;
    const uint8_t* outer = r->map + off;
    uint64_t outer_size = next - off;
    if ((h.specific_bitset >> 2) & 1) {  // accessor_internals_new.hpp:857-886
        std::vector<uint8_t>& buf = r->inflated[b];
        if (buf.empty()) {
            if (!Zstd::get().ok()) return XSI_E_ZSTD;
            uint64_t cs, us;
            const uint8_t* frame;
            if (h.version >= 5) { cs = rd_u64(outer); us = rd_u64(outer + 8); frame = outer + 16; }
            else { cs = rd_u32(outer); us = rd_u32(outer + 4); frame = outer + 8; }
            if (frame + cs > r->map + r->size) return XSI_E_FORMAT;
            buf.resize(us);
            const size_t res = Zstd::get().decompress(buf.data(), us, frame, cs);
            if (Zstd::get().isError(res) || res != us) { buf.clear(); return XSI_E_ZSTD; }
        }
        outer = buf.data();
        outer_size = buf.size();
    }
    // outer dictionary -> KEY_GT_ENTRY (accessor_internals_new.hpp:830-843, interfaces.hpp:77-90)
    if (outer_size < 8 || rd_u32(outer) != 0xFFFFFFFFu) return XSI_E_FORMAT;
    const uint32_t n = rd_u32(outer + 4);
    if (8 + (uint64_t)n * 8 > outer_size) return XSI_E_FORMAT;
    uint32_t gt_off = VAL_UNDEFINED;
    for (uint32_t i = 0; i < n; ++i) if (rd_u32(outer + 8 + 8 * (size_t)i) == KEY_GT_ENTRY) gt_off = rd_u32(outer + 12 + 8 * (size_t)i);
    if (gt_off == VAL_UNDEFINED || gt_off >= outer_size) return XSI_E_FORMAT;
    *ptr = outer + gt_off;
    *size = outer_size - gt_off;  // includes up to 3 bytes of alignment padding for plain files
    return XSI_OK;
}
