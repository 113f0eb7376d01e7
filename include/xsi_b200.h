/* include/xsi_b200.h -- C ABI of the B200-native xSqueezeIt genotype encode/decode path.
 *
 * This is the drop-in boundary: plain pointers and sizes, int return codes, no C++ or torch
 * types, no exceptions across the boundary.  Every entry point names the reference interface
 * (rwk-unil/xSqueezeIt @55ad8c7, paths under /root/reference) it replaces.  INTEGRATION.md
 * shows the reference-side binding a maintainer would add.
 *
 * There is NO CPU fallback: every call that does genotype work needs a CUDA device and fails
 * with XSI_E_CUDA when there is none.
 *
 * Genotype values use the htslib int32 encoding of bcf_get_genotypes (htslib/htslib/vcf.h:892-898,
 * 1324,1329): (allele+1)<<1 | phased, 0/1 = missing, INT32_MIN = bcf_int32_missing,
 * INT32_MIN+1 = bcf_int32_vector_end.  gt_elem_bytes = 1 takes the raw BCF FORMAT/GT int8
 * payload instead (same encoding, 0x80 = missing, 0x81 = vector end; bcf_fmt_t.p, vcf.h:152-158).
 */
#ifndef XSI_B200_H
#define XSI_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
    XSI_OK = 0,
    XSI_E_CUDA = -1,        /* no device / CUDA runtime error (see xsi_last_error) */
    XSI_E_ARG = -2,         /* bad argument */
    XSI_E_ALLELE = -3,      /* reference: throw "Unknown allele error !"  (gt_block.hpp:259-265) */
    XSI_E_PLOIDY = -4,      /* reference: "Ploidy higher than 2 is not yet supported" (gt_compressor_new.hpp:118-120) */
    XSI_E_UNSUPPORTED = -5, /* shapes the reference itself cannot round-trip (e.g. 32768..65535 samples) */
    XSI_E_FORMAT = -6,      /* malformed .xsi / GT block */
    XSI_E_NOMEM = -7,
    XSI_E_IO = -8,
    XSI_E_ZSTD = -9,        /* libzstd not loadable / (de)compression error */
};

typedef struct xsi_ctx xsi_ctx; /* one per (thread, device); owns a CUDA stream and scratch pools */

int  xsi_create(int device, xsi_ctx** out);
void xsi_destroy(xsi_ctx* ctx);
const char* xsi_last_error(const xsi_ctx* ctx); /* never NULL */
const char* xsi_version(void);
/* CUDA stream of the context as a cudaStream_t cast to void* (for event timing by the caller) */
void* xsi_stream(xsi_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t xsi_kernel_launches(const xsi_ctx* ctx);

/* Per-kernel timing with CUDA events on the context stream (measurement aid, off by default).
 * xsi_profile_read waits for the stream and returns "name launches total_ms\n" lines for
 * everything launched since the previous read (string owned by ctx).                          */
int xsi_profile(xsi_ctx* ctx, int on);
const char* xsi_profile_read(xsi_ctx* ctx);

/* Page-locked host memory (cudaHostAlloc, portable) for bindings that stage rows for xsi_encode_launch or
 * receive rows from xsi_decode_records*: pinned buffers cross PCIe by DMA without a bounce copy. */
int  xsi_host_alloc(void** p, uint64_t bytes);
void xsi_host_free(void* p);
/* Device memory on the context's GPU for callers without a CUDA runtime of their own (rows handed from xsi_decode_records*
 * with out_on_device = 1 to xsi_encode_launch[_strided] with gt_on_device = 1). */
int  xsi_device_alloc(xsi_ctx* ctx, void** p, uint64_t bytes);
void xsi_device_free(xsi_ctx* ctx, void* p);

/* ------------------------------------------------------------------------------------------
 * ENCODE  -- replaces GtBlock<A_T,uint16_t>::encode_line + write_to_stream
 *            (include/gt_block.hpp:185-204,279-406), i.e. the IWritableBCFLineEncoder the
 *            reference registers under KEY_GT_ENTRY (include/xsi_factory.hpp:419-433),
 *            for a batch of whole blocks at a time.
 * ------------------------------------------------------------------------------------------ */
typedef struct {
    uint64_t n_records;        /* records in this batch; blocks are cut every block_len records          */
    uint32_t n_samples;        /* bcf_hdr_nsamples                                                          */
    uint32_t block_len;        /* --variant-block-length (xsqueezeit.hpp:113, default 8192)                 */
    uint64_t mac_threshold;    /* (size_t)((double)N_HAPS*MAF), gt_compressor_new.hpp:98-99                 */
    int32_t  default_phasing;  /* seek_default_phased, xcf.cpp:811-836                                      */
    int32_t  gt_elem_bytes;    /* 4: int32 rows (bcf_get_genotypes); 1: raw BCF int8 rows                   */
    int32_t  gt_on_device;     /* 0: gt is host memory (copied H2D inside the call); 1: device pointer      */
    int32_t  wah_encode_missing; /* --wah-encode-missing (xsqueezeit.hpp:58, gt_block.hpp:174-176): missing and
                                  end-of-vector lines as natural-order WAH (WS_WAH) instead of index lists  */
    const void*     gt;        /* rows back to back: row r starts at element sum(ploidy[0..r))*n_samples    */
    const uint32_t* n_allele;  /* host, per record: bcf1_t::n_allele                                        */
    const uint8_t*  ploidy;    /* host, per record: ngt / n_samples (1 or 2); NULL = all 2                  */
} xsi_encode_desc;

/* Asynchronous part: uploads (if needed), runs every encode kernel on the context stream.  */
int xsi_encode_launch(xsi_ctx* ctx, const xsi_encode_desc* desc);
/* The same launch for DEVICE rows that are not back to back: row r starts at element r * row_stride (row_stride >= 2 * n_samples;
 * a haploid row uses the first n_samples elements).  This is the layout xsi_decode_records* write with out_on_device = 1, so the
 * extractor's XSI -> XSI path (-Ox with -s/-S: decode, select samples, encode again; XsiFactoryExt::append of the selected row,
 * include/gt_decompressor_new.hpp:241-273) runs without the rows leaving the device.                                        */
int xsi_encode_launch_strided(xsi_ctx* ctx, const xsi_encode_desc* desc, uint64_t row_stride);
/* on = 1: xsi_encode_launch of DEVICE rows returns at once; the batch is encoded by a thread of the library on a stream of
 * its own and xsi_encode_collect waits for it (errors of the launch are reported there).  The caller may decode another batch
 * on the same context meanwhile (xsi_decode_*): one host thread then keeps the PBWT chain of batch i+1 and the HBM-bound
 * decode kernels of batch i on the device together, and the host-side work of either call overlaps the other's kernels.
 * While a launch is in flight the descriptor's arrays and rows must stay valid, and no other xsi_encode_* call may be made
 * before xsi_encode_collect.  The blocks a collect returned stay valid until the launch AFTER the next one.               */
int xsi_encode_async(xsi_ctx* ctx, int on);
/* Waits, brings the encoded sections back and assembles the byte-exact GT blocks.
 * n_blocks_out blocks; block b is blocks_out[b] .. +sizes_out[b] (memory owned by ctx, valid
 * until the next xsi_encode_collect on this ctx).  Each is exactly what the reference's
 * GtBlock::write_to_stream writes: [u32 -1][u32 n][dictionary][sections].                   */
int xsi_encode_collect(xsi_ctx* ctx, uint32_t* n_blocks_out, const uint8_t* const** blocks_out,
                       const uint64_t** sizes_out);
/* Per-block payload byte counts of the last launch, available without assembling (what a
 * multi-GPU writer all-gathers to build the global offset table).  Valid after collect.    */
int xsi_encode_block_sizes(xsi_ctx* ctx, uint32_t* n_blocks_out, const uint64_t** sizes_out);
/* max ploidy seen in the last batch (header.ploidy, xsi_factory.hpp:548) */
int xsi_encode_max_ploidy(const xsi_ctx* ctx);
/* Line statistics of the last collected batch: binary lines (KEY_BINARY_LINES summed over the
 * blocks) and how many of them were PBWT+WAH encoded (the rest are sparse).  Valid after collect. */
int xsi_encode_line_counts(const xsi_ctx* ctx, uint64_t* n_binary_lines, uint64_t* n_wah_lines);

/* ------------------------------------------------------------------------------------------
 * DECODE  -- replaces DecompressPointerGTBlock<A_T,uint16_t> (seek + fill_genotype_array_advance,
 *            include/accessor_internals_new.hpp:154-384) behind AccessorInternals::fill_genotype_array
 *            (include/accessor_internals.hpp:399-413).
 * ------------------------------------------------------------------------------------------ */
/* Stage 1: upload the GT blocks (host pointers to the [u32 -1][n][dict].. payloads, i.e. what
 * set_gt_block_ptr yields, accessor_internals_new.hpp:830-843), expand every WAH line and undo
 * the PBWT order on the device.  num_samples / aet_bytes come from header_t
 * (include/compression.hpp:40-104).  Replaces any previously loaded set.                    */
int xsi_decode_load_blocks(xsi_ctx* ctx, uint32_t n_blocks, const uint8_t* const* gt_blocks,
                           const uint64_t* sizes, uint64_t num_samples, int32_t aet_bytes);
/* The same load with a LAZY inverse-PBWT chain: every WAH line is expanded (parallel work), but the sequential chain that
 * puts lines back into sample order stops after the WAH lines among the first `initial_lines` binary lines of each block (0:
 * none at all).  xsi_decode_records* continue the chain on demand up to the last line they are asked for plus a window
 * (XSI_LAZY_WINDOW, default 1024 lines), so a record near the start of a block no longer costs the whole block -- the
 * reference's seek replays every line before the one it wants (accessor_internals_new.hpp:154-196), and so does this, but
 * it never goes further than needed.  xsi_decode_allele_counts needs no chain at all.  Blocks that the continuing kernel
 * does not serve (more than 65,534 haplotypes, all-haploid lines) are loaded in full.                                       */
int xsi_decode_load_blocks_lazy(xsi_ctx* ctx, uint32_t n_blocks, const uint8_t* const* gt_blocks, const uint64_t* sizes,
                                uint64_t num_samples, int32_t aet_bytes, uint32_t initial_lines);
/* Continues the chain of loaded block `block_index` so that its binary lines [0, line_end) are final. */
int xsi_decode_extend(xsi_ctx* ctx, uint32_t block_index, uint32_t line_end);
/* Binary lines [0, *lines_ready) of the block are final (the whole block once its chain is complete). */
int xsi_decode_lines_ready(const xsi_ctx* ctx, uint32_t block_index, uint32_t* lines_ready);
/* Stage 2: materialise records. Record i is the one whose first binary line is line_offset[i]
 * (the low 15 bits of BM) in loaded block block_index[i] (index into the loaded set), with
 * n_alleles[i] alleles.  Row i is written at out + i*out_stride (int32 elements); entries
 * past the returned length are left untouched.  out_on_device: 1 = device pointer.
 * n_filled[i] (host, may be NULL) receives what fill_genotype_array returns
 * (CURRENT_N_HAPS: num_samples for an all-haploid line, else 2*num_samples).
 * allele_counts (host, may be NULL): row i holds n_alleles[i] counts at allele_counts + i*counts_stride
 * (AccessorInternals::get_allele_counts).                                                    */
int xsi_decode_records(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                       const uint32_t* n_alleles, int32_t* out, uint64_t out_stride, int32_t out_on_device,
                       uint32_t* n_filled, uint64_t* allele_counts, uint32_t counts_stride);
/* Same, but row i is the record's raw BCF FORMAT/GT payload (int8: the bytes bcf_update_genotypes +
 * bcf_write1 put into the record, htslib vcf.h:152-158; vector end = 0x81), so the extract side
 * (gt_decompressor_new.hpp:275-320) can splice rows into BCF records without narrowing int32 on the
 * CPU.  out_stride in int8 elements.  XSI_E_UNSUPPORTED when a record has more than 63 alleles
 * (BCF itself switches to int16 there).                                                        */
int xsi_decode_records_i8(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                          const uint32_t* n_alleles, int8_t* out, uint64_t out_stride, int32_t out_on_device,
                          uint32_t* n_filled, uint64_t* allele_counts, uint32_t counts_stride);
/* Counts only, no row is materialised: replaces AccessorInternals::fill_allele_counts
 * (accessor_internals.hpp:404, accessor_internals_new.hpp:407-440,747-752; the AC/AN recompute of af_stats).
 * Row i of allele_counts (host) receives n_alleles[i] counts.  As in the reference, count[0] is
 * CURRENT_N_HAPS - sum(ALT counts): missing and end-of-vector entries are NOT subtracted here
 * (fill_genotype_array / xsi_decode_records does subtract them).                               */
int xsi_decode_allele_counts(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                             const uint32_t* n_alleles, uint64_t* allele_counts, uint32_t counts_stride);
/* Dot products on the ENCODED lines: the compressive-access consumer of the reference (dot_prod/dot_prod.hpp:113-245 over
 * InternalGtAccess, accessor_internals_new.hpp:444-471).  For record i and ALT allele a, out[i*out_stride + a-1] receives the
 * sum of y[sample] over the entries of the record that carry allele a (y: num_samples doubles, host or device; out: host).
 * WAH lines are read as 1-bit-per-genotype rows, sparse lines as their index lists; no genotype row is materialised except
 * for lists of REF carriers, which the reference decompresses too (dot_prod.hpp:405-412).  Accumulation is in double, in a
 * fixed but different order than a sequential loop: results agree with a CPU sum to rounding (tests use rel. 1e-10). */
int xsi_decode_dot_products(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                            const uint32_t* n_alleles, const double* y, int32_t y_on_device, double* out, uint32_t out_stride);
/* InternalGtAccess (include/accessor_internals.hpp:374-397, filled by DecompressPointerGTBlock::get_internal_access,
 * accessor_internals_new.hpp:444-471): for the record at (block_index, line_offset), where each of its n_alleles-1 encoded
 * lines lies inside the GT block the caller loaded (byte_offset from the block start; a WAH line of n_entries 16-bit words, or
 * a sparse list whose first A_T word is its count with the MSB meaning "lists REF carriers"), the default allele of the first
 * line, and -- when `a` is not NULL -- the PBWT arrangement in force at the record's last line (what the reference's live `a`
 * holds when it returns): a[j] = haplotype at position j, the order
 * WAH lines are written in (aet_bytes per entry, 2*num_samples entries, host).  The arrangement needs a block loaded with
 * xsi_decode_load_blocks_lazy whose chain has not yet passed the record (XSI_E_UNSUPPORTED otherwise: load it again). */
typedef struct {
    uint32_t is_sparse;
    uint32_t reserved;
    uint64_t byte_offset;
    uint64_t n_entries;
} xsi_line_access;
int xsi_decode_internal_access(xsi_ctx* ctx, uint32_t block_index, uint32_t line_offset, uint32_t n_alleles, xsi_line_access* lines,
                               int32_t* default_allele, void* a);
/* Sample subset: what the extractor's -s/-S does per record (fill_selected_genotypes,
 * include/gt_decompressor_new.hpp:208-238, sample list from enable_select_samples :324-365).  Row i of `out`
 * holds the entries of samples_to_use[0..n_sel) in that order (n_sel * ploidy values, ploidy 1 for an all-haploid
 * record), n_filled[i] that length, and ac + i*ac_stride the selected carriers of ALT allele 1..n_alleles[i]-1
 * (ac_s; ac may be NULL).  Host or device `out` as for xsi_decode_records; sample indices must be < num_samples. */
int xsi_decode_records_subset(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                              const uint32_t* n_alleles, const uint32_t* samples_to_use, uint32_t n_sel,
                              int32_t* out, uint64_t out_stride, int32_t out_on_device, uint32_t* n_filled,
                              uint32_t* ac, uint32_t ac_stride);
/* KEY_BCF_LINES / KEY_BINARY_LINES of loaded block `block_index` (gt_block.hpp:466-467): what a reader that
 * decodes ahead of its caller needs to bound its window (bindings/accessor_internals_b200.hpp). */
int xsi_decode_block_info(const xsi_ctx* ctx, uint32_t block_index, uint32_t* bcf_lines, uint32_t* binary_lines);
/* Blocks until everything queued on the context stream is done. */
int xsi_sync(xsi_ctx* ctx);

/* ------------------------------------------------------------------------------------------
 * Host rows over PCIe.  bcf_get_genotypes widens the record's int8 FORMAT/GT payload to int32
 * (htslib/vcf.c:4728-4795) and fill_genotype_array returns int32 (accessor_internals.hpp:399-413).
 * With HOST int32 buffers, xsi_encode_launch / xsi_decode_records move the rows across the bus in
 * their BCF int8 encoding and convert on the host (worker pool, beside the DMA) whenever every
 * value has one (at most 63 alleles); otherwise int32 moves as is.  Transport only: results are
 * identical either way.  PINNED host buffers get a second route beside it: whole chunks cross as int32
 * by the DMA engine alone and are converted by a device kernel; each chunk takes whichever route is
 * free.  Environment: XSI_HOST_NARROW=0 disables the int8 transport, XSI_HOST_DMA=0 the second route,
 * XSI_HOST_THREADS sets the pool size.  The two conversions are exported for tests and for callers that stage rows themselves.
 * ------------------------------------------------------------------------------------------ */
/* returns 1 when every value was representable, 0 otherwise (dst then holds garbage) */
int  xsi_host_narrow_i32_i8(const int32_t* src, int8_t* dst, uint64_t n);
/* row r: dst[r*dst_stride + j] = int32 form of src[r*src_stride + j] for j < len[r] */
void xsi_host_widen_i8_i32(const int8_t* src, uint64_t src_stride, int32_t* dst, uint64_t dst_stride,
                           const uint32_t* len, uint64_t n_rows);
uint32_t xsi_host_threads(void);
/* bytes this context has moved host->device / device->host in the int8 transport encoding */
void xsi_transport_stats(const xsi_ctx* ctx, uint64_t* narrowed_h2d_bytes, uint64_t* narrowed_d2h_bytes);

/* ------------------------------------------------------------------------------------------
 * Host container layer (no GPU work by itself): the .xsi file, byte-compatible with
 * XsiFactoryExt (include/xsi_factory.hpp:435-639) and readable like Accessor (include/accessor.hpp).
 * ------------------------------------------------------------------------------------------ */
typedef struct xsi_writer xsi_writer;
/* sample_names: n_samples NUL-terminated strings back to back. zstd_level used when zstd_on. */
int xsi_writer_open(const char* path, uint32_t n_samples, const char* sample_names, uint32_t block_len,
                    uint64_t mac_threshold, int32_t default_phasing, int32_t zstd_on, int32_t zstd_level,
                    xsi_writer** out);
/* Appends finished GT blocks (from xsi_encode_collect) in order; n_records / n_variants are the
 * BCF lines and sum(n_allele-1) they cover (xsi_factory.hpp:518-519).                          */
int xsi_writer_add_blocks(xsi_writer* w, uint32_t n_blocks, const uint8_t* const* blocks, const uint64_t* sizes,
                          uint64_t n_records, uint64_t n_variants);
/* finalize_file (xsi_factory.hpp:543-606): index, sample names, header rewrite. */
int xsi_writer_close(xsi_writer* w, int32_t max_ploidy);

/* One file written by several ranks (blocks shard across GPUs): after every rank has put its blocks at the
 * offsets of the all-gathered table, ONE rank adds the index, the sample names and the header (the code of xsi_writer_close).
 * indices[b]: absolute file offset of block b; end_of_blocks: first byte after the last block. */
int xsi_writer_finalize_sharded(const char* path, uint32_t n_samples, const char* sample_names, uint32_t block_len,
                                uint64_t mac_threshold, int32_t default_phasing, int32_t max_ploidy, uint32_t n_blocks,
                                const uint64_t* indices, uint64_t end_of_blocks, uint64_t n_records, uint64_t n_variants);

typedef struct xsi_reader xsi_reader;
int  xsi_reader_open(const char* path, xsi_reader** out); /* mmap + header checks (accessor.cpp:26-82) */
void xsi_reader_close(xsi_reader* r);
int  xsi_reader_info(const xsi_reader* r, uint64_t* num_samples, uint64_t* hap_samples, uint32_t* ploidy,
                     uint32_t* aet_bytes, uint32_t* n_blocks, uint32_t* block_len, uint64_t* xcf_entries,
                     uint64_t* num_variants, int32_t* zstd, uint64_t* rare_threshold, int32_t* default_phased);
const char* xsi_reader_sample_name(const xsi_reader* r, uint64_t i);
/* GT block payload of block b (inflated into reader-owned memory when the file is zstd'ed). */
int xsi_reader_gt_block(xsi_reader* r, uint32_t b, const uint8_t** ptr, uint64_t* size);

#ifdef __cplusplus
}
#endif
#endif /* XSI_B200_H */
