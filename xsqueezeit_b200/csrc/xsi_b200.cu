// xsi_b200.cu -- CUDA context behind include/xsi_b200.h: buffer pools, kernel launches, and the
// host-side assembly of byte-exact GT blocks.  No CPU fallback lives here: every genotype
// operation is a kernel from encode_kernels.cuh / decode_kernels.cuh.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/xsi_b200.h"
#include "decode_kernels.cuh"
#include "encode_kernels.cuh"
#include "host_narrow.hpp"
#include "host_util.hpp"

using namespace xsi;

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};
struct PinBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

template <typename T>
size_t vec_bytes(const std::vector<T>& v) { return v.size() * sizeof(T); }

}  // namespace

struct xsi_ctx {
    int device = 0;
    cudaStream_t stream = nullptr, stream2 = nullptr;
    // encode stream: the context stream, or (xsi_encode_async) a stream of its own so that an encode running on the
    // library's worker thread and a decode issued by the caller overlap on the device
    cudaStream_t es = nullptr, stream_enc = nullptr;
    cudaEvent_t ev_side = nullptr;
    std::string err = "";
    std::atomic<uint64_t> launches{0};
    // asynchronous encode (xsi_encode_async): xsi_encode_launch hands the batch to this thread, xsi_encode_collect joins it
    bool async_encode = false;
    std::thread enc_thread;
    int enc_rc = XSI_OK;
    bool enc_pending = false;  // a launch was handed to enc_thread and its result not yet reported by xsi_encode_collect
    xsi_encode_desc enc_desc;
    int prio_hi = 0;          // greatest stream / launch priority of the device
    bool perm_priority = false;  // launch the PBWT cluster kernel at prio_hi (XSI_PERMUTE_PRIORITY=1, or the ordered overlap below)
    // Ordered overlap inside one context (xsi_encode_async; XSI_OVERLAP_ORDER=0 turns it off): the decode of batch i and the
    // encode of batch i+1 are on the device together.  The PBWT cluster kernel of the encode owns 128 whole SMs for most of
    // the batch's time and leaves 20 idle, while the HBM-bound compose kernels of the decode need few SMs to move their bytes:
    // so the compose kernels wait for the encode's scan (both want every SM and all of HBM) and start WITH the cluster kernel,
    // which is launched at the greatest priority and takes its SMs first; the compose CTAs fill what is left.
    bool overlap_order = true;
    cudaEvent_t ev_scan = nullptr;                 // recorded on the encode stream right after the scan kernel of an asynchronous launch
    std::atomic<uint64_t> enc_seq{0}, scan_seq{0};  // launches handed to the encode thread / launches whose scan is enqueued (or that ended)
    uint64_t enc_row_stride = 0;   // xsi_encode_launch_strided: element distance between rows of the launch being set up (0: back to back)
    std::mutex prof_m;
    int sm_count = 148;
    size_t smem_optin = 0;
    // optional per-kernel timing (CUDA events on the launching stream)
    bool profile = false;
    struct Span { const char* name; cudaEvent_t a, b; };
    std::vector<Span> spans;
    std::string profile_text;
    std::map<std::string, std::pair<int, double>> host_spans;  // host-side phases (wall clock), reported as "host:<name>"

    // pinned ring for host int32 rows that cross PCIe as int8 (host_narrow.cpp): slots of RING_BYTES
    static constexpr int RING_SLOTS = 3;
    static constexpr size_t RING_BYTES = 16u << 20;
    PinBuf ring;
    cudaEvent_t ring_ev[RING_SLOTS] = {nullptr, nullptr, nullptr};
    bool ring_busy[RING_SLOTS] = {false, false, false};
    uint64_t narrowed_h2d = 0, narrowed_d2h = 0;  // bytes that crossed the bus narrowed (statistics)
    // second route for PINNED host int32 rows: plain DMA of int32 chunks (no host core involved), converted by a
    // device kernel, running beside the host conversion of other chunks (whichever route is free takes the next chunk)
    static constexpr int DMA_SLOTS = 8;                                 // staging slots of the DMA route (XSI_DMA_SLOTS uses fewer)
    static constexpr size_t DMA_ELEMS = RING_BYTES / 4;                 // genotypes per slot: copies as large as the int8 ring's (16 MB)
    DevBuf dma_stage, dma_flag;
    cudaEvent_t dma_ev[DMA_SLOTS] = {};
    bool dma_busy[DMA_SLOTS] = {};
    PinBuf dma_flag_host;
    uint64_t dma_h2d = 0, dma_d2h = 0;  // int32 bytes moved by that route (statistics)

    // ---------------- encode ----------------
    struct {
        bool launched = false;
        uint64_t R = 0, L = 0;
        uint32_t nb = 0, n_samples = 0, block_len = 0, WS = 0, SLOTW = 0, aet = 2;
        int32_t default_phasing = 0;
        int max_ploidy = 0;
        uint32_t aux_cap = 0, phase_cap = 0;
        std::vector<uint32_t> h_nallele, h_ngt, h_line0, h_line_rec, h_blk_line0, h_blk_rec0;
        std::vector<uint64_t> h_goff;
        DevBuf gt, tables, bitrows, auxrows, phrows, counters, rec_aux, line_u32, line_flags, rec_u32, rec_flags,
            wah_list, blk_nwah, wahslots, phslots, offs, blkoffs, scanjobs, scansums, out_wah, out_sparse, out_miss, out_eov, out_phase,
            auxslots, out_missw, out_eovw,
            a_pool;
        PinBuf h_small, h_offs, h_out, h_flags;
        uint64_t tot_sparse = 0, tot_miss = 0, tot_eov = 0, tot_wah = 0, tot_phase = 0, tot_missw = 0, tot_eovw = 0;
        bool wah_missing = false;  // --wah-encode-missing (WS_WAH)
        // finished GT blocks, back to back (16-byte aligned starts).  Two arenas alternate from launch to launch: the blocks
        // a collect returned stay valid while the NEXT launch runs (a caller that decodes batch i while batch i+1 encodes)
        PinBuf arena[2];
        int gen = 0;                        // arena / result set of the last launch
        std::vector<uint64_t> block_at;     // offset of every block in the arena
        std::vector<const uint8_t*> block_ptrs[2];
        std::vector<uint64_t> block_sizes[2];
        bool collected = false;
        bool any_haploid = false;
        std::vector<uint8_t> h_blk_hap;    // [nb] block holds an all-haploid record
        std::vector<uint32_t> h_blk_map;   // diploid-only blocks, then the others (PBWT launch lists)
        DevBuf blk_map;
        uint64_t n_wah_lines = 0;
    } enc;

    // ---------------- decode ----------------
    struct {
        bool loaded = false;
        uint32_t nb = 0, n_samples = 0, aet = 2, WS = 0, NJ = 0, n_gt_jobs = 0;
        uint64_t Lt = 0;
        std::vector<DecBlock> h_blocks;
        std::vector<uint32_t> h_bin_lines, h_bcf_lines;
        // lazy inverse-PBWT chain (xsi_decode_load_blocks_lazy): WAH lines of block b in order, how many of them are back in
        // sample order, and the launch shape of the chain kernel so that a continuation matches the first launch
        std::vector<std::vector<uint16_t>> h_wah_lines;
        std::vector<uint32_t> h_wah_done;
        bool lazy_ok = false;
        uint32_t v3_kh = 0, v3_nc = 0, v3_slices = 0, ps_stride = 0;
        size_t v3_smem = 0, m_blk = 0;
        DevBuf pos_state;
        // host views of the tables of the loaded set (inside h_meta, valid until the next load): for xsi_decode_internal_access
        const DecBlock* hv_blocks = nullptr; const DecSeg* hv_segs = nullptr; const uint32_t* hv_job_seg = nullptr;
        const uint32_t* hv_dl_ord = nullptr; const uint8_t* hv_dl_flags = nullptr;
        std::vector<uint64_t> h_blob_off, h_blk_size;
        DevBuf blob, meta, rows, job_u32, job_hap, tile_u32, dline, lists, err, a_pool, x_pool, req, out, scratch, counts,
            seg_total, tabs;
        PinBuf h_stage, h_meta, h_req;
        cudaEvent_t ev_req = nullptr;  // the last copy out of h_req
        DecDev dev;
    } dec;
};

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                        \
            return XSI_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)
// PROF(name) { launch; }  brackets a launch with events when profiling is on
struct ProfScope {
    xsi_ctx* c; bool on; cudaEvent_t a = nullptr, b = nullptr; const char* name; cudaStream_t st;
    ProfScope(xsi_ctx* ctx, const char* n, cudaStream_t s) : c(ctx), on(ctx->profile), name(n), st(s) {
        if (on) { cudaEventCreate(&a); cudaEventCreate(&b); cudaEventRecord(a, st); }
    }
    ~ProfScope() { if (on) { cudaEventRecord(b, st); std::lock_guard<std::mutex> g(c->prof_m); c->spans.push_back({name, a, b}); } }
};
// XSI_STREAM names the stream of the code being compiled: the context stream for decode, ctx->es inside the encode path
#define XSI_STREAM ctx->stream
#define PROF(name) ProfScope prof_scope__(ctx, name, XSI_STREAM)
// HOSTSPAN("host:parse");  wall-clock time of a host-side phase, to the end of the enclosing scope (profiling only)
struct HostSpan {
    xsi_ctx* c; const char* name; std::chrono::steady_clock::time_point t0;
    HostSpan(xsi_ctx* ctx, const char* n) : c(ctx), name(n), t0(std::chrono::steady_clock::now()) {}
    ~HostSpan() {
        if (!c->profile) return;
        std::lock_guard<std::mutex> g(c->prof_m);
        auto& s = c->host_spans[name];
        s.first++;
        s.second += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
};
#define HOSTSPAN_CAT2(a, b) a##b
#define HOSTSPAN_CAT(a, b) HOSTSPAN_CAT2(a, b)
#define HOSTSPAN(name) HostSpan HOSTSPAN_CAT(host_span__, __LINE__)(ctx, name)
// XSI_DEBUG_SYNC=1 waits for every kernel right after its launch, so that a device fault names its launch site
static const bool g_debug_sync = getenv("XSI_DEBUG_SYNC") != nullptr;
#define CKL()                                                                                      \
    do {                                                                                           \
        ctx->launches++;                                                                           \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ == cudaSuccess && g_debug_sync) e__ = cudaStreamSynchronize(XSI_STREAM);           \
        if (e__ != cudaSuccess) {                                                                  \
            ctx->err = std::string("kernel launch (xsi_b200.cu:") + std::to_string(__LINE__) + "): " + cudaGetErrorString(e__); \
            return XSI_E_CUDA;                                                                     \
        }                                                                                          \
    } while (0)

extern "C" const char* xsi_version(void) { return "xsi-b200 0.1 (sm_100a)"; }

extern "C" int xsi_create(int device, xsi_ctx** out) {
    if (!out) return XSI_E_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return XSI_E_CUDA;
    xsi_ctx* ctx = new xsi_ctx();
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return XSI_E_CUDA; }
    {
        int lo = 0, hi = 0;
        if (cudaDeviceGetStreamPriorityRange(&lo, &hi) == cudaSuccess) ctx->prio_hi = hi;
        if (const char* s_ = getenv("XSI_PERMUTE_PRIORITY")) ctx->perm_priority = atoi(s_) != 0;
        if (const char* s_ = getenv("XSI_OVERLAP_ORDER")) ctx->overlap_order = atoi(s_) != 0;
    }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->stream2, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming) != cudaSuccess) {
        delete ctx;
        return XSI_E_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) {
        ctx->sm_count = prop.multiProcessorCount;
        ctx->smem_optin = prop.sharedMemPerBlockOptin;
    }
    ctx->es = ctx->stream;
    *out = ctx;
    return XSI_OK;
}

extern "C" void xsi_destroy(xsi_ctx* ctx) {
    if (!ctx) return;
    if (ctx->enc_thread.joinable()) ctx->enc_thread.join();
    cudaSetDevice(ctx->device);
    if (ctx->stream_enc) { cudaStreamSynchronize(ctx->stream_enc); cudaStreamDestroy(ctx->stream_enc); }
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->stream2);
    auto& e = ctx->enc;
    for (DevBuf* b : {&e.gt, &e.tables, &e.bitrows, &e.auxrows, &e.phrows, &e.counters, &e.rec_aux, &e.line_u32,
                      &e.line_flags, &e.rec_u32, &e.rec_flags, &e.wah_list, &e.blk_nwah, &e.wahslots, &e.phslots,
                      &e.offs, &e.scanjobs, &e.out_wah, &e.out_sparse, &e.out_miss, &e.out_eov, &e.out_phase, &e.a_pool,
                      &e.auxslots, &e.out_missw, &e.out_eovw, &e.blk_map, &e.scansums, &e.blkoffs})
        b->release();
    for (PinBuf* b : {&e.h_small, &e.h_offs, &e.h_out, &e.h_flags, &e.arena[0], &e.arena[1]}) b->release();
    auto& d = ctx->dec;
    for (DevBuf* b : {&d.blob, &d.meta, &d.rows, &d.job_u32, &d.job_hap, &d.tile_u32, &d.dline, &d.lists, &d.err,
                      &d.a_pool, &d.x_pool, &d.req, &d.out, &d.scratch, &d.counts, &d.seg_total, &d.tabs, &d.pos_state})
        b->release();
    d.h_stage.release();
    d.h_meta.release();
    d.h_req.release();
    if (d.ev_req) cudaEventDestroy(d.ev_req);
    ctx->ring.release();
    ctx->dma_stage.release(); ctx->dma_flag.release(); ctx->dma_flag_host.release();
    for (cudaEvent_t ev : ctx->ring_ev) if (ev) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->dma_ev) if (ev) cudaEventDestroy(ev);
    cudaEventDestroy(ctx->ev_side);
    if (ctx->ev_scan) cudaEventDestroy(ctx->ev_scan);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->stream2);
    delete ctx;
}

// pinned host buffers for bindings that stage rows themselves (bindings/gt_block_b200.hpp, accessor_internals_b200.hpp)
extern "C" int xsi_host_alloc(void** p, uint64_t bytes) {
    if (!p) return XSI_E_ARG;
    *p = nullptr;
    if (bytes == 0) return XSI_OK;
    const cudaError_t e = cudaHostAlloc(p, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; return e == cudaErrorMemoryAllocation ? XSI_E_NOMEM : XSI_E_CUDA; }
    return XSI_OK;
}
extern "C" void xsi_host_free(void* p) { if (p) cudaFreeHost(p); }
// device rows for C / C++ callers that keep a decode -> encode hand-over on the device (bindings/xsi_b200_bcf.cpp `subset`)
extern "C" int xsi_device_alloc(xsi_ctx* ctx, void** p, uint64_t bytes) {
    if (!ctx || !p) return XSI_E_ARG;
    *p = nullptr;
    if (bytes == 0) return XSI_OK;
    if (cudaSetDevice(ctx->device) != cudaSuccess) { cudaGetLastError(); return XSI_E_CUDA; }
    const cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); *p = nullptr; ctx->err = "device allocation failed"; return e == cudaErrorMemoryAllocation ? XSI_E_NOMEM : XSI_E_CUDA; }
    return XSI_OK;
}
extern "C" void xsi_device_free(xsi_ctx* ctx, void* p) {
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFree(p);
}

extern "C" const char* xsi_last_error(const xsi_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
extern "C" void* xsi_stream(xsi_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t xsi_kernel_launches(const xsi_ctx* ctx) { return ctx ? ctx->launches.load() : (uint64_t)0; }
extern "C" int xsi_profile(xsi_ctx* ctx, int on) {
    if (!ctx) return XSI_E_ARG;
    ctx->profile = on != 0;
    return XSI_OK;
}
// "name count total_ms\n" per kernel since the last read; clears the record
extern "C" const char* xsi_profile_read(xsi_ctx* ctx) {
    if (!ctx) return "";
    if (ctx->enc_thread.joinable()) ctx->enc_thread.join();  // (its result is still reported by xsi_encode_collect)
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->es != ctx->stream) cudaStreamSynchronize(ctx->es);
    std::lock_guard<std::mutex> g(ctx->prof_m);
    std::map<std::string, std::pair<int, double>> agg;
    std::vector<std::string> order;
    // XSI_TIMELINE=<file>: start / end of every profiled kernel relative to the first one of this read, on whichever stream it ran
    // (how the encode of one batch and the decode of another actually interleave on the device)
    FILE* tl = nullptr;
    if (const char* tf = getenv("XSI_TIMELINE")) if (!ctx->spans.empty()) tl = fopen(tf, "a");
    if (tl) fprintf(tl, "# read\n");
    const cudaEvent_t first = ctx->spans.empty() ? nullptr : ctx->spans.front().a;  // destroyed after the loop: every span is measured against it
    for (auto& sp : ctx->spans) {
        float ms = 0;
        cudaEventElapsedTime(&ms, sp.a, sp.b);
        if (tl) {
            float t0 = 0;
            cudaEventElapsedTime(&t0, first, sp.a);
            fprintf(tl, "%-18s %10.3f %10.3f\n", sp.name, t0, t0 + ms);
        }
        if (!agg.count(sp.name)) order.push_back(sp.name);
        agg[sp.name].first++;
        agg[sp.name].second += ms;
        if (sp.a != first) cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    if (tl) fclose(tl);
    if (first) cudaEventDestroy(first);
    ctx->spans.clear();
    ctx->profile_text.clear();
    char line[256];
    for (auto& n : order) {
        snprintf(line, sizeof line, "%s %d %.6f\n", n.c_str(), agg[n].first, agg[n].second);
        ctx->profile_text += line;
    }
    for (auto& kv : ctx->host_spans) {
        snprintf(line, sizeof line, "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        ctx->profile_text += line;
    }
    ctx->host_spans.clear();
    return ctx->profile_text.c_str();
}
extern "C" int xsi_sync(xsi_ctx* ctx) {
    if (!ctx) return XSI_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return XSI_OK;
}

extern "C" int xsi_host_narrow_i32_i8(const int32_t* src, int8_t* dst, uint64_t n) {
    return (src && dst) ? (narrow_i32_to_i8(src, dst, n) ? 1 : 0) : 0;
}
extern "C" void xsi_host_widen_i8_i32(const int8_t* src, uint64_t src_stride, int32_t* dst, uint64_t dst_stride,
                                      const uint32_t* len, uint64_t n_rows) {
    if (src && dst && len) widen_rows_i8_to_i32(src, src_stride, dst, dst_stride, len, n_rows);
}
extern "C" uint32_t xsi_host_threads(void) { return host_threads(); }
extern "C" void xsi_transport_stats(const xsi_ctx* ctx, uint64_t* h2d, uint64_t* d2h) {
    if (h2d) *h2d = ctx ? ctx->narrowed_h2d : 0;
    if (d2h) *d2h = ctx ? ctx->narrowed_d2h : 0;
}

// =================================================================================================
// Host rows over PCIe as int8 (transport only, see host_narrow.cpp).  XSI_HOST_NARROW=0 moves int32.
// =================================================================================================
namespace {

bool host_narrow_on() {
    const char* s = getenv("XSI_HOST_NARROW");
    return !(s && s[0] == '0');
}

int ring_prepare(xsi_ctx* ctx) {
    CK(ctx->ring.ensure(xsi_ctx::RING_SLOTS * xsi_ctx::RING_BYTES));
    for (int i = 0; i < xsi_ctx::RING_SLOTS; ++i)
        if (!ctx->ring_ev[i]) CK(cudaEventCreateWithFlags(&ctx->ring_ev[i], cudaEventDisableTiming));
    return XSI_OK;
}
int ring_wait(xsi_ctx* ctx, int slot) {
    if (ctx->ring_busy[slot]) { CK(cudaEventSynchronize(ctx->ring_ev[slot])); ctx->ring_busy[slot] = false; }
    return XSI_OK;
}

bool host_ptr_is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}
bool dma_route_on() {
    const char* s = getenv("XSI_HOST_DMA");
    return !(s && s[0] == '0');
}
int dma_prepare(xsi_ctx* ctx) {
    CK(ctx->dma_stage.ensure((size_t)xsi_ctx::DMA_SLOTS * xsi_ctx::DMA_ELEMS * 4));
    CK(ctx->dma_flag.ensure(16));
    CK(ctx->dma_flag_host.ensure(16));
    for (int i = 0; i < xsi_ctx::DMA_SLOTS; ++i)
        if (!ctx->dma_ev[i]) CK(cudaEventCreateWithFlags(&ctx->dma_ev[i], cudaEventDisableTiming));
    return XSI_OK;
}
// index of a free staging slot of the DMA route, or -1
int dma_free_slot(xsi_ctx* ctx) {
    static const int n_slots = [] {
        int v = 4;
        if (const char* s = getenv("XSI_DMA_SLOTS")) v = atoi(s);
        return std::max(1, std::min(v, (int)xsi_ctx::DMA_SLOTS));
    }();
    for (int i = 0; i < n_slots; ++i) {
        if (ctx->dma_busy[i] && cudaEventQuery(ctx->dma_ev[i]) == cudaSuccess) ctx->dma_busy[i] = false;
        if (!ctx->dma_busy[i]) return i;
    }
    cudaGetLastError();  // cudaErrorNotReady is not an error
    return -1;
}

// Uploads n host int32 genotypes as int8 into dst (device).  Chunks of 16 Mi genotypes take one of two routes,
// whichever is free: (a) narrowed by the worker pool into a pinned ring slot and copied as int8, (b) when the
// source is pinned, copied as int32 by the DMA engine alone into a device staging slot and narrowed by a kernel
// (side stream).  Returns 1 when done, 0 when a value has no int8 encoding (nothing usable was uploaded; the
// caller moves int32 instead), < 0 on error.
int upload_narrowed(xsi_ctx* ctx, const int32_t* src, size_t n, int8_t* dst) {
    int rc = ring_prepare(ctx);
    if (rc) return rc;
    const size_t CH = xsi_ctx::RING_BYTES;
    const bool dma = dma_route_on() && n >= 4 * CH && host_ptr_is_pinned(src) && reinterpret_cast<uintptr_t>(src) % 16 == 0 &&
                     reinterpret_cast<uintptr_t>(dst) % 16 == 0;
    if (dma) {
        if ((rc = dma_prepare(ctx))) return rc;
        CK(cudaMemsetAsync(ctx->dma_flag.p, 0, 4, ctx->stream2));
    }
    int slot = 0;
    bool used_dma = false;
    size_t by_host = 0;
    const size_t DE = xsi_ctx::DMA_ELEMS;
    for (size_t a = 0; a < n;) {
        // top up the DMA route first (it needs no host core): units as large in bytes as the ring's int8 copies, so
        // that the two routes share the bus evenly and the host route never waits long for its next chunk
        while (dma && n - a >= DE) {
            const int ds = dma_free_slot(ctx);
            if (ds < 0) break;
            int32_t* st = ctx->dma_stage.as<int32_t>() + (size_t)ds * DE;
            CK(cudaMemcpyAsync(st, src + a, DE * 4, cudaMemcpyHostToDevice, ctx->stream2));
            narrow_i32_i8_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream2>>>(reinterpret_cast<const uint4*>(st), reinterpret_cast<uint4*>(dst + a),
                                                                              DE / 16, ctx->dma_flag.as<uint32_t>());
            CKL();
            CK(cudaEventRecord(ctx->dma_ev[ds], ctx->stream2));
            ctx->dma_busy[ds] = true;
            ctx->dma_h2d += DE * 4;
            used_dma = true;
            a += DE;
        }
        if (a >= n) break;
        const size_t m = std::min(CH, n - a);
        if ((rc = ring_wait(ctx, slot))) return rc;
        int8_t* stage = ctx->ring.as<int8_t>() + (size_t)slot * CH;
        if (!narrow_i32_to_i8(src + a, stage, m)) { if (used_dma) cudaStreamSynchronize(ctx->stream2); return 0; }
        CK(cudaMemcpyAsync(dst + a, stage, m, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaEventRecord(ctx->ring_ev[slot], ctx->stream));
        ctx->ring_busy[slot] = true;
        slot = (slot + 1) % xsi_ctx::RING_SLOTS;
        by_host += m;
        a += m;
    }
    if (used_dma) {
        // join the side stream: its chunks are part of the rows the kernels on the main stream will read
        uint32_t* hf = ctx->dma_flag_host.as<uint32_t>();
        CK(cudaMemcpyAsync(hf, ctx->dma_flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream2));
        CK(cudaStreamSynchronize(ctx->stream2));
        for (int i = 0; i < xsi_ctx::DMA_SLOTS; ++i) ctx->dma_busy[i] = false;
        if (*hf) return 0;
    }
    ctx->narrowed_h2d += by_host;
    return 1;
}

}  // namespace

// =================================================================================================
// ENCODE
// =================================================================================================
namespace {

// words-per-warp of the sequential PBWT kernels: smallest power of two that covers the row with
// <= 32 warps; XSI_PBWT_WPW overrides it (tuning).  WPW 128 runs 512 threads with 128 registers.
#undef XSI_STREAM
#define XSI_STREAM ctx->es
int choose_wpw(uint32_t W) {
    int wpw = 2;
    while ((W + wpw - 1) / wpw > 32) wpw *= 2;
    if (const char* s = getenv("XSI_PBWT_WPW")) {
        const int v = atoi(s);
        if ((v == 2 || v == 4 || v == 8 || v == 16 || v == 32 || v == 64 || v == 128) && (W + v - 1) / v <= (v == 128 ? 16u : 32u)) wpw = v;
    }
    return wpw;
}

template <int WPW, int MAXT>
int launch_permute(xsi_ctx* ctx, const EncDev& p, uint32_t NW, size_t smem) {
    CK(cudaFuncSetAttribute(pbwt_permute_smem_kernel<WPW, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { PROF("pbwt_permute"); pbwt_permute_smem_kernel<WPW, MAXT><<<p.nb, NW * 32, smem, ctx->es>>>(p); }
    CKL();
    return XSI_OK;
}

template <int C, int KH>
int launch_permute_v4(xsi_ctx* ctx, const EncDev& p, const PermV4Cfg& cfg, uint32_t NT, size_t smem, bool probe_only, int* max_clusters) {
    CK(cudaFuncSetAttribute(pbwt_permute_v4_kernel<C, KH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(p.nb * C); lc.blockDim = dim3(NT); lc.dynamicSmemBytes = smem; lc.stream = ctx->es;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = C > 1 ? 1 : 0;
    if (probe_only) {
        *max_clusters = 1 << 30;
        if (C > 1) CK(cudaOccupancyMaxActiveClusters(max_clusters, pbwt_permute_v4_kernel<C, KH>, &lc));
        return XSI_OK;
    }
    // The chain kernel owns whole SMs (1024 threads, all registers) and is the critical path of a batch: when a decode of
    // the previous batch runs beside it (xsi_encode_async, several contexts), its clusters should get the SMs that free up
    // before the HBM-bound kernels' CTAs do, which then fill the SMs the clusters cannot use (148 - 4 * 32 = 20).
    if ((ctx->perm_priority || (ctx->async_encode && ctx->overlap_order)) && C > 1) {
        at[1].id = cudaLaunchAttributePriority;
        at[1].val.priority = ctx->prio_hi;
        lc.numAttrs = 2;
    }
    { PROF("pbwt_permute"); CK(cudaLaunchKernelEx(&lc, pbwt_permute_v4_kernel<C, KH>, p, cfg)); }
    CKL();
    return XSI_OK;
}

template <int C>
int launch_permute_v4_kh(xsi_ctx* ctx, const EncDev& p, const PermV4Cfg& cfg, uint32_t KH, uint32_t NT, size_t smem, bool probe_only, int* maxc) {
    switch (KH) {
        case 8: return launch_permute_v4<C, 8>(ctx, p, cfg, NT, smem, probe_only, maxc);
        case 16: return launch_permute_v4<C, 16>(ctx, p, cfg, NT, smem, probe_only, maxc);
        case 32: return launch_permute_v4<C, 32>(ctx, p, cfg, NT, smem, probe_only, maxc);
        default: return launch_permute_v4<C, 64>(ctx, p, cfg, NT, smem, probe_only, maxc);
    }
}

int launch_permute_v4_c(xsi_ctx* ctx, const EncDev& p, uint32_t C, const PermV4Cfg& cfg, uint32_t KH, uint32_t NT, size_t smem, bool probe_only, int* maxc) {
    switch (C) {
        case 8: return launch_permute_v4_kh<8>(ctx, p, cfg, KH, NT, smem, probe_only, maxc);
        case 4: return launch_permute_v4_kh<4>(ctx, p, cfg, KH, NT, smem, probe_only, maxc);
        case 2: return launch_permute_v4_kh<2>(ctx, p, cfg, KH, NT, smem, probe_only, maxc);
        default: return launch_permute_v4_kh<1>(ctx, p, cfg, KH, NT, smem, probe_only, maxc);
    }
}

// Picks the cluster size C (the largest whose clusters are all co-resident, so that a batch with fewer
// blocks than SMs still fills the GPU) and the haplotypes per thread KH; XSI_PBWT_CLUSTER / XSI_PBWT_KH override.
int run_permute_v4(xsi_ctx* ctx, const EncDev& p, uint32_t W, bool* done) {
    *done = false;
    int forced = 0, forced_kh = 0;
    if (const char* s = getenv("XSI_PBWT_CLUSTER")) forced = atoi(s);
    if (const char* s = getenv("XSI_PBWT_KH")) forced_kh = atoi(s);
    for (uint32_t C : {8u, 4u, 2u, 1u}) {
        if (forced && (uint32_t)forced != C) continue;
        PermV4Cfg cfg;
        uint32_t per = (W + C - 1) / C, wsl = 32, sh = 10;
        while (wsl < per) { wsl *= 2; ++sh; }
        cfg.WSL = wsl; cfg.SH = sh;
        const uint32_t HS = wsl * 32;
        if ((uint64_t)C * HS > 65536) continue;
        if (!forced && C > 1 && wsl * (C / 2) >= W) continue;  // half the cluster would already cover the row
        // one CTA per SM: two 512-thread CTAs of a C=8 cluster sharing an SM lose to C=4 (20.6 vs 17.3 ms at 32 HRC blocks)
        if (!forced && C > 1 && (uint64_t)p.nb * C > (uint64_t)ctx->sm_count) continue;
        uint32_t KH = std::max<uint32_t>(8, HS / 1024);
        if (HS / KH > 512 && C == 8 && KH < 64) KH *= 2;  // 512-thread CTAs: two per SM
        // one CTA per block and more blocks than SMs (1KGP3 shape: 220 blocks of 5,008 haplotypes): 512-thread CTAs, two per
        // SM, so that the whole batch is resident and the chains cover each other's barriers (15.0 -> 12.4 ms, r02g)
        if (C == 1 && KH == 8 && HS / KH > 512 && p.nb > (uint32_t)ctx->sm_count) KH = 16;
        if (forced_kh == 8 || forced_kh == 16 || forced_kh == 32 || forced_kh == 64) KH = (uint32_t)forced_kh;
        const uint32_t NT = HS / KH;
        if (NT > 1024 || NT < 32 || NT < wsl / (KH == 64 ? 2 : 1)) continue;
        const size_t WT = (size_t)C * wsl;
        const size_t smem = (4 * WT + 8 + 32) * 4 + 16;
        if (smem > ctx->smem_optin) continue;
        int maxc = 0;
        int rc = launch_permute_v4_c(ctx, p, C, cfg, KH, NT, smem, true, &maxc);
        if (rc) return rc;
        if (!forced && C > 1 && (uint32_t)maxc < p.nb) continue;  // the clusters would not all be resident at once
        rc = launch_permute_v4_c(ctx, p, C, cfg, KH, NT, smem, false, &maxc);
        *done = rc == XSI_OK;
        return rc;
    }
    return XSI_OK;
}

// ---- v5: two WAH lines per exchange round ----
template <int C, int KH>
int launch_permute_v5(xsi_ctx* ctx, const EncDev& p, const PermV4Cfg& cfg, uint32_t NT, size_t smem, bool probe_only, int* max_clusters) {
    CK(cudaFuncSetAttribute(pbwt_permute_v5_kernel<C, KH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(p.nb * C); lc.blockDim = dim3(NT); lc.dynamicSmemBytes = smem; lc.stream = ctx->es;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = C; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    lc.attrs = at; lc.numAttrs = C > 1 ? 1 : 0;
    if (probe_only) {
        *max_clusters = 1 << 30;
        if (C > 1) CK(cudaOccupancyMaxActiveClusters(max_clusters, pbwt_permute_v5_kernel<C, KH>, &lc));
        return XSI_OK;
    }
    { PROF("pbwt_permute"); CK(cudaLaunchKernelEx(&lc, pbwt_permute_v5_kernel<C, KH>, p, cfg)); }
    CKL();
    return XSI_OK;
}
template <int C>
int launch_permute_v5_kh(xsi_ctx* ctx, const EncDev& p, const PermV4Cfg& cfg, uint32_t KH, uint32_t NT, size_t smem, bool probe_only, int* maxc) {
    switch (KH) {
        case 8: return launch_permute_v5<C, 8>(ctx, p, cfg, NT, smem, probe_only, maxc);
        case 16: return launch_permute_v5<C, 16>(ctx, p, cfg, NT, smem, probe_only, maxc);
        default: return launch_permute_v5<C, 32>(ctx, p, cfg, NT, smem, probe_only, maxc);
    }
}
int launch_permute_v5_c(xsi_ctx* ctx, const EncDev& p, uint32_t C, const PermV4Cfg& cfg, uint32_t KH, uint32_t NT, size_t smem, bool probe_only, int* maxc) {
    switch (C) {
        case 8: return launch_permute_v5_kh<8>(ctx, p, cfg, KH, NT, smem, probe_only, maxc);
        case 4: return launch_permute_v5_kh<4>(ctx, p, cfg, KH, NT, smem, probe_only, maxc);
        case 2: return launch_permute_v5_kh<2>(ctx, p, cfg, KH, NT, smem, probe_only, maxc);
        default: return launch_permute_v5_kh<1>(ctx, p, cfg, KH, NT, smem, probe_only, maxc);
    }
}
// Same cluster-size choice as v4 (the largest C whose clusters are all co-resident); additionally a slice may hold at most
// 32768 positions (16-bit class counts) and the CTA needs one thread per row word of its slice (KH <= 32).
int run_permute_v5(xsi_ctx* ctx, const EncDev& p, uint32_t W, bool* done) {
    *done = false;
    int forced = 0, forced_kh = 0;
    if (const char* s = getenv("XSI_PBWT_CLUSTER")) forced = atoi(s);
    if (const char* s = getenv("XSI_PBWT_KH")) forced_kh = atoi(s);
    for (uint32_t C : {8u, 4u, 2u, 1u}) {
        if (forced && (uint32_t)forced != C) continue;
        PermV4Cfg cfg;
        uint32_t per = (W + C - 1) / C, wsl = 32, sh = 10;
        while (wsl < per) { wsl *= 2; ++sh; }
        cfg.WSL = wsl; cfg.SH = sh;
        const uint32_t HS = wsl * 32;
        if ((uint64_t)C * HS > 65536 || HS > 32768) continue;
        if (!forced && C > 1 && wsl * (C / 2) >= W) continue;
        if (!forced && C > 1 && (uint64_t)p.nb * C > (uint64_t)ctx->sm_count) continue;
        uint32_t KH = std::max<uint32_t>(8, HS / 1024);
        if (C == 1 && KH == 8 && HS / KH > 512 && p.nb > (uint32_t)ctx->sm_count) KH = 16;
        if (forced_kh == 8 || forced_kh == 16 || forced_kh == 32) KH = (uint32_t)forced_kh;
        if (KH > 32) continue;
        const uint32_t NT = HS / KH;
        if (NT > 1024 || NT < 32 || NT < wsl) continue;
        const size_t WT = (size_t)C * wsl;
        const size_t smem = (14 * WT + 32) * 4 + 34 * 8;
        if (smem > ctx->smem_optin) continue;
        int maxc = 0;
        int rc = launch_permute_v5_c(ctx, p, C, cfg, KH, NT, smem, true, &maxc);
        if (rc) return rc;
        if (!forced && C > 1 && (uint32_t)maxc < p.nb) continue;
        rc = launch_permute_v5_c(ctx, p, C, cfg, KH, NT, smem, false, &maxc);
        *done = rc == XSI_OK;
        return rc;
    }
    return XSI_OK;
}

template <int KH>
int launch_permute_grid(xsi_ctx* ctx, const EncDev& p, const PermGridCfg& cfg, size_t smem, bool probe_only, int* per_sm) {
    CK(cudaFuncSetAttribute(pbwt_permute_grid_kernel<KH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (probe_only) {
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, pbwt_permute_grid_kernel<KH>, 1024, smem));
        return XSI_OK;
    }
    void* args[] = {const_cast<EncDev*>(&p), const_cast<PermGridCfg*>(&cfg)};
    { PROF("pbwt_permute"); CK(cudaLaunchCooperativeKernel((const void*)pbwt_permute_grid_kernel<KH>, dim3(ctx->sm_count), dim3(1024), args, smem, ctx->es)); }
    CKL();
    return XSI_OK;
}
int launch_permute_grid_kh(xsi_ctx* ctx, uint32_t KH, const EncDev& p, const PermGridCfg& cfg, size_t smem, bool probe_only, int* per_sm) {
    switch (KH) {
        case 8: return launch_permute_grid<8>(ctx, p, cfg, smem, probe_only, per_sm);
        case 16: return launch_permute_grid<16>(ctx, p, cfg, smem, probe_only, per_sm);
        default: return launch_permute_grid<32>(ctx, p, cfg, smem, probe_only, per_sm);
    }
}

// > 65,534 haplotypes: the whole GPU advances a group of PBWT blocks line by line (cooperative launch).
// Groups are as many blocks as the positions-in-registers budget allows (sm_count CTAs of 1024 threads, KH <= 32).
int run_permute_grid(xsi_ctx* ctx, const EncDev& p, bool* done) {
    *done = false;
    auto& e = ctx->enc;
    const uint32_t N = 2 * p.n_samples;
    const uint32_t ctas = (uint32_t)ctx->sm_count;
    const uint32_t WSP = (p.WS + 31) & ~31u, NG = WSP / 4;  // NG a multiple of 8
    const size_t smem = (size_t)NG * 4;
    if (smem > ctx->smem_optin) return XSI_OK;  // more than ~7 M haplotypes: generic kernel
    auto cpb_of = [&](uint32_t kh) { return (uint32_t)((((uint64_t)N + kh - 1) / kh + 1023) / 1024); };  // CTAs per PBWT block
    if (cpb_of(32) > ctas) return XSI_OK;      // more than ~4.8 M haplotypes: generic kernel
    int coop = 0;
    CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
    if (!coop) return XSI_OK;
    const uint32_t max_group = ctas / cpb_of(32);
    const size_t per_block = (size_t)3 * WSP + (size_t)3 * NG;
    CK(e.a_pool.ensure(((size_t)max_group * per_block + 4) * 4));
    for (uint32_t b0 = 0; b0 < p.nb;) {
        const uint32_t rem = p.nb - b0;
        uint32_t KH = 32;
        if (const char* s = getenv("XSI_PBWT_KH")) { const int v = atoi(s); if (v == 8 || v == 16 || v == 32) KH = (uint32_t)v; }
        else for (uint32_t kh : {8u, 16u}) if (ctas / cpb_of(kh) >= rem) { KH = kh; break; }
        PermGridCfg cfg;
        cfg.TPB = cpb_of(KH) * 1024u;
        cfg.b0 = b0;
        cfg.nbg = std::min<uint32_t>(rem, ctas / cpb_of(KH));
        if (cfg.nbg == 0) return XSI_OK;
        cfg.WSP = WSP;
        uint32_t* base = e.a_pool.as<uint32_t>();
        cfg.Y = base;
        cfg.CNT = cfg.Y + (size_t)cfg.nbg * 3 * WSP;
        cfg.bar = cfg.CNT + (size_t)cfg.nbg * 3 * NG;
        int per_sm = 0;
        int rc = launch_permute_grid_kh(ctx, KH, p, cfg, smem, true, &per_sm);
        if (rc) return rc;
        if (per_sm < 1) return XSI_OK;
        CK(cudaMemsetAsync(base, 0, ((size_t)cfg.nbg * per_block + 4) * 4, ctx->es));
        rc = launch_permute_grid_kh(ctx, KH, p, cfg, smem, false, &per_sm);
        if (rc) return rc;
        b0 += cfg.nbg;
    }
    *done = true;
    return XSI_OK;
}

int run_permute(xsi_ctx* ctx, const EncDev& p) {
    const uint32_t N = 2 * p.n_samples;
    const uint32_t W = (N + 31) / 32;
    // XSI_PBWT_V (tests, A/B runs): 4 = one WAH line per exchange round (default), 5 = two lines per round (measured slower:
    // 21.6 vs 17.3 ms at 32 HRC blocks, see the note at pbwt_permute_v5_kernel), 1 = the general kernels (a[] in shared /
    // global memory) that also serve blocks with an all-haploid record
    int ver = 4;
    if (const char* s = getenv("XSI_PBWT_V")) ver = atoi(s);
    // Mixed batch: blocks with an all-haploid record need the general kernel (haploid lines are encoded in the order of
    // haploid_rearrangement_from_diploid, interfaces.hpp:318-333); the other blocks of the batch keep the cluster kernel.
    if (ctx->enc.any_haploid && N <= 65534 && ver >= 4 && p.blk_map == nullptr) {
        auto& e = ctx->enc;
        uint32_t ndip = 0;
        e.h_blk_map.clear();
        for (uint32_t b = 0; b < p.nb; ++b) if (!e.h_blk_hap[b]) { e.h_blk_map.push_back(b); ++ndip; }
        for (uint32_t b = 0; b < p.nb; ++b) if (e.h_blk_hap[b]) e.h_blk_map.push_back(b);
        if (ndip > 0 && ndip < p.nb) {
            CK(e.blk_map.ensure((size_t)p.nb * 4));
            CK(cudaMemcpyAsync(e.blk_map.p, e.h_blk_map.data(), (size_t)p.nb * 4, cudaMemcpyHostToDevice, ctx->es));
            EncDev q = p;
            q.blk_map = e.blk_map.as<uint32_t>(); q.nb = ndip;
            e.any_haploid = false;
            int rc = run_permute(ctx, q);  // the diploid-only blocks
            e.any_haploid = true;
            if (rc) return rc;
            q.blk_map = e.blk_map.as<uint32_t>() + ndip; q.nb = p.nb - ndip;
            return run_permute(ctx, q);     // the blocks with an all-haploid record (general kernel)
        }
    }
    if (!ctx->enc.any_haploid && N > 65534 && ver != 1) {
        bool done = false;
        const int rc = run_permute_grid(ctx, p, &done);
        if (rc || done) return rc;
    }
    if (!ctx->enc.any_haploid && N <= 65534 && ver >= 5) {
        bool done = false;
        const int rc = run_permute_v5(ctx, p, W, &done);
        if (rc || done) return rc;
    }
    if (!ctx->enc.any_haploid && N <= 65534 && ver >= 4) {
        bool done = false;
        const int rc = run_permute_v4(ctx, p, W, &done);
        if (rc || done) return rc;
    }
    const size_t smem = ((size_t)N * 2 + 15) / 16 * 16 + (size_t)4 * p.WS * 4 + 64 * 4 + 16;
    if (N <= 65536 && smem <= ctx->smem_optin) {
        const int wpw = choose_wpw(W);
        const uint32_t NW = (W + wpw - 1) / wpw;
        switch (wpw) {
            case 2: return launch_permute<2, 1024>(ctx, p, NW, smem);
            case 4: return launch_permute<4, 1024>(ctx, p, NW, smem);
            case 8: return launch_permute<8, 1024>(ctx, p, NW, smem);
            case 16: return launch_permute<16, 1024>(ctx, p, NW, smem);
            case 32: return launch_permute<32, 1024>(ctx, p, NW, smem);
            case 64: return launch_permute<64, 1024>(ctx, p, NW, smem);
            default: return launch_permute<128, 512>(ctx, p, NW, smem);
        }
    }
    // > 65536 haplotypes: a[] lives in global memory (two uint32 copies per block + y / even-mask words)
    auto& e = ctx->enc;
    CK(e.a_pool.ensure(((size_t)p.nb * 2 * N + (size_t)p.nb * 2 * p.WS) * 4));
    { PROF("pbwt_permute"); pbwt_permute_gmem_kernel<<<p.nb, 1024, 0, ctx->es>>>(p, e.a_pool.as<uint32_t>()); }
    CKL();
    return XSI_OK;
}

}  // namespace

static int xsi_encode_launch_impl(xsi_ctx* ctx, const xsi_encode_desc* d, uint64_t row_stride = 0) {
    if (!ctx || !d) return XSI_E_ARG;
    auto& e = ctx->enc;
    if (row_stride) {
        if (!d->gt_on_device) { ctx->err = "strided rows must be device rows"; return XSI_E_ARG; }
        if (row_stride < 2ull * d->n_samples) { ctx->err = "row_stride smaller than 2*n_samples"; return XSI_E_ARG; }
    }
    e.launched = false; e.collected = false;
    if (!d->gt || !d->n_allele || d->n_records == 0 || d->n_samples == 0 || d->block_len == 0) { ctx->err = "bad encode descriptor"; return XSI_E_ARG; }
    if (d->gt_elem_bytes != 4 && d->gt_elem_bytes != 1) { ctx->err = "gt_elem_bytes must be 1 or 4"; return XSI_E_ARG; }
    if (d->n_records >= (1ull << 31)) { ctx->err = "batch too large"; return XSI_E_ARG; }
    if (d->n_samples > 32767 && d->n_samples <= 65535) {
        ctx->err = "32768..65535 samples: the reference pairs a uint16 permutation with >65535 haplotypes (xsi_factory.hpp:425 vs gt_compressor_new.hpp:182)";
        return XSI_E_UNSUPPORTED;
    }
    CK(cudaSetDevice(ctx->device));
    const uint64_t R = d->n_records;
    const uint32_t S = d->n_samples;
    e.R = R; e.n_samples = S; e.block_len = d->block_len; e.default_phasing = d->default_phasing ? 1 : 0;
    e.aet = S <= 65535 ? 2 : 4;
    e.wah_missing = d->wah_encode_missing != 0;
    e.nb = (uint32_t)((R + d->block_len - 1) / d->block_len);
    // ---- host tables ----
    std::chrono::steady_clock::time_point t_tables = std::chrono::steady_clock::now();
    // two passes over chunks of records on the worker pool: per-chunk sums, serial prefix, fill
    e.h_nallele.resize(R); e.h_ngt.resize(R); e.h_line0.resize(R); e.h_goff.resize(R);
    e.h_blk_line0.assign(e.nb + 1, 0); e.h_blk_rec0.assign(e.nb + 1, 0);
    e.h_blk_hap.assign(e.nb, 0);
    constexpr uint64_t TCH = 1u << 16;
    const size_t nch = (size_t)((R + TCH - 1) / TCH);
    struct ChunkSum { uint64_t g = 0, l = 0; int rc = XSI_OK; const char* err = nullptr; int max_pl = 0; bool hap = false, al4 = true, al1 = true; };
    std::vector<ChunkSum> cs(nch);
    host_parallel_for(nch, [&](size_t c) {
        ChunkSum& k = cs[c];
        const uint64_t r1 = std::min<uint64_t>(R, (c + 1) * TCH);
        for (uint64_t r = c * TCH; r < r1; ++r) {
            const uint32_t pl = d->ploidy ? d->ploidy[r] : 2;
            const uint32_t na = d->n_allele[r];
            if (pl > 2) { k.err = "Ploidy higher than 2 is not yet supported"; k.rc = XSI_E_PLOIDY; return; }
            if (pl == 0) { k.err = "record with ploidy 0"; k.rc = XSI_E_ARG; return; }
            // 2..254: what xsi_decode_records can read back.  A record without an ALT allele (n_allele 1) has no binary line;
            // the reference then writes per-record flags that no longer line up with the per-line sections
            // (gt_block.hpp:650-666) and cannot decode the block either, so it is refused here rather than written.
            if (na < 2 || na > 254) { k.err = "n_allele out of range (2..254)"; k.rc = XSI_E_UNSUPPORTED; return; }
            if ((int)pl > k.max_pl) k.max_pl = (int)pl;
            if (pl == 1) { k.hap = true; e.h_blk_hap[r / d->block_len] = 1; }
            k.g += (uint64_t)S * pl;
            k.l += na - 1;
        }
    });
    uint64_t goff = 0, L = 0;
    bool al4 = true, al1 = true;  // every row starts and ends on a 16-byte boundary (int32 / int8 elements) -> TMA-fed scan
    e.max_ploidy = 0;
    e.any_haploid = false;
    for (size_t c = 0; c < nch; ++c) {
        ChunkSum& k = cs[c];
        if (k.rc) { ctx->err = k.err; return k.rc; }
        const uint64_t g = k.g, l = k.l;
        k.g = goff; k.l = L;  // now: the chunk's first element / line
        goff += g; L += l;
        if (row_stride) k.g = (uint64_t)c * TCH * row_stride;
        if (L >= (1ull << 31)) { ctx->err = "too many binary lines in batch"; return XSI_E_ARG; }
        e.max_ploidy = std::max(e.max_ploidy, k.max_pl);
        e.any_haploid |= k.hap;
    }
    e.L = L;
    e.h_line_rec.resize(L ? L : 1);
    host_parallel_for(nch, [&](size_t c) {
        ChunkSum& k = cs[c];
        uint64_t g = k.g, l = k.l;
        const uint64_t r1 = std::min<uint64_t>(R, (c + 1) * TCH);
        for (uint64_t r = c * TCH; r < r1; ++r) {
            const uint32_t pl = d->ploidy ? d->ploidy[r] : 2;
            const uint32_t na = d->n_allele[r];
            if (r % d->block_len == 0) { e.h_blk_line0[r / d->block_len] = (uint32_t)l; e.h_blk_rec0[r / d->block_len] = (uint32_t)r; }
            e.h_nallele[r] = na;
            e.h_ngt[r] = S * pl;
            e.h_goff[r] = g;
            if ((g * 4) % 16 || ((uint64_t)S * pl * 4) % 16) k.al4 = false;
            if (g % 16 || ((uint64_t)S * pl) % 16) k.al1 = false;
            e.h_line0[r] = (uint32_t)l;
            for (uint32_t a = 1; a < na; ++a) e.h_line_rec[l + a - 1] = (uint32_t)r;
            g += row_stride ? row_stride : (uint64_t)S * pl;
            l += na - 1;
        }
    });
    for (size_t c = 0; c < nch; ++c) { al4 &= cs[c].al4; al1 &= cs[c].al1; }
    e.h_blk_line0[e.nb] = (uint32_t)L; e.h_blk_rec0[e.nb] = (uint32_t)R;
    for (uint32_t b = 0; b < e.nb; ++b)
        if (e.h_blk_line0[b + 1] - e.h_blk_line0[b] >= 32768) { ctx->err = "block with >= 32768 binary lines (BM offset is 15 bits)"; return XSI_E_UNSUPPORTED; }
    const uint32_t N2 = 2 * S;
    e.WS = ((N2 + 31) / 32 + 3) / 4 * 4;
    e.SLOTW = ((N2 + 14) / 15 + 2 + 7) / 8 * 8;
    const uint64_t Lp = L ? L : 1;

    if (ctx->profile) {
        std::lock_guard<std::mutex> g_(ctx->prof_m);
        auto& sp = ctx->host_spans["host:encode_tables"];
        sp.first++;
        sp.second += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_tables).count();
    }
    // ---- device buffers ----
    int elem = d->gt_elem_bytes;  // element size of the rows on the DEVICE
    const void* dgt = d->gt;
    if (!d->gt_on_device) {
        bool moved = false;
        if (elem == 4 && host_narrow_on()) {
            // int32 rows from bcf_get_genotypes cross the bus in their BCF int8 encoding (host_narrow.cpp)
            CK(e.gt.ensure(goff));
            const int rc = upload_narrowed(ctx, static_cast<const int32_t*>(d->gt), goff, e.gt.as<int8_t>());
            if (rc < 0) return rc;
            if (rc == 1) { moved = true; elem = 1; }
        }
        if (!moved) {
            const size_t gt_bytes = goff * (size_t)elem;
            CK(e.gt.ensure(gt_bytes));
            CK(cudaMemcpyAsync(e.gt.p, d->gt, gt_bytes, cudaMemcpyHostToDevice, ctx->es));
        }
        dgt = e.gt.p;
    }
    bool rows_aligned16 = elem == 4 ? al4 : al1;
    if (reinterpret_cast<uintptr_t>(dgt) % 16) rows_aligned16 = false;
    // tables: goff[R] u64 | ngt[R] | nallele[R] | line0[R] | line_rec[L] | blk_line0[nb+1]
    const size_t t_goff = 0, t_ngt = t_goff + R * 8, t_nal = t_ngt + R * 4, t_l0 = t_nal + R * 4, t_lr = t_l0 + R * 4,
                 t_bl = t_lr + Lp * 4, t_end = t_bl + (e.nb + 1) * 4;
    CK(e.tables.ensure(t_end));
    CK(e.h_small.ensure(t_end));
    {
        uint8_t* h = e.h_small.as<uint8_t>();
        memcpy(h + t_goff, e.h_goff.data(), R * 8);
        memcpy(h + t_ngt, e.h_ngt.data(), R * 4);
        memcpy(h + t_nal, e.h_nallele.data(), R * 4);
        memcpy(h + t_l0, e.h_line0.data(), R * 4);
        memcpy(h + t_lr, e.h_line_rec.data(), Lp * 4);
        memcpy(h + t_bl, e.h_blk_line0.data(), (e.nb + 1) * 4);
        CK(cudaMemcpyAsync(e.tables.p, h, t_end, cudaMemcpyHostToDevice, ctx->es));
    }
    CK(e.bitrows.ensure(Lp * e.WS * 4));
    CK(e.wahslots.ensure(Lp * e.SLOTW * 2));
    CK(e.counters.ensure(16));
    CK(e.rec_aux.ensure(R * 3 * 4));
    CK(e.line_u32.ensure(Lp * 3 * 4));  // cnt | sparse_n | wah_n
    CK(e.line_flags.ensure(Lp));
    CK(e.rec_u32.ensure(R * 5 * 4));    // miss_n | eov_n | phase_n | missw_n | eovw_n (the last two: WS_WAH only)
    CK(e.rec_flags.ensure(R));
    CK(e.wah_list.ensure(Lp * 4));
    CK(e.blk_nwah.ensure(e.nb * 4));
    // offsets: line_sparse_off[L+1] | rec_miss_off[R+1] | rec_eov_off[R+1] | line_wah_off[L+1] | rec_phase_off[R+1]
    const size_t o_sp = 0, o_ms = o_sp + (L + 1) * 8, o_ev = o_ms + (R + 1) * 8, o_wh = o_ev + (R + 1) * 8,
                 o_ph = o_wh + (L + 1) * 8, o_mw = o_ph + (R + 1) * 8, o_ew = o_mw + (R + 1) * 8, o_end = o_ew + (R + 1) * 8;
    CK(e.offs.ensure(o_end));
    CK(e.scanjobs.ensure(8 * sizeof(ScanJob)));
    if (e.aux_cap == 0) e.aux_cap = (uint32_t)std::max<uint64_t>(256, R / 16);
    if (e.phase_cap == 0) e.phase_cap = (uint32_t)std::max<uint64_t>(256, R / 16);

    for (int attempt = 0; attempt < 3; ++attempt) {
        CK(e.auxrows.ensure((size_t)e.aux_cap * e.WS * 4));
        CK(e.phrows.ensure((size_t)e.phase_cap * e.WS * 4));
        CK(e.phslots.ensure((size_t)e.phase_cap * e.SLOTW * 2));
        if (e.wah_missing) CK(e.auxslots.ensure((size_t)e.aux_cap * e.SLOTW * 2));
        EncDev p;
        uint8_t* tb = e.tables.as<uint8_t>();
        p.gt = dgt;
        p.rec_goff = reinterpret_cast<const uint64_t*>(tb + t_goff);
        p.rec_ngt = reinterpret_cast<const uint32_t*>(tb + t_ngt);
        p.rec_nallele = reinterpret_cast<const uint32_t*>(tb + t_nal);
        p.rec_line0 = reinterpret_cast<const uint32_t*>(tb + t_l0);
        p.line_rec = reinterpret_cast<const uint32_t*>(tb + t_lr);
        p.blk_line0 = reinterpret_cast<const uint32_t*>(tb + t_bl);
        p.n_samples = S; p.R = (uint32_t)R; p.L = (uint32_t)L; p.nb = e.nb;
        p.WS = e.WS; p.SLOTW = e.SLOTW; p.aux_cap = e.aux_cap; p.phase_cap = e.phase_cap;
        p.mac_thr = d->mac_threshold; p.default_phasing = e.default_phasing;
        p.bitrows = e.bitrows.as<uint32_t>(); p.auxrows = e.auxrows.as<uint32_t>(); p.phrows = e.phrows.as<uint32_t>();
        p.counters = e.counters.as<uint32_t>(); p.rec_aux = e.rec_aux.as<int32_t>();
        p.line_cnt = e.line_u32.as<uint32_t>(); p.line_sparse_n = p.line_cnt + Lp; p.line_wah_n = p.line_sparse_n + Lp;
        p.line_flags = e.line_flags.as<uint8_t>();
        p.rec_miss_n = e.rec_u32.as<uint32_t>(); p.rec_eov_n = p.rec_miss_n + R; p.rec_phase_n = p.rec_eov_n + R;
        p.rec_flags = e.rec_flags.as<uint8_t>();
        p.wah_list = e.wah_list.as<uint32_t>(); p.blk_nwah = e.blk_nwah.as<uint32_t>();
        p.wahslots = e.wahslots.as<uint16_t>(); p.phslots = e.phslots.as<uint16_t>();
        p.wah_missing = e.wah_missing ? 1u : 0u;
        p.blk_map = nullptr;
        p.auxslots = e.auxslots.as<uint16_t>();
        p.rec_missw_n = p.rec_phase_n + R; p.rec_eovw_n = p.rec_missw_n + R;
        if (e.wah_missing) CK(cudaMemsetAsync(p.rec_missw_n, 0, R * 2 * 4, ctx->es));

        CK(cudaMemsetAsync(e.counters.p, 0, 16, ctx->es));
        if (rows_aligned16 && !getenv("XSI_SCAN_V1")) {
            // TMA-fed persistent scan: CTAs walk records blockIdx.x, +gridDim.x, ...; the CTA is as wide as the rows are long
            // (in words, rounded up to a warp) up to 256 threads, 2 to 6 CTAs per SM
            const uint32_t words = (2 * e.n_samples + 31) / 32;
            uint32_t nt = 256;
            if (const char* s_ = getenv("XSI_SCAN_NT")) nt = (uint32_t)atoi(s_);
            else if (words <= 128) nt = 128;
            else if (words <= 160) nt = 160;
            else if (words <= 192) nt = 192;
            if (nt != 128 && nt != 160 && nt != 192) nt = 256;
            const uint32_t grid = (uint32_t)std::min<uint64_t>(R, (uint64_t)ctx->sm_count * s2_ctas_per_sm((int)nt));
            const size_t stages = (size_t)s2_stages(elem, (int)nt);
            const size_t smem = stages * s2_tile((int)nt) * elem + 2 * stages * 8;
            PROF("scan_rows");
            auto go = [&](auto kern) -> cudaError_t {
                cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e_ != cudaSuccess) return e_;
                kern<<<grid, nt, smem, ctx->es>>>(p);
                return cudaSuccess;
            };
            if (elem == 4) {
                if (nt == 128) CK(go(scan_rows_v2_kernel<4, 128>));
                else if (nt == 160) CK(go(scan_rows_v2_kernel<4, 160>));
                else if (nt == 192) CK(go(scan_rows_v2_kernel<4, 192>));
                else CK(go(scan_rows_v2_kernel<4, 256>));
            } else {
                if (nt == 128) CK(go(scan_rows_v2_kernel<1, 128>));
                else if (nt == 160) CK(go(scan_rows_v2_kernel<1, 160>));
                else if (nt == 192) CK(go(scan_rows_v2_kernel<1, 192>));
                else CK(go(scan_rows_v2_kernel<1, 256>));
            }
        } else {
            PROF("scan_rows");
            if (elem == 4) scan_rows_kernel<4><<<(uint32_t)R, E1_THREADS, 0, ctx->es>>>(p);
            else scan_rows_kernel<1><<<(uint32_t)R, E1_THREADS, 0, ctx->es>>>(p);
        }
        CKL();
        if (ctx->async_encode && ctx->overlap_order && ctx->ev_scan && ctx->es == ctx->stream_enc) {
            CK(cudaEventRecord(ctx->ev_scan, ctx->es));
            ctx->scan_seq.store(ctx->enc_seq.load());
        }
        if (L) {
            { PROF("build_wah_lists"); build_wah_lists_kernel<<<e.nb, 32, 0, ctx->es>>>(p); }
            CKL();
            int rc = run_permute(ctx, p);
            if (rc) return rc;
        }
        {
            const uint64_t jobs = L + (e.wah_missing ? 3 : 1) * R;
            { PROF("wah_encode_rows"); wah_encode_rows_kernel<<<(uint32_t)((jobs + E4_WARPS - 1) / E4_WARPS), E4_WARPS * 32, 0, ctx->es>>>(p); }
            CKL();
        }
        uint8_t* ob = e.offs.as<uint8_t>();
        ScanJob jobs[7] = {{p.line_sparse_n, reinterpret_cast<uint64_t*>(ob + o_sp), (uint32_t)L, 0},
                           {p.rec_miss_n, reinterpret_cast<uint64_t*>(ob + o_ms), (uint32_t)R, 0},
                           {p.rec_eov_n, reinterpret_cast<uint64_t*>(ob + o_ev), (uint32_t)R, 0},
                           {p.line_wah_n, reinterpret_cast<uint64_t*>(ob + o_wh), (uint32_t)L, 0},
                           {p.rec_phase_n, reinterpret_cast<uint64_t*>(ob + o_ph), (uint32_t)R, 0},
                           {p.rec_missw_n, reinterpret_cast<uint64_t*>(ob + o_mw), (uint32_t)R, 0},
                           {p.rec_eovw_n, reinterpret_cast<uint64_t*>(ob + o_ew), (uint32_t)R, 0}};
        const uint32_t n_scans = e.wah_missing ? 7 : 5;
        CK(cudaMemcpyAsync(e.scanjobs.p, jobs, sizeof(jobs), cudaMemcpyHostToDevice, ctx->es));
        {
            const uint32_t max_tiles = scan_tiles((uint32_t)std::max<uint64_t>(L, R));
            CK(e.scansums.ensure((size_t)n_scans * max_tiles * 8));
            PROF("scan_u32");
            scan_u32_sums_kernel<<<dim3(max_tiles, n_scans), SCAN_THREADS, 0, ctx->es>>>(e.scanjobs.as<ScanJob>(), e.scansums.as<uint64_t>(), max_tiles);
            scan_u32_kernel<<<dim3(max_tiles, n_scans), SCAN_THREADS, 0, ctx->es>>>(e.scanjobs.as<ScanJob>(), e.scansums.as<uint64_t>(), max_tiles);
        }
        CKL();
        // section offsets at the block boundaries + counters + flags back, then size the outputs and lay the blocks out
        const size_t nbp = (size_t)e.nb + 1, c_end = 7 * nbp * 8;
        CK(e.blkoffs.ensure(c_end));
        {
            GatherOffs g;
            const size_t at[7] = {o_sp, o_ms, o_ev, o_wh, o_ph, o_mw, o_ew};
            const uint32_t by_line[7] = {1, 0, 0, 1, 0, 0, 0};
            for (int a = 0; a < 7; ++a) { g.src[a] = reinterpret_cast<const uint64_t*>(ob + at[a]); g.by_line[a] = by_line[a]; }
            g.n_arrays = n_scans; g.nb = e.nb; g.block_len = e.block_len; g.R = (uint32_t)R;
            g.blk_line0 = p.blk_line0; g.dst = e.blkoffs.as<uint64_t>();
            gather_block_offsets_kernel<<<(uint32_t)((n_scans * nbp + 255) / 256), 256, 0, ctx->es>>>(g);
            CKL();
        }
        CK(e.h_offs.ensure(c_end + 64));
        uint8_t* ho = e.h_offs.as<uint8_t>();
        memset(ho, 0, c_end);
        CK(cudaMemcpyAsync(ho, e.blkoffs.p, n_scans * nbp * 8, cudaMemcpyDeviceToHost, ctx->es));
        CK(cudaMemcpyAsync(ho + c_end, e.counters.p, 16, cudaMemcpyDeviceToHost, ctx->es));
        CK(e.h_flags.ensure(Lp + R + 16));
        CK(cudaMemcpyAsync(e.h_flags.p, e.line_flags.p, Lp, cudaMemcpyDeviceToHost, ctx->es));
        CK(cudaMemcpyAsync(e.h_flags.as<uint8_t>() + Lp, e.rec_flags.p, R, cudaMemcpyDeviceToHost, ctx->es));
        CK(cudaStreamSynchronize(ctx->es));
        const uint32_t* cnt = reinterpret_cast<const uint32_t*>(ho + c_end);
        if (cnt[2] & ERR_ALLELE) { ctx->err = "Unknown allele error !"; return XSI_E_ALLELE; }
        if (cnt[2] & (ERR_AUX_OVERFLOW | ERR_PHASE_OVERFLOW)) {
            e.aux_cap = std::max(e.aux_cap, cnt[0] + cnt[0] / 8 + 16);
            e.phase_cap = std::max(e.phase_cap, cnt[1] + cnt[1] / 8 + 16);
            continue;  // rerun the batch with big enough row pools
        }
        // compact offsets: array a (sparse, missing, eov, wah, phase, missing-wah, eov-wah) at block b = hoff[a * (nb+1) + b]
        const uint64_t* hoff = reinterpret_cast<const uint64_t*>(ho);
        e.tot_sparse = hoff[0 * nbp + e.nb];
        e.tot_miss = hoff[1 * nbp + e.nb];
        e.tot_eov = hoff[2 * nbp + e.nb];
        e.tot_wah = hoff[3 * nbp + e.nb];
        e.tot_phase = hoff[4 * nbp + e.nb];
        e.tot_missw = e.wah_missing ? hoff[5 * nbp + e.nb] : 0;
        e.tot_eovw = e.wah_missing ? hoff[6 * nbp + e.nb] : 0;
        CK(e.out_sparse.ensure(e.tot_sparse * e.aet + 16));
        CK(e.out_miss.ensure(e.tot_miss * e.aet + 16));
        CK(e.out_eov.ensure(e.tot_eov * e.aet + 16));
        CK(e.out_wah.ensure(e.tot_wah * 2 + 16));
        CK(e.out_phase.ensure(e.tot_phase * 2 + 16));
        if (e.wah_missing) { CK(e.out_missw.ensure(e.tot_missw * 2 + 16)); CK(e.out_eovw.ensure(e.tot_eovw * 2 + 16)); }
        EmitDev em;
        em.line_sparse_off = reinterpret_cast<const uint64_t*>(ob + o_sp);
        em.rec_miss_off = reinterpret_cast<const uint64_t*>(ob + o_ms);
        em.rec_eov_off = reinterpret_cast<const uint64_t*>(ob + o_ev);
        em.line_wah_off = reinterpret_cast<const uint64_t*>(ob + o_wh);
        em.rec_phase_off = reinterpret_cast<const uint64_t*>(ob + o_ph);
        em.out_sparse = e.out_sparse.p; em.out_miss = e.out_miss.p; em.out_eov = e.out_eov.p;
        em.out_wah = e.out_wah.as<uint16_t>(); em.out_phase = e.out_phase.as<uint16_t>();
        em.rec_missw_off = reinterpret_cast<const uint64_t*>(ob + o_mw); em.rec_eovw_off = reinterpret_cast<const uint64_t*>(ob + o_ew);
        em.out_missw = e.out_missw.as<uint16_t>(); em.out_eovw = e.out_eovw.as<uint16_t>();
        {
            const uint64_t jobs5 = L + 2 * R;
            const uint32_t grid = (uint32_t)((jobs5 + E5_WARPS - 1) / E5_WARPS);
            {
                PROF("sparse_emit");
                if (e.aet == 2) sparse_emit_kernel<uint16_t><<<grid, E5_WARPS * 32, 0, ctx->es>>>(p, em);
                else sparse_emit_kernel<uint32_t><<<grid, E5_WARPS * 32, 0, ctx->es>>>(p, em);
            }
            CKL();
            const uint64_t jobs6 = L + (e.wah_missing ? 3 : 1) * R;
            { PROF("pack_wah"); pack_wah_kernel<<<(uint32_t)((jobs6 + E6_WARPS - 1) / E6_WARPS), E6_WARPS * 32, 0, ctx->es>>>(p, em); }
            CKL();
        }
        // ---- while those run: lay out every GT block (dictionary, per-line bool vectors, section offsets) in the
        //      pinned arena, then queue the device->host copies of the sections straight into their final place ----
        {
            const uint64_t *off_sp = hoff, *off_ms = hoff + nbp, *off_ev = hoff + 2 * nbp, *off_wh = hoff + 3 * nbp, *off_ph = hoff + 4 * nbp,
                           *off_mw = hoff + 5 * nbp, *off_ew = hoff + 6 * nbp;  // indexed by BLOCK: [b] first entry of block b, [b+1] its end
            const uint8_t* lflags = e.h_flags.as<uint8_t>();
            const uint8_t* rflags = lflags + Lp;
            HOSTSPAN("host:encode_layout");
            struct Copy { uint64_t dst; const void* src; uint64_t bytes; };  // dst relative to the block start
            struct Tail { uint64_t at; std::vector<uint8_t> bytes; };       // later bool vectors (missing / eov / phase / haploid)
            struct Layout { std::vector<uint8_t> head; std::vector<Tail> tails; std::vector<Copy> copies; uint64_t size = 0; uint64_t n_wah = 0; };
            std::vector<Layout> lay(e.nb);
            // one task per block on the worker pool; the arena offsets are a prefix sum afterwards
            host_parallel_for(e.nb, [&](size_t bb) {
                const uint32_t b = (uint32_t)bb;
                Layout& ly = lay[b];
                const uint32_t r0 = e.h_blk_rec0[b], r1 = e.h_blk_rec0[b + 1], l0 = e.h_blk_line0[b], l1 = e.h_blk_line0[b + 1];
                const uint32_t nrec = r1 - r0, nlines = l1 - l0;
                bool any_missing = false, any_eov = false, any_phase = false, any_hap = false;
                uint32_t max_pl = 1;  // gt_block.hpp:168
                std::vector<uint8_t> v_wah(nlines), v_miss(nlines, 0), v_eov(nlines, 0), v_phase(nlines, 0), v_hap(nrec, 0);
                for (uint32_t l = l0; l < l1; ++l) { v_wah[l - l0] = (lflags[l] & LF_WAH) ? 1 : 0; ly.n_wah += v_wah[l - l0]; }
                for (uint32_t r = r0; r < r1; ++r) {
                    const uint8_t f = rflags[r];
                    const uint32_t pl = e.h_ngt[r] / e.n_samples;
                    max_pl = std::max(max_pl, pl);
                    if (f & RF_HAPLOID) { any_hap = true; v_hap[r - r0] = 1; }
                    any_missing |= (f & RF_MISSING) != 0; any_eov |= (f & RF_EOV) != 0; any_phase |= (f & RF_PHASE) != 0;
                    // re-index record flags to the record's first binary line (gt_block.hpp:650-666)
                    const uint32_t bl = e.h_line0[r] - l0;
                    if (bl < nlines) {
                        if (f & RF_MISSING) v_miss[bl] = 1;
                        if (f & RF_EOV) v_eov[bl] = 1;
                        if (f & RF_PHASE) v_phase[bl] = 1;
                    }
                }
                // dictionary, insertion sequence of GtBlock::fill_dictionary (gt_block.hpp:464-510)
                RefDictOrder ord;
                std::map<uint32_t, uint32_t> val;
                auto ins = [&](uint32_t k, uint32_t v) { ord.insert(k); val[k] = v; };
                ins(KEY_BCF_LINES, nrec); ins(KEY_BINARY_LINES, nlines); ins(KEY_MAX_LINE_PLOIDY, max_pl);
                ins(KEY_DEFAULT_PHASING, (uint32_t)e.default_phasing); ins(KEY_WEIRDNESS_STRATEGY, e.wah_missing ? WS_WAH : WS_SPARSE);
                ins(KEY_LINE_SORT, VAL_UNDEFINED); ins(KEY_LINE_SELECT, VAL_UNDEFINED); ins(KEY_MATRIX_WAH, VAL_UNDEFINED);
                ins(KEY_MATRIX_SPARSE, VAL_UNDEFINED);
                if (any_missing) { ins(KEY_LINE_MISSING, VAL_UNDEFINED); ins(KEY_MATRIX_MISSING, VAL_UNDEFINED); ins(KEY_MATRIX_MISSING_SPARSE, VAL_UNDEFINED); }
                if (any_eov) { ins(KEY_LINE_END_OF_VECTORS, VAL_UNDEFINED); ins(KEY_MATRIX_END_OF_VECTORS, VAL_UNDEFINED); ins(KEY_MATRIX_END_OF_VECTORS_SPARSE, VAL_UNDEFINED); }
                if (any_phase) { ins(KEY_LINE_NON_UNIFORM_PHASING, VAL_UNDEFINED); ins(KEY_MATRIX_NON_UNIFORM_PHASING, VAL_UNDEFINED); }
                if (any_hap) ins(KEY_LINE_HAPLOID, VAL_UNDEFINED);
                const std::vector<uint32_t> order = ord.order();

                std::vector<uint8_t>& head = ly.head;
                put_u32(head, 0xFFFFFFFFu);
                put_u32(head, (uint32_t)order.size());
                const size_t dict_at = head.size();
                head.resize(head.size() + order.size() * 8);
                // write_writables, gt_block.hpp:512-647
                val[KEY_LINE_SORT] = val[KEY_LINE_SELECT] = (uint32_t)head.size();
                wah16_encode_bools(v_wah, head);
                uint64_t pos = head.size();  // running size of the block
                auto section = [&](uint32_t key, const void* dev, uint64_t first, uint64_t last, uint32_t unit) {
                    val[key] = (uint32_t)pos;
                    const uint64_t bytes = (last - first) * unit;
                    if (bytes) ly.copies.push_back({pos, (const uint8_t*)dev + first * unit, bytes});
                    pos += bytes;
                };
                auto boolvec = [&](uint32_t key, const std::vector<uint8_t>& v) {
                    val[key] = (uint32_t)pos;
                    Tail t; t.at = pos;
                    wah16_encode_bools(v, t.bytes);
                    pos += t.bytes.size();
                    ly.tails.push_back(std::move(t));
                };
                section(KEY_MATRIX_WAH, e.out_wah.p, off_wh[b], off_wh[b + 1], 2);
                section(KEY_MATRIX_SPARSE, e.out_sparse.p, off_sp[b], off_sp[b + 1], e.aet);
                if (any_missing) {
                    boolvec(KEY_LINE_MISSING, v_miss);
                    if (e.wah_missing) section(KEY_MATRIX_MISSING, e.out_missw.p, off_mw[b], off_mw[b + 1], 2);  // gt_block.hpp:574-576
                    else section(KEY_MATRIX_MISSING_SPARSE, e.out_miss.p, off_ms[b], off_ms[b + 1], e.aet);
                }
                if (any_eov) {
                    boolvec(KEY_LINE_END_OF_VECTORS, v_eov);
                    if (e.wah_missing) section(KEY_MATRIX_END_OF_VECTORS, e.out_eovw.p, off_ew[b], off_ew[b + 1], 2);  // :596-598
                    else section(KEY_MATRIX_END_OF_VECTORS_SPARSE, e.out_eov.p, off_ev[b], off_ev[b + 1], e.aet);
                }
                if (any_phase) { boolvec(KEY_LINE_NON_UNIFORM_PHASING, v_phase); section(KEY_MATRIX_NON_UNIFORM_PHASING, e.out_phase.p, off_ph[b], off_ph[b + 1], 2); }
                if (any_hap) boolvec(KEY_LINE_HAPLOID, v_hap);  // one bit per BCF line, gt_block.hpp:219-224,639-642
                for (size_t i = 0; i < order.size(); ++i) {
                    memcpy(head.data() + dict_at + 8 * i, &order[i], 4);
                    const uint32_t v = val[order[i]];
                    memcpy(head.data() + dict_at + 8 * i + 4, &v, 4);
                }
                ly.size = pos;
            });
            const int gen = e.gen ^ 1;  // the other arena: what the previous collect handed out stays intact
            e.block_at.assign(e.nb, 0);
            e.block_sizes[gen].assign(e.nb, 0);
            e.n_wah_lines = 0;
            uint64_t arena_size = 0;
            for (uint32_t b = 0; b < e.nb; ++b) {
                e.block_at[b] = arena_size;
                e.block_sizes[gen][b] = lay[b].size;
                e.n_wah_lines += lay[b].n_wah;
                arena_size = (arena_size + lay[b].size + 15) / 16 * 16;
            }
            CK(e.arena[gen].ensure(arena_size + 16));
            uint8_t* ar = e.arena[gen].as<uint8_t>();
            e.block_ptrs[gen].assign(e.nb, nullptr);
            e.gen = gen;
            for (uint32_t b = 0; b < e.nb; ++b) {
                uint8_t* at = ar + e.block_at[b];
                e.block_ptrs[gen][b] = at;
                memcpy(at, lay[b].head.data(), lay[b].head.size());
                for (const Tail& t : lay[b].tails) memcpy(at + t.at, t.bytes.data(), t.bytes.size());
                for (const Copy& c : lay[b].copies) CK(cudaMemcpyAsync(at + c.dst, c.src, c.bytes, cudaMemcpyDeviceToHost, ctx->es));
            }
        }
        e.launched = true;
        return XSI_OK;
    }
    ctx->err = "row pool overflow persisted";
    return XSI_E_NOMEM;
}

static int xsi_encode_collect_impl(xsi_ctx* ctx, uint32_t* n_blocks_out, const uint8_t* const** blocks_out,
                                  const uint64_t** sizes_out) {
    if (!ctx) return XSI_E_ARG;
    auto& e = ctx->enc;
    if (!e.launched) { ctx->err = "xsi_encode_collect without a successful xsi_encode_launch"; return XSI_E_ARG; }
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->es));  // the sections have landed in the arena (laid out in xsi_encode_launch)
    e.collected = true;
    if (n_blocks_out) *n_blocks_out = e.nb;
    if (blocks_out) *blocks_out = e.block_ptrs[e.gen].data();
    if (sizes_out) *sizes_out = e.block_sizes[e.gen].data();
    return XSI_OK;
}

extern "C" int xsi_encode_block_sizes(xsi_ctx* ctx, uint32_t* n_blocks_out, const uint64_t** sizes_out) {
    if (!ctx || !ctx->enc.collected) return XSI_E_ARG;
    if (n_blocks_out) *n_blocks_out = ctx->enc.nb;
    if (sizes_out) *sizes_out = ctx->enc.block_sizes[ctx->enc.gen].data();
    return XSI_OK;
}
extern "C" int xsi_encode_max_ploidy(const xsi_ctx* ctx) { return ctx ? ctx->enc.max_ploidy : 0; }
extern "C" int xsi_encode_line_counts(const xsi_ctx* ctx, uint64_t* n_binary_lines, uint64_t* n_wah_lines) {
    if (!ctx || !ctx->enc.collected) return XSI_E_ARG;
    if (n_binary_lines) *n_binary_lines = ctx->enc.L;
    if (n_wah_lines) *n_wah_lines = ctx->enc.n_wah_lines;
    return XSI_OK;
}

// =================================================================================================
// DECODE
// =================================================================================================
namespace {

#undef XSI_STREAM
#define XSI_STREAM ctx->stream
struct ParsedBlock {
    std::map<uint32_t, uint32_t> dict;
    uint32_t bcf_lines = 0, bin_lines = 0, default_phasing = 0;
    uint32_t ws = WS_SPARSE;  // KEY_WEIRDNESS_STRATEGY: WS_SPARSE, or WS_WAH for --wah-encode-missing files
    std::vector<uint8_t> is_wah, has_missing, has_eov, has_phase, haploid;
    bool p_missing = false, p_eov = false, p_phase = false;
};

// end of the section that starts at `off`: the next larger dictionary offset, else the block end
uint64_t section_end(const ParsedBlock& pb, uint32_t off, uint64_t size) {
    uint64_t end = size;
    for (const auto& kv : pb.dict)
        if (kv.first >= 0x10 && kv.second != VAL_UNDEFINED && kv.second > off && kv.second < end) end = kv.second;
    return end;
}

int parse_block(std::string& err, const uint8_t* p, uint64_t size, ParsedBlock& pb) {
    if (size < 8 || rd_u32(p) != 0xFFFFFFFFu) { err = "GT block: missing dictionary marker"; return XSI_E_FORMAT; }
    const uint32_t n = rd_u32(p + 4);
    if (8 + (uint64_t)n * 8 > size) { err = "GT block: truncated dictionary"; return XSI_E_FORMAT; }
    for (uint32_t i = 0; i < n; ++i) pb.dict[rd_u32(p + 8 + 8 * (size_t)i)] = rd_u32(p + 12 + 8 * (size_t)i);
    auto need = [&](uint32_t k, uint32_t& v) { auto it = pb.dict.find(k); if (it == pb.dict.end()) return false; v = it->second; return true; };
    uint32_t dp = 0, ws = 0;
    if (!need(KEY_BCF_LINES, pb.bcf_lines) || !need(KEY_BINARY_LINES, pb.bin_lines) || !need(KEY_DEFAULT_PHASING, dp)) { err = "GT block: required key missing"; return XSI_E_FORMAT; }
    // BM addresses a binary line with 15 bits (accessor_internals.hpp:411) and every line costs at least a bit of the
    // LINE_SELECT vector: anything else is not a GT block (and would size the flag vectors from unvalidated input)
    if (pb.bin_lines >= 32768 || pb.bcf_lines > pb.bin_lines + 1 || (uint64_t)pb.bin_lines > size * 8) { err = "GT block: implausible line counts"; return XSI_E_FORMAT; }
    pb.default_phasing = dp == 1 ? 1 : 0;  // accessor_internals_new.hpp:77-81
    // WS_SPARSE is the writer's default, WS_WAH what --wah-encode-missing selects (gt_block.hpp:174-176); WS_PBWT_WAH (a
    // second PBWT over the weirdness lines, the v4 default) cannot be produced by the v5 command line
    if (!need(KEY_WEIRDNESS_STRATEGY, ws) || (ws != WS_SPARSE && ws != WS_WAH)) { err = "GT block: unsupported weirdness strategy (a version-4 PBWT+WAH file?)"; return XSI_E_UNSUPPORTED; }
    pb.ws = ws;
    auto vec = [&](uint32_t key, std::vector<uint8_t>& v, bool& present) {
        present = false;
        auto it = pb.dict.find(key);
        if (it == pb.dict.end() || it->second == VAL_UNDEFINED || it->second >= size) return;
        wah16_decode_bools(p + it->second, p + size, pb.bin_lines, v);  // accessor_internals_new.hpp:591-604
        present = true;
    };
    bool pw = false, ps = false, ph = false;
    vec(KEY_LINE_SELECT, pb.is_wah, pw);
    if (!pw) { err = "GT block: no LINE_SELECT vector"; return XSI_E_FORMAT; }
    std::vector<uint8_t> sort;
    vec(KEY_LINE_SORT, sort, ps);
    if (ps && sort != pb.is_wah) { err = "GT block: LINE_SORT differs from LINE_SELECT (not produced by the v5 writer)"; return XSI_E_UNSUPPORTED; }
    vec(KEY_LINE_MISSING, pb.has_missing, pb.p_missing);
    vec(KEY_LINE_END_OF_VECTORS, pb.has_eov, pb.p_eov);
    vec(KEY_LINE_NON_UNIFORM_PHASING, pb.has_phase, pb.p_phase);
    vec(KEY_LINE_HAPLOID, pb.haploid, ph);  // read per BINARY line although written per BCF line (reference quirk)
    // The writer emits LINE_HAPLOID with one bit per BCF line (gt_block.hpp:219-224,639-642), the reader takes one bit per
    // BINARY line (accessor_internals_new.hpp:116,165,265-271).  With a multi-allelic record in the block the two disagree:
    // the reference's Accessor then corrupts its heap on such a block (probed: "munmap_chunk(): invalid pointer").  The file
    // is written byte for byte like the reference's, but it is not decoded: the caller gets XSI_E_UNSUPPORTED instead.
    if (ph && pb.bcf_lines != pb.bin_lines) {
        bool any = false;
        for (uint8_t f : pb.haploid) any |= f != 0;
        if (any) { err = "GT block holds all-haploid and multi-allelic records: the reference cannot decode it either (LINE_HAPLOID is per BCF line)"; return XSI_E_UNSUPPORTED; }
    }
    if (!ph) pb.haploid.assign(pb.bin_lines, 0);
    if (!pb.p_missing) pb.has_missing.assign(pb.bin_lines, 0);
    if (!pb.p_eov) pb.has_eov.assign(pb.bin_lines, 0);
    if (!pb.p_phase) pb.has_phase.assign(pb.bin_lines, 0);
    return XSI_OK;
}

template <int WPW, int MAXT>
int launch_unpermute(xsi_ctx* ctx, const DecDev& dd, const uint8_t* job_hap, uint32_t NW, size_t smem) {
    CK(cudaFuncSetAttribute(pbwt_unpermute_smem_kernel<WPW, MAXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    { PROF("pbwt_unpermute"); pbwt_unpermute_smem_kernel<WPW, MAXT><<<dd.nb, NW * 32, smem, ctx->stream>>>(dd, job_hap); }
    CKL();
    return XSI_OK;
}

}  // namespace

// D2 v3 over blocks [b0, b0 + nb): the WAH lines [wah_done, wah_todo) the device block table names
static int launch_unpermute_v3(xsi_ctx* ctx, const DecDev& dd, uint32_t b0, uint32_t nb) {
    auto& d = ctx->dec;
    const dim3 g(nb, d.v3_slices);
    uint16_t* ps = d.pos_state.as<uint16_t>();
    const int fence = getenv("XSI_UNPERM_FENCE") ? atoi(getenv("XSI_UNPERM_FENCE")) : 0;
    if (d.v3_kh == 32) {
        CK(cudaFuncSetAttribute(pbwt_unpermute_v3_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.v3_smem));
        PROF("pbwt_unpermute"); pbwt_unpermute_v3_kernel<32><<<g, d.v3_nc + 32, d.v3_smem, ctx->stream>>>(dd, b0, ps, d.ps_stride, fence);
    } else if (d.v3_kh == 16) {
        CK(cudaFuncSetAttribute(pbwt_unpermute_v3_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.v3_smem));
        PROF("pbwt_unpermute"); pbwt_unpermute_v3_kernel<16><<<g, d.v3_nc + 32, d.v3_smem, ctx->stream>>>(dd, b0, ps, d.ps_stride, fence);
    } else {
        CK(cudaFuncSetAttribute(pbwt_unpermute_v3_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)d.v3_smem));
        PROF("pbwt_unpermute"); pbwt_unpermute_v3_kernel<8><<<g, d.v3_nc + 32, d.v3_smem, ctx->stream>>>(dd, b0, ps, d.ps_stride, fence);
    }
    CKL();
    return XSI_OK;
}

// Makes the binary lines [0, line_end) of loaded block b final (continues its inverse-PBWT chain as far as needed).
static int extend_chain(xsi_ctx* ctx, uint32_t b, uint32_t line_end) {
    auto& d = ctx->dec;
    if (!d.lazy_ok || b >= d.nb) return XSI_OK;
    const auto& wl = d.h_wah_lines[b];
    const uint32_t todo = (uint32_t)(std::lower_bound(wl.begin(), wl.end(), (uint16_t)std::min<uint32_t>(line_end, 65535u)) - wl.begin());
    if (todo <= d.h_wah_done[b]) return XSI_OK;
    const uint32_t range[2] = {d.h_wah_done[b], todo};  // DecBlock::wah_done, wah_todo (adjacent)
    uint8_t* dev_blk = d.meta.as<uint8_t>() + d.m_blk + (size_t)b * sizeof(DecBlock) + offsetof(DecBlock, wah_done);
    CK(cudaMemcpyAsync(dev_blk, range, sizeof range, cudaMemcpyHostToDevice, ctx->stream));  // pageable source: staged before the call returns
    const int rc = launch_unpermute_v3(ctx, d.dev, b, 1);
    if (rc) return rc;
    d.h_wah_done[b] = todo;
    return XSI_OK;
}

// blocks loaded lazily: continue their inverse-PBWT chains up to the last line the requests need (plus a window)
static int ensure_lines(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset, const uint32_t* n_alleles) {
    auto& d = ctx->dec;
    if (!d.lazy_ok) return XSI_OK;
    uint32_t window = 1024;
    if (const char* s_ = getenv("XSI_LAZY_WINDOW")) window = (uint32_t)std::max(0, atoi(s_));
    std::vector<uint32_t> need(d.nb, 0);
    for (uint64_t i = 0; i < n; ++i) need[block_index[i]] = std::max(need[block_index[i]], line_offset[i] + n_alleles[i] - 1);
    for (uint32_t b = 0; b < d.nb; ++b) {
        if (!need[b] || d.h_wah_done[b] >= d.h_wah_lines[b].size()) continue;
        const int rc = extend_chain(ctx, b, std::min<uint32_t>(d.h_bin_lines[b], need[b] + window));
        if (rc) return rc;
    }
    return XSI_OK;
}

static int xsi_decode_load_blocks_impl(xsi_ctx* ctx, uint32_t n_blocks, const uint8_t* const* gt_blocks,
                                      const uint64_t* sizes, uint64_t num_samples, int32_t aet_bytes, uint32_t lazy_lines = 0xFFFFFFFFu) {
    if (!ctx || !gt_blocks || !sizes || n_blocks == 0) return XSI_E_ARG;
    if (aet_bytes != 2 && aet_bytes != 4) { ctx->err = "Unsupported access type"; return XSI_E_ARG; }
    if (num_samples == 0 || num_samples >= (1ull << 30)) { ctx->err = "bad num_samples"; return XSI_E_ARG; }
    CK(cudaSetDevice(ctx->device));
    auto& d = ctx->dec;
    d.loaded = false;
    const uint32_t S = (uint32_t)num_samples, N = 2 * S;
    d.n_samples = S; d.aet = (uint32_t)aet_bytes; d.nb = n_blocks;
    d.WS = ((N + 31) / 32 + 3) / 4 * 4;

    std::vector<ParsedBlock> pbs(n_blocks);
    std::vector<uint64_t> blob_off(n_blocks);
    uint64_t blob_size = 0;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        blob_off[b] = blob_size;
        blob_size += (sizes[b] + 15) / 16 * 16;
    }
    // the payload upload runs while the host parses the dictionaries and flag vectors
    CK(d.blob.ensure(blob_size + 64));
    for (uint32_t b = 0; b < n_blocks; ++b)
        CK(cudaMemcpyAsync(d.blob.as<uint8_t>() + blob_off[b], gt_blocks[b], sizes[b], cudaMemcpyHostToDevice, ctx->stream));

    // ---- pass A (worker pool, one task per block): parse, count what the block contributes ----
    struct BlkPlan {
        int rc = XSI_OK; std::string err;
        uint32_t n_wah = 0, n_sp = 0, n_ms = 0, n_ev = 0, n_ph = 0;
        uint32_t n_msw = 0, n_evw = 0;         // missing / end-of-vector lines stored as WAH (WS_WAH blocks): expand-only jobs
        uint32_t wah_words = 0, ph_words = 0;  // u16 words of the WAH / phase matrix
        uint32_t wah_val = 0, ph_val = 0;
        uint32_t msw_words = 0, evw_words = 0, msw_val = 0, evw_val = 0;
        uint32_t msw0 = 0, evw0 = 0, seg_m = 0, seg_e = 0, tile_m = 0, tile_e = 0;
        bool any_hap = false;
        // prefix sums (filled serially)
        uint64_t line0 = 0;
        uint32_t wah0 = 0, sp0 = 0, ms0 = 0, ev0 = 0, ph0 = 0, seg_w = 0, seg_p = 0, tile_w = 0, tile_p = 0;
    };
    std::vector<BlkPlan> plan(n_blocks);
    d.h_wah_lines.resize(n_blocks);
    d.h_wah_done.assign(n_blocks, 0);
    {
        HOSTSPAN("host:decode_parse");
        host_parallel_for(n_blocks, [&](size_t b) {
            BlkPlan& pl = plan[b];
            ParsedBlock& pb = pbs[b];
            pl.rc = parse_block(pl.err, gt_blocks[b], sizes[b], pb);
            if (pl.rc) return;
            d.h_wah_lines[b].clear();
            for (uint32_t l = 0; l < pb.bin_lines; ++l) {
                if (pb.is_wah[l]) { pl.n_wah++; d.h_wah_lines[b].push_back((uint16_t)l); if (pb.haploid[l]) pl.any_hap = true; } else pl.n_sp++;
                if (pb.ws == WS_WAH) { pl.n_msw += pb.has_missing[l] != 0; pl.n_evw += pb.has_eov[l] != 0; }
                else { pl.n_ms += pb.has_missing[l] != 0; pl.n_ev += pb.has_eov[l] != 0; }
                pl.n_ph += pb.has_phase[l] != 0;
            }
            auto matrix = [&](uint32_t key, uint32_t& val, uint32_t& words, const char* what) {
                auto it = pb.dict.find(key);
                if (it == pb.dict.end() || it->second == VAL_UNDEFINED || it->second >= sizes[b]) { pl.err = what; pl.rc = XSI_E_FORMAT; return; }
                val = it->second;
                words = (uint32_t)((section_end(pb, it->second, sizes[b]) - it->second) / 2);
            };
            if (pl.n_wah) matrix(KEY_MATRIX_WAH, pl.wah_val, pl.wah_words, "GT block: WAH lines without a WAH matrix");
            if (pl.n_ph && !pl.rc) matrix(KEY_MATRIX_NON_UNIFORM_PHASING, pl.ph_val, pl.ph_words, "GT block: phase lines without their matrix");
            if (pl.n_msw && !pl.rc) matrix(KEY_MATRIX_MISSING, pl.msw_val, pl.msw_words, "GT block: WAH missing lines without their matrix");
            if (pl.n_evw && !pl.rc) matrix(KEY_MATRIX_END_OF_VECTORS, pl.evw_val, pl.evw_words, "GT block: WAH end-of-vector lines without their matrix");
        });
    }
    for (uint32_t b = 0; b < n_blocks; ++b)
        if (plan[b].rc) { ctx->err = plan[b].err; return plan[b].rc; }
    // ---- serial prefix over blocks: GT WAH jobs of every block first (contiguous per block), then the phase
    //      lines as extra expand-only jobs, one segment per block that has any ----
    uint64_t Lt = 0;
    uint32_t njobs = 0, nsp = 0, nms = 0, nev = 0, nseg = 0, ntiles = 0;
    bool any_hap_job = false;
    auto tiles_of = [](uint32_t words) { return (words + D0_TILE - 1) / D0_TILE; };
    for (uint32_t b = 0; b < n_blocks; ++b) {
        BlkPlan& pl = plan[b];
        pl.line0 = Lt; Lt += pbs[b].bin_lines;
        pl.wah0 = njobs; njobs += pl.n_wah;
        pl.sp0 = nsp; nsp += pl.n_sp;
        pl.ms0 = nms; nms += pl.n_ms;
        pl.ev0 = nev; nev += pl.n_ev;
        if (pl.n_wah) { pl.seg_w = nseg++; pl.tile_w = ntiles; ntiles += tiles_of(pl.wah_words); }
        any_hap_job |= pl.any_hap;
    }
    d.n_gt_jobs = njobs;
    for (uint32_t b = 0; b < n_blocks; ++b) {
        BlkPlan& pl = plan[b];
        if (!pl.n_ph) continue;
        pl.ph0 = njobs; njobs += pl.n_ph;
        pl.seg_p = nseg++; pl.tile_p = ntiles; ntiles += tiles_of(pl.ph_words);
    }
    for (uint32_t b = 0; b < n_blocks; ++b) {  // WS_WAH blocks: their missing / end-of-vector lines, one segment each
        BlkPlan& pl = plan[b];
        if (pl.n_msw) { pl.msw0 = njobs; njobs += pl.n_msw; pl.seg_m = nseg++; pl.tile_m = ntiles; ntiles += tiles_of(pl.msw_words); }
        if (pl.n_evw) { pl.evw0 = njobs; njobs += pl.n_evw; pl.seg_e = nseg++; pl.tile_e = ntiles; ntiles += tiles_of(pl.evw_words); }
    }
    d.Lt = Lt;
    d.NJ = njobs;
    // The chain kernel that can stop and continue is D2 v3 (positions in registers, <= 65534 haplotypes, no haploid lines);
    // anything else undoes the PBWT order of whole blocks at once.
    const bool chain_v3 = d.n_gt_jobs && !any_hap_job && 2 * S <= 65534 && !(getenv("XSI_PBWT_V") && atoi(getenv("XSI_PBWT_V")) == 1) &&
                          (size_t)D3_STAGES * (2 * d.WS + 4) * 4 + 2 * D3_STAGES * 8 <= ctx->smem_optin;
    d.lazy_ok = chain_v3;
    const bool lazy = chain_v3 && lazy_lines != 0xFFFFFFFFu;
    // ---- one pinned staging area for everything the kernels index ----
    const uint32_t NJp = njobs ? njobs : 1, ntp = ntiles ? ntiles : 1;
    const uint64_t Ltp = Lt ? Lt : 1;
    size_t at = 0;
    auto take = [&](size_t bytes) { const size_t o = at; at = (at + bytes + 15) / 16 * 16; return o; };
    const size_t s_blk = take(sizeof(DecBlock) * n_blocks), s_seg = take(sizeof(DecSeg) * (nseg ? nseg : 1));
    const size_t s_job = take((size_t)NJp * 3 * 4), s_hap = take(NJp), s_tile = take((size_t)ntp * 2 * 4), s_dl = take(Ltp * 17);
    CK(d.h_meta.ensure(at));
    uint8_t* hm = d.h_meta.as<uint8_t>();
    DecBlock* blocks = reinterpret_cast<DecBlock*>(hm + s_blk);
    DecSeg* segs = reinterpret_cast<DecSeg*>(hm + s_seg);
    uint32_t* job_gcum = reinterpret_cast<uint32_t*>(hm + s_job), *job_nbits = job_gcum + NJp, *job_seg = job_nbits + NJp;
    uint8_t* job_hap = hm + s_hap;
    uint32_t* tile_seg = reinterpret_cast<uint32_t*>(hm + s_tile), *tile_word0 = tile_seg + ntp;
    uint32_t* dl_ord = reinterpret_cast<uint32_t*>(hm + s_dl), *dl_mord = dl_ord + Ltp, *dl_eord = dl_mord + Ltp, *dl_pord = dl_eord + Ltp;
    uint8_t* dl_flags = reinterpret_cast<uint8_t*>(dl_pord + Ltp);
    d.hv_blocks = blocks; d.hv_segs = segs; d.hv_job_seg = job_seg; d.hv_dl_ord = dl_ord; d.hv_dl_flags = dl_flags;
    d.h_blob_off = blob_off; d.h_blk_size.assign(sizes, sizes + n_blocks);
    // ---- pass B (worker pool): every block fills its own slices ----
    {
        HOSTSPAN("host:decode_tables");
        host_parallel_for(n_blocks, [&](size_t b) {
            const BlkPlan& pl = plan[b];
            const ParsedBlock& pb = pbs[b];
            DecBlock& bk = blocks[b];
            memset(&bk, 0, sizeof(bk));
            bk.blob_off = blob_off[b];
            bk.line0 = (uint32_t)pl.line0; bk.n_lines = pb.bin_lines;
            bk.default_phasing = pb.default_phasing;
            auto moff = [&](uint32_t key) -> uint64_t {
                auto it = pb.dict.find(key);
                if (it == pb.dict.end() || it->second == VAL_UNDEFINED || it->second >= sizes[b]) return ~0ull;
                return blob_off[b] + it->second;
            };
            bk.sparse_off = moff(KEY_MATRIX_SPARSE); bk.miss_off = moff(KEY_MATRIX_MISSING_SPARSE); bk.eov_off = moff(KEY_MATRIX_END_OF_VECTORS_SPARSE);
            // entries each index-list matrix may hold (to the next section or the block end): the list walk is bounded by it
            auto mend = [&](uint32_t key) -> uint64_t {
                auto it = pb.dict.find(key);
                if (it == pb.dict.end() || it->second == VAL_UNDEFINED || it->second >= sizes[b]) return 0;
                return (section_end(pb, it->second, sizes[b]) - it->second) / d.aet;
            };
            bk.sp_end = mend(KEY_MATRIX_SPARSE); bk.ms_end = mend(KEY_MATRIX_MISSING_SPARSE); bk.ev_end = mend(KEY_MATRIX_END_OF_VECTORS_SPARSE);
            bk.wah0 = pl.wah0; bk.sp0 = pl.sp0; bk.ms0 = pl.ms0; bk.ev0 = pl.ev0;
            bk.n_wah = pl.n_wah; bk.n_sp = pl.n_sp; bk.n_ms = pl.n_ms; bk.n_ev = pl.n_ev;
            bk.wah_done = 0;
            bk.wah_todo = pl.n_wah;
            if (lazy) {  // WAH lines among the first lazy_lines binary lines
                const auto& wl = d.h_wah_lines[b];
                bk.wah_todo = (uint32_t)(std::lower_bound(wl.begin(), wl.end(), (uint16_t)std::min<uint32_t>(lazy_lines, 65535u)) - wl.begin());
            }
            d.h_wah_done[b] = bk.wah_todo;
            uint32_t jw = pl.wah0, jp = pl.ph0, sp = pl.sp0, ms = pl.ms0, ev = pl.ev0, gc = 0, gcp = 0;
            uint32_t jm = pl.msw0, je = pl.evw0, gcm = 0, gce = 0;
            const bool weird_wah = pb.ws == WS_WAH;
            for (uint32_t l = 0; l < pb.bin_lines; ++l) {
                const uint64_t gl = pl.line0 + l;
                uint8_t f = 0;
                const bool hap = pb.haploid[l] != 0;
                const uint32_t nbits = hap ? S : N;
                if (hap) f |= DL_HAPLOID;
                if (pb.is_wah[l]) {
                    f |= DL_WAH;
                    dl_ord[gl] = jw;
                    job_gcum[jw] = gc; job_nbits[jw] = nbits; job_seg[jw] = pl.seg_w; job_hap[jw] = hap ? 1 : 0;
                    gc += (nbits + 14) / 15;
                    ++jw;
                } else {
                    dl_ord[gl] = sp++;
                }
                dl_mord[gl] = 0; dl_eord[gl] = 0; dl_pord[gl] = 0;
                if (pb.has_missing[l] && weird_wah) {
                    f |= DL_MISSING | DL_WEIRD_WAH;
                    dl_mord[gl] = jm;
                    job_gcum[jm] = gcm; job_nbits[jm] = nbits; job_seg[jm] = pl.seg_m; job_hap[jm] = 0;
                    gcm += (nbits + 14) / 15;
                    ++jm;
                } else if (pb.has_missing[l]) { f |= DL_MISSING; dl_mord[gl] = ms++; }
                if (pb.has_eov[l] && weird_wah) {
                    f |= DL_EOV | DL_WEIRD_WAH;
                    dl_eord[gl] = je;
                    job_gcum[je] = gce; job_nbits[je] = nbits; job_seg[je] = pl.seg_e; job_hap[je] = 0;
                    gce += (nbits + 14) / 15;
                    ++je;
                } else if (pb.has_eov[l]) { f |= DL_EOV; dl_eord[gl] = ev++; }
                if (pb.has_phase[l]) {
                    f |= DL_PHASE;
                    dl_pord[gl] = jp;
                    job_gcum[jp] = gcp; job_nbits[jp] = nbits; job_seg[jp] = pl.seg_p; job_hap[jp] = 0;
                    gcp += (nbits + 14) / 15;
                    ++jp;
                }
                dl_flags[gl] = f;
            }
            auto segment = [&](uint32_t si, uint32_t val, uint32_t words, uint32_t job0, uint32_t nj, uint32_t tile0) {
                DecSeg& sg = segs[si];
                sg.byte_off = blob_off[b] + val; sg.n_words = words; sg.job0 = job0; sg.njobs = nj; sg.tile0 = tile0;
                for (uint32_t w = 0, t = tile0; w < words; w += D0_TILE, ++t) { tile_seg[t] = si; tile_word0[t] = w; }
            };
            if (pl.n_wah) segment(pl.seg_w, pl.wah_val, pl.wah_words, pl.wah0, pl.n_wah, pl.tile_w);
            if (pl.n_ph) segment(pl.seg_p, pl.ph_val, pl.ph_words, pl.ph0, pl.n_ph, pl.tile_p);
            if (pl.n_msw) segment(pl.seg_m, pl.msw_val, pl.msw_words, pl.msw0, pl.n_msw, pl.tile_m);
            if (pl.n_evw) segment(pl.seg_e, pl.evw_val, pl.evw_words, pl.evw0, pl.n_evw, pl.tile_e);
        });
    }
    for (uint32_t b = 0; b < n_blocks; ++b) {
        const DecBlock& bk = blocks[b];
        if ((bk.n_sp && bk.sparse_off == ~0ull) || (bk.n_ms && bk.miss_off == ~0ull) || (bk.n_ev && bk.eov_off == ~0ull)) { ctx->err = "GT block: index lists without their matrix"; return XSI_E_FORMAT; }
    }
    d.h_blocks.assign(blocks, blocks + n_blocks);
    d.h_bin_lines.resize(n_blocks);
    d.h_bcf_lines.resize(n_blocks);
    for (uint32_t b = 0; b < n_blocks; ++b) { d.h_bin_lines[b] = pbs[b].bin_lines; d.h_bcf_lines[b] = pbs[b].bcf_lines; }

    // ---- upload (from the pinned staging area: asynchronous) ----
    // meta: blocks | segs
    const size_t m_blk = 0, m_seg = m_blk + sizeof(DecBlock) * n_blocks, m_end = m_seg + sizeof(DecSeg) * (nseg ? nseg : 1);
    d.m_blk = m_blk;
    CK(d.meta.ensure(m_end));
    CK(cudaMemcpyAsync(d.meta.as<uint8_t>() + m_blk, blocks, sizeof(DecBlock) * n_blocks, cudaMemcpyHostToDevice, ctx->stream));
    if (nseg) CK(cudaMemcpyAsync(d.meta.as<uint8_t>() + m_seg, segs, sizeof(DecSeg) * nseg, cudaMemcpyHostToDevice, ctx->stream));
    // job arrays: gcum | nbits | seg | word0 | ones
    CK(d.job_u32.ensure((size_t)NJp * 5 * 4));
    uint32_t* ju = d.job_u32.as<uint32_t>();
    if (njobs) {
        CK(cudaMemcpyAsync(ju, job_gcum, (size_t)NJp * 3 * 4, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaMemsetAsync(ju + 3 * NJp, 0xFF, (size_t)njobs * 4, ctx->stream));
    }
    CK(d.job_hap.ensure(NJp));
    if (njobs) CK(cudaMemcpyAsync(d.job_hap.p, job_hap, njobs, cudaMemcpyHostToDevice, ctx->stream));
    CK(d.tile_u32.ensure((size_t)ntp * 3 * 4));
    uint32_t* tu = d.tile_u32.as<uint32_t>();
    if (ntiles) CK(cudaMemcpyAsync(tu, tile_seg, (size_t)ntp * 2 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(d.dline.ensure(Ltp * 17));
    uint8_t* dlb = d.dline.as<uint8_t>();
    if (Lt) CK(cudaMemcpyAsync(dlb, dl_ord, Ltp * 17, cudaMemcpyHostToDevice, ctx->stream));
    CK(d.lists.ensure(((size_t)nsp + nms + nev + 3) * 8));
    CK(d.err.ensure(64));
    CK(cudaMemsetAsync(d.err.p, 0, 64, ctx->stream));
    CK(d.rows.ensure((size_t)NJp * d.WS * 4));
    CK(d.seg_total.ensure((size_t)(nseg ? nseg : 1) * 4));

    DecDev& dd = d.dev;
    dd.blob = d.blob.as<uint8_t>();
    dd.segs = reinterpret_cast<const DecSeg*>(d.meta.as<uint8_t>() + m_seg); dd.nseg = nseg;
    dd.blocks = reinterpret_cast<const DecBlock*>(d.meta.as<uint8_t>() + m_blk); dd.nb = n_blocks;
    dd.tile_seg = tu; dd.tile_word0 = tu + ntp; dd.tile_sum = tu + 2 * ntp; dd.ntiles = ntiles;
    dd.job_gcum = ju; dd.job_nbits = ju + NJp; dd.job_seg = ju + 2 * NJp; dd.job_word0 = ju + 3 * NJp; dd.job_ones = ju + 4 * NJp;
    dd.NJ = njobs; dd.rows = d.rows.as<uint32_t>(); dd.WS = d.WS; dd.n_samples = S; dd.aet = d.aet;
    dd.dline_ord = reinterpret_cast<const uint32_t*>(dlb); dd.dline_mord = reinterpret_cast<const uint32_t*>(dlb + Ltp * 4);
    dd.dline_eord = reinterpret_cast<const uint32_t*>(dlb + Ltp * 8); dd.dline_pord = reinterpret_cast<const uint32_t*>(dlb + Ltp * 12);
    dd.dline_flags = dlb + Ltp * 16;
    dd.sp_off = d.lists.as<uint64_t>(); dd.ms_off = dd.sp_off + nsp + 1; dd.ev_off = dd.ms_off + nms + 1;
    dd.err = d.err.as<uint32_t>();
    // D2 v2 (barrier-free, sliced over haplotypes) needs the per-line tables; all-haploid lines use v1
    const uint32_t TWv2 = 2 * d.WS + 4;
    // XSI_PBWT_V=1 forces the general kernels (a[] in shared / global memory), which also serve blocks with haploid lines
    const bool cluster_ok = d.n_gt_jobs && !any_hap_job && !(getenv("XSI_PBWT_V") && atoi(getenv("XSI_PBWT_V")) == 1);
    const bool use_v2 = cluster_ok && N <= 65534;
    // D2 v3: positions in registers, tables through a TMA ring
    uint32_t v3_kh = 0, v3_nc = 0, v3_slices = 0;
    const size_t v3_smem = (size_t)D3_STAGES * TWv2 * 4 + 2 * D3_STAGES * 8;
    if (use_v2 && v3_smem <= ctx->smem_optin) {
        // haplotypes per thread: the largest of 32/16/8 that still gives every SM ~768 consumer threads
        const uint64_t want = (uint64_t)ctx->sm_count * 768;
        v3_kh = 8;
        for (uint32_t kh : {32u, 16u}) if ((uint64_t)n_blocks * ((N + kh - 1) / kh) >= want) { v3_kh = kh; break; }
        if (const char* sw = getenv("XSI_UNPERM_KH")) { const int v = atoi(sw); if (v == 8 || v == 16 || v == 32) v3_kh = (uint32_t)v; }
        const uint32_t thr_total = ((N + v3_kh - 1) / v3_kh + 31) / 32 * 32;
        v3_nc = std::min<uint32_t>(512, thr_total);
        // short rows (1KGP3 / chrX shapes: 5,008 haplotypes = 640 threads at KH = 8): four CTAs of a quarter of the row each
        // instead of one wide and one nearly empty CTA per block (r02o: 6.8 -> 5.7 ms at 220 blocks, 2.3 -> 1.3 ms at 24)
        if (v3_kh == 8 && thr_total <= 1024) v3_nc = std::max<uint32_t>(128, (thr_total / 4 + 31) / 32 * 32);
        if (const char* sw = getenv("XSI_UNPERM_NC")) { const int v = atoi(sw); if (v >= 32 && v <= 512 && v % 32 == 0) v3_nc = std::min<uint32_t>(thr_total, (uint32_t)v); }
        v3_slices = (thr_total + v3_nc - 1) / v3_nc;
    }
    const bool v2_ok = v3_kh != 0;
    // D2 wide: more than 65,534 haplotypes (tables stay in global memory / L2)
    const bool use_wide = cluster_ok && N > 65534;
    uint32_t wide_kh = 0, wide_slices = 0;
    if (use_wide) {
        // haplotypes per thread: the largest of 32/16/8 that still gives every SM ~1024 threads
        const uint64_t want = (uint64_t)ctx->sm_count * 1024;
        wide_kh = 8;
        for (uint32_t kh : {32u, 16u}) if ((uint64_t)n_blocks * ((N + kh - 1) / kh) >= want) { wide_kh = kh; break; }
        if (const char* sw = getenv("XSI_UNPERM_KH")) { const int v = atoi(sw); if (v == 8 || v == 16 || v == 32) wide_kh = (uint32_t)v; }
        wide_slices = (uint32_t)(((uint64_t)d.WS * 32 + 256ull * wide_kh - 1) / (256ull * wide_kh));
    }
    dd.tabs = nullptr; dd.TW = TWv2; dd.n_gt_jobs = d.n_gt_jobs; dd.tab_inv = v3_kh ? 0xFFFFu : 0u; dd.tab_wide = use_wide ? 1u : 0u;
    if (v2_ok || use_wide) {
        CK(d.tabs.ensure((size_t)d.n_gt_jobs * TWv2 * 4));
        dd.tabs = d.tabs.as<uint32_t>();
    }

    // ---- kernels ----
    if (getenv("XSI_DEBUG_UPLOAD_SYNC")) CK(cudaStreamSynchronize(ctx->stream));
    const bool side = !getenv("XSI_NO_SIDE_STREAM");
    if ((nsp + nms + nev) && !side) {
        sparse_index_kernel<<<(n_blocks * 3 + 63) / 64, 64, 0, ctx->stream>>>(dd);
        CKL();
    }
    if ((nsp + nms + nev) && side) {
        // latency-bound list walk: runs beside the WAH pipeline on the side stream, joined before the error read
        CK(cudaEventRecord(ctx->ev_side, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_side, 0));
        sparse_index_kernel<<<(n_blocks * 3 + 63) / 64, 64, 0, ctx->stream2>>>(dd);
        CKL();
        CK(cudaEventRecord(ctx->ev_side, ctx->stream2));
    }
    if (njobs) {
        { PROF("wah_tile_sum"); wah_tile_sum_kernel<<<ntiles, D0_THREADS, 0, ctx->stream>>>(dd); }
        CKL();
        { PROF("wah_tile_base"); wah_tile_base_kernel<<<(nseg + 3) / 4, 128, 0, ctx->stream>>>(dd, d.seg_total.as<uint32_t>()); }
        CKL();
        { PROF("wah_find_lines"); wah_find_lines_kernel<<<ntiles, D0_THREADS, 0, ctx->stream>>>(dd); }
        CKL();
        if (use_wide) {
            // long lines: one CTA per line, every warp expands one segment of the row (a multiple of 15 words)
            const uint32_t SEGW = ((d.WS + D1W_WARPS - 1) / D1W_WARPS + 14) / 15 * 15, SEGG = SEGW * 32 / 15, SEGT = (SEGG + 1 + 31) / 32 + 1;
            const size_t per_warp_w = ((size_t)(SEGG + 8) * 2 + 3) / 4 * 4 + (size_t)SEGT * 4;
            const size_t smem_w = per_warp_w * D1W_WARPS;
            if (smem_w > ctx->smem_optin) { ctx->err = "haplotype count too large for the WAH expand kernel"; return XSI_E_UNSUPPORTED; }
            CK(cudaFuncSetAttribute(wah_expand_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
            { PROF("wah_expand"); wah_expand_wide_kernel<<<njobs, D1W_WARPS * 32, smem_w, ctx->stream>>>(dd, SEGW, SEGT); }
            CKL();
        } else {
        const uint32_t G = (N + 14) / 15;
        const uint32_t Gpad = (G + 4 + 7) / 8 * 8, Tpad = (G + 1 + 31) / 32 + 1;
        const size_t per_warp = (size_t)Gpad * 2 + (size_t)Tpad * 4;
        uint32_t wpc = (uint32_t)std::min<size_t>(4, std::max<size_t>(1, (ctx->smem_optin - 1024) / per_warp));
        if (per_warp > ctx->smem_optin) { ctx->err = "haplotype count too large for the WAH expand kernel"; return XSI_E_UNSUPPORTED; }
        const size_t smem = per_warp * wpc;
        CK(cudaFuncSetAttribute(wah_expand_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        { PROF("wah_expand"); wah_expand_kernel<<<(njobs + wpc - 1) / wpc, wpc * 32, smem, ctx->stream>>>(dd, wpc, Gpad, Tpad); }
        CKL();
        }
        if (v3_kh) {
            d.v3_kh = v3_kh; d.v3_nc = v3_nc; d.v3_slices = v3_slices; d.v3_smem = v3_smem;
            d.ps_stride = v3_slices * v3_nc * v3_kh;
            CK(d.pos_state.ensure((size_t)n_blocks * d.ps_stride * 2));
            int rc = launch_unpermute_v3(ctx, dd, 0, n_blocks);
            if (rc) return rc;
        } else if (use_wide) {
            const dim3 g(n_blocks, wide_slices);
            uint32_t max_nwah = 0;
            for (uint32_t b = 0; b < n_blocks; ++b) max_nwah = std::max(max_nwah, plan[b].n_wah);
            // lines per launch: the tables of a window (all blocks) stay inside ~half of L2
            uint64_t win = (48ull << 20) / ((uint64_t)n_blocks * TWv2 * 4);
            win = std::max<uint64_t>(8, std::min<uint64_t>(512, win));
            if (const char* sw = getenv("XSI_UNPERM_WINDOW")) { const int v = atoi(sw); if (v >= 1) win = (uint64_t)v; }
            CK(d.x_pool.ensure((size_t)n_blocks * wide_kh * wide_slices * 256 * 4));
            uint32_t* ps = d.x_pool.as<uint32_t>();
            PROF("pbwt_unpermute");
            for (uint32_t k0 = 0; k0 < max_nwah; k0 += (uint32_t)win) {
                const uint32_t k1 = (uint32_t)std::min<uint64_t>(max_nwah, k0 + win);
                if (wide_kh == 32) pbwt_unpermute_wide_kernel<32><<<g, 256, 0, ctx->stream>>>(dd, k0, k1, ps);
                else if (wide_kh == 16) pbwt_unpermute_wide_kernel<16><<<g, 256, 0, ctx->stream>>>(dd, k0, k1, ps);
                else pbwt_unpermute_wide_kernel<8><<<g, 256, 0, ctx->stream>>>(dd, k0, k1, ps);
                CKL();
            }
        } else if (d.n_gt_jobs) {
            const uint32_t W = (N + 31) / 32;
            const size_t smem2 = ((size_t)N * 2 + 15) / 16 * 16 + (size_t)2 * d.WS * 4 + ((size_t)N + 32 + 15) / 16 * 16 + 64 * 4 + 16;
            if (N <= 65536 && smem2 <= ctx->smem_optin) {
                const int wpw = choose_wpw(W);
                const uint32_t NW = (W + wpw - 1) / wpw;
                const uint8_t* jh = d.job_hap.as<uint8_t>();
                int rc;
                switch (wpw) {
                    case 2: rc = launch_unpermute<2, 1024>(ctx, dd, jh, NW, smem2); break;
                    case 4: rc = launch_unpermute<4, 1024>(ctx, dd, jh, NW, smem2); break;
                    case 8: rc = launch_unpermute<8, 1024>(ctx, dd, jh, NW, smem2); break;
                    case 16: rc = launch_unpermute<16, 1024>(ctx, dd, jh, NW, smem2); break;
                    case 32: rc = launch_unpermute<32, 1024>(ctx, dd, jh, NW, smem2); break;
                    case 64: rc = launch_unpermute<64, 1024>(ctx, dd, jh, NW, smem2); break;
                    default: rc = launch_unpermute<128, 512>(ctx, dd, jh, NW, smem2); break;
                }
                if (rc) return rc;
            } else {
                CK(d.a_pool.ensure(((size_t)n_blocks * 2 * N + (size_t)n_blocks * d.WS) * 4));
                CK(d.x_pool.ensure((size_t)n_blocks * (N + 32)));
                { PROF("pbwt_unpermute"); pbwt_unpermute_gmem_kernel<<<n_blocks, 1024, 0, ctx->stream>>>(dd, d.job_hap.as<uint8_t>(), d.a_pool.as<uint32_t>(), d.x_pool.as<uint8_t>()); }
                CKL();
            }
        }
    }
    if ((nsp + nms + nev) && side) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
    uint32_t herr = 0;
    CK(cudaMemcpyAsync(&herr, d.err.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (herr) { ctx->err = "malformed WAH / index stream in GT block (device check)"; return XSI_E_FORMAT; }
    d.loaded = true;
    return XSI_OK;
}

extern "C" int xsi_decode_block_info(const xsi_ctx* ctx, uint32_t block_index, uint32_t* bcf_lines, uint32_t* binary_lines) {
    if (!ctx || !ctx->dec.loaded || block_index >= ctx->dec.nb) return XSI_E_ARG;
    if (bcf_lines) *bcf_lines = ctx->dec.h_bcf_lines[block_index];
    if (binary_lines) *binary_lines = ctx->dec.h_bin_lines[block_index];
    return XSI_OK;
}

template <typename OT>
static int decode_records_impl(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                               const uint32_t* n_alleles, OT* out, uint64_t out_stride, int32_t out_on_device,
                               uint32_t* n_filled, uint64_t* allele_counts, uint32_t counts_stride) {
    if (!ctx || !block_index || !line_offset || !n_alleles || !out) return XSI_E_ARG;
    auto& d = ctx->dec;
    if (!d.loaded) { ctx->err = "xsi_decode_records without loaded blocks"; return XSI_E_ARG; }
    if (n == 0) return XSI_OK;
    CK(cudaSetDevice(ctx->device));
    const uint32_t N = 2 * d.n_samples;
    if (out_stride < N) { ctx->err = "out_stride smaller than 2*num_samples"; return XSI_E_ARG; }
    uint32_t max_all = 2;
    {   // argument checks, in chunks on the worker pool (millions of records per call on short-row files)
        constexpr uint64_t VCH = 1u << 16;
        const size_t nvc = (size_t)((n + VCH - 1) / VCH);
        struct Chk { int bad = 0; uint32_t max_all = 2; };
        std::vector<Chk> chk(nvc);
        host_parallel_for(nvc, [&](size_t c) {
            Chk& k = chk[c];
            const uint64_t i1 = std::min<uint64_t>(n, (c + 1) * VCH);
            for (uint64_t i = c * VCH; i < i1; ++i) {
                if (block_index[i] >= d.nb) { k.bad = 1; return; }
                if (n_alleles[i] < 2 || n_alleles[i] > 254) { k.bad = 2; return; }
                // BCF keeps FORMAT/GT as int8 only while (allele+1)<<1|1 <= 127 (htslib vcf.c bcf_update_format: wider types above)
                if (sizeof(OT) == 1 && n_alleles[i] > 63) { k.bad = 3; return; }
                if ((uint64_t)line_offset[i] + n_alleles[i] - 1 > d.h_bin_lines[block_index[i]]) { k.bad = 4; return; }
                k.max_all = std::max(k.max_all, n_alleles[i]);
            }
        });
        for (const Chk& k : chk) {
            if (k.bad == 1) { ctx->err = "block index out of range"; return XSI_E_ARG; }
            if (k.bad == 2) { ctx->err = "n_alleles out of range (2..254)"; return XSI_E_UNSUPPORTED; }
            if (k.bad == 3) { ctx->err = "int8 genotype rows need n_alleles <= 63"; return XSI_E_UNSUPPORTED; }
            if (k.bad == 4) { ctx->err = "record runs past the end of its block"; return XSI_E_ARG; }
            max_all = std::max(max_all, k.max_all);
        }
    }
    const bool want_counts = allele_counts != nullptr;
    if (want_counts && counts_stride < max_all) { ctx->err = "counts_stride too small"; return XSI_E_ARG; }
    { const int rc_ = ensure_lines(ctx, n, block_index, line_offset, n_alleles); if (rc_) return rc_; }
    const uint32_t Npad = (N + 63) / 64 * 64;
    // uploads the requests of records [c0, c0+cn) and composes their rows (element type DT) at dev_out, stride in elements
    auto compose_chunk = [&](auto* dev_out, uint64_t stride, uint64_t c0, uint64_t cn, ReqDev& q) -> int {
        using DT = std::remove_pointer_t<decltype(dev_out)>;
        if (ctx->async_encode && ctx->overlap_order && ctx->enc_pending && ctx->ev_scan) {
            // ordered overlap (see xsi_ctx::overlap_order): compose after the scan of the batch that is being encoded beside us
            const uint64_t want = ctx->enc_seq.load();
            for (int spin = 0; ctx->scan_seq.load() < want && spin < 100000; ++spin) std::this_thread::sleep_for(std::chrono::microseconds(20));
            if (ctx->scan_seq.load() >= want) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_scan, 0));
        }
        CK(d.req.ensure(cn * 4 * 4));
        uint32_t* rq = d.req.as<uint32_t>();
        if (cn >= (1u << 16)) {
            // the three request arrays are the caller's pageable memory: packed into one pinned buffer by the worker pool and sent
            // as ONE asynchronous copy (three staged pageable copies of 7 MB each cost several ms per call at 1.8 M records)
            if (!d.ev_req) CK(cudaEventCreateWithFlags(&d.ev_req, cudaEventDisableTiming));
            else CK(cudaEventSynchronize(d.ev_req));  // the previous copy out of h_req has finished
            CK(d.h_req.ensure(cn * 3 * 4));
            uint32_t* hq = d.h_req.as<uint32_t>();
            constexpr uint64_t PCH = 1u << 17;
            host_parallel_for((size_t)((cn + PCH - 1) / PCH), [&](size_t c) {
                const uint64_t i0 = c * PCH, m = std::min<uint64_t>(PCH, cn - i0);
                memcpy(hq + i0, block_index + c0 + i0, m * 4);
                memcpy(hq + cn + i0, line_offset + c0 + i0, m * 4);
                memcpy(hq + 2 * cn + i0, n_alleles + c0 + i0, m * 4);
            });
            CK(cudaMemcpyAsync(rq, hq, cn * 3 * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaEventRecord(d.ev_req, ctx->stream));
        } else {
            CK(cudaMemcpyAsync(rq, block_index + c0, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(rq + cn, line_offset + c0, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(rq + 2 * cn, n_alleles + c0, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        const uint32_t grid = (uint32_t)std::min<uint64_t>(cn, (uint64_t)ctx->sm_count * 8);
        CK(d.scratch.ensure((size_t)grid * 2 * Npad));
        q.blk = rq; q.line = rq + cn; q.nall = rq + 2 * cn; q.n = (uint32_t)cn;
        q.filled = rq + 3 * cn;
        q.out_stride = stride;
        q.out = dev_out;
        q.counts = nullptr; q.counts_stride = counts_stride;
        if (want_counts) { CK(d.counts.ensure(cn * counts_stride * 4)); q.counts = d.counts.as<uint32_t>(); }
        q.scratch = d.scratch.as<uint8_t>(); q.Npad = Npad;
        // rows that start and end on 16-byte boundaries take the TMA-store fast path for simple records
        // (records whose own length is not, e.g. all-haploid rows of an odd sample count, stay with compose_records)
        const bool fast = !getenv("XSI_COMPOSE_V1") && (reinterpret_cast<uintptr_t>(q.out) % 16 == 0) && ((stride * sizeof(DT)) % 16 == 0);
        if (fast) {
            const uint32_t words = (N + 31) / 32;
            uint32_t nt = 256;
            if (const char* s_ = getenv("XSI_COMPOSE_NT")) nt = (uint32_t)atoi(s_);
            else if (words <= 160) nt = 160;
            else if (words <= 192) nt = 192;
            if (nt != 160 && nt != 192) nt = 256;
            const size_t smem = (size_t)2 * nt * 32 * sizeof(DT);
            const uint32_t g2 = (uint32_t)std::min<uint64_t>(cn, (uint64_t)ctx->sm_count * d5_ctas_per_sm((int)nt, (int)sizeof(DT)));
            auto go = [&](auto kern) -> cudaError_t {
                cudaError_t e_ = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                if (e_ != cudaSuccess) return e_;
                kern<<<g2, nt, smem, ctx->stream>>>(d.dev, q);
                return cudaSuccess;
            };
            {
                PROF("compose_simple");
                if (nt == 160) CK(go(compose_simple_kernel<DT, 160>));
                else if (nt == 192) CK(go(compose_simple_kernel<DT, 192>));
                else CK(go(compose_simple_kernel<DT, 256>));
            }
            CKL();
        }
        const size_t rec_smem = (size_t)2 * Npad <= (48u << 10) ? (size_t)2 * Npad : 0;
        { PROF("compose_records"); compose_records_kernel<DT><<<grid, D4_THREADS, rec_smem, ctx->stream>>>(d.dev, q, fast ? 1 : 0, rec_smem ? 1 : 0); }
        CKL();
        return XSI_OK;
    };
    // counts of chunk [c0, c0+cn) to the caller (synchronises the stream)
    auto fetch_counts = [&](const ReqDev& q, uint64_t c0, uint64_t cn) -> int {
        CK(d.h_stage.ensure(cn * counts_stride * 4));
        CK(cudaMemcpyAsync(d.h_stage.p, q.counts, cn * counts_stride * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        const uint32_t* hc = d.h_stage.as<uint32_t>();
        for (uint64_t i = 0; i < cn; ++i)
            for (uint32_t k = 0; k < n_alleles[c0 + i]; ++k) allele_counts[(c0 + i) * counts_stride + k] = hc[i * counts_stride + k];
        return XSI_OK;
    };

    // ---- host int32 rows: composed and moved as BCF int8, widened on the host beside the DMA (host_narrow.cpp) ----
    const uint64_t stride8 = ((uint64_t)N + 15) / 16 * 16;
    if (sizeof(OT) == 4 && !out_on_device && max_all <= 63 && stride8 <= xsi_ctx::RING_BYTES && host_narrow_on()) {
        int rc = ring_prepare(ctx);
        if (rc) return rc;
        const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(n, (1ull << 30) / stride8));
        const uint64_t rps = xsi_ctx::RING_BYTES / stride8;  // rows per ring slot
        std::vector<uint32_t> filled;
        for (uint64_t c0 = 0; c0 < n; c0 += chunk) {
            const uint64_t cn = std::min(chunk, n - c0);
            CK(d.out.ensure(cn * stride8));
            ReqDev q;
            if ((rc = compose_chunk(d.out.as<int8_t>(), stride8, c0, cn, q))) return rc;
            filled.resize(cn);
            CK(cudaMemcpyAsync(filled.data(), q.filled, cn * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (want_counts) { if ((rc = fetch_counts(q, c0, cn))) return rc; }
            else CK(cudaStreamSynchronize(ctx->stream));
            if (n_filled) memcpy(n_filled + c0, filled.data(), cn * 4);
            // Sub-chunks of `rps` rows take one of two routes, whichever is free: (a) copied as int8 into a pinned ring
            // slot and widened by the worker pool, (b) when `out` is pinned and the rows are full diploid rows, widened
            // by a device kernel into a staging slot and copied as int32 by the DMA engine alone (side stream).
            const uint64_t nsub = (cn + rps - 1) / rps;
            const uint64_t du = std::max<uint64_t>(1, xsi_ctx::DMA_ELEMS / N);  // rows per unit of the DMA route (16 MB of int32)
            const bool dma = dma_route_on() && stride8 == N && nsub >= 4 && host_ptr_is_pinned(out) && reinterpret_cast<uintptr_t>(out) % 16 == 0 &&
                             (out_stride * 4) % 16 == 0 && (uint64_t)N <= xsi_ctx::DMA_ELEMS;
            if (dma) {
                if ((rc = dma_prepare(ctx))) return rc;
                CK(cudaEventRecord(ctx->ev_side, ctx->stream));          // rows composed (the stream was synchronised above,
                CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_side, 0));  // this only keeps the dependency explicit)
            }
            bool used_dma = false;
            int slot = 0, pend_slot = -1;
            uint64_t pend_r0 = 0, pend_rn = 0, by_host = 0;
            auto widen_pending = [&]() -> int {
                if (pend_slot < 0) return XSI_OK;
                int rr = ring_wait(ctx, pend_slot);
                if (rr) return rr;
                widen_rows_i8_to_i32(ctx->ring.as<int8_t>() + (size_t)pend_slot * xsi_ctx::RING_BYTES, stride8,
                                     reinterpret_cast<int32_t*>(out) + (c0 + pend_r0) * out_stride, out_stride, filled.data() + pend_r0, pend_rn);
                pend_slot = -1;
                return XSI_OK;
            };
            for (uint64_t r0 = 0; r0 < cn;) {
                // top up the DMA route first (no host core involved): units as large in bytes as the ring's int8 copies
                while (dma && cn - r0 >= du) {
                    bool full = true;
                    for (uint64_t i = 0; i < du && full; ++i) full = filled[r0 + i] == N;
                    const int ds = full ? dma_free_slot(ctx) : -1;
                    if (ds < 0) break;
                    int32_t* st = ctx->dma_stage.as<int32_t>() + (size_t)ds * xsi_ctx::DMA_ELEMS;
                    widen_rows_i8_i32_kernel<<<ctx->sm_count * 2, 256, 0, ctx->stream2>>>(d.out.as<int8_t>() + r0 * stride8, stride8, st, N, (uint32_t)du);
                    CKL();
                    int32_t* dsth = reinterpret_cast<int32_t*>(out) + (c0 + r0) * out_stride;
                    if (out_stride == N) CK(cudaMemcpyAsync(dsth, st, du * (size_t)N * 4, cudaMemcpyDeviceToHost, ctx->stream2));
                    else CK(cudaMemcpy2DAsync(dsth, out_stride * 4, st, (size_t)N * 4, (size_t)N * 4, du, cudaMemcpyDeviceToHost, ctx->stream2));
                    CK(cudaEventRecord(ctx->dma_ev[ds], ctx->stream2));
                    ctx->dma_busy[ds] = true;
                    ctx->dma_d2h += du * (uint64_t)N * 4;
                    used_dma = true;
                    r0 += du;
                }
                if (r0 >= cn) break;
                const uint64_t rn = std::min(rps, cn - r0);
                if ((rc = ring_wait(ctx, slot))) return rc;
                CK(cudaMemcpyAsync(ctx->ring.as<int8_t>() + (size_t)slot * xsi_ctx::RING_BYTES, d.out.as<int8_t>() + r0 * stride8,
                                   rn * stride8, cudaMemcpyDeviceToHost, ctx->stream));
                CK(cudaEventRecord(ctx->ring_ev[slot], ctx->stream));
                ctx->ring_busy[slot] = true;
                if ((rc = widen_pending())) return rc;  // the previous sub-chunk of this route, while this one's copy runs
                pend_slot = slot; pend_r0 = r0; pend_rn = rn;
                slot = (slot + 1) % xsi_ctx::RING_SLOTS;
                by_host += rn * stride8;
                r0 += rn;
            }
            if ((rc = widen_pending())) return rc;
            if (used_dma) {
                CK(cudaStreamSynchronize(ctx->stream2));
                for (int i = 0; i < xsi_ctx::DMA_SLOTS; ++i) ctx->dma_busy[i] = false;
            }
            ctx->narrowed_d2h += by_host;
        }
        return XSI_OK;
    }

    // chunk so that the staging buffers stay bounded
    const uint64_t row_bytes = out_stride * sizeof(OT);
    // host output is staged through a bounded device buffer; device output needs no chunking
    const uint64_t chunk = out_on_device ? std::min<uint64_t>(n, 1ull << 30)
                                         : std::max<uint64_t>(1, std::min<uint64_t>(n, (1ull << 30) / row_bytes));
    for (uint64_t c0 = 0; c0 < n; c0 += chunk) {
        const uint64_t cn = std::min(chunk, n - c0);
        ReqDev q;
        OT* dev_out;
        if (out_on_device) dev_out = out + c0 * out_stride;
        else { CK(d.out.ensure(cn * row_bytes)); dev_out = d.out.as<OT>(); }
        int rc = compose_chunk(dev_out, out_stride, c0, cn, q);
        if (rc) return rc;
        if (!out_on_device) CK(cudaMemcpyAsync(out + c0 * out_stride, q.out, cn * row_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if (n_filled) CK(cudaMemcpyAsync(n_filled + c0, q.filled, cn * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (want_counts) {
            if ((rc = fetch_counts(q, c0, cn))) return rc;
        } else if (!out_on_device || n_filled || c0 + chunk < n) {
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    return XSI_OK;
}

static int xsi_decode_records_impl(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                                  const uint32_t* n_alleles, int32_t* out, uint64_t out_stride, int32_t out_on_device,
                                  uint32_t* n_filled, uint64_t* allele_counts, uint32_t counts_stride) {
    return decode_records_impl<int32_t>(ctx, n, block_index, line_offset, n_alleles, out, out_stride, out_on_device, n_filled,
                                        allele_counts, counts_stride);
}

static int xsi_decode_records_subset_impl(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                                         const uint32_t* n_alleles, const uint32_t* samples_to_use, uint32_t n_sel,
                                         int32_t* out, uint64_t out_stride, int32_t out_on_device, uint32_t* n_filled,
                                         uint32_t* ac, uint32_t ac_stride) {
    if (!ctx || !block_index || !line_offset || !n_alleles || !samples_to_use || !out || n_sel == 0) return XSI_E_ARG;
    auto& d = ctx->dec;
    if (!d.loaded) { ctx->err = "xsi_decode_records_subset without loaded blocks"; return XSI_E_ARG; }
    if (n == 0) return XSI_OK;
    CK(cudaSetDevice(ctx->device));
    const uint32_t N = 2 * d.n_samples;
    if (out_stride < 2ull * n_sel) { ctx->err = "out_stride smaller than 2*n_sel"; return XSI_E_ARG; }
    uint32_t max_all = 2;
    for (uint64_t i = 0; i < n; ++i) max_all = std::max(max_all, n_alleles[i]);
    if (ac && ac_stride + 1 < max_all) { ctx->err = "ac_stride too small"; return XSI_E_ARG; }
    for (uint32_t i = 0; i < n_sel; ++i)
        if (samples_to_use[i] >= d.n_samples) { ctx->err = "sample index out of range"; return XSI_E_ARG; }
    // full rows of a chunk of records on the device (the ordinary decode), then the gather
    const uint64_t full_stride = ((uint64_t)N + 3) / 4 * 4;
    const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(n, (1ull << 30) / (full_stride * 4)));
    DevBuf& full = d.a_pool;      // scratch pools that are idle between xsi_decode_load_blocks calls
    DevBuf& aux = d.x_pool;
    CK(full.ensure(chunk * full_stride * 4));
    // aux: sel[n_sel] | filled_full[chunk] | nall[chunk] | filled_sub[chunk] | ac[chunk*ac_stride] | sub rows (host output only)
    const size_t a_sel = 0, a_ff = a_sel + (size_t)n_sel * 4, a_na = a_ff + chunk * 4, a_fs = a_na + chunk * 4, a_ac = a_fs + chunk * 4,
                 a_rows = (a_ac + (ac ? chunk * ac_stride * 4 : 0) + 15) / 16 * 16,
                 a_end = a_rows + (out_on_device ? 0 : chunk * out_stride * 4);
    CK(aux.ensure(a_end));
    uint8_t* ab = aux.as<uint8_t>();
    CK(cudaMemcpyAsync(ab + a_sel, samples_to_use, (size_t)n_sel * 4, cudaMemcpyHostToDevice, ctx->stream));
    for (uint64_t c0 = 0; c0 < n; c0 += chunk) {
        const uint64_t cn = std::min(chunk, n - c0);
        int rc = xsi_decode_records(ctx, cn, block_index + c0, line_offset + c0, n_alleles + c0, full.as<int32_t>(), full_stride, 1,
                                    nullptr, nullptr, 0);
        if (rc) return rc;
        // row lengths of the full rows (CURRENT_N_HAPS) were left in the request buffer by the compose kernels
        const uint32_t* filled_full = d.req.as<uint32_t>() + 3 * cn;
        CK(cudaMemcpyAsync(ab + a_na, n_alleles + c0, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
        int32_t* dst = out_on_device ? out + c0 * out_stride : reinterpret_cast<int32_t*>(ab + a_rows);
        uint32_t* dac = ac ? reinterpret_cast<uint32_t*>(ab + a_ac) : nullptr;
        if (dac) CK(cudaMemsetAsync(dac, 0, cn * ac_stride * 4, ctx->stream));
        if (!out_on_device) CK(cudaMemsetAsync(dst, 0, cn * out_stride * 4, ctx->stream));
        { PROF("select_samples");
          select_samples_kernel<<<(uint32_t)std::min<uint64_t>(cn, (uint64_t)ctx->sm_count * 8), 256, 0, ctx->stream>>>(
              full.as<int32_t>(), full_stride, filled_full, reinterpret_cast<const uint32_t*>(ab + a_na), d.n_samples,
              reinterpret_cast<const uint32_t*>(ab + a_sel), n_sel, dst, out_stride, reinterpret_cast<uint32_t*>(ab + a_fs), dac, ac_stride, (uint32_t)cn); }
        CKL();
        if (!out_on_device) CK(cudaMemcpyAsync(out + c0 * out_stride, dst, cn * out_stride * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (n_filled) CK(cudaMemcpyAsync(n_filled + c0, ab + a_fs, cn * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (ac) CK(cudaMemcpyAsync(ac + c0 * ac_stride, dac, cn * ac_stride * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return XSI_OK;
}

static int xsi_decode_allele_counts_impl(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                                        const uint32_t* n_alleles, uint64_t* allele_counts, uint32_t counts_stride) {
    if (!ctx || !block_index || !line_offset || !n_alleles || !allele_counts) return XSI_E_ARG;
    auto& d = ctx->dec;
    if (!d.loaded) { ctx->err = "xsi_decode_allele_counts without loaded blocks"; return XSI_E_ARG; }
    if (n == 0) return XSI_OK;
    if (n >= (1ull << 31)) { ctx->err = "too many records in one call"; return XSI_E_ARG; }
    CK(cudaSetDevice(ctx->device));
    uint32_t max_all = 2;
    for (uint64_t i = 0; i < n; ++i) {
        if (block_index[i] >= d.nb) { ctx->err = "block index out of range"; return XSI_E_ARG; }
        if (n_alleles[i] < 2 || n_alleles[i] > 254) { ctx->err = "n_alleles out of range (2..254)"; return XSI_E_UNSUPPORTED; }
        if ((uint64_t)line_offset[i] + n_alleles[i] - 1 > d.h_bin_lines[block_index[i]]) { ctx->err = "record runs past the end of its block"; return XSI_E_ARG; }
        max_all = std::max(max_all, n_alleles[i]);
    }
    if (counts_stride < max_all) { ctx->err = "counts_stride too small"; return XSI_E_ARG; }
    CK(d.req.ensure(n * 4 * 4));
    uint32_t* rq = d.req.as<uint32_t>();
    CK(cudaMemcpyAsync(rq, block_index, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(rq + n, line_offset, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(rq + 2 * n, n_alleles, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ReqDev q = {};
    q.blk = rq; q.line = rq + n; q.nall = rq + 2 * n; q.n = (uint32_t)n;
    CK(d.counts.ensure(n * counts_stride * 4));
    q.counts = d.counts.as<uint32_t>(); q.counts_stride = counts_stride;
    { PROF("allele_counts"); allele_counts_kernel<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->stream>>>(d.dev, q); }
    CKL();
    CK(d.h_stage.ensure(n * counts_stride * 4));
    CK(cudaMemcpyAsync(d.h_stage.p, q.counts, n * counts_stride * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const uint32_t* hc = d.h_stage.as<uint32_t>();
    for (uint64_t i = 0; i < n; ++i)
        for (uint32_t k = 0; k < n_alleles[i]; ++k) allele_counts[i * counts_stride + k] = hc[i * counts_stride + k];
    return XSI_OK;
}


static int xsi_decode_dot_products_impl(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                                        const uint32_t* n_alleles, const double* y, int32_t y_on_device, double* out, uint32_t out_stride) {
    if (!ctx || !block_index || !line_offset || !n_alleles || !y || !out) return XSI_E_ARG;
    auto& d = ctx->dec;
    if (!d.loaded) { ctx->err = "xsi_decode_dot_products without loaded blocks"; return XSI_E_ARG; }
    if (n == 0) return XSI_OK;
    if (n >= (1ull << 31)) { ctx->err = "too many records in one call"; return XSI_E_ARG; }
    CK(cudaSetDevice(ctx->device));
    uint32_t max_all = 2;
    for (uint64_t i = 0; i < n; ++i) {
        if (block_index[i] >= d.nb) { ctx->err = "block index out of range"; return XSI_E_ARG; }
        if (n_alleles[i] < 2 || n_alleles[i] > 63) { ctx->err = "n_alleles out of range (2..63)"; return XSI_E_UNSUPPORTED; }
        if ((uint64_t)line_offset[i] + n_alleles[i] - 1 > d.h_bin_lines[block_index[i]]) { ctx->err = "record runs past the end of its block"; return XSI_E_ARG; }
        max_all = std::max(max_all, n_alleles[i]);
    }
    if (out_stride + 1 < max_all) { ctx->err = "out_stride too small"; return XSI_E_ARG; }
    { const int rc_ = ensure_lines(ctx, n, block_index, line_offset, n_alleles); if (rc_) return rc_; }
    const uint32_t S = d.n_samples;
    // scratch: y[S] f64 | out[n*stride] f64 | fallback[n] u32   (d.counts is idle here)
    const size_t o_y = 0, o_out = o_y + (size_t)S * 8, o_fb = o_out + (size_t)n * out_stride * 8, o_end = o_fb + (size_t)n * 4;
    CK(d.counts.ensure(o_end));
    uint8_t* sb = d.counts.as<uint8_t>();
    DotDev t;
    if (y_on_device) t.y = y;
    else { CK(cudaMemcpyAsync(sb + o_y, y, (size_t)S * 8, cudaMemcpyHostToDevice, ctx->stream)); t.y = reinterpret_cast<const double*>(sb + o_y); }
    t.out = reinterpret_cast<double*>(sb + o_out); t.out_stride = out_stride; t.fallback = reinterpret_cast<uint32_t*>(sb + o_fb);
    CK(cudaMemsetAsync(t.out, 0, (size_t)n * out_stride * 8, ctx->stream));
    CK(d.req.ensure(n * 4 * 4));
    uint32_t* rq = d.req.as<uint32_t>();
    CK(cudaMemcpyAsync(rq, block_index, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(rq + n, line_offset, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(rq + 2 * n, n_alleles, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ReqDev q = {};
    q.blk = rq; q.line = rq + n; q.nall = rq + 2 * n; q.n = (uint32_t)n;
    { PROF("dot_lines"); dot_lines_kernel<<<(uint32_t)((n + 3) / 4), 128, 0, ctx->stream>>>(d.dev, q, t); }
    CKL();
    std::vector<uint32_t> fb(n);
    CK(cudaMemcpyAsync(fb.data(), t.fallback, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<uint32_t> idx;
    for (uint64_t i = 0; i < n; ++i) if (fb[i]) idx.push_back((uint32_t)i);
    if (!idx.empty()) {  // records with a negated sparse line: from their composed rows, in chunks
        const uint64_t stride8 = ((uint64_t)2 * S + 15) / 16 * 16;
        const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(idx.size(), (256ull << 20) / stride8));
        std::vector<uint32_t> b2, l2, a2;
        for (uint64_t c0 = 0; c0 < idx.size(); c0 += chunk) {
            const uint64_t cn = std::min<uint64_t>(chunk, idx.size() - c0);
            b2.resize(cn); l2.resize(cn); a2.resize(cn);
            for (uint64_t k = 0; k < cn; ++k) { b2[k] = block_index[idx[c0 + k]]; l2[k] = line_offset[idx[c0 + k]]; a2[k] = n_alleles[idx[c0 + k]]; }
            CK(d.x_pool.ensure(cn * stride8 + cn * 8 + (size_t)n * 4));
            int8_t* rows8 = d.x_pool.as<int8_t>();
            uint32_t* dfilled = reinterpret_cast<uint32_t*>(rows8 + cn * stride8);
            uint32_t* didx = dfilled + cn;
            uint32_t* dnall = didx + cn;
            std::vector<uint32_t> filled(cn);
            int rc = decode_records_impl<int8_t>(ctx, cn, b2.data(), l2.data(), a2.data(), rows8, stride8, 1, filled.data(), nullptr, 0);
            if (rc) return rc;
            CK(cudaMemcpyAsync(dfilled, filled.data(), cn * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(didx, idx.data() + c0, cn * 4, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(dnall, n_alleles, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
            { PROF("dot_rows"); dot_rows_kernel<<<(uint32_t)((cn + 3) / 4), 128, 0, ctx->stream>>>(rows8, stride8, dfilled, didx, dnall, (uint32_t)cn, S, t); }
            CKL();
            CK(cudaStreamSynchronize(ctx->stream));  // filled / idx staging of this chunk is reused by the next
        }
    }
    CK(cudaMemcpyAsync(out, t.out, (size_t)n * out_stride * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return XSI_OK;
}

// InternalGtAccess (accessor_internals.hpp:374-397; filled by DecompressPointerGTBlock::get_internal_access,
// accessor_internals_new.hpp:444-471): where the encoded lines of a record lie inside its GT block, the default allele of its
// first line, and the PBWT arrangement in force at the record (a[j] = haplotype at position j, the order WAH lines are in).
static int xsi_decode_internal_access_impl(xsi_ctx* ctx, uint32_t b, uint32_t line, uint32_t n_alleles, xsi_line_access* lines,
                                           int32_t* default_allele, void* a_out) {
    if (!ctx || !lines) return XSI_E_ARG;
    auto& d = ctx->dec;
    if (!d.loaded || b >= d.nb) { ctx->err = "xsi_decode_internal_access: block not loaded"; return XSI_E_ARG; }
    if (n_alleles < 2 || (uint64_t)line + n_alleles - 1 > d.h_bin_lines[b]) { ctx->err = "record runs past the end of its block"; return XSI_E_ARG; }
    CK(cudaSetDevice(ctx->device));
    const DecBlock& bk = d.hv_blocks[b];
    const uint32_t msb = d.aet == 2 ? 0x8000u : 0x80000000u;
    const uint32_t nl = n_alleles - 1;
    // what the device located: first word of every WAH line inside its matrix, first entry of every sparse list
    std::vector<uint32_t> w0(nl + 1, 0);
    std::vector<uint64_t> e0(nl, 0);
    for (uint32_t k = 0; k < nl; ++k) {
        const uint32_t gl = bk.line0 + line + k;
        const uint32_t ord = d.hv_dl_ord[gl];
        if (d.hv_dl_flags[gl] & DL_WAH) {
            CK(cudaMemcpyAsync(&w0[k], d.dev.job_word0 + ord, 4, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            CK(cudaMemcpyAsync(&e0[k], d.dev.sp_off + ord, 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    std::vector<uint32_t> w1(nl, 0);  // word after the line: the next WAH job of the block's matrix, or its end
    for (uint32_t k = 0; k < nl; ++k) {
        const uint32_t gl = bk.line0 + line + k;
        if (!(d.hv_dl_flags[gl] & DL_WAH)) continue;
        const uint32_t ord = d.hv_dl_ord[gl];
        const DecSeg& sg = d.hv_segs[d.hv_job_seg[ord]];
        if (ord + 1 < sg.job0 + sg.njobs) CK(cudaMemcpyAsync(&w1[k], d.dev.job_word0 + ord + 1, 4, cudaMemcpyDeviceToHost, ctx->stream));
        else w1[k] = sg.n_words;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint32_t k = 0; k < nl; ++k) {
        const uint32_t gl = bk.line0 + line + k;
        const uint32_t ord = d.hv_dl_ord[gl];
        if (d.hv_dl_flags[gl] & DL_WAH) {
            const DecSeg& sg = d.hv_segs[d.hv_job_seg[ord]];
            lines[k].is_sparse = 0;
            lines[k].byte_offset = (sg.byte_off - bk.blob_off) + (uint64_t)w0[k] * 2;
            lines[k].n_entries = w1[k] - w0[k];
        } else {
            lines[k].is_sparse = 1;
            lines[k].byte_offset = (bk.sparse_off - bk.blob_off) + e0[k] * d.aet;
            lines[k].n_entries = 0;  // the list's own header word says (count, MSB = lists REF carriers)
        }
    }
    if (default_allele) {  // accessor_internals_new.hpp:456-463
        *default_allele = 0;
        if (lines[0].is_sparse) {
            uint32_t hdr = 0;
            CK(cudaMemcpy(&hdr, d.blob.as<uint8_t>() + bk.blob_off + lines[0].byte_offset, d.aet, cudaMemcpyDeviceToHost));
            *default_allele = (hdr & msb) ? 1 : 0;
        }
    }
    if (a_out) {
        if (!d.lazy_ok) { ctx->err = "the arrangement is only kept by the lazy chain (at most 65534 haplotypes, no all-haploid lines)"; return XSI_E_UNSUPPORTED; }
        const auto& wl = d.h_wah_lines[b];
        // the reference hands out its live `a` after seeking to the record's LAST line (the loop at :456-468 seeks line by line)
        const uint32_t at = line + nl - 1;
        const uint32_t k = (uint32_t)(std::lower_bound(wl.begin(), wl.end(), (uint16_t)at) - wl.begin());  // WAH lines before it
        if (d.h_wah_done[b] > k) { ctx->err = "the chain of this block is already past that line: load it lazily again"; return XSI_E_UNSUPPORTED; }
        if (d.h_wah_done[b] < k) { const int rc = extend_chain(ctx, b, at); if (rc) return rc; }
        const uint32_t N = 2 * d.n_samples;
        std::vector<uint16_t> pos(N);
        if (k == 0) for (uint32_t i = 0; i < N; ++i) pos[i] = (uint16_t)i;  // identity at block start (gt_block.hpp:179)
        else {
            CK(cudaMemcpyAsync(pos.data(), d.pos_state.as<uint16_t>() + (size_t)b * d.ps_stride, (size_t)N * 2, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
        if (d.aet == 2) { uint16_t* a = static_cast<uint16_t*>(a_out); for (uint32_t i = 0; i < N; ++i) a[pos[i]] = (uint16_t)i; }
        else { uint32_t* a = static_cast<uint32_t*>(a_out); for (uint32_t i = 0; i < N; ++i) a[pos[i]] = i; }
    }
    return XSI_OK;
}

static int xsi_decode_records_i8_impl(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset,
                                     const uint32_t* n_alleles, int8_t* out, uint64_t out_stride, int32_t out_on_device,
                                     uint32_t* n_filled, uint64_t* allele_counts, uint32_t counts_stride) {
    return decode_records_impl<int8_t>(ctx, n, block_index, line_offset, n_alleles, out, out_stride, out_on_device, n_filled,
                                       allele_counts, counts_stride);
}

// ---- the C ABI promises no exception across the boundary: host containers (std::vector, std::map) may throw ----
template <typename F>
static int guarded(xsi_ctx* ctx, F&& f) {
    try {
        return f();
    } catch (const std::bad_alloc&) {
        if (ctx) ctx->err = "out of host memory";
        return XSI_E_NOMEM;
    } catch (const std::exception& ex) {
        if (ctx) ctx->err = std::string("internal error: ") + ex.what();
        return XSI_E_ARG;
    } catch (...) {
        if (ctx) ctx->err = "internal error";
        return XSI_E_ARG;
    }
}
extern "C" int xsi_encode_async(xsi_ctx* ctx, int on) {
    if (!ctx) return XSI_E_ARG;
    if (ctx->enc_thread.joinable()) ctx->enc_thread.join();
    if (on && !ctx->stream_enc) {
        if (cudaSetDevice(ctx->device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream_enc, cudaStreamNonBlocking) != cudaSuccess) {
            ctx->err = "cannot create the encode stream";
            return XSI_E_CUDA;
        }
    }
    if (on && !ctx->ev_scan && cudaEventCreateWithFlags(&ctx->ev_scan, cudaEventDisableTiming) != cudaSuccess) {
        ctx->ev_scan = nullptr;
        ctx->err = "cannot create the scan event";
        return XSI_E_CUDA;
    }
    if (ctx->es) cudaStreamSynchronize(ctx->es);
    ctx->async_encode = on != 0;
    ctx->es = on ? ctx->stream_enc : ctx->stream;
    return XSI_OK;
}

extern "C" int xsi_encode_launch(xsi_ctx* ctx, const xsi_encode_desc* d) {
    if (ctx && ctx->enc_thread.joinable()) ctx->enc_thread.join();  // a launch that was never collected
    if (ctx) ctx->enc_pending = false;
    // asynchronous mode: rows already on the device only (host rows share the context's transport rings with the decode side)
    if (ctx && d && ctx->async_encode && d->gt_on_device) {
        ctx->enc_desc = *d;
        ctx->enc_rc = XSI_OK;
        ctx->enc_pending = true;
        const uint64_t stride = ctx->enc_row_stride;
        const uint64_t seq = ctx->enc_seq.fetch_add(1) + 1;
        try {
            ctx->enc_thread = std::thread([ctx, stride, seq] {
                ctx->enc_rc = guarded(ctx, [&] { return xsi_encode_launch_impl(ctx, &ctx->enc_desc, stride); });
                if (ctx->scan_seq.load() < seq) ctx->scan_seq.store(seq);  // a launch that failed before its scan: nobody waits for it
            });
        } catch (...) {
            ctx->err = "cannot start the encode thread";
            return XSI_E_NOMEM;
        }
        return XSI_OK;
    }
    const uint64_t stride = ctx ? ctx->enc_row_stride : 0;
    return guarded(ctx, [&] { return xsi_encode_launch_impl(ctx, d, stride); });
}
extern "C" int xsi_encode_launch_strided(xsi_ctx* ctx, const xsi_encode_desc* d, uint64_t row_stride) {
    if (!ctx) return XSI_E_ARG;
    if (ctx->enc_thread.joinable()) ctx->enc_thread.join();
    ctx->enc_row_stride = row_stride;
    const int rc = xsi_encode_launch(ctx, d);
    ctx->enc_row_stride = 0;
    return rc;
}
extern "C" int xsi_encode_collect(xsi_ctx* ctx, uint32_t* n_blocks_out, const uint8_t* const** blocks_out, const uint64_t** sizes_out) {
    if (ctx && ctx->enc_thread.joinable()) ctx->enc_thread.join();
    if (ctx && ctx->enc_pending) {
        ctx->enc_pending = false;
        if (ctx->enc_rc != XSI_OK) return ctx->enc_rc;
    }
    return guarded(ctx, [&] { return xsi_encode_collect_impl(ctx, n_blocks_out, blocks_out, sizes_out); });
}
extern "C" int xsi_decode_load_blocks(xsi_ctx* ctx, uint32_t n_blocks, const uint8_t* const* gt_blocks, const uint64_t* sizes, uint64_t num_samples, int32_t aet_bytes) {
    return guarded(ctx, [&] { return xsi_decode_load_blocks_impl(ctx, n_blocks, gt_blocks, sizes, num_samples, aet_bytes); });
}
extern "C" int xsi_decode_load_blocks_lazy(xsi_ctx* ctx, uint32_t n_blocks, const uint8_t* const* gt_blocks, const uint64_t* sizes, uint64_t num_samples,
                                           int32_t aet_bytes, uint32_t initial_lines) {
    return guarded(ctx, [&] { return xsi_decode_load_blocks_impl(ctx, n_blocks, gt_blocks, sizes, num_samples, aet_bytes, initial_lines == 0xFFFFFFFFu ? 0xFFFFFFFEu : initial_lines); });
}
extern "C" int xsi_decode_extend(xsi_ctx* ctx, uint32_t block_index, uint32_t line_end) {
    if (!ctx || !ctx->dec.loaded || block_index >= ctx->dec.nb) return XSI_E_ARG;
    return guarded(ctx, [&]() -> int {
        CK(cudaSetDevice(ctx->device));
        return extend_chain(ctx, block_index, std::min<uint32_t>(line_end, ctx->dec.h_bin_lines[block_index]));
    });
}
extern "C" int xsi_decode_lines_ready(const xsi_ctx* ctx, uint32_t block_index, uint32_t* lines_ready) {
    if (!ctx || !ctx->dec.loaded || block_index >= ctx->dec.nb || !lines_ready) return XSI_E_ARG;
    const auto& d = ctx->dec;
    const auto& wl = d.h_wah_lines[block_index];
    *lines_ready = (!d.lazy_ok || d.h_wah_done[block_index] >= wl.size()) ? d.h_bin_lines[block_index] : wl[d.h_wah_done[block_index]];
    return XSI_OK;
}
extern "C" int xsi_decode_records(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset, const uint32_t* n_alleles, int32_t* out, uint64_t out_stride, int32_t out_on_device, uint32_t* n_filled, uint64_t* allele_counts, uint32_t counts_stride) {
    return guarded(ctx, [&] { return xsi_decode_records_impl(ctx, n, block_index, line_offset, n_alleles, out, out_stride, out_on_device, n_filled, allele_counts, counts_stride); });
}
extern "C" int xsi_decode_records_i8(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset, const uint32_t* n_alleles, int8_t* out, uint64_t out_stride, int32_t out_on_device, uint32_t* n_filled, uint64_t* allele_counts, uint32_t counts_stride) {
    return guarded(ctx, [&] { return xsi_decode_records_i8_impl(ctx, n, block_index, line_offset, n_alleles, out, out_stride, out_on_device, n_filled, allele_counts, counts_stride); });
}
extern "C" int xsi_decode_records_subset(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset, const uint32_t* n_alleles, const uint32_t* samples_to_use, uint32_t n_sel, int32_t* out, uint64_t out_stride, int32_t out_on_device, uint32_t* n_filled, uint32_t* ac, uint32_t ac_stride) {
    return guarded(ctx, [&] { return xsi_decode_records_subset_impl(ctx, n, block_index, line_offset, n_alleles, samples_to_use, n_sel, out, out_stride, out_on_device, n_filled, ac, ac_stride); });
}
extern "C" int xsi_decode_allele_counts(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset, const uint32_t* n_alleles, uint64_t* allele_counts, uint32_t counts_stride) {
    return guarded(ctx, [&] { return xsi_decode_allele_counts_impl(ctx, n, block_index, line_offset, n_alleles, allele_counts, counts_stride); });
}
extern "C" int xsi_decode_dot_products(xsi_ctx* ctx, uint64_t n, const uint32_t* block_index, const uint32_t* line_offset, const uint32_t* n_alleles,
                                       const double* y, int32_t y_on_device, double* out, uint32_t out_stride) {
    return guarded(ctx, [&] { return xsi_decode_dot_products_impl(ctx, n, block_index, line_offset, n_alleles, y, y_on_device, out, out_stride); });
}
extern "C" int xsi_decode_internal_access(xsi_ctx* ctx, uint32_t block_index, uint32_t line_offset, uint32_t n_alleles, xsi_line_access* lines,
                                          int32_t* default_allele, void* a) {
    return guarded(ctx, [&] { return xsi_decode_internal_access_impl(ctx, block_index, line_offset, n_alleles, lines, default_allele, a); });
}
