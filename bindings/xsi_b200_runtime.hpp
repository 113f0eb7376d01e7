// bindings/xsi_b200_runtime.hpp -- what the two reference-side adapters share: one xsi_ctx per host thread,
// pinned staging buffers that outlive the per-block adapter objects, and the mapping of C-ABI return codes to
// the reference's error convention (`throw const char*`, e.g. gt_block.hpp:259-265, accessor.cpp:29-50).
//
// This header is compiled INTO the reference tree's translation units (see bindings/Makefile); it only needs
// include/xsi_b200.h and the C++ standard library.
#ifndef XSI_B200_RUNTIME_HPP
#define XSI_B200_RUNTIME_HPP
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "xsi_b200.h"

namespace xsi_b200 {

inline int device_from_env() {
    const char* e = getenv("XSI_B200_DEVICE");
    return e ? atoi(e) : 0;
}

// A C-ABI failure becomes the reference's kind of exception.  The text is kept in a thread-local string because
// xsi_last_error's storage belongs to a context that the unwinding may destroy.
[[noreturn]] inline void raise(const xsi_ctx* ctx, int rc, const char* what) {
    static thread_local std::string msg;
    msg = std::string(what) + ": " + (ctx ? xsi_last_error(ctx) : "no context") + " (xsi_b200 rc " + std::to_string(rc) + ")";
    fprintf(stderr, "%s\n", msg.c_str());
    // same literals as the reference where it has one, so that callers that compare or print them see no change
    if (rc == XSI_E_ALLELE) throw "Unknown allele error !";                       // gt_block.hpp:259-265
    if (rc == XSI_E_PLOIDY) throw "Ploidy higher than 2 is not yet supported";     // gt_compressor_new.hpp:118-120
    throw msg.c_str();
}

// pinned host buffer that only grows
struct Pinned {
    void* p = nullptr;
    size_t cap = 0;
    void reserve(size_t bytes, size_t keep = 0) {
        if (bytes <= cap) return;
        size_t want = cap ? cap : (size_t)1 << 20;
        while (want < bytes) want *= 2;
        void* q = nullptr;
        const int rc = xsi_host_alloc(&q, want);
        if (rc != XSI_OK) raise(nullptr, rc, "xsi_host_alloc");
        if (keep) memcpy(q, p, keep);
        xsi_host_free(p);
        p = q; cap = want;
    }
    ~Pinned() { xsi_host_free(p); }
    template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

// One encode context and one row staging area per host thread: GtBlockB200 objects come and go with every
// block (xsi_factory.hpp:537), their device pools and pinned rows should not.
struct ThreadState {
    xsi_ctx* ctx = nullptr;
    Pinned rows;
    ~ThreadState() { if (ctx) xsi_destroy(ctx); }
    xsi_ctx* context() {
        if (!ctx) {
            const int rc = xsi_create(device_from_env(), &ctx);
            if (rc != XSI_OK) { ctx = nullptr; raise(nullptr, rc, "xsi_create (no CUDA device? there is no CPU fallback)"); }
        }
        return ctx;
    }
};
inline ThreadState& thread_state() {
    static thread_local ThreadState ts;
    return ts;
}

}  // namespace xsi_b200
#endif
