// common.cuh -- shared device helpers for the sm_100a genotype kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define XSI_FULL 0xFFFFFFFFu
#define XSI_I32_MISSING ((int32_t)0x80000000)     // bcf_int32_missing, htslib/vcf.h:1324
#define XSI_I32_VECTOR_END ((int32_t)0x80000001)  // bcf_int32_vector_end, htslib/vcf.h:1329

// line_flags bits (one byte per binary line)
#define LF_WAH 1u      // line is PBWT+WAH encoded (else sparse)            gt_block.hpp:299-303
#define LF_NEGATED 2u  // sparse line lists allele==0 carriers, MSB set    gt_block.hpp:318-324
#define LF_HAPLOID 4u  // line belongs to an all-haploid record (ngt == n_samples)
// rec_flags bits (one byte per BCF record)
#define RF_MISSING 1u
#define RF_EOV 2u
#define RF_PHASE 4u
#define RF_HAPLOID 8u

namespace xsi {

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) -------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; bytes % 16 == 0, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- distributed shared memory (thread-block cluster) ------------------------------------------
__device__ __forceinline__ uint32_t mapa_u32(uint32_t cta_smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(cta_smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t cluster_addr) {
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(cluster_addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t cluster_addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(cluster_addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_v2(uint32_t cluster_addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(cluster_addr), "r"(a), "r"(b) : "memory");
}

// arrive on an mbarrier that lives in another CTA of the cluster (address from mapa_u32), release at cluster scope
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_mbar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_mbar_addr) : "memory");
}
// wait on a local mbarrier whose arrivals come from the whole cluster (acquire at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAITC_%=:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONEC_%=;\n"
        "bra WAITC_%=;\n"
        "DONEC_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// Asynchronous stores into another CTA's shared memory that complete transaction bytes on an mbarrier of
// that CTA (both addresses from mapa_u32).  The receiver only waits on its own mbarrier: no release /
// acquire fence at cluster scope (which ptxas turns into MEMBAR.ALL.GPU + CCTL.IVALL) is involved.
__device__ __forceinline__ void st_async_b32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr), "r"(v),
                 "r"(cluster_mbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_v2(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t cluster_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(cluster_addr),
                 "r"(a), "r"(b), "r"(cluster_mbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint4 v, uint32_t cluster_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                     cluster_addr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_mbar)
                 : "memory");
}
// predicated OR into a shared-memory word (32-bit shared-window address, no generic-address arithmetic, no branch)
__device__ __forceinline__ void red_or_shared_if(uint32_t test, uint32_t saddr, uint32_t val) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %0, 0;\n"
        "@p red.shared::cta.or.b32 [%1], %2;\n"
        "}\n" ::"r"(test),
        "r"(saddr), "r"(val)
        : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {  // plain shared load from a 32-bit shared-window address
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
// barrier among the first `nthreads` threads of the CTA (a multiple of 32), hardware barrier 1
__device__ __forceinline__ void named_bar_sync1(uint32_t nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// ---- genotype value helpers (htslib/vcf.h:892-898) ------------------------------------------
template <int ELEM>
__device__ __forceinline__ int32_t load_gt(const void* base, uint64_t idx) {
    if (ELEM == 4) {
        return __ldg(reinterpret_cast<const int32_t*>(base) + idx);
    } else {
        int32_t b = (int32_t)__ldg(reinterpret_cast<const signed char*>(base) + idx);
        // raw BCF int8: 0x80 = missing, 0x81 = end of vector (vcf.h bcf_int8_missing / bcf_int8_vector_end)
        return b == -128 ? XSI_I32_MISSING : (b == -127 ? XSI_I32_VECTOR_END : b);
    }
}
__device__ __forceinline__ bool gt_is_missing(int32_t v) { return ((v >> 1) == 0) || v == XSI_I32_MISSING; }

// ---- int32 <-> BCF int8 transport encoding on the device (the host side is csrc/host_narrow.cpp) ----------
// n a multiple of 16 (the caller pads), both pointers 16-byte aligned; *bad |= 1 when a value has no int8 encoding
__global__ void __launch_bounds__(256) narrow_i32_i8_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, uint64_t n16,
                                                             uint32_t* bad) {
    uint32_t lost = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n16; i += (uint64_t)gridDim.x * blockDim.x) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint4 v = __ldcs(src + 4 * i + k);
            const uint32_t u[4] = {v.x, v.y, v.z, v.w};
            uint32_t w = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                w |= ((u[b] & 0x7Fu) | ((u[b] >> 24) & 0x80u)) << (8 * b);
                lost |= (u[b] & 0x7FFFFF80u) | ((uint32_t)((int32_t)u[b] >> 31) & u[b] & 0x7Eu);
            }
            o[k] = w;
        }
        dst[i] = make_uint4(o[0], o[1], o[2], o[3]);
    }
    if (lost) atomicOr(bad, 1u);
}
// rows of `len` int8 genotypes (stride8 apart) -> rows of int32 (len apart); len a multiple of 16, 16-byte aligned rows
__global__ void __launch_bounds__(256) widen_rows_i8_i32_kernel(const int8_t* __restrict__ src, uint64_t stride8, int32_t* __restrict__ dst,
                                                                 uint32_t len, uint32_t rows) {
    const uint32_t per_row = len >> 4;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (uint64_t)rows * per_row; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = (uint32_t)(i / per_row), c = (uint32_t)(i - (uint64_t)r * per_row);
        const uint4 v = *reinterpret_cast<const uint4*>(src + r * stride8 + 16ull * c);
        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
        uint4* d4 = reinterpret_cast<uint4*>(dst + (uint64_t)r * len + 16ull * c);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t o[4];
#pragma unroll
            for (int b = 0; b < 4; ++b) { const uint32_t x = (u[k] >> (8 * b)) & 0xFFu; o[b] = (x & 0x7Fu) | ((x & 0x80u) << 24); }
            __stcs(d4 + k, make_uint4(o[0], o[1], o[2], o[3]));
        }
    }
}

}  // namespace xsi
