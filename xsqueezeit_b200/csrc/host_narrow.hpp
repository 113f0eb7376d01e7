// host_narrow.hpp -- int32 <-> BCF int8 genotype transport conversion on the host (see host_narrow.cpp)
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace xsi {

// worker threads the conversions run on (XSI_HOST_THREADS, default: the cores this process may use, <= 32)
unsigned host_threads();

// fn(0) .. fn(n_tasks-1) on the worker pool (the caller takes part); one job at a time per process
void host_parallel_for(size_t n_tasks, const std::function<void(size_t)>& fn);

// dst[i] = BCF int8 encoding of src[i].  Returns false when some value has no int8 encoding (allele index
// above 62, or a negative value other than bcf_int32_missing / bcf_int32_vector_end); dst is then garbage.
bool narrow_i32_to_i8(const int32_t* src, int8_t* dst, size_t n);

// Row r: dst[r*dst_stride + j] = int32 encoding of src[r*src_stride + j] for j < len[r]; the rest of the
// destination row is left untouched.
void widen_rows_i8_to_i32(const int8_t* src, size_t src_stride, int32_t* dst, size_t dst_stride, const uint32_t* len,
                          size_t n_rows);

}  // namespace xsi
