// Shared helpers for libqsv (host error plumbing, launch counter, complex arithmetic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/qsv.h"

namespace qsv {

void set_error(const char *fmt, ...);
void count_launch(unsigned n = 1);

#define QSV_CHECK_ARG(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) { ::qsv::set_error(__VA_ARGS__); return 1; }  \
    } while (0)

#define QSV_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            ::qsv::set_error("%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,  \
                             __LINE__);                                                         \
            return 2;                                                                           \
        }                                                                                       \
    } while (0)

#define QSV_LAUNCH_CHECK()                                                                      \
    do {                                                                                        \
        cudaError_t e_ = cudaGetLastError();                                                    \
        if (e_ != cudaSuccess) {                                                                \
            ::qsv::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_),        \
                             __FILE__, __LINE__);                                               \
            return 2;                                                                           \
        }                                                                                       \
        ::qsv::count_launch();                                                                  \
    } while (0)

// complex128 helpers on double2 (x = re, y = im)
__device__ __forceinline__ double2 cmul(double2 a, double2 b) {
    return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cfma(double2 a, double2 b, double2 c) {  // a*b + c
    return make_double2(fma(a.x, b.x, fma(-a.y, b.y, c.x)), fma(a.x, b.y, fma(a.y, b.x, c.y)));
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) {
    return make_double2(a.x + b.x, a.y + b.y);
}
// |a|^2 with ONE fixed evaluation order (re * re + round(im * im), fused): the generated pass kernels compute the same
// weights in their sums epilogue (csrc/qsv_jit.cu, qsv_weight) and must agree with the sampler to the last bit
__device__ __forceinline__ double cnorm2(double2 a) { return __fma_rn(a.x, a.x, __dmul_rn(a.y, a.y)); }

// Insert a zero bit at position p of v (bits >= p move up by one).
__device__ __forceinline__ uint64_t insert_zero(uint64_t v, unsigned p) {
    return ((v >> p) << (p + 1)) | (v & ((1ull << p) - 1ull));
}
__device__ __forceinline__ uint32_t insert_zero32(uint32_t v, unsigned p) {
    return ((v >> p) << (p + 1)) | (v & ((1u << p) - 1u));
}

struct PosList {           // small by-value list of bit positions for the wide kernels
    int32_t n;
    uint8_t pos[63];
};

}  // namespace qsv
