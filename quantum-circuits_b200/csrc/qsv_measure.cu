// Measurement: probabilities, block sums and the sampling scan.
//
// Replaces main/circuit.py:329-337 (output relabel P1^T * psi, weights |psi|^2) and the
// random.choices draw of main/circuit.py:360-364 (sequential accumulate + bisect_right with
// hi = N-1).  The output relabel is folded into the read: output index z reads physical index
// y(z) whose bit src_pos[b] is bit b of z.
#include "qsv_common.cuh"

namespace qsv {

struct BitMap {
    int32_t n;          // number of bits
    int32_t identity;   // 1: y == z
    uint8_t src[64];    // physical position feeding bit b of z
};

__device__ __forceinline__ uint64_t map_index(uint64_t z, const BitMap &m) {
    if (m.identity) return z;
    uint64_t y = 0;
    for (int b = 0; b < m.n; ++b) y |= ((z >> b) & 1ull) << m.src[b];
    return y;
}

__global__ void k_probs(const double2 *__restrict__ state, uint64_t n_amps, const __grid_constant__ BitMap m,
                        double *__restrict__ out) {
    for (uint64_t z = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; z < n_amps;
         z += (uint64_t)gridDim.x * blockDim.x)
        out[z] = cnorm2(state[map_index(z, m)]);
}

// One CTA per block of 2^block_bits output indices.  Deterministic: every thread accumulates a
// fixed strided subset in ascending order, then a fixed-shape warp-shuffle + shared-memory tree.
constexpr int kSumThreads = 256;
__global__ void __launch_bounds__(kSumThreads)
k_block_sums(const double2 *__restrict__ state, const __grid_constant__ BitMap m, int block_bits,
             uint64_t n_blocks, double *__restrict__ sums) {
    __shared__ double warp_part[kSumThreads / 32];
    const uint64_t len = 1ull << block_bits;
    for (uint64_t blk = blockIdx.x; blk < n_blocks; blk += gridDim.x) {
        const uint64_t z0 = blk << block_bits;
        double acc = 0.0;
        for (uint64_t w = threadIdx.x; w < len; w += kSumThreads) acc += cnorm2(state[map_index(z0 + w, m)]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < kSumThreads / 32; ++w) s += warp_part[w];
            sums[blk] = s;
        }
        __syncthreads();
    }
}

// Sequential scan in reference order.  One warp: lanes prefetch 32 weights at a time, lane 0 does
// the left-to-right accumulation so rounding matches itertools.accumulate exactly.
__global__ void k_sample_scan(const double2 *__restrict__ state, const __grid_constant__ BitMap m, uint64_t z_start,
                              uint64_t z_last, double prefix, double target, unsigned long long *result) {
    const unsigned lane = threadIdx.x;
    double cum = prefix;
    unsigned long long found = ~0ull;
    for (uint64_t z0 = z_start; z0 <= z_last && found == ~0ull; z0 += 32) {
        const uint64_t z = z0 + lane;
        const double w = (z <= z_last) ? cnorm2(state[map_index(z, m)]) : 0.0;
        for (int l = 0; l < 32; ++l) {
            const double wl = __shfl_sync(0xffffffffu, w, l);
            if (found == ~0ull && z0 + l <= z_last) {
                cum += wl;
                if (cum > target) found = z0 + l;
            }
        }
    }
    if (lane == 0) {
        result[0] = (found == ~0ull) ? (unsigned long long)z_last : (found < z_last ? found : z_last);
        reinterpret_cast<double *>(result)[1] = cum;   // running sum where the scan stopped
    }
}

static int make_map(BitMap &m, int n_local, const int *src_pos, const char *who) {
    QSV_CHECK_ARG(n_local >= 0 && n_local <= 40, "%s: n_local=%d out of range", who, n_local);
    m.n = n_local;
    m.identity = 1;
    uint64_t seen = 0;
    for (int b = 0; b < n_local; ++b) {
        const int s = src_pos ? src_pos[b] : b;
        QSV_CHECK_ARG(s >= 0 && s < n_local && !((seen >> s) & 1ull), "%s: src_pos is not a permutation", who);
        seen |= 1ull << s;
        m.src[b] = (uint8_t)s;
        if (s != b) m.identity = 0;
    }
    return 0;
}

}  // namespace qsv

using namespace qsv;

extern "C" int qsv_probs(const void *state_dev, int n_local, const int *src_pos, void *out_dev, void *stream) {
    QSV_CHECK_ARG(state_dev && out_dev, "qsv_probs: NULL pointer");
    BitMap m;
    if (int rc = make_map(m, n_local, src_pos, "qsv_probs")) return rc;
    const uint64_t n = 1ull << n_local;
    uint64_t g = (n + 255) / 256;
    if (g > 148ull * 32ull) g = 148ull * 32ull;
    k_probs<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const double2 *)state_dev, n, m, (double *)out_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_block_sums(const void *state_dev, int n_local, const int *src_pos, int block_bits,
                              void *sums_dev, void *stream) {
    QSV_CHECK_ARG(state_dev && sums_dev, "qsv_block_sums: NULL pointer");
    QSV_CHECK_ARG(block_bits >= 0 && block_bits <= n_local, "qsv_block_sums: block_bits=%d invalid", block_bits);
    BitMap m;
    if (int rc = make_map(m, n_local, src_pos, "qsv_block_sums")) return rc;
    const uint64_t n_blocks = 1ull << (n_local - block_bits);
    const unsigned g = (unsigned)(n_blocks < 148ull * 8ull ? n_blocks : 148ull * 8ull);
    k_block_sums<<<g, kSumThreads, 0, (cudaStream_t)stream>>>((const double2 *)state_dev, m, block_bits, n_blocks,
                                                              (double *)sums_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_sample_scan(const void *state_dev, int n_local, const int *src_pos, uint64_t z_start,
                               uint64_t z_last, double prefix, double target, void *result_dev, void *stream) {
    QSV_CHECK_ARG(state_dev && result_dev, "qsv_sample_scan: NULL pointer");
    BitMap m;
    if (int rc = make_map(m, n_local, src_pos, "qsv_sample_scan")) return rc;
    QSV_CHECK_ARG(z_start <= z_last && z_last < (1ull << n_local), "qsv_sample_scan: bad range");
    k_sample_scan<<<1, 32, 0, (cudaStream_t)stream>>>((const double2 *)state_dev, m, z_start, z_last, prefix, target,
                                                      (unsigned long long *)result_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}
