// Feynman path sum of Circuit.run_method1 (reference main/circuit.py:221-276) on the device.
//
// The reference fixes the input wires to in_v, then for every assignment `out` of the n output wires sums, over all
// 2^W assignments of the W internal wires, the product over the gates of gate.matrix[i][j] with i = the gate's
// in-wire values and j = its out-wire values (first wire = most significant bit), and returns |sum|^2.  A wire that
// runs from a circuit input straight to a circuit output lets only out == in through.
//
// Here: one CTA per output assignment, its threads stride over the internal assignments (the product over the gates
// is a handful of table lookups), a CTA-wide reduction gives the amplitude.  The work is the reference's own
// 2^(n + W) * gates - the algorithm is kept, not replaced by a state-vector evolution (that is run_method2) - so it is
// meant for the circuits the reference's cost model sends to method 1 (few internal wires).
#include "qsv_common.cuh"
#include <stdint.h>

namespace qsv {

constexpr int kPathThreads = 256;

// wire value: kind 0 = constant (idx = the bit), 1 = output wire (bit idx of the out index, idx = n-1-i for output
// position i), 2 = internal wire (bit idx of the assign index, idx = W-1-j for the j-th internal wire)
struct PathWire { uint8_t kind, idx; };

__global__ void __launch_bounds__(kPathThreads)
k_path_sum(const PathWire *__restrict__ wires, const uint32_t *__restrict__ gate_hdr,   // per gate: k, matrix offset, port offset
           const uint32_t *__restrict__ ports,                                          // per gate: k in-wire ids, k out-wire ids
           const double2 *__restrict__ mats, int n_gates, int n_internal,
           const uint32_t *__restrict__ direct, int n_direct,                           // (in bit, out bit position) pairs
           uint64_t out0, double *__restrict__ weights) {
    const uint64_t out = out0 + blockIdx.x;
    __shared__ double red_r[kPathThreads], red_i[kPathThreads];
    bool dead = false;
    for (int d = 0; d < n_direct; ++d)
        if (((out >> direct[2 * d + 1]) & 1ull) != (uint64_t)direct[2 * d]) dead = true;
    double sr = 0.0, si = 0.0;
    if (!dead) {
        const uint64_t n_assign = 1ull << n_internal;
        for (uint64_t a = threadIdx.x; a < n_assign; a += kPathThreads) {
            double pr = 1.0, pi = 0.0;
            bool zero = false;
            for (int g = 0; g < n_gates && !zero; ++g) {
                const uint32_t k = gate_hdr[3 * g], moff = gate_hdr[3 * g + 1], poff = gate_hdr[3 * g + 2];
                uint32_t i = 0, j = 0;
                for (uint32_t p = 0; p < k; ++p) {
                    const PathWire wi = wires[ports[poff + p]], wo = wires[ports[poff + k + p]];
                    const uint32_t vi = wi.kind == 0 ? wi.idx : (uint32_t)(((wi.kind == 1 ? out : a) >> wi.idx) & 1ull);
                    const uint32_t vo = wo.kind == 0 ? wo.idx : (uint32_t)(((wo.kind == 1 ? out : a) >> wo.idx) & 1ull);
                    i = (i << 1) | vi;
                    j = (j << 1) | vo;
                }
                const double2 m = mats[moff + ((uint64_t)i << k) + j];
                if (m.x == 0.0 && m.y == 0.0) { zero = true; break; }
                const double t = pr * m.x - pi * m.y;
                pi = pr * m.y + pi * m.x;
                pr = t;
            }
            if (!zero) { sr += pr; si += pi; }
        }
    }
    red_r[threadIdx.x] = sr;
    red_i[threadIdx.x] = si;
    __syncthreads();
    for (int s = kPathThreads / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            red_r[threadIdx.x] += red_r[threadIdx.x + s];
            red_i[threadIdx.x] += red_i[threadIdx.x + s];
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) weights[out] = red_r[0] * red_r[0] + red_i[0] * red_i[0];
}

}  // namespace qsv

using namespace qsv;

extern "C" int qsv_path_sum(const void *wires_dev, int n_wires, const void *gate_hdr_dev, const void *ports_dev,
                            const void *mats_dev, int n_gates, int n_outputs, int n_internal, const void *direct_dev,
                            int n_direct, void *weights_dev, void *stream) {
    QSV_CHECK_ARG(wires_dev && gate_hdr_dev && ports_dev && mats_dev && weights_dev, "qsv_path_sum: NULL pointer");
    QSV_CHECK_ARG(n_wires >= 1 && n_gates >= 0 && n_outputs >= 1 && n_outputs <= 30 && n_internal >= 0 && n_internal <= 40,
                  "qsv_path_sum: sizes out of range (outputs %d, internal wires %d)", n_outputs, n_internal);
    QSV_CHECK_ARG(n_direct == 0 || direct_dev, "qsv_path_sum: NULL direct-wire list");
    static_assert(sizeof(PathWire) == 2, "PathWire is two bytes");
    const uint64_t n_out = 1ull << n_outputs;
    for (uint64_t o = 0; o < n_out; o += 1u << 20) {          // grid.x chunks
        const uint64_t cnt = n_out - o < (1u << 20) ? n_out - o : (1u << 20);
        k_path_sum<<<(unsigned)cnt, kPathThreads, 0, (cudaStream_t)stream>>>(
            (const PathWire *)wires_dev, (const uint32_t *)gate_hdr_dev, (const uint32_t *)ports_dev,
            (const double2 *)mats_dev, n_gates, n_internal, (const uint32_t *)direct_dev, n_direct, o,
            (double *)weights_dev);
        QSV_LAUNCH_CHECK();
    }
    return 0;
}
