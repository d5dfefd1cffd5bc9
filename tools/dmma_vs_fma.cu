// The north star's tensor-core question, measured: does a FUSED DENSE 2^k x 2^k block on the FP64 tensor path (DMMA,
// mma.sync.m8n8k4.f64) beat the bandwidth path?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o tools/dmma_vs_fma tools/dmma_vs_fma.cu
//   ./tools/dmma_vs_fma [log2 amplitudes, default 28]
//
// Kernels
//   peak_dfma / peak_dmma   register-only loops: the FP64 FMA rate and the FP64 mma.sync rate of the device
//   dense4_dmma             psi -> (U (x) I) psi for a dense complex 16 x 16 block U on the 4 LOWEST qubits, as the real
//                           32 x 32 GEMM  Y = M X  (M = [[Re U, -Im U], [Im U, Re U]] in interleaved re/im order, one
//                           column of X = the 16 amplitudes of one group) on mma.sync.aligned.m8n8k4.f64
//   dense4_fma              the same block with plain FMAs (one group of 16 amplitudes per thread)
//   copy                    read + write of the state (the HBM floor of any gate pass)
// A k = 4 block costs 16 complex MACs = 128 real flop per amplitude whatever the four gates inside it were; the
// bandwidth path applies the same four 1-qubit gates with 4 x (4..14) flop per amplitude.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, const double a, const double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void peak_dfma(double *out, int iters) {
    double a[16];
    for (int i = 0; i < 16; ++i) a[i] = 1.0 + threadIdx.x * 1e-9 + i;
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], m, c);
    double s = 0;
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void peak_dmma(double *out, int iters) {
    double c[16];
    for (int i = 0; i < 16; ++i) c[i] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    double s = 0;
    for (int i = 0; i < 16; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Y = M X, M real 32 x 32 (row-major in global memory), X = [32 rows][n_groups columns], column g = the 32 doubles
// (16 complex amplitudes) at state + 32 g.  One warp handles 8 columns per iteration.
__global__ void __launch_bounds__(256) dense4_dmma(double *__restrict__ state, const double *__restrict__ M, uint64_t n_groups) {
    const int lane = threadIdx.x & 31;
    const int r4 = lane & 3, q8 = lane >> 2;
    double a[4][8];                                 // A fragments: M[8 mi + q8][4 kk + r4]
#pragma unroll
    for (int mi = 0; mi < 4; ++mi)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) a[mi][kk] = M[(8 * mi + q8) * 32 + 4 * kk + r4];
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t w = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    for (uint64_t g0 = w * 8; g0 < n_groups; g0 += warps * 8) {
        double *col = state + (g0 + q8) * 32;       // this lane's B column n = q8
        double b[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) b[kk] = col[4 * kk + r4];
        double c[4][2];
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            c[mi][0] = c[mi][1] = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) dmma884(c[mi][0], c[mi][1], a[mi][kk], b[kk]);
        }
        __syncwarp();                               // every lane has read its inputs before anybody overwrites them
        // C fragment: row 8 mi + q8, columns n = 2 r4 + {0, 1}
#pragma unroll
        for (int mi = 0; mi < 4; ++mi) {
            state[(g0 + 2 * r4) * 32 + 8 * mi + q8] = c[mi][0];
            state[(g0 + 2 * r4 + 1) * 32 + 8 * mi + q8] = c[mi][1];
        }
    }
}

// out of place (the 32 outputs of a group are written as they are produced; same traffic as in place)
__global__ void __launch_bounds__(256) dense4_fma(const double *__restrict__ state, double *__restrict__ out,
                                                  const double *__restrict__ M, uint64_t n_groups) {
    __shared__ double sm[32 * 32];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = M[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += stride) {
        const double2 *p = reinterpret_cast<const double2 *>(state + g * 32);
        double2 *q = reinterpret_cast<double2 *>(out + g * 32);
        double x[32];
#pragma unroll
        for (int i = 0; i < 16; ++i) { const double2 v = p[i]; x[2 * i] = v.x; x[2 * i + 1] = v.y; }
#pragma unroll 1
        for (int r = 0; r < 16; ++r) {
            double s0 = 0.0, s1 = 0.0;
            const double *m0 = sm + (2 * r) * 32, *m1 = m0 + 32;
#pragma unroll
            for (int k = 0; k < 32; ++k) { s0 = fma(m0[k], x[k], s0); s1 = fma(m1[k], x[k], s1); }
            q[r] = make_double2(s0, s1);
        }
    }
}

__global__ void __launch_bounds__(256) copy_rw(double2 *__restrict__ state, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        double2 v = state[i];
        v.x += 0.0;
        state[i] = v;
    }
}

template <class F> static float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); f();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / reps;
}

int main(int argc, char **argv) {
    const int nq = argc > 1 ? atoi(argv[1]) : 28;
    const uint64_t N = 1ull << nq, n_groups = N / 16;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    double *state, *Mdev, *scratch, *state2;
    CK(cudaMalloc(&state, N * 16));
    CK(cudaMalloc(&state2, N * 16));
    CK(cudaMalloc(&Mdev, 1024 * 8));
    CK(cudaMalloc(&scratch, (size_t)sms * 8 * 256 * 8));
    // U = H (x) T.H (x) H (x) S.H : a dense unitary 16 x 16 (so the state keeps its norm over the repetitions)
    std::vector<double> Mh(1024, 0.0);
    {
        const double s = sqrt(0.5);
        double ur[16][16], ui[16][16];
        const double g[4][2][2][2] = {
            {{{s, 0}, {s, 0}}, {{s, 0}, {-s, 0}}},
            {{{s, 0}, {s, 0}}, {{0.5, 0.5}, {-0.5, -0.5}}},
            {{{s, 0}, {s, 0}}, {{s, 0}, {-s, 0}}},
            {{{s, 0}, {s, 0}}, {{0, s}, {0, -s}}}};
        for (int r = 0; r < 16; ++r)
            for (int c = 0; c < 16; ++c) {
                double re = 1, im = 0;
                for (int q = 0; q < 4; ++q) {
                    const double *e = g[q][(r >> q) & 1][(c >> q) & 1];
                    const double nr = re * e[0] - im * e[1], ni = re * e[1] + im * e[0];
                    re = nr; im = ni;
                }
                ur[r][c] = re; ui[r][c] = im;
            }
        for (int r = 0; r < 16; ++r)
            for (int c = 0; c < 16; ++c) {
                Mh[(2 * r) * 32 + 2 * c] = ur[r][c];      Mh[(2 * r) * 32 + 2 * c + 1] = -ui[r][c];
                Mh[(2 * r + 1) * 32 + 2 * c] = ui[r][c];  Mh[(2 * r + 1) * 32 + 2 * c + 1] = ur[r][c];
            }
    }
    CK(cudaMemcpy(Mdev, Mh.data(), 1024 * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(state, 0, N * 16));
    {
        const double one[2] = {1.0, 0.0};
        CK(cudaMemcpy(state + 2 * 12345, one, 16, cudaMemcpyHostToDevice));
    }
    const int iters = 4096;
    const float t_fma = time_ms([&] { peak_dfma<<<sms * 8, 256>>>(scratch, iters); }, 5);
    const float t_mma = time_ms([&] { peak_dmma<<<sms * 8, 256>>>(scratch, iters); }, 5);
    const double fl_fma = (double)sms * 8 * 256 * iters * 16 * 2, fl_mma = (double)sms * 8 * 8 * iters * 8 * 512.0;
    const float t_copy = time_ms([&] { copy_rw<<<sms * 16, 256>>>((double2 *)state, N); }, 5);
    const float t_d = time_ms([&] { dense4_dmma<<<sms * 8, 256>>>(state, Mdev, n_groups); }, 5);
    const float t_f = time_ms([&] { dense4_fma<<<sms * 8, 256>>>(state, state2, Mdev, n_groups); }, 5);
    // correctness of the two dense kernels against each other: same input, compare outputs on a sample
    std::vector<double> in(32 * 64), o1(32 * 64), o2(32 * 64);
    for (size_t i = 0; i < in.size(); ++i) in[i] = sin(0.37 * (double)i) * 0.1;
    CK(cudaMemcpy(state, in.data(), in.size() * 8, cudaMemcpyHostToDevice));
    dense4_dmma<<<1, 256>>>(state, Mdev, 64);
    CK(cudaMemcpy(o1.data(), state, in.size() * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(state, in.data(), in.size() * 8, cudaMemcpyHostToDevice));
    dense4_fma<<<1, 256>>>(state, state2, Mdev, 64);
    CK(cudaMemcpy(o2.data(), state2, in.size() * 8, cudaMemcpyDeviceToHost));
    double err = 0, ref_err = 0;
    for (int g = 0; g < 64; ++g)
        for (int r = 0; r < 32; ++r) {
            double s = 0;
            for (int k = 0; k < 32; ++k) s += Mh[r * 32 + k] * in[g * 32 + k];
            err = fmax(err, fabs(o1[g * 32 + r] - o2[g * 32 + r]));
            ref_err = fmax(ref_err, fabs(o2[g * 32 + r] - s));
        }
    const double gb = 32.0 * (double)N / 1e9, flops = 128.0 * (double)N;
    printf("{\"device\": \"%s\", \"sms\": %d, \"log2_amplitudes\": %d,\n", prop.name, sms, nq);
    printf(" \"peak_dfma_tflops\": %.2f, \"peak_dmma_m8n8k4_tflops\": %.2f,\n", fl_fma / t_fma / 1e9, fl_mma / t_mma / 1e9);
    printf(" \"copy_ms\": %.3f, \"copy_GBps\": %.0f,\n", t_copy, gb / t_copy * 1e3);
    printf(" \"dense4_dmma_ms\": %.3f, \"dense4_dmma_GBps\": %.0f, \"dense4_dmma_tflops\": %.2f,\n", t_d, gb / t_d * 1e3, flops / t_d / 1e9);
    printf(" \"dense4_fma_ms\": %.3f, \"dense4_fma_GBps\": %.0f, \"dense4_fma_tflops\": %.2f,\n", t_f, gb / t_f * 1e3, flops / t_f / 1e9);
    printf(" \"dmma_vs_fma_max_abs_diff\": %.3g, \"fma_vs_host_max_abs_diff\": %.3g}\n", err, ref_err);
    return (err < 1e-12 && ref_err < 1e-12) ? 0 : 2;
}
