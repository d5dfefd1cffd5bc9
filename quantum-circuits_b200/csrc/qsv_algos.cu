// Algorithm-level kernels: Grover oracle + diffusion, Shor modular-exponentiation involution.
#include "qsv_common.cuh"

namespace qsv {

// ---- Grover oracle: main/Grover.py:14-22.  F is the identity except that it exchanges the basis
// states 2x and 2x+1 (X on the ancilla when the search register equals x): an O(1) exact swap.
__global__ void k_swap_pair(double2 *state, uint64_t i0, uint64_t i1) {
    const double2 a = state[i0], b = state[i1];
    state[i0] = b;
    state[i1] = a;
}

// ---- Grover diffusion: main/Grover.py:24-31,40-44.  (H^n (x) I)(U0 (x) I)(H^n (x) I) with
// U0 = diag(+1,-1,...,-1) = 2|0><0| - I is 2|s><s| - I on the search register, separately for every
// value a of the low `anc_bits` bits:  psi'[i,a] = 2*mean_a - psi[i,a].
// Pass 1 (read N): per-CTA partial sums per ancilla value, fixed order -> deterministic.
constexpr int kRedThreads = 256;
constexpr int kMaxAnc = 3;   // up to 8 ancilla values

__global__ void __launch_bounds__(kRedThreads)
k_diff_partial(const double2 *__restrict__ state, uint64_t n_amps, int anc_bits, double2 *__restrict__ partial) {
    // thread t handles elements e = chunk*kRedThreads + t: with kRedThreads a multiple of 2^anc_bits the
    // ancilla value of a thread is fixed = t & (A-1).
    __shared__ double2 sh[kRedThreads];
    const uint64_t per_cta = (n_amps + gridDim.x - 1) / gridDim.x;
    const uint64_t per_cta_al = (per_cta + kRedThreads - 1) / kRedThreads * kRedThreads;
    const uint64_t lo = blockIdx.x * per_cta_al;
    uint64_t hi = lo + per_cta_al;
    if (hi > n_amps) hi = n_amps;
    double2 acc = make_double2(0.0, 0.0);
    for (uint64_t e = lo + threadIdx.x; e < hi; e += kRedThreads) acc = cadd(acc, state[e]);
    sh[threadIdx.x] = acc;
    __syncthreads();
    const int A = 1 << anc_bits;
    for (int s = kRedThreads / 2; s >= A; s >>= 1) {
        if ((int)threadIdx.x < s) sh[threadIdx.x] = cadd(sh[threadIdx.x], sh[threadIdx.x + s]);
        __syncthreads();
    }
    if ((int)threadIdx.x < A) partial[(size_t)blockIdx.x * A + threadIdx.x] = sh[threadIdx.x];
}

__global__ void k_diff_final(const double2 *__restrict__ partial, int n_parts, int anc_bits, double2 *sums) {
    const int a = threadIdx.x;
    const int A = 1 << anc_bits;
    if (a >= A) return;
    double2 acc = make_double2(0.0, 0.0);
    for (int p = 0; p < n_parts; ++p) acc = cadd(acc, partial[(size_t)p * A + a]);
    sums[a] = acc;
}

// Pass 2 (read + write N).
__global__ void k_diff_update(double2 *__restrict__ state, uint64_t n_amps, int anc_bits,
                              const double2 *__restrict__ sums, double scale) {
    __shared__ double2 two_mean[1 << kMaxAnc];
    const int A = 1 << anc_bits;
    if ((int)threadIdx.x < A) {
        const double2 s = sums[threadIdx.x];
        two_mean[threadIdx.x] = make_double2(s.x * scale, s.y * scale);
    }
    __syncthreads();
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < n_amps;
         e += (uint64_t)gridDim.x * blockDim.x) {
        const double2 m = two_mean[e & (uint64_t)(A - 1)];
        const double2 v = state[e];
        state[e] = make_double2(m.x - v.x, m.y - v.y);
    }
}

// ---- Shor modular exponentiation: main/Shor.py:54-83 (same loops as main/Shor_example.py:11-40).
// The reference builds a 2^q x 2^q 0/1 matrix; restated as an index involution it exchanges, for every
// x with f = a^x mod N,   (x,0) <-> (x,f)   and   (x,2^h-1) <-> (x,(2^h-1) & ~f)   and fixes the rest.
__device__ __forceinline__ uint64_t mulmod(uint64_t a, uint64_t b, uint64_t m) {
    return (uint64_t)(((unsigned __int128)a * b) % m);
}
__device__ __forceinline__ uint64_t powmod(uint64_t a, uint64_t e, uint64_t m) {
    uint64_t r = 1 % m;
    a %= m;
    while (e) {
        if (e & 1ull) r = mulmod(r, a, m);
        a = mulmod(a, a, m);
        e >>= 1;
    }
    return r;
}

struct ModexpGeom {
    int32_t nx, h, n_local, n_rest;
    uint8_t x_pos[40];     // x_pos[0] = MSB of x; positions >= n_local come from rank_bits
    uint8_t y_pos[40];     // y_pos[0] = MSB of y (all local)
    uint8_t rest_pos[40];  // ascending local positions that are neither x nor y
};

__global__ void k_modexp(double2 *__restrict__ state, const __grid_constant__ ModexpGeom g, uint64_t a, uint64_t N,
                         uint64_t rank_bits, uint64_t n_work) {
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_work;
         w += (uint64_t)gridDim.x * blockDim.x) {
        // w enumerates (local x bits, rest bits): low n_rest bits = rest, then the local x bits (LSB first)
        uint64_t base = 0;
        uint64_t v = w;
        for (int r = 0; r < g.n_rest; ++r, v >>= 1) base |= (v & 1ull) << g.rest_pos[r];
        uint64_t x = 0;
        for (int p = g.nx - 1; p >= 0; --p) {   // p = nx-1 is the LSB of x
            const int pos = g.x_pos[p];
            uint64_t bit;
            if (pos >= g.n_local) bit = (rank_bits >> pos) & 1ull;
            else { bit = v & 1ull; v >>= 1; base |= bit << pos; }
            x |= bit << (g.nx - 1 - p);
        }
        const uint64_t f = powmod(a, x, N);
        const uint64_t ones = (1ull << g.h) - 1ull;
        const uint64_t nf = ones & ~f;
        uint64_t o_f = 0, o_ones = 0, o_nf = 0;
        for (int p = 0; p < g.h; ++p) {
            const int sh = g.h - 1 - p;
            o_f |= ((f >> sh) & 1ull) << g.y_pos[p];
            o_nf |= ((nf >> sh) & 1ull) << g.y_pos[p];
            o_ones |= 1ull << g.y_pos[p];
        }
        if (f != 0) {
            const double2 u = state[base], t = state[base | o_f];
            state[base] = t;
            state[base | o_f] = u;
        }
        if (f != 0 && f != ones) {   // f == ones: the reference writes the same two entries twice (one swap)
            const double2 u = state[base | o_ones], t = state[base | o_nf];
            state[base | o_ones] = t;
            state[base | o_nf] = u;
        }
    }
}

static unsigned grid_for(uint64_t n, int block) {
    uint64_t g = (n + block - 1) / block;
    const uint64_t cap = 148ull * 16ull;
    return (unsigned)(g < cap ? (g ? g : 1) : cap);
}
static int diff_parts(int n_local) {
    const uint64_t n = 1ull << n_local;
    uint64_t p = n / (kRedThreads * 8);
    if (p < 1) p = 1;
    if (p > 148 * 8) p = 148 * 8;
    return (int)p;
}

}  // namespace qsv

using namespace qsv;

extern "C" int qsv_grover_oracle(void *state_dev, uint64_t idx0, uint64_t idx1, void *stream) {
    QSV_CHECK_ARG(state_dev, "qsv_grover_oracle: NULL state");
    if (idx0 == idx1) return 0;
    k_swap_pair<<<1, 1, 0, (cudaStream_t)stream>>>((double2 *)state_dev, idx0, idx1);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" size_t qsv_diffusion_scratch_bytes(int n_local, int anc_bits) {
    if (n_local < 0 || anc_bits < 0 || anc_bits > kMaxAnc) return 0;
    return (size_t)diff_parts(n_local) * (1u << anc_bits) * sizeof(double2);
}

extern "C" int qsv_diffusion_sums(const void *state_dev, int n_local, int anc_bits, void *scratch_dev,
                                  void *sums_dev, void *stream) {
    QSV_CHECK_ARG(state_dev && scratch_dev && sums_dev, "qsv_diffusion_sums: NULL pointer");
    QSV_CHECK_ARG(anc_bits >= 0 && anc_bits <= kMaxAnc && anc_bits <= n_local,
                  "qsv_diffusion_sums: anc_bits=%d invalid", anc_bits);
    const int parts = diff_parts(n_local);
    k_diff_partial<<<parts, kRedThreads, 0, (cudaStream_t)stream>>>((const double2 *)state_dev, 1ull << n_local,
                                                                    anc_bits, (double2 *)scratch_dev);
    QSV_LAUNCH_CHECK();
    k_diff_final<<<1, 32, 0, (cudaStream_t)stream>>>((const double2 *)scratch_dev, parts, anc_bits,
                                                     (double2 *)sums_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_diffusion_update(void *state_dev, int n_local, int anc_bits, const void *sums_dev,
                                    int n_search_total, void *stream) {
    QSV_CHECK_ARG(state_dev && sums_dev, "qsv_diffusion_update: NULL pointer");
    QSV_CHECK_ARG(anc_bits >= 0 && anc_bits <= kMaxAnc && anc_bits <= n_local,
                  "qsv_diffusion_update: anc_bits=%d invalid", anc_bits);
    QSV_CHECK_ARG(n_search_total >= 0 && n_search_total < 1000, "qsv_diffusion_update: bad n_search_total");
    const uint64_t n = 1ull << n_local;
    const double scale = ldexp(2.0, -n_search_total);   // 2 / 2^n_search, exact
    k_diff_update<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((double2 *)state_dev, n, anc_bits,
                                                                      (const double2 *)sums_dev, scale);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_modexp_involution(void *state_dev, int n_local, int nx, const int *x_pos, int h,
                                     const int *y_pos, uint64_t a, uint64_t N, uint64_t rank_bits, void *stream) {
    QSV_CHECK_ARG(state_dev && x_pos && y_pos, "qsv_modexp_involution: NULL pointer");
    QSV_CHECK_ARG(nx >= 1 && nx <= 40 && h >= 1 && h <= 40, "qsv_modexp_involution: bad register sizes");
    QSV_CHECK_ARG(N >= 2 && N <= (1ull << h), "qsv_modexp_involution: N must satisfy 2 <= N <= 2^h");
    ModexpGeom g;
    g.nx = nx; g.h = h; g.n_local = n_local;
    uint64_t used = 0;
    int n_x_local = 0;
    for (int p = 0; p < nx; ++p) {
        QSV_CHECK_ARG(x_pos[p] >= 0 && x_pos[p] < 64, "qsv_modexp_involution: x position out of range");
        g.x_pos[p] = (uint8_t)x_pos[p];
        if (x_pos[p] < n_local) {
            QSV_CHECK_ARG(!((used >> x_pos[p]) & 1ull), "qsv_modexp_involution: duplicate position");
            used |= 1ull << x_pos[p];
            ++n_x_local;
        }
    }
    for (int p = 0; p < h; ++p) {
        QSV_CHECK_ARG(y_pos[p] >= 0 && y_pos[p] < n_local, "qsv_modexp_involution: y register must be local");
        QSV_CHECK_ARG(!((used >> y_pos[p]) & 1ull), "qsv_modexp_involution: duplicate position");
        used |= 1ull << y_pos[p];
        g.y_pos[p] = (uint8_t)y_pos[p];
    }
    g.n_rest = 0;
    for (int b = 0; b < n_local; ++b)
        if (!((used >> b) & 1ull)) g.rest_pos[g.n_rest++] = (uint8_t)b;
    const uint64_t n_work = 1ull << (n_x_local + g.n_rest);
    k_modexp<<<grid_for(n_work, 128), 128, 0, (cudaStream_t)stream>>>((double2 *)state_dev, g, a, N, rank_bits,
                                                                      n_work);
    QSV_LAUNCH_CHECK();
    return 0;
}
