// Whole-state kernels that do not go through the shared-memory tile: basis-state preparation,
// wide (k > 5) dense / permutation / diagonal gates for the reference's full-width operators, and
// the slice pack/unpack used by the chunked multi-GPU qubit swap.
#include "qsv_common.cuh"

namespace qsv {

// |in_v>: replaces main/circuit.py:290-301 (tensor product of e0/e1 columns).
__global__ void k_set_one(double2 *state, uint64_t idx) { state[idx] = make_double2(1.0, 0.0); }

// All 2^n basis inputs of an n-qubit circuit at once (get_subcircuit, main/circuit.py:122-157): a register of 2n
// qubits whose high half holds the column index c and whose low half starts in |c ^ flip>.
__global__ void k_set_columns(double2 *state, int n, uint64_t flip) {
    const uint64_t c = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (c < (1ull << n)) state[(c << n) | (c ^ flip)] = make_double2(1.0, 0.0);
}

__device__ __forceinline__ uint64_t scatter_bits(uint32_t c, const PosList &t) {
    // bit (k-1-p) of c goes to physical position t.pos[p]  (t.pos[0] = gate MSB = port 0)
    uint64_t o = 0;
    for (int p = 0; p < t.n; ++p) o |= (uint64_t)((c >> (t.n - 1 - p)) & 1u) << t.pos[p];
    return o;
}
__device__ __forceinline__ uint32_t gather_bits(uint64_t i, const PosList &t) {
    uint32_t c = 0;
    for (int p = 0; p < t.n; ++p) c |= (uint32_t)((i >> t.pos[p]) & 1ull) << (t.n - 1 - p);
    return c;
}

// psi'[hi, r, lo] = sum_c M[r][c] psi[hi, c, lo]  — main/circuit.py:325-327 / main/matrix.py:60-70,
// summed over columns in ascending order like the reference.
__global__ void k_dense_wide(const double2 *__restrict__ in, double2 *__restrict__ out, uint64_t n_amps,
                             const __grid_constant__ PosList t, uint64_t tmask, const double2 *__restrict__ M) {
    const uint32_t D = 1u << t.n;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_amps;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t r = gather_bits(i, t);
        const uint64_t base = i & ~tmask;
        const double2 *row = M + (size_t)r * D;
        double2 acc = make_double2(0.0, 0.0);
        for (uint32_t c = 0; c < D; ++c) {
            const double2 m = __ldg(row + c);
            if (m.x == 0.0 && m.y == 0.0) continue;   // exact zeros contribute exactly 0 in the reference too
            acc = cfma(m, in[base | scatter_bits(c, t)], acc);
        }
        out[i] = acc;
    }
}

// psi'[.. perm[c] ..] = psi[.. c ..] — exact data movement for permutation matrices
// (Grover F main/Grover.py:14-22, Shor U main/Shor.py:54-83, X/CNot/Toffoli tensor powers).
__global__ void k_perm_wide(const double2 *__restrict__ in, double2 *__restrict__ out, uint64_t n_amps,
                            const __grid_constant__ PosList t, uint64_t tmask, const uint32_t *__restrict__ perm) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_amps;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = gather_bits(i, t);
        out[(i & ~tmask) | scatter_bits(__ldg(perm + c), t)] = in[i];
    }
}

__global__ void k_diag_wide(double2 *__restrict__ state, uint64_t n_amps, const __grid_constant__ PosList t,
                            const double2 *__restrict__ d) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_amps;
         i += (uint64_t)gridDim.x * blockDim.x)
        state[i] = cmul(state[i], __ldg(d + gather_bits(i, t)));
}

// slice pack/unpack for the chunked all-to-all (16-byte vector copies, fully coalesced)
__global__ void k_pack(const double2 *__restrict__ state, double2 *__restrict__ staging, uint64_t n_chunks,
                       uint64_t stride, uint64_t offset, uint64_t chunk, bool unpack, double2 *state_w) {
    const uint64_t total = n_chunks * chunk;
    for (uint64_t e = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; e < total;
         e += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = e / chunk, w = e - j * chunk;
        if (unpack) state_w[j * stride + offset + w] = staging[e];
        else staging[e] = state[j * stride + offset + w];
    }
}

static unsigned grid_for(uint64_t n, int block) {
    uint64_t g = (n + block - 1) / block;
    const uint64_t cap = 148ull * 32ull;
    return (unsigned)(g < cap ? (g ? g : 1) : cap);
}

static int fill_pos(PosList &t, int n_local, int k, const int *targets, uint64_t *mask, const char *who) {
    QSV_CHECK_ARG(k >= 1 && k <= 30 && k <= n_local, "%s: k=%d invalid for n_local=%d", who, k, n_local);
    t.n = k;
    *mask = 0;
    for (int p = 0; p < k; ++p) {
        QSV_CHECK_ARG(targets[p] >= 0 && targets[p] < n_local, "%s: target %d out of range", who, targets[p]);
        QSV_CHECK_ARG(!((*mask >> targets[p]) & 1ull), "%s: duplicate target %d", who, targets[p]);
        *mask |= 1ull << targets[p];
        t.pos[p] = (uint8_t)targets[p];
    }
    return 0;
}

}  // namespace qsv

using namespace qsv;

extern "C" int qsv_init_basis(void *state_dev, int n_local, uint64_t index, uint64_t rank, void *stream) {
    QSV_CHECK_ARG(state_dev, "qsv_init_basis: NULL state");
    QSV_CHECK_ARG(n_local >= 0 && n_local <= 40, "qsv_init_basis: n_local=%d out of range", n_local);
    cudaStream_t st = (cudaStream_t)stream;
    QSV_CUDA(cudaMemsetAsync(state_dev, 0, (size_t)16 << n_local, st));
    if ((index >> n_local) == rank) {
        k_set_one<<<1, 1, 0, st>>>((double2 *)state_dev, index & ((1ull << n_local) - 1ull));
        QSV_LAUNCH_CHECK();
    }
    return 0;
}

// |index> WITHOUT clearing the register: only the basis amplitude is written.  The passes that follow must run in the
// zero-fill form (qsv_run_pass_rounds_fresh) until one of them has touched every tile.
extern "C" int qsv_init_basis_sparse(void *state_dev, int n_local, uint64_t index, uint64_t rank, void *stream) {
    QSV_CHECK_ARG(state_dev, "qsv_init_basis_sparse: NULL state");
    QSV_CHECK_ARG(n_local >= 0 && n_local <= 40, "qsv_init_basis_sparse: n_local=%d out of range", n_local);
    if ((index >> n_local) == rank) {
        k_set_one<<<1, 1, 0, (cudaStream_t)stream>>>((double2 *)state_dev, index & ((1ull << n_local) - 1ull));
        QSV_LAUNCH_CHECK();
    }
    return 0;
}

extern "C" int qsv_init_columns(void *state_dev, int n, uint64_t flip, void *stream) {
    QSV_CHECK_ARG(state_dev, "qsv_init_columns: NULL state");
    QSV_CHECK_ARG(n >= 1 && n <= 15, "qsv_init_columns: n=%d out of range (the batch register has 2n qubits)", n);
    QSV_CHECK_ARG(flip < (1ull << n), "qsv_init_columns: flip does not fit n bits");
    cudaStream_t st = (cudaStream_t)stream;
    QSV_CUDA(cudaMemsetAsync(state_dev, 0, (size_t)16 << (2 * n), st));
    const uint64_t cols = 1ull << n;
    k_set_columns<<<(unsigned)((cols + 255) / 256), 256, 0, st>>>((double2 *)state_dev, n, flip);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_apply_dense_wide(const void *state_in_dev, void *state_out_dev, int n_local, int k,
                                    const int *targets, const void *matrix_dev, void *stream) {
    QSV_CHECK_ARG(state_in_dev && state_out_dev && matrix_dev && targets, "qsv_apply_dense_wide: NULL pointer");
    QSV_CHECK_ARG(state_in_dev != state_out_dev, "qsv_apply_dense_wide: must be out of place");
    PosList t; uint64_t mask;
    if (int rc = fill_pos(t, n_local, k, targets, &mask, "qsv_apply_dense_wide")) return rc;
    const uint64_t n = 1ull << n_local;
    k_dense_wide<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const double2 *)state_in_dev,
        (double2 *)state_out_dev, n, t, mask, (const double2 *)matrix_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_apply_perm_wide(const void *state_in_dev, void *state_out_dev, int n_local, int k,
                                   const int *targets, const uint32_t *perm_dev, void *stream) {
    QSV_CHECK_ARG(state_in_dev && state_out_dev && perm_dev && targets, "qsv_apply_perm_wide: NULL pointer");
    QSV_CHECK_ARG(state_in_dev != state_out_dev, "qsv_apply_perm_wide: must be out of place");
    PosList t; uint64_t mask;
    if (int rc = fill_pos(t, n_local, k, targets, &mask, "qsv_apply_perm_wide")) return rc;
    const uint64_t n = 1ull << n_local;
    k_perm_wide<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const double2 *)state_in_dev,
        (double2 *)state_out_dev, n, t, mask, perm_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_apply_diag_wide(void *state_dev, int n_local, int k, const int *targets, const void *diag_dev,
                                   void *stream) {
    QSV_CHECK_ARG(state_dev && diag_dev && targets, "qsv_apply_diag_wide: NULL pointer");
    PosList t; uint64_t mask;
    if (int rc = fill_pos(t, n_local, k, targets, &mask, "qsv_apply_diag_wide")) return rc;
    const uint64_t n = 1ull << n_local;
    k_diag_wide<<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((double2 *)state_dev, n, t,
                                                                     (const double2 *)diag_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_pack_slices(const void *state_dev, void *staging_dev, uint64_t n_chunks, uint64_t stride_elems,
                               uint64_t offset_elems, uint64_t chunk_elems, void *stream) {
    QSV_CHECK_ARG(state_dev && staging_dev, "qsv_pack_slices: NULL pointer");
    QSV_CHECK_ARG(chunk_elems > 0 && offset_elems + chunk_elems <= stride_elems, "qsv_pack_slices: bad slice");
    k_pack<<<grid_for(n_chunks * chunk_elems, 256), 256, 0, (cudaStream_t)stream>>>(
        (const double2 *)state_dev, (double2 *)staging_dev, n_chunks, stride_elems, offset_elems, chunk_elems, false,
        nullptr);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_unpack_slices(void *state_dev, const void *staging_dev, uint64_t n_chunks, uint64_t stride_elems,
                                 uint64_t offset_elems, uint64_t chunk_elems, void *stream) {
    QSV_CHECK_ARG(state_dev && staging_dev, "qsv_unpack_slices: NULL pointer");
    QSV_CHECK_ARG(chunk_elems > 0 && offset_elems + chunk_elems <= stride_elems, "qsv_unpack_slices: bad slice");
    k_pack<<<grid_for(n_chunks * chunk_elems, 256), 256, 0, (cudaStream_t)stream>>>(
        nullptr, (double2 *)staging_dev, n_chunks, stride_elems, offset_elems, chunk_elems, true,
        (double2 *)state_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}
