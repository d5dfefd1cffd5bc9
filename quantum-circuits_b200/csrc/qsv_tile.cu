// Tiled gate pass: the one kernel every gate on the hot path goes through.
//
// Replaces the per-gate body of the reference's Circuit.run_method2 (main/circuit.py:303-327:
// build permutation matrices P1/P2, Kronecker-expand the gate, three dense mat-vecs) with one
// in-place sweep over the complex128 state.  Each CTA owns one TILE of 2^T amplitudes: the
// 2^L lowest index bits (contiguous 16*2^L-byte runs in HBM) plus T-L freely chosen higher bits.
// The tile is brought into shared memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier),
// a whole program of ops acting on the tile's qubits is applied there, and the tile is written
// back with bulk stores.  Controls and diagonal factors on bits OUTSIDE the tile are constants of
// the tile; a tile for which every op is switched off by its external controls is never loaded.
#include "qsv_common.cuh"
#include <string.h>

namespace qsv {

constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy, completion signalled on an mbarrier (TMA, UBLKCP in SASS)
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy, tracked by the bulk async-group
__device__ __forceinline__ void bulk_s2g(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Spread the bits of j over the positions NOT listed in ins (ascending), leaving zeros there.
__device__ __forceinline__ uint32_t expand_free(uint32_t j, uint64_t ins_lo, uint64_t ins_hi, int n_ins) {
#pragma unroll 1
    for (int q = 0; q < n_ins; ++q) {
        unsigned p = (unsigned)(((q < 8) ? (ins_lo >> (8 * q)) : (ins_hi >> (8 * (q - 8)))) & 0xffu);
        j = insert_zero32(j, p);
    }
    return j;
}

struct OpView {
    const qsv_op_t *op;
    uint64_t ins_lo, ins_hi, tgt;
    uint32_t k, n_ins, ctrl, flags;
    const double2 *payload;
};

__device__ __forceinline__ OpView load_op(const uint8_t *p) {
    OpView v;
    v.op = reinterpret_cast<const qsv_op_t *>(p);
    v.k = __ldg(&v.op->k);
    v.n_ins = __ldg(&v.op->n_ins);
    v.flags = __ldg(&v.op->flags);
    v.ctrl = __ldg(&v.op->loc_ctrl_val);
    const uint64_t *ins64 = reinterpret_cast<const uint64_t *>(v.op->ins);
    v.ins_lo = __ldg(ins64);
    v.ins_hi = __ldg(ins64 + 1);
    v.tgt = __ldg(reinterpret_cast<const uint64_t *>(v.op->tgt));
    v.payload = reinterpret_cast<const double2 *>(p + __ldg(&v.op->payload_off));
    return v;
}

// ---- 1-qubit dense: matrix in registers; low targets reorder the two accesses so that every
// quarter-warp (8 x 16 B) touches 8 distinct 16-byte bank groups.
__device__ __forceinline__ void tile_dense1(double2 *tile, const OpView &v, int T, int tid) {
    const double2 m00 = __ldg(v.payload + 0), m01 = __ldg(v.payload + 1);
    const double2 m10 = __ldg(v.payload + 2), m11 = __ldg(v.payload + 3);
    const unsigned b = (unsigned)(v.tgt & 0xffu);
    const uint32_t bit = 1u << b;
    const uint32_t cnt = 1u << (T - (int)v.n_ins);
    const bool real = (v.flags & QSV_OPF_REAL_MATRIX) != 0;
    for (uint32_t j = tid; j < cnt; j += kThreads) {
        const uint32_t i0 = expand_free(j, v.ins_lo, v.ins_hi, v.n_ins) | v.ctrl;
        const uint32_t i1 = i0 | bit;
        const bool sw = (b < 3u) && (j & 4u);
        const uint32_t ia = sw ? i1 : i0, ib = sw ? i0 : i1;
        const double2 xa = tile[ia], xb = tile[ib];
        const double2 x0 = sw ? xb : xa, x1 = sw ? xa : xb;
        double2 y0, y1;
        if (real) {
            y0 = make_double2(fma(m01.x, x1.x, m00.x * x0.x), fma(m01.x, x1.y, m00.x * x0.y));
            y1 = make_double2(fma(m11.x, x1.x, m10.x * x0.x), fma(m11.x, x1.y, m10.x * x0.y));
        } else {
            y0 = cfma(m01, x1, cmul(m00, x0));
            y1 = cfma(m11, x1, cmul(m10, x0));
        }
        tile[ia] = sw ? y1 : y0;
        tile[ib] = sw ? y0 : y1;
    }
}

// ---- k-qubit dense, k = 2..5: gather 2^k amplitudes into registers, row-by-row mat-vec with the
// matrix read through the read-only path (warp-uniform addresses: one broadcast per load).
template <int K>
__device__ __forceinline__ void tile_dense(double2 *tile, const OpView &v, int T, int tid) {
    constexpr int D = 1 << K;
    uint32_t off[D];
#pragma unroll
    for (int c = 0; c < D; ++c) {
        uint32_t o = 0;
#pragma unroll
        for (int p = 0; p < K; ++p)
            if ((c >> (K - 1 - p)) & 1) o |= 1u << (unsigned)((v.tgt >> (8 * p)) & 0xffu);
        off[c] = o;
    }
    const uint32_t cnt = 1u << (T - (int)v.n_ins);
    for (uint32_t j = tid; j < cnt; j += kThreads) {
        const uint32_t base = expand_free(j, v.ins_lo, v.ins_hi, v.n_ins) | v.ctrl;
        double2 x[D];
#pragma unroll
        for (int c = 0; c < D; ++c) x[c] = tile[base | off[c]];
#pragma unroll
        for (int r = 0; r < D; ++r) {
            double2 acc = cmul(__ldg(v.payload + r * D), x[0]);
#pragma unroll
            for (int c = 1; c < D; ++c) acc = cfma(__ldg(v.payload + r * D + c), x[c], acc);
            tile[base | off[r]] = acc;   // safe: every input of this group is already in x[]
        }
    }
}

__device__ __forceinline__ void tile_mcx(double2 *tile, const OpView &v, int T, int tid) {
    const uint32_t bit = 1u << (unsigned)(v.tgt & 0xffu);
    const uint32_t cnt = 1u << (T - (int)v.n_ins);
    for (uint32_t j = tid; j < cnt; j += kThreads) {
        const uint32_t i0 = expand_free(j, v.ins_lo, v.ins_hi, v.n_ins) | v.ctrl;
        const uint32_t i1 = i0 | bit;
        const double2 a = tile[i0], b = tile[i1];
        tile[i0] = b;
        tile[i1] = a;
    }
}

__device__ __forceinline__ void tile_swap(double2 *tile, const OpView &v, int T, int tid) {
    const uint32_t ba = 1u << (unsigned)(v.tgt & 0xffu);
    const uint32_t bb = 1u << (unsigned)((v.tgt >> 8) & 0xffu);
    const uint32_t cnt = 1u << (T - (int)v.n_ins);
    for (uint32_t j = tid; j < cnt; j += kThreads) {
        const uint32_t base = expand_free(j, v.ins_lo, v.ins_hi, v.n_ins) | v.ctrl;
        const double2 a = tile[base | ba], b = tile[base | bb];
        tile[base | ba] = b;
        tile[base | bb] = a;
    }
}

__device__ __forceinline__ void tile_phase(double2 *tile, const OpView &v, int T, int tid) {
    const double2 s = make_double2(__ldg(&v.op->scalar[0]), __ldg(&v.op->scalar[1]));
    const uint32_t cnt = 1u << (T - (int)v.n_ins);
    for (uint32_t j = tid; j < cnt; j += kThreads) {
        const uint32_t i = expand_free(j, v.ins_lo, v.ins_hi, v.n_ins) | v.ctrl;
        tile[i] = cmul(tile[i], s);
    }
}

__device__ __forceinline__ void tile_diag(double2 *tile, const OpView &v, int T, int tid, uint64_t gbase) {
    const int k = (int)v.k;
    uint32_t ext_idx = 0;
    for (int p = 0; p < k; ++p) {
        const unsigned t = (unsigned)((v.tgt >> (8 * p)) & 0xffu);
        if (t & 0x80u) ext_idx |= (uint32_t)((gbase >> (t & 0x7fu)) & 1ull) << (k - 1 - p);
    }
    const uint32_t cnt = 1u << (T - (int)v.n_ins);
    for (uint32_t j = tid; j < cnt; j += kThreads) {
        const uint32_t i = expand_free(j, v.ins_lo, v.ins_hi, v.n_ins) | v.ctrl;
        uint32_t di = ext_idx;
        for (int p = 0; p < k; ++p) {
            const unsigned t = (unsigned)((v.tgt >> (8 * p)) & 0xffu);
            if (!(t & 0x80u)) di |= ((i >> t) & 1u) << (k - 1 - p);
        }
        tile[i] = cmul(tile[i], __ldg(v.payload + di));
    }
}

struct TileFix {          // pinned bits of the tile index, ascending; bit k of `val` is the value of entry k
    int32_t n;
    uint8_t bit[40];
    uint64_t val;
};

template <int KMAX>
__global__ void __launch_bounds__(kThreads, (KMAX <= 3) ? 3 : 1)
k_pass(double2 *__restrict__ state, const __grid_constant__ qsv_pass_t P, const uint8_t *__restrict__ prog,
       const uint32_t *__restrict__ op_off, int n_ops, uint64_t n_tiles, const __grid_constant__ TileFix fix) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint64_t s_hi_off[256];

    const int tid = threadIdx.x;
    const int T = P.tile_bits, L = P.low_bits, n_hi = P.n_hi;
    const uint32_t n_runs = 1u << n_hi;
    const uint32_t run_bytes = 16u << L;

    for (uint32_t r = tid; r < n_runs; r += kThreads) {
        uint64_t o = 0;
        for (int j = 0; j < n_hi; ++j)
            if ((r >> j) & 1u) o |= 1ull << P.hi_pos[j];
        s_hi_off[r] = o;
    }
    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    uint32_t parity = 0;
    // `n_tiles` counts the CANDIDATE tiles: tile-index bits listed in `fix` are pinned (support of the state and/or a
    // control shared by every op), the grid enumerates the others only - no CTA is spent on a tile known to be skipped
    for (uint64_t c = blockIdx.x; c < n_tiles; c += gridDim.x) {
        uint64_t t = c;
        for (int k = 0; k < fix.n; ++k)
            t = insert_zero(t, fix.bit[k]) | (((fix.val >> k) & 1ull) << fix.bit[k]);
        uint64_t base = t << L;
        for (int j = 0; j < n_hi; ++j) base = insert_zero(base, P.hi_pos[j]);
        const uint64_t gbase = base | P.rank_bits;

        // skip tiles that no op touches (external controls all unsatisfied)
        int act = 0;
        for (int i = tid; i < n_ops; i += kThreads) {
            const qsv_op_t *op = reinterpret_cast<const qsv_op_t *>(prog + __ldg(op_off + i));
            act |= ((gbase & __ldg(&op->ext_ctrl_mask)) == __ldg(&op->ext_ctrl_val));
        }
        if (!__syncthreads_or(act)) continue;

        if (tid == 0) mbar_arrive_expect_tx(&mbar, 16u << T);
        if ((tid & 31) == 0)        // one lane per warp issues the bulk copies: operands stay warp-uniform
            for (uint32_t r = tid >> 5; r < n_runs; r += kThreads / 32)
                bulk_g2s(tile + ((size_t)r << L), state + base + s_hi_off[r], run_bytes, &mbar);
        mbar_wait(&mbar, parity);
        parity ^= 1u;

        for (int i = 0; i < n_ops; ++i) {
            const uint8_t *p = prog + __ldg(op_off + i);
            const qsv_op_t *op = reinterpret_cast<const qsv_op_t *>(p);
            if ((gbase & __ldg(&op->ext_ctrl_mask)) != __ldg(&op->ext_ctrl_val)) continue;
            const OpView v = load_op(p);
            switch (__ldg(&op->kind)) {
                case QSV_OP_DENSE:
                    if (v.k == 1) tile_dense1(tile, v, T, tid);
                    else if (v.k == 2) tile_dense<2>(tile, v, T, tid);
                    else if (v.k == 3) tile_dense<3>(tile, v, T, tid);
                    else if (KMAX >= 4 && v.k == 4) tile_dense<(KMAX >= 4 ? 4 : 1)>(tile, v, T, tid);
                    else if (KMAX >= 5 && v.k == 5) tile_dense<(KMAX >= 5 ? 5 : 1)>(tile, v, T, tid);
                    break;
                case QSV_OP_MCX: tile_mcx(tile, v, T, tid); break;
                case QSV_OP_PHASE: tile_phase(tile, v, T, tid); break;
                case QSV_OP_DIAG: tile_diag(tile, v, T, tid, gbase); break;
                case QSV_OP_SWAP: tile_swap(tile, v, T, tid); break;
                default: break;
            }
            __syncthreads();
        }

        fence_proxy_async();   // generic-proxy writes to the tile -> visible to the bulk-copy engine
        __syncthreads();
        if ((tid & 31) == 0)
            for (uint32_t r = tid >> 5; r < n_runs; r += kThreads / 32)
                bulk_s2g(state + base + s_hi_off[r], tile + ((size_t)r << L), run_bytes);
        bulk_commit();
        bulk_wait_read0();     // the tile may be overwritten / the CTA may exit once the reads are done
    }
}

static int validate_pass(const qsv_pass_t *P) {
    QSV_CHECK_ARG(P != nullptr, "qsv_run_pass: pass is NULL");
    QSV_CHECK_ARG(P->n_local >= 1 && P->n_local <= 40, "qsv_run_pass: n_local=%d out of range", P->n_local);
    QSV_CHECK_ARG(P->tile_bits >= 1 && P->tile_bits <= QSV_MAX_TILE_T && P->tile_bits <= P->n_local,
                  "qsv_run_pass: tile_bits=%d invalid for n_local=%d", P->tile_bits, P->n_local);
    QSV_CHECK_ARG(P->low_bits >= 0 && P->low_bits <= P->tile_bits, "qsv_run_pass: low_bits=%d invalid", P->low_bits);
    QSV_CHECK_ARG(P->n_hi == P->tile_bits - P->low_bits && P->n_hi <= 8,
                  "qsv_run_pass: n_hi=%d must equal tile_bits-low_bits and be <= 8", P->n_hi);
    int prev = P->low_bits - 1;
    for (int j = 0; j < P->n_hi; ++j) {
        QSV_CHECK_ARG((int)P->hi_pos[j] > prev && (int)P->hi_pos[j] < P->n_local,
                      "qsv_run_pass: hi_pos[%d]=%d not ascending in [low_bits, n_local)", j, (int)P->hi_pos[j]);
        prev = P->hi_pos[j];
    }
    return 0;
}

int launch_pass(void *state_dev, const qsv_pass_t *P, const void *prog_dev, const uint32_t *op_off_dev, int n_ops,
                bool wide_k, cudaStream_t st, uint64_t fixed_mask, uint64_t fixed_val) {
    if (int rc = validate_pass(P)) return rc;
    QSV_CHECK_ARG(state_dev && prog_dev && op_off_dev, "qsv_run_pass: NULL device pointer");
    if (n_ops <= 0) return 0;
    // index bits the caller pins (support hint / controls common to all ops): global ones decide whether this shard
    // has any work at all, local ones outside the tile shrink the grid; bits inside the tile cannot skip anything
    TileFix fix;
    memset(&fix, 0, sizeof(fix));
    const uint64_t local_mask = (1ull << P->n_local) - 1ull;
    if ((fixed_mask & ~local_mask) && ((P->rank_bits ^ fixed_val) & fixed_mask & ~local_mask)) return 0;
    {
        int idx_bit = 0, j = 0;
        for (int p = P->low_bits; p < P->n_local; ++p) {
            if (j < P->n_hi && P->hi_pos[j] == p) { ++j; continue; }
            if ((fixed_mask >> p) & 1ull) {
                fix.bit[fix.n] = (uint8_t)idx_bit;
                fix.val |= ((fixed_val >> p) & 1ull) << fix.n;
                ++fix.n;
            }
            ++idx_bit;
        }
    }
    const uint64_t n_tiles = 1ull << (P->n_local - P->tile_bits - fix.n);
    const size_t smem = (size_t)16 << P->tile_bits;
    const unsigned grid = (unsigned)(n_tiles < (1ull << 30) ? n_tiles : (1ull << 30));
    static bool attr_done[2] = {false, false};
    if (!attr_done[wide_k]) {
        if (wide_k)
            QSV_CUDA(cudaFuncSetAttribute(k_pass<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          16 << QSV_MAX_TILE_T));
        else
            QSV_CUDA(cudaFuncSetAttribute(k_pass<3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          16 << QSV_MAX_TILE_T));
        attr_done[wide_k] = true;
    }
    if (wide_k)
        k_pass<5><<<grid, kThreads, smem, st>>>((double2 *)state_dev, *P, (const uint8_t *)prog_dev, op_off_dev,
                                                 n_ops, n_tiles, fix);
    else
        k_pass<3><<<grid, kThreads, smem, st>>>((double2 *)state_dev, *P, (const uint8_t *)prog_dev, op_off_dev,
                                                 n_ops, n_tiles, fix);
    QSV_LAUNCH_CHECK();
    return 0;
}

}  // namespace qsv
