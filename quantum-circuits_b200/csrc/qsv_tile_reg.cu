// Register-resident tiled gate pass (k_pass_reg): the fused-circuit workhorse.
//
// Same tile geometry and TMA bulk load/store as k_pass (qsv_tile.cu), but between load and store the
// tile lives in REGISTERS: 256 threads x 16 complex128.  A pass is split by the host into rounds; in a
// round, four tile bits are "register bits" (each thread owns the 16 amplitudes that differ in them), so
// 1-qubit dense gates, (multi-)controlled X and phases whose target is a register bit are straight-line
// register arithmetic: no shared-memory traffic, no barrier.  Between rounds the tile is re-distributed
// through shared memory (one store + one load per amplitude).  In the shared-memory kernel every op costs
// a full tile read+write of shared memory and a barrier, which made 30-gate passes 6x slower than the HBM
// time of the pass; here the cost is per round, not per op.
#include "qsv_common.cuh"
#include <string.h>
#include <stdlib.h>

namespace qsv {

#ifndef QSV_REG_CTAS
#define QSV_REG_CTAS 2
#endif

__device__ __forceinline__ uint32_t smem_u32r(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init_r(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32r(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_r(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32r(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_r(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done) : "r"(smem_u32r(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s_r(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32r(smem_dst)), "l"(gmem_src), "r"(bytes), "r"(smem_u32r(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g_r(void *gmem_dst, const void *smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst),
                 "r"(smem_u32r(smem_src)), "r"(bytes) : "memory");
}

// ---- register ops: everything is unrolled over the 16 register slots, r is a compile-time constant.
// The common cases are STRAIGHT-LINE code selected by the op's kind byte (no per-slot predicates, no
// selects): the first profile of this kernel showed the ALU pipe saturated by FSEL/ISETP/LOP3 that the
// compiler generated for dynamic per-slot conditions.  Encodings (see qsv/schedule.py encode_rounds):
//   0x10+J        dense 1-qubit, complex matrix, target register bit J, no register-bit condition
//   0x14+J        same, real matrix
//   0x80          (multi-)controlled X: NOT a register op.  It is an index permutation, applied to the
//                 shared-memory addresses when a round loads / stores its registers (zero data movement;
//                 register swaps cost ~100 MOV per op and thread in the first version of this kernel)
//   0x60          phase with no register-bit dependence: folded into the thread's pending scalar
//   0x70+2J+V     phase on the slots whose register bit J == V
//   0xF3/0xF4     generic predicated phase / diagonal table (any rmask/rval)

// In-place complex scale x <- x * s (5 FP64 instructions, one temporary, no register moves).
__device__ __forceinline__ void cscale_inplace(double2 &x, const double2 s) {
    const double t = s.y * x.x;
    x.x = x.x * s.x;
    x.x = fma(-s.y, x.y, x.x);
    x.y = x.y * s.x;
    x.y = x.y + t;
}

// 2x2 update written so that every result lands in the register that held its last operand: the
// interpreter's switch forces x[] into fixed registers at every merge point, and the natural
// "compute y0,y1 then assign" form cost 6-8 MOV per pair (60 % of all issued instructions in the first
// profile).   y0 = a x0 + b x1,  y1 = c x0 + d x1.
template <int J, bool REAL, int NS>
__device__ __forceinline__ void f_dense1(double2 (&x)[NS], const qsv_rop_t *op) {
    const double2 a = make_double2(op->m[0], op->m[1]), b = make_double2(op->m[2], op->m[3]);
    const double2 c = make_double2(op->m[4], op->m[5]), d = make_double2(op->m[6], op->m[7]);
#pragma unroll
    for (int r = 0; r < NS; ++r) {
        if ((r >> J) & 1) continue;
        double2 &x0 = x[r], &x1 = x[r | (1 << J)];
        if (REAL) {
            const double tr = b.x * x1.x, ti = b.x * x1.y;     // b*x1 before x1 is overwritten
            x1.x = d.x * x1.x;
            x1.y = d.x * x1.y;
            x1.x = fma(c.x, x0.x, x1.x);
            x1.y = fma(c.x, x0.y, x1.y);
            x0.x = fma(a.x, x0.x, tr);
            x0.y = fma(a.x, x0.y, ti);
        } else {
            double tr = b.x * x1.x, ti = b.x * x1.y;
            tr = fma(-b.y, x1.y, tr);
            ti = fma(b.y, x1.x, ti);
            cscale_inplace(x1, d);
            x1.x = fma(c.x, x0.x, x1.x);
            x1.x = fma(-c.y, x0.y, x1.x);
            x1.y = fma(c.x, x0.y, x1.y);
            x1.y = fma(c.y, x0.x, x1.y);
            const double w = a.y * x0.x;
            x0.x = fma(a.x, x0.x, tr);
            x0.x = fma(-a.y, x0.y, x0.x);
            x0.y = fma(a.x, x0.y, ti);
            x0.y = x0.y + w;
        }
    }
}

template <int J, int V, int NS>
__device__ __forceinline__ void f_phase1(double2 (&x)[NS], const qsv_rop_t *op) {
    const double2 s = make_double2(op->m[0], op->m[1]);
#pragma unroll
    for (int r = 0; r < NS; ++r)
        if (((r >> J) & 1) == V) cscale_inplace(x[r], s);
}

// generic predicated forms (register-bit conditions of any shape)
template <int NS>
__device__ __forceinline__ void r_phase(double2 (&x)[NS], const qsv_rop_t *op) {
    const uint32_t rmask = op->rmask, rval = op->rval;
    const double2 s = make_double2(op->m[0], op->m[1]);
#pragma unroll
    for (int r = 0; r < NS; ++r)
        if ((r & rmask) == rval) x[r] = cmul(x[r], s);
}

template <int NS>
__device__ __forceinline__ void r_diag(double2 (&x)[NS], const qsv_rop_t *op, const uint8_t *sprog, uint32_t tb,
                                       uint64_t gbase) {
    const int k = (int)op->k;
    uint32_t di = 0;
    for (int p = 0; p < k; ++p) {
        const uint32_t s = op->dsrc[p];
        if (s == 0xffu) continue;
        const uint32_t bit = (s & 0x80u) ? (uint32_t)((gbase >> (s & 0x7fu)) & 1ull) : ((tb >> s) & 1u);
        di |= bit << (k - 1 - p);
    }
    const double2 *tab = reinterpret_cast<const double2 *>(sprog + op->table_off);
    const uint32_t c0 = op->rc[0], c1 = op->rc[1], c2 = op->rc[2], c3 = op->rc[3];
    const uint32_t rmask = op->rmask, rval = op->rval;
#pragma unroll
    for (int r = 0; r < NS; ++r) {
        if ((r & rmask) != rval) continue;
        const uint32_t d = di | ((r & 1) ? c0 : 0u) | ((r & 2) ? c1 : 0u) | ((r & 4) ? c2 : 0u) | ((r & 8) ? c3 : 0u);
        cscale_inplace(x[r], __ldg(tab + d));
    }
}

// kinds 0..17 are contiguous; register bits >= log2(NS) do not exist in the 8-slot variant
#define QSV_CASE_J(base, J_, CALL) \
    case (base) + (J_): if constexpr ((1 << (J_)) < NS) { constexpr int J = (J_); CALL; } break;
#define QSV_CASE_J4(base, CALL) \
    QSV_CASE_J(base, 0, CALL) QSV_CASE_J(base, 1, CALL) QSV_CASE_J(base, 2, CALL) QSV_CASE_J(base, 3, CALL)
#define QSV_CASE_PH1(J_) \
    case QSV_ROP_PHASE1 + 2 * (J_) + 0: if constexpr ((1 << (J_)) < NS) f_phase1<J_, 0, NS>(x, op); break; \
    case QSV_ROP_PHASE1 + 2 * (J_) + 1: if constexpr ((1 << (J_)) < NS) f_phase1<J_, 1, NS>(x, op); break;

template <int NS>
__device__ __forceinline__ void apply_rop(double2 (&x)[NS], const qsv_rop_t *op, uint32_t kind,
                                          const uint8_t *sprog, uint32_t tb, uint64_t gbase) {
    switch (kind) {
        QSV_CASE_J4(QSV_ROP_DENSE1, (f_dense1<J, false, NS>(x, op)))
        QSV_CASE_J4(QSV_ROP_DENSE1_REAL, (f_dense1<J, true, NS>(x, op)))
        QSV_CASE_PH1(0) QSV_CASE_PH1(1) QSV_CASE_PH1(2) QSV_CASE_PH1(3)
        case QSV_ROP_GEN_PHASE: r_phase<NS>(x, op); break;
        case QSV_ROP_DIAG: r_diag<NS>(x, op, sprog, tb, gbase); break;
        default: break;
    }
}

// ---- address permutations at round boundaries -----------------------------------------------------
// The 16 slot addresses of a thread are kept in XOR-linear form  address(r) = base ^ XOR_{q in r} d_q.
// A (multi-)controlled X whose condition does not depend on the slot flips `base`; one whose condition
// reads a single register-dependent bit flips d_q (and maybe base): 4-10 instructions per op and thread
// instead of 3 per op and SLOT.  The host proves linearity (qsv/schedule.py linearize_perms); anything
// else runs the per-slot general form.
struct AddrState { uint32_t base, d0, d1, d2, d3; };

__device__ __forceinline__ void perm_linear(AddrState &A, const qsv_rop_t *op, uint64_t gbase) {
    // branch-free so that an unrolled list can issue the loads of several ops ahead of the (serial) updates
    const uint32_t cm = op->tmask, cv = op->tval, xm = op->table_off, q = op->j;
    const uint32_t pbit = op->flags;                    // the one condition bit that may vary with r_q (0: none)
    const bool ext_ok = !(op->flags2 & 1u) || (gbase & op->ext_mask) == op->ext_val;
    const uint32_t cmu = cm & ~pbit;
    const uint32_t f = (ext_ok && (A.base & cmu) == (cv & cmu)) ? xm : 0u;
    const uint32_t dq = (q == 0u) ? A.d0 : (q == 1u) ? A.d1 : (q == 2u) ? A.d2 : (q == 3u) ? A.d3 : 0u;
    const uint32_t fd = (dq & pbit) ? f : 0u;           // slots with r_q = 1 see the bit toggled
    A.base ^= (((A.base ^ cv) & pbit) == 0u) ? f : 0u;  // the r_q = 0 slots (or all slots) satisfy the condition
    A.d0 ^= (q == 0u) ? fd : 0u;
    A.d1 ^= (q == 1u) ? fd : 0u;
    A.d2 ^= (q == 2u) ? fd : 0u;
    A.d3 ^= (q == 3u) ? fd : 0u;
}

template <int NS>
__device__ __forceinline__ void slot_addresses(uint32_t (&a)[NS], const AddrState &A) {
    a[0] = A.base;
#pragma unroll
    for (int r = 1; r < NS; ++r) {
        const int low = r & -r;                          // lowest set bit of r
        a[r] = a[r & (r - 1)] ^ (low == 1 ? A.d0 : low == 2 ? A.d1 : low == 4 ? A.d2 : A.d3);
    }
}

template <int NS>
__device__ __forceinline__ void perm_general(uint32_t (&a)[NS], const qsv_rop_t *op, uint64_t gbase) {
    if ((gbase & op->ext_mask) != op->ext_val) return;
    const uint32_t cm = op->tmask, cv = op->tval, xm = op->table_off;
#pragma unroll
    for (int r = 0; r < NS; ++r) a[r] ^= ((a[r] & cm) == cv) ? xm : 0u;
}

// The pass program travels as a __grid_constant__ kernel parameter: it lives in the constant bank, op
// fields and matrix entries are fetched through the uniform datapath (LDCU/ULDC -> uniform registers or
// direct constant operands), which takes ~25 vector registers out of the interpreter's hot loop.
struct RProgParam {
    qsv_rprog_hdr_t hdr;
    qsv_round_t rounds[QSV_RPROG_MAX_ROUNDS];
    qsv_rop_t ops[QSV_RPROG_MAX_OPS];
};

template <int RB>
__global__ void __launch_bounds__(4096 >> RB, QSV_REG_CTAS)
k_pass_reg(double2 *__restrict__ state, const __grid_constant__ qsv_pass_t P,
           const __grid_constant__ RProgParam prog, const uint8_t *__restrict__ tables, uint64_t n_tiles) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double2 *tile = reinterpret_cast<double2 *>(smem_raw);
    __shared__ __align__(8) uint64_t mbar;
    __shared__ uint64_t s_hi_off[256];

    constexpr int NS = 1 << RB;                    // amplitudes per thread
    constexpr int kRThreads = 4096 >> RB;          // 256 threads x 16 slots, or 512 x 8
    const int tid = threadIdx.x;
    const int T = P.tile_bits, L = P.low_bits, n_hi = P.n_hi;
    const uint8_t *sprog = tables;      // diagonal tables stay in global memory (read-only path)
    // The program arrives in the constant bank (kernel parameter); every CTA copies the used part to
    // shared memory once: indexed constant loads (LDC c[0x0][R+imm]) cost a dependent ~100+ cycle round
    // trip per field, three to four of them per op, which made every op cost ~600 cycles per tile.
    const uint32_t n_runs = 1u << n_hi;
    const uint32_t run_bytes = 16u << L;
    const uint32_t n_active = 1u << (T - RB);      // threads that own amplitudes

    for (uint32_t r = tid; r < n_runs; r += kRThreads) {
        uint64_t o = 0;
        for (int j = 0; j < n_hi; ++j)
            if ((r >> j) & 1u) o |= 1ull << P.hi_pos[j];
        s_hi_off[r] = o;
    }
    if (tid == 0) {
        mbar_init_r(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int n_rounds = (int)prog.hdr.n_rounds, n_ops = (int)prog.hdr.n_ops;
    const qsv_round_t *rounds = prog.rounds;
    const qsv_rop_t *ops = prog.ops;

    uint32_t parity = 0;
    for (uint64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        uint64_t base = t << L;
        for (int j = 0; j < n_hi; ++j) base = insert_zero(base, P.hi_pos[j]);
        const uint64_t gbase = base | P.rank_bits;

        int act = 0;
        for (int i = tid; i < n_ops; i += kRThreads) act |= ((gbase & ops[i].ext_mask) == ops[i].ext_val);
        if (!__syncthreads_or(act)) continue;

        if (tid == 0) mbar_arrive_expect_tx_r(&mbar, 16u << T);
        if ((tid & 31) == 0)        // one lane per warp issues the bulk copies: operands stay warp-uniform
            for (uint32_t r = tid >> 5; r < n_runs; r += kRThreads / 32)
                bulk_g2s_r(tile + ((size_t)r << L), state + base + s_hi_off[r], run_bytes, &mbar);
        mbar_wait_r(&mbar, parity);
        parity ^= 1u;

        for (int rd = 0; rd < n_rounds; ++rd) {
            const qsv_round_t R = rounds[rd];
            uint32_t tb = (uint32_t)tid;
#pragma unroll
            for (int q = 0; q < RB; ++q) tb = insert_zero32(tb, R.rb[q]);
            const uint32_t n_pre = R.n_pre & 0x7fffu, n_post = R.n_post & 0x7fffu;
            const uint32_t n_reg = R.n_ops & 0xffffu;            // ordered register ops
            const uint32_t n_pt = (R.n_ops >> 16) & 0x3fffu;     // thread-level phases (commute with everything)
            if (rd > 0) __syncthreads();           // the previous round's stores are visible
            double2 x[NS];
            // Threads beyond 2^(T-4) own no amplitudes (tiles smaller than 4096): they run the same
            // UNIFORM control flow on zeros and only their loads/stores are predicated off.  Keeping the
            // op loop out of any thread-dependent branch lets the compiler keep the loop counter, the op
            // pointer and the op fields in uniform registers (ULDC) instead of chains of indexed LDC.
            const bool owner = (uint32_t)tid < n_active;
            {
                // ---- load: slot r <- tile[address(r)], address(r) = base ^ XOR_{q in r} d_q.  Leading
                // X / CNot / Toffoli ops permute the addresses: new[a] = old[p_1(p_2(...p_k(a)))].
                uint32_t a[NS];
                AddrState A = {tb, 1u << R.rb[0], 1u << R.rb[1], 1u << R.rb[2], RB > 3 ? 1u << R.rb[3] : 0u};
                if (!(R.n_pre & 0x8000u)) {
#pragma unroll 4
                    for (int i = (int)n_pre - 1; i >= 0; --i) perm_linear(A, ops + R.first_op + i, gbase);
                    slot_addresses<NS>(a, A);
                } else {
                    slot_addresses<NS>(a, A);
                    for (int i = (int)n_pre - 1; i >= 0; --i) perm_general<NS>(a, ops + R.first_op + i, gbase);
                }
#pragma unroll
                for (int r = 0; r < NS; ++r) x[r] = owner ? tile[a[r]] : make_double2(0.0, 0.0);
            }
            if (n_pt) {
                // thread-level phases: one pending scalar per thread, applied once to all 16 slots
                // two independent accumulators and loads issued ahead of their use: the loop is a chain of
                // dependent shared/constant loads otherwise (~300 cycles per op with 2 warps per scheduler)
                double2 S = make_double2(1.0, 0.0), S2 = make_double2(1.0, 0.0);
                const qsv_rop_t *pt = ops + R.first_op + n_pre;
                uint32_t i = 0;
                for (; i + 1 < n_pt; i += 2) {
                    const qsv_rop_t *o0 = pt + i, *o1 = pt + i + 1;
                    const uint32_t tm0 = o0->tmask, tv0 = o0->tval, f0 = o0->flags2;
                    const uint32_t tm1 = o1->tmask, tv1 = o1->tval, f1 = o1->flags2;
                    const double2 m0 = make_double2(o0->m[0], o0->m[1]);
                    const double2 m1 = make_double2(o1->m[0], o1->m[1]);
                    const bool e0 = !(f0 & 1u) || (gbase & o0->ext_mask) == o0->ext_val;
                    const bool e1 = !(f1 & 1u) || (gbase & o1->ext_mask) == o1->ext_val;
                    if (e0 && (tb & tm0) == tv0) S = cmul(S, m0);
                    if (e1 && (tb & tm1) == tv1) S2 = cmul(S2, m1);
                }
                if (i < n_pt) {
                    const qsv_rop_t *o0 = pt + i;
                    const bool e0 = !(o0->flags2 & 1u) || (gbase & o0->ext_mask) == o0->ext_val;
                    if (e0 && (tb & o0->tmask) == o0->tval) S = cmul(S, make_double2(o0->m[0], o0->m[1]));
                }
                S = cmul(S, S2);
#pragma unroll
                for (int r = 0; r < NS; ++r) cscale_inplace(x[r], S);
            }
            {
                const qsv_rop_t *rg = ops + R.first_op + n_pre + n_pt;
                // software pipelining of the op decode: the header of op i+1 is fetched while op i runs
                uint4 hn = make_uint4(0u, 0u, 0u, 0u);
                if (n_reg) hn = *reinterpret_cast<const uint4 *>(rg);
                for (uint32_t i = 0; i < n_reg; ++i) {
                    const uint4 h = hn;
                    if (i + 1 < n_reg) hn = *reinterpret_cast<const uint4 *>(rg + i + 1);
                    if ((h.w & 1u) && (gbase & rg[i].ext_mask) != rg[i].ext_val) continue;     // uniform
                    if ((tb & h.y) == h.z) apply_rop<NS>(x, rg + i, h.x & 0xffu, sprog, tb, gbase);
                }
            }
            if (n_pre + n_post) __syncthreads();   // permuted stores land in other threads' slots: all loads first
            {
                // ---- store: tile[p_k(...p_1(address(r)))] <- slot r
                uint32_t a[NS];
                AddrState A = {tb, 1u << R.rb[0], 1u << R.rb[1], 1u << R.rb[2], RB > 3 ? 1u << R.rb[3] : 0u};
                const qsv_rop_t *post = ops + R.first_op + n_pre + n_pt + n_reg;
                if (!(R.n_post & 0x8000u)) {
#pragma unroll 4
                    for (uint32_t i = 0; i < n_post; ++i) perm_linear(A, post + i, gbase);
                    slot_addresses<NS>(a, A);
                } else {
                    slot_addresses<NS>(a, A);
                    for (uint32_t i = 0; i < n_post; ++i) perm_general<NS>(a, post + i, gbase);
                }
                if (owner) {
#pragma unroll
                    for (int r = 0; r < NS; ++r) tile[a[r]] = x[r];
                }
            }
        }

        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if ((tid & 31) == 0)
            for (uint32_t r = tid >> 5; r < n_runs; r += kRThreads / 32)
                bulk_s2g_r(state + base + s_hi_off[r], tile + ((size_t)r << L), run_bytes);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

}  // namespace qsv

using namespace qsv;

namespace qsv { namespace jit {
bool wanted(const qsv_pass_t *P, const void *prog_host);
int launch(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes, const void *tables_dev,
           cudaStream_t st, unsigned grid, unsigned long long sp_mask, unsigned long long sp_val, unsigned zf_mask,
           unsigned zf_val);
int run_sums(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes, const void *tables_dev,
             cudaStream_t st, unsigned grid, unsigned long long sp_mask, unsigned long long sp_val, unsigned zf_mask,
             unsigned zf_val, int n_bits, const int *positions, void *bk_dev);
} }

static_assert(sizeof(qsv_rop_t) == 128, "qsv_rop_t must be 128 bytes");
static_assert(sizeof(qsv_round_t) == 16 && sizeof(qsv_rprog_hdr_t) == 16, "round/program headers are 16 bytes");
static_assert(sizeof(qsv_op_t) == 88 && sizeof(qsv_pass_t) == 48, "struct layout mirrored in qsv/schedule.py");

static int sm_count() {
    static int cached[64] = {0};           // per device ordinal (benign race: every writer stores the same value)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < 64 && cached[dev] > 0) return cached[dev];
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < 64) cached[dev] = n;
    return n;
}

static int run_pass_rounds(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes,
                           const void *tables_dev, void *stream, uint64_t sp_mask, uint64_t sp_val,
                           uint32_t zf_mask = 0, uint32_t zf_val = 0);

extern "C" int qsv_run_pass_rounds(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes,
                                   const void *tables_dev, void *stream) {
    return run_pass_rounds(state_dev, P, prog_host, prog_bytes, tables_dev, stream, 0, 0);
}

extern "C" int qsv_run_pass_rounds_sparse(void *state_dev, const qsv_pass_t *P, const void *prog_host,
                                          uint32_t prog_bytes, const void *tables_dev, uint64_t fixed_mask,
                                          uint64_t fixed_val, void *stream) {
    QSV_CHECK_ARG((fixed_val & ~fixed_mask) == 0, "qsv_run_pass_rounds_sparse: fixed_val has bits outside fixed_mask");
    return run_pass_rounds(state_dev, P, prog_host, prog_bytes, tables_dev, stream, fixed_mask, fixed_val);
}

// Zero-fill form (qsv_init_basis_sparse): the register was NOT cleared, only the basis amplitude was written.  Memory
// outside the support of the state is uninitialised; the pass reads the tiles of the support (fixed_mask / fixed_val over
// the positions outside the tile, as in the sparse form) and takes every amplitude whose TILE-LOCAL index disagrees with
// (zf_mask, zf_val) - the support's fixed bits inside the tile - as zero.  Every tile it touches is written in full,
// so after the first pass that touches all tiles the register is an ordinary dense state.  Replaces the 16 * 2^n byte
// memset of main/circuit.py:290-301's initial vector (17 GB = 2.6 ms of a 63 ms step at 30 qubits).  Only the
// pass-specialised kernel implements it.
extern "C" int qsv_run_pass_rounds_fresh(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes,
                                         const void *tables_dev, uint64_t fixed_mask, uint64_t fixed_val,
                                         uint32_t zf_mask, uint32_t zf_val, void *stream) {
    QSV_CHECK_ARG((fixed_val & ~fixed_mask) == 0, "qsv_run_pass_rounds_fresh: fixed_val has bits outside fixed_mask");
    QSV_CHECK_ARG((zf_val & ~zf_mask) == 0 && P && zf_mask < (1u << P->tile_bits),
                  "qsv_run_pass_rounds_fresh: zf_val / zf_mask must be tile-local bit masks");
    QSV_CHECK_ARG(P->reserved == 0 || P->reserved == 4, "qsv_run_pass_rounds_fresh: needs 4 register bits per round");
    QSV_CHECK_ARG(prog_host && jit::wanted(P, prog_host),
                  "qsv_run_pass_rounds_fresh: only the pass-specialised kernel zero-fills (QSV_JIT / register size)");
    return run_pass_rounds(state_dev, P, prog_host, prog_bytes, tables_dev, stream, fixed_mask, fixed_val, zf_mask | 0x80000000u,
                           zf_val);
}

static unsigned jit_grid(const qsv_pass_t *P) {
    const uint64_t n_tiles_j = 1ull << (P->n_local - P->tile_bits);
    // persistent CTAs, 2 resident per SM; the grid holds QSV_JIT_WAVES (default 16; measured 4 -> 16: single-gate sweep 5.45 -> 5.2 ms) times as many, so that the
    // hardware scheduler evens out SMs that run at different speeds (the tail of a static split is one CTA long)
    static const unsigned waves = [] {
        const char *e = getenv("QSV_JIT_WAVES");
        const int v = e ? atoi(e) : 16;
        return (unsigned)(v >= 1 && v <= 4096 ? v : 16);
    }();
    const uint64_t resident_j = (uint64_t)sm_count() * 2ull * waves;
    return (unsigned)(n_tiles_j < resident_j ? n_tiles_j : resident_j);
}

// The LAST pass of a run that ends in a sample: qsv_run_pass_rounds_sparse / _fresh (zf_mask = 0: plain sparse form) that
// also adds every tile's total weight - exact 2^-96 fixed point, the sampler's arithmetic - to bucket
// hb = index bits at bucket_pos[0..n_bits) of `buckets_dev` (4 << n_bits uint64 limbs, zeroed by this call): the first
// level of qsv_sample_level's radix descent without its read of the whole state.  bucket_pos are physical positions
// outside the pass's tile (global positions allowed: they come from pass->rank_bits).  Pass-specialised kernel only.
extern "C" int qsv_run_pass_rounds_sums(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes,
                                        const void *tables_dev, uint64_t fixed_mask, uint64_t fixed_val, uint32_t zf_mask,
                                        uint32_t zf_val, int n_bits, const int *bucket_pos, void *buckets_dev,
                                        void *stream) {
    QSV_CHECK_ARG(state_dev && P && prog_host && buckets_dev && bucket_pos, "qsv_run_pass_rounds_sums: NULL pointer");
    QSV_CHECK_ARG((fixed_val & ~fixed_mask) == 0 && (zf_val & ~zf_mask) == 0, "qsv_run_pass_rounds_sums: value bits outside mask");
    QSV_CHECK_ARG(P->n_local >= 12 && P->n_local <= 40 && P->tile_bits == 12, "qsv_run_pass_rounds_sums: needs 12-bit tiles");
    QSV_CHECK_ARG((P->reserved == 0 || P->reserved == 4) && prog_bytes >= sizeof(qsv_rprog_hdr_t) && jit::wanted(P, prog_host),
                  "qsv_run_pass_rounds_sums: only the pass-specialised kernel sums (QSV_JIT / register size)");
    QSV_CHECK_ARG(n_bits >= 1 && n_bits <= 12, "qsv_run_pass_rounds_sums: n_bits=%d out of range", n_bits);
    QSV_CUDA(cudaMemsetAsync(buckets_dev, 0, (size_t)32 << n_bits, (cudaStream_t)stream));
    return jit::run_sums(state_dev, P, prog_host, prog_bytes, tables_dev, (cudaStream_t)stream, jit_grid(P), fixed_mask,
                         fixed_val, zf_mask, zf_val, n_bits, bucket_pos, buckets_dev);
}

static int run_pass_rounds(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes,
                           const void *tables_dev, void *stream, uint64_t sp_mask, uint64_t sp_val, uint32_t zf_mask,
                           uint32_t zf_val) {
    const bool fresh = (zf_mask & 0x80000000u) != 0;      // no fallback to the interpreter, which cannot zero-fill
    zf_mask &= 0x7fffffffu;
    QSV_CHECK_ARG(state_dev && P && prog_host, "qsv_run_pass_rounds: NULL pointer");
    QSV_CHECK_ARG(P->n_local >= 1 && P->n_local <= 40, "qsv_run_pass_rounds: n_local=%d out of range", P->n_local);
    const int RB = P->reserved ? P->reserved : 4;             // register bits per round (3 or 4)
    QSV_CHECK_ARG(RB == 3 || RB == 4, "qsv_run_pass_rounds: register bits (pass->reserved) must be 3 or 4");
    QSV_CHECK_ARG(P->tile_bits >= RB && P->tile_bits <= 12 && P->tile_bits <= P->n_local,
                  "qsv_run_pass_rounds: tile_bits=%d must be in [%d, 12] and <= n_local", P->tile_bits, RB);
    QSV_CHECK_ARG(P->low_bits >= 0 && P->low_bits <= P->tile_bits && P->n_hi == P->tile_bits - P->low_bits &&
                      P->n_hi <= 8, "qsv_run_pass_rounds: bad tile geometry");
    int prev = P->low_bits - 1;
    for (int j = 0; j < P->n_hi; ++j) {
        QSV_CHECK_ARG((int)P->hi_pos[j] > prev && (int)P->hi_pos[j] < P->n_local,
                      "qsv_run_pass_rounds: hi_pos[%d]=%d not ascending in [low_bits, n_local)", j, (int)P->hi_pos[j]);
        prev = P->hi_pos[j];
    }
    QSV_CHECK_ARG(prog_bytes >= sizeof(qsv_rprog_hdr_t), "qsv_run_pass_rounds: program too short");
    // Large states run the pass-specialised kernel (qsv_jit.cu): same tile geometry and round structure,
    // generated as straight-line code for this program and cached by its bytes.  It has its own (much
    // larger) program capacity, so it is tried before the interpreter's kernel-parameter limits apply.
    if (RB == 4 && jit::wanted(P, prog_host)) {
        const unsigned grid_j = jit_grid(P);
        const int rc = jit::launch(state_dev, P, prog_host, prog_bytes, tables_dev, (cudaStream_t)stream, grid_j,
                                   sp_mask, sp_val, zf_mask, zf_val);
        const char *mode = getenv("QSV_JIT");
        if (rc == 0 || fresh || (mode && !strcmp(mode, "force"))) return rc;
        static bool warned = false;
        if (!warned) {
            warned = true;
            fprintf(stderr, "libqsv: specialised kernel unavailable (%s); using the interpreting kernel k_pass_reg\n",
                    qsv_last_error());
        }
    }
    const qsv_rprog_hdr_t *hdr = static_cast<const qsv_rprog_hdr_t *>(prog_host);
    QSV_CHECK_ARG(hdr->n_rounds >= 1 && hdr->n_rounds <= QSV_RPROG_MAX_ROUNDS && hdr->n_ops <= QSV_RPROG_MAX_OPS,
                  "qsv_run_pass_rounds: %u rounds / %u ops exceed the limits (%d / %d)", hdr->n_rounds, hdr->n_ops,
                  QSV_RPROG_MAX_ROUNDS, QSV_RPROG_MAX_OPS);
    const size_t need = sizeof(qsv_rprog_hdr_t) + sizeof(qsv_round_t) * hdr->n_rounds + sizeof(qsv_rop_t) * hdr->n_ops;
    QSV_CHECK_ARG(prog_bytes >= need, "qsv_run_pass_rounds: prog_bytes=%u smaller than the %zu bytes the header implies",
                  prog_bytes, need);
    // repacked: fixed-capacity arrays inside the kernel parameter.  Per thread (not a shared static): host threads may
    // launch passes concurrently, each on its own stream; the launch copies the parameter before it returns.
    static thread_local RProgParam param;
    memset(&param, 0, sizeof(param));
    param.hdr = *hdr;
    const uint8_t *src = static_cast<const uint8_t *>(prog_host) + sizeof(qsv_rprog_hdr_t);
    memcpy(param.rounds, src, sizeof(qsv_round_t) * hdr->n_rounds);
    memcpy(param.ops, src + sizeof(qsv_round_t) * hdr->n_rounds, sizeof(qsv_rop_t) * hdr->n_ops);
    bool has_diag = false;
    for (uint32_t i = 0; i < hdr->n_ops; ++i) {
        has_diag = has_diag || param.ops[i].kind == QSV_ROP_DIAG;
        QSV_CHECK_ARG(param.ops[i].kind != QSV_ROP_MCX_REG,
                      "qsv_run_pass_rounds: the program holds in-register permutations (QSV_ROP_MCX_REG), which only "
                      "the pass-specialised kernel implements, and that kernel is unavailable: %s", qsv_last_error());
        for (int q = 0; q < 4; ++q) (void)q;
    }
    for (uint32_t r = 0; r < hdr->n_rounds; ++r) {
        const qsv_round_t &R = param.rounds[r];
        QSV_CHECK_ARG(R.rb[0] < R.rb[1] && R.rb[1] < R.rb[2] && (RB == 3 || R.rb[2] < R.rb[3]) &&
                          R.rb[RB - 1] < P->tile_bits,
                      "qsv_run_pass_rounds: round %u register bits not ascending inside the tile", r);
        QSV_CHECK_ARG((uint64_t)R.first_op + (R.n_pre & 0x7fffu) + (R.n_ops & 0xffffu) + ((R.n_ops >> 16) & 0x3fffu) +
                              (R.n_post & 0x7fffu) <= hdr->n_ops, "qsv_run_pass_rounds: round %u op range", r);
    }
    QSV_CHECK_ARG(!has_diag || tables_dev, "qsv_run_pass_rounds: diagonal op without a table buffer");
    const uint64_t n_tiles = 1ull << (P->n_local - P->tile_bits);
    const size_t smem = (size_t)16 << P->tile_bits;
    // persistent CTAs: 2 resident per SM, 4x oversubscribed so that skipped tiles do not unbalance the
    // SMs; the per-CTA prologue (program copy, tables, mbarrier) is paid once per ~50+ tiles, not per tile
    const uint64_t resident = (uint64_t)sm_count() * 2ull * 4ull;
    const unsigned grid = (unsigned)(n_tiles < resident ? n_tiles : resident);
    if (RB == 4) {
        QSV_CUDA(cudaFuncSetAttribute(k_pass_reg<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 12));
        k_pass_reg<4><<<grid, 256, smem, (cudaStream_t)stream>>>((double2 *)state_dev, *P, param,
                                                                 (const uint8_t *)tables_dev, n_tiles);
    } else {
        QSV_CUDA(cudaFuncSetAttribute(k_pass_reg<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 << 12));
        k_pass_reg<3><<<grid, 512, smem, (cudaStream_t)stream>>>((double2 *)state_dev, *P, param,
                                                                 (const uint8_t *)tables_dev, n_tiles);
    }
    QSV_LAUNCH_CHECK();
    return 0;
}
