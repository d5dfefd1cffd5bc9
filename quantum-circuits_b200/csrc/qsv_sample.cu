// Exact sampling on large and sharded registers: the random.choices draw of main/circuit.py:360-364,
//     index = bisect_right(accumulate(weights), u * total, 0, N - 1),
// evaluated in OUTPUT order with EXACT arithmetic, so that the outcome does not depend on how the register is
// sharded, tiled or summed.
//
// The reference accumulates 2^n Python floats one after the other; that is a serial chain no device can follow at
// 2^30 outcomes, and any parallel float summation rounds differently from it (and from itself on another shard
// count).  Here every weight w = re^2 + im^2 is converted to 128-bit fixed point with 2^-96 resolution - exact for
// every double >= 2^-44, truncated below - and all sums are integer sums: associative, hence identical for any
// partition of the state over threads, CTAs and GPUs.  The outcome is the first output index whose exact cumulative
// weight exceeds u * total; it can differ from the float rule only when u * total falls within one rounding error of
// a cumulative sum.  Registers of up to 2^20 outcomes on one GPU keep the sequential float scan (qsv_sample_scan /
// qsv_sample_sequential), which is bit-exact with CPython.
//
// Search = radix descent over the output index, most significant bits first.  A LEVEL fixes up to 12 output bits:
// one sweep over the elements still compatible with the bits fixed so far accumulates 2^bits bucket sums
// (qsv_sample_level), the ranks' bucket arrays are summed (all-reduce of int64 limbs, done by the caller), and one
// small kernel picks the bucket that holds the target and subtracts the weight in front of it (qsv_sample_select).
// The first level reads the whole state once (16 B per amplitude, HBM-bound); every later level reads at most 1/4 of
// the previous one.  The last <= 12 bits are resolved by scattering the remaining weights into an array in output
// order (qsv_sample_leaf) and the same select kernel.  Nothing but the final 64-byte state goes back to the host.
#include "qsv_common.cuh"
#include <string.h>

namespace qsv {

struct U128 { unsigned long long lo, hi; };

__device__ __forceinline__ void u128_add(U128 &a, const U128 b) {
    a.lo += b.lo;
    a.hi += b.hi + (a.lo < b.lo ? 1ull : 0ull);
}
__device__ __forceinline__ bool u128_gt(const U128 a, const U128 b) { return a.hi > b.hi || (a.hi == b.hi && a.lo > b.lo); }
__device__ __forceinline__ U128 u128_sub(const U128 a, const U128 b) {
    U128 r;
    r.lo = a.lo - b.lo;
    r.hi = a.hi - b.hi - (a.lo < b.lo ? 1ull : 0ull);
    return r;
}

// w >= 0 as fixed point in units of 2^-96 (floor).  Exact whenever w >= 2^-44.
__device__ __forceinline__ U128 weight_fix(const double w) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(w);
    const int e = (int)((b >> 52) & 0x7ffull);
    U128 r = {0ull, 0ull};
    if (e == 0) return r;                                   // zero / subnormal: below the resolution
    const unsigned long long m = (b & ((1ull << 52) - 1ull)) | (1ull << 52);
    const int sh = e - 1075 + 96;                           // w = m * 2^(e-1075)
    if (sh >= 75) { r.hi = ~0ull; r.lo = ~0ull; }           // > 2^31: saturate (never a probability)
    else if (sh >= 64) r.hi = m << (sh - 64);
    else if (sh > 0) { r.lo = m << sh; r.hi = m >> (64 - sh); }
    else if (sh == 0) r.lo = m;
    else if (sh > -64) r.lo = m >> (-sh);
    return r;
}

__device__ __forceinline__ uint64_t deposit_bits(uint64_t v, uint64_t mask) {     // software pdep
    uint64_t r = 0;
    while (mask && v) {
        const uint64_t low = mask & (0ull - mask);
        if (v & 1ull) r |= low;
        v >>= 1;
        mask ^= low;
    }
    return r;
}

struct LevelParams {
    int32_t n_local, n_bits, n_low, run_bits;
    uint64_t rank_bits;             // rank << n_local
    uint64_t fixed_local;           // local positions pinned by earlier levels
    uint64_t fixed_global;          // global positions pinned by earlier levels (pre-shifted)
    uint64_t run_mask;              // the run_bits lowest free local positions
    uint64_t high_free;             // the other free local positions
    uint8_t pos[12];                // physical position feeding bucket bit k (k = 0 least significant)
    uint8_t low_bit[2];             // bucket bits whose position lies inside the run
};

struct SampleState {                // qsv_sample_state_t
    unsigned long long fixed_val, z, target_lo, target_hi, total_lo, total_hi, status, levels_done;
};

constexpr int kLevelThreads = 256;

template <int NLOW>
__global__ void __launch_bounds__(kLevelThreads)
k_sample_level(const double2 *__restrict__ state, const __grid_constant__ LevelParams P,
               const SampleState *__restrict__ sel, unsigned long long *__restrict__ buckets) {
    __shared__ uint64_t tab_k[32];
    const uint64_t fixed_val = sel->fixed_val;
    if ((fixed_val ^ P.rank_bits) & P.fixed_global) return;          // the chosen prefix lives on other ranks
    const unsigned lane = threadIdx.x & 31u, warp_in_cta = threadIdx.x >> 5;
    const int lane_bits = P.run_bits < 5 ? P.run_bits : 5;
    const int k_bits = P.run_bits - lane_bits;
    // low lane_bits run positions carry the lane index, the next k_bits the iteration index
    uint64_t lane_mask = 0, m = P.run_mask;
    for (int i = 0; i < lane_bits; ++i) { const uint64_t low = m & (0ull - m); lane_mask |= low; m ^= low; }
    const uint64_t k_mask = m;
    if (threadIdx.x < 32) tab_k[threadIdx.x] = deposit_bits(threadIdx.x, k_mask);
    __syncthreads();
    const bool lane_on = lane < (1u << lane_bits);
    const uint64_t y_lane = deposit_bits(lane, lane_mask) | (fixed_val & P.fixed_local);

    int n_free_hi = 0;
    for (uint64_t t = P.high_free; t; t &= t - 1) ++n_free_hi;
    const uint64_t n_runs = 1ull << n_free_hi;
    const uint64_t warps = (uint64_t)gridDim.x * (kLevelThreads / 32);
    const uint64_t w = (uint64_t)blockIdx.x * (kLevelThreads / 32) + warp_in_cta;
    const uint64_t r0 = n_runs * w / warps, r1 = n_runs * (w + 1) / warps;      // n_runs, warps < 2^32

    // bucket bits outside the run are constant per run
    uint64_t hi_pos_mask = 0;
    for (int k = 0; k < P.n_bits; ++k)
        if (!((P.run_mask >> P.pos[k]) & 1ull)) hi_pos_mask |= 1ull << k;

    U128 acc[1 << NLOW];
#pragma unroll
    for (int a = 0; a < (1 << NLOW); ++a) acc[a] = {0ull, 0ull};
    uint32_t cur = 0xffffffffu;

    auto flush = [&](uint32_t hb) {
#pragma unroll
        for (int a = 0; a < (1 << NLOW); ++a) {
            U128 v = acc[a];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                U128 t;
                t.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
                t.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
                u128_add(v, t);
            }
            if (lane == 0 && (v.lo | v.hi)) {
                uint32_t idx = hb;
                if (NLOW >= 1) idx |= (uint32_t)(a & 1) << P.low_bit[0];
                if (NLOW >= 2) idx |= (uint32_t)((a >> 1) & 1) << P.low_bit[1];
                unsigned long long *b = buckets + 4ull * idx;      // four 32-bit limbs in 64-bit words: no carries
                atomicAdd(b + 0, v.lo & 0xffffffffull);
                atomicAdd(b + 1, v.lo >> 32);
                atomicAdd(b + 2, v.hi & 0xffffffffull);
                atomicAdd(b + 3, v.hi >> 32);
            }
            acc[a] = {0ull, 0ull};
        }
    };

    for (uint64_t r = r0; r < r1; ++r) {
        const uint64_t y_hi = deposit_bits(r, P.high_free);
        const uint64_t g = y_hi | fixed_val | P.rank_bits;          // every bit outside the run
        uint32_t hb = 0;
        for (int k = 0; k < P.n_bits; ++k)
            if ((hi_pos_mask >> k) & 1ull) hb |= (uint32_t)((g >> P.pos[k]) & 1ull) << k;
        if (hb != cur) {
            if (cur != 0xffffffffu) flush(cur);
            cur = hb;
        }
        const uint64_t y0 = y_hi | y_lane;
        if (lane_on) {
#pragma unroll 4
            for (int k = 0; k < (1 << k_bits); ++k) {
                const uint64_t y = y0 | tab_k[k];
                const U128 q = weight_fix(cnorm2(state[y]));
                if (NLOW == 0) {
                    u128_add(acc[0], q);
                } else {
                    const unsigned b0 = (unsigned)((y >> P.pos[P.low_bit[0]]) & 1ull);
                    const unsigned b1 = NLOW >= 2 ? (unsigned)((y >> P.pos[P.low_bit[1]]) & 1ull) : 0u;
                    const unsigned a_sel = b0 | (b1 << 1);
#pragma unroll
                    for (int a = 0; a < (1 << NLOW); ++a) {
                        const U128 qa = {a_sel == (unsigned)a ? q.lo : 0ull, a_sel == (unsigned)a ? q.hi : 0ull};
                        u128_add(acc[a], qa);
                    }
                }
            }
        }
    }
    if (cur != 0xffffffffu) flush(cur);
}

// Leaf: every remaining element goes to its own slot (leaf index = its remaining output bits): plain stores.
struct LeafParams {
    int32_t n_local, n_bits;
    uint64_t rank_bits, fixed_local, fixed_global, free_local;
    uint8_t pos[12];
};

__global__ void k_sample_leaf(const double2 *__restrict__ state, const __grid_constant__ LeafParams P,
                              const SampleState *__restrict__ sel, unsigned long long *__restrict__ buckets) {
    const uint64_t fixed_val = sel->fixed_val;
    if ((fixed_val ^ P.rank_bits) & P.fixed_global) return;
    int n_free = 0;
    for (uint64_t t = P.free_local; t; t &= t - 1) ++n_free;
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= (1ull << n_free)) return;
    const uint64_t y = deposit_bits(i, P.free_local) | (fixed_val & P.fixed_local);
    const uint64_t g = y | P.rank_bits;
    uint32_t idx = 0;
    for (int k = 0; k < P.n_bits; ++k) idx |= (uint32_t)((g >> P.pos[k]) & 1ull) << k;
    const U128 q = weight_fix(cnorm2(state[y]));
    unsigned long long *b = buckets + 4ull * idx;
    b[0] = q.lo & 0xffffffffull;
    b[1] = q.lo >> 32;
    b[2] = q.hi & 0xffffffffull;
    b[3] = q.hi >> 32;
}

__device__ __forceinline__ U128 limbs_value(const unsigned long long *b) {
    // b[0] + b[1] 2^32 + b[2] 2^64 + b[3] 2^96 (each word holds a sum of 32-bit pieces, so it may exceed 32 bits)
    U128 v;
    v.lo = b[0];
    v.hi = b[2];
    const unsigned long long t = b[1] << 32;
    v.lo += t;
    v.hi += (b[1] >> 32) + (v.lo < t ? 1ull : 0ull) + (b[3] << 32);
    return v;
}

struct SelectParams {
    int32_t n_bits, first;          // first = 1: this is level 1 - compute total and the target from u
    uint64_t u_mant;                // u = u_mant * 2^-u_shift, u_mant < 2^53
    int32_t u_shift, pad;
    uint8_t pos[12];                // physical position of bucket bit k
    uint8_t out_bit[12];            // output-index bit of bucket bit k
};

constexpr int kSelThreads = 256;
__global__ void __launch_bounds__(kSelThreads)
k_sample_select(const unsigned long long *__restrict__ buckets, const __grid_constant__ SelectParams P, SampleState *sel) {
    __shared__ U128 part[kSelThreads];
    const uint32_t nb = 1u << P.n_bits;
    const uint32_t per = (nb + kSelThreads - 1) / kSelThreads;
    const uint32_t b0 = threadIdx.x * per;
    U128 s = {0ull, 0ull};
    for (uint32_t b = b0; b < b0 + per && b < nb; ++b) u128_add(s, limbs_value(buckets + 4ull * b));
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x != 0) return;
    U128 target = {sel->target_lo, sel->target_hi};
    if (P.first) {
        U128 total = {0ull, 0ull};
        for (int t = 0; t < kSelThreads; ++t) u128_add(total, part[t]);
        sel->total_lo = total.lo;
        sel->total_hi = total.hi;
        if (!(total.lo | total.hi)) { sel->status = 1ull; return; }
        // target = floor(total * u_mant / 2^u_shift): 128 x 53 bit product, 192 bits, shifted right
        const unsigned long long m = P.u_mant;
        const unsigned long long l0 = total.lo * m, h0 = __umul64hi(total.lo, m);
        const unsigned long long l1 = total.hi * m, h1 = __umul64hi(total.hi, m);
        unsigned long long p0 = l0, p1 = h0 + l1, p2 = h1 + (p1 < l1 ? 1ull : 0ull);
        int sh = P.u_shift;
        while (sh >= 64) { p0 = p1; p1 = p2; p2 = 0ull; sh -= 64; }
        if (sh > 0) {
            p0 = (p0 >> sh) | (p1 << (64 - sh));
            p1 = (p1 >> sh) | (p2 << (64 - sh));
        }
        target.lo = p0;
        target.hi = p1;
    } else if (sel->status) {
        return;
    }
    // first bucket whose inclusive cumulative weight exceeds the target (bisect_right)
    U128 cum = {0ull, 0ull};
    int t = 0;
    for (; t < kSelThreads - 1; ++t) {
        U128 nxt = cum;
        u128_add(nxt, part[t]);
        if (u128_gt(nxt, target)) break;
        cum = nxt;
    }
    uint32_t b = (uint32_t)t * per, found = nb - 1;
    bool hit = false;
    for (; b < nb; ++b) {
        U128 nxt = cum;
        u128_add(nxt, limbs_value(buckets + 4ull * b));
        if (u128_gt(nxt, target)) { found = b; hit = true; break; }
        cum = nxt;
    }
    // no bucket exceeds the target: cannot happen with u < 1 (target < total).  bisect's hi = N-1 clamp: the LAST
    // index, at this level and (target saturated) at every level below
    U128 rest = u128_sub(target, cum);
    if (!hit) { found = nb - 1; rest = {~0ull, ~0ull}; }
    sel->target_lo = rest.lo;
    sel->target_hi = rest.hi;
    unsigned long long fv = sel->fixed_val, z = sel->z;
    for (int k = 0; k < P.n_bits; ++k) {
        const unsigned long long bit = (found >> k) & 1ull;
        fv |= bit << P.pos[k];
        z |= bit << P.out_bit[k];
    }
    sel->fixed_val = fv;
    sel->z = z;
    sel->levels_done += 1ull;
}

// Sequential float scan of a small register in ONE launch: total first, then the search, both in reference order
// (itertools.accumulate adds left to right; random.choices multiplies u by the last partial sum).
struct SeqMap { int32_t n, identity; uint8_t src[64]; };
__global__ void k_sample_sequential(const double2 *__restrict__ state, const __grid_constant__ SeqMap M, double u,
                                    unsigned long long *result) {
    const unsigned lane = threadIdx.x;
    const int n_local = M.n;
    const uint64_t n = 1ull << n_local;
    auto weight = [&](uint64_t z) {
        uint64_t y = z;
        if (!M.identity) {
            y = 0;
            for (int b = 0; b < n_local; ++b) y |= ((z >> b) & 1ull) << M.src[b];
        }
        return cnorm2(state[y]);
    };
    double total = 0.0;
    for (uint64_t z0 = 0; z0 < n; z0 += 32) {
        const double w = (z0 + lane < n) ? weight(z0 + lane) : 0.0;
        for (int l = 0; l < 32; ++l) {
            const double wl = __shfl_sync(0xffffffffu, w, l);
            if (z0 + l < n) total += wl;
        }
    }
    const double target = u * total;
    double cum = 0.0;
    unsigned long long found = ~0ull;
    for (uint64_t z0 = 0; z0 < n && found == ~0ull; z0 += 32) {
        const double w = (z0 + lane < n) ? weight(z0 + lane) : 0.0;
        for (int l = 0; l < 32; ++l) {
            const double wl = __shfl_sync(0xffffffffu, w, l);
            if (found == ~0ull && z0 + l < n) {
                cum += wl;
                if (cum > target) found = z0 + l;
            }
        }
    }
    if (lane == 0) {
        result[0] = (found == ~0ull || found > n - 1) ? (unsigned long long)(n - 1) : found;
        reinterpret_cast<double *>(result)[1] = total;
    }
}

static int popc64(uint64_t v) { int c = 0; for (; v; v &= v - 1) ++c; return c; }

}  // namespace qsv

using namespace qsv;

extern "C" int qsv_sample_begin(void *sel_dev, void *stream) {
    QSV_CHECK_ARG(sel_dev, "qsv_sample_begin: NULL pointer");
    static_assert(sizeof(SampleState) == sizeof(qsv_sample_state_t), "qsv_sample_state_t layout");
    QSV_CUDA(cudaMemsetAsync(sel_dev, 0, sizeof(SampleState), (cudaStream_t)stream));
    return 0;
}

static int check_positions(const char *who, int n_bits, const int *pos, int n_local) {
    QSV_CHECK_ARG(n_bits >= 0 && n_bits <= QSV_SAMPLE_MAX_LEVEL_BITS, "%s: n_bits=%d out of range", who, n_bits);
    QSV_CHECK_ARG(n_bits == 0 || pos, "%s: NULL position list", who);
    for (int k = 0; k < n_bits; ++k)
        QSV_CHECK_ARG(pos[k] >= 0 && pos[k] < 48 && pos[k] < n_local + 8, "%s: position %d out of range", who, pos[k]);
    return 0;
}

extern "C" int qsv_sample_level(const void *state_dev, int n_local, uint64_t rank_bits, uint64_t fixed_mask, int n_bits,
                                const int *bucket_pos, const void *sel_dev, void *buckets_dev, void *stream) {
    QSV_CHECK_ARG(state_dev && sel_dev && buckets_dev, "qsv_sample_level: NULL pointer");
    QSV_CHECK_ARG(n_local >= 1 && n_local <= 40, "qsv_sample_level: n_local=%d out of range", n_local);
    if (int rc = check_positions("qsv_sample_level", n_bits, bucket_pos, n_local)) return rc;
    QSV_CHECK_ARG(n_bits >= 1, "qsv_sample_level: a level needs at least one bit");
    const uint64_t local_mask = (1ull << n_local) - 1ull;
    LevelParams P;
    memset(&P, 0, sizeof(P));
    P.n_local = n_local;
    P.n_bits = n_bits;
    P.rank_bits = rank_bits;
    P.fixed_local = fixed_mask & local_mask;
    P.fixed_global = fixed_mask & ~local_mask;
    const uint64_t free_local = local_mask & ~P.fixed_local;
    const int n_free = popc64(free_local);
    P.run_bits = n_free < 10 ? n_free : 10;
    uint64_t m = free_local;
    for (int i = 0; i < P.run_bits; ++i) { const uint64_t low = m & (0ull - m); P.run_mask |= low; m ^= low; }
    P.high_free = m;
    for (int k = 0; k < n_bits; ++k) {
        QSV_CHECK_ARG(!((fixed_mask >> bucket_pos[k]) & 1ull), "qsv_sample_level: position %d is already fixed", bucket_pos[k]);
        P.pos[k] = (uint8_t)bucket_pos[k];
        if ((P.run_mask >> bucket_pos[k]) & 1ull) {
            QSV_CHECK_ARG(P.n_low < 2, "qsv_sample_level: more than two bucket bits inside the run (positions below "
                          "the 10 lowest free ones); split the level");
            P.low_bit[P.n_low++] = (uint8_t)k;
        }
    }
    QSV_CUDA(cudaMemsetAsync(buckets_dev, 0, (size_t)32 << n_bits, (cudaStream_t)stream));
    const uint64_t n_runs = 1ull << (n_free - P.run_bits);
    uint64_t grid = (n_runs + (kLevelThreads / 32) - 1) / (kLevelThreads / 32);
    const uint64_t cap = 148ull * 8ull;
    if (grid > cap) grid = cap;
    const auto *st = (const double2 *)state_dev;
    const auto *sel = (const SampleState *)sel_dev;
    auto *bk = (unsigned long long *)buckets_dev;
    if (P.n_low == 0) k_sample_level<0><<<(unsigned)grid, kLevelThreads, 0, (cudaStream_t)stream>>>(st, P, sel, bk);
    else if (P.n_low == 1) k_sample_level<1><<<(unsigned)grid, kLevelThreads, 0, (cudaStream_t)stream>>>(st, P, sel, bk);
    else k_sample_level<2><<<(unsigned)grid, kLevelThreads, 0, (cudaStream_t)stream>>>(st, P, sel, bk);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_sample_leaf(const void *state_dev, int n_local, uint64_t rank_bits, uint64_t fixed_mask, int n_bits,
                               const int *leaf_pos, const void *sel_dev, void *buckets_dev, void *stream) {
    QSV_CHECK_ARG(state_dev && sel_dev && buckets_dev, "qsv_sample_leaf: NULL pointer");
    QSV_CHECK_ARG(n_local >= 1 && n_local <= 40, "qsv_sample_leaf: n_local=%d out of range", n_local);
    if (int rc = check_positions("qsv_sample_leaf", n_bits, leaf_pos, n_local)) return rc;
    const uint64_t local_mask = (1ull << n_local) - 1ull;
    LeafParams P;
    memset(&P, 0, sizeof(P));
    P.n_local = n_local;
    P.n_bits = n_bits;
    P.rank_bits = rank_bits;
    P.fixed_local = fixed_mask & local_mask;
    P.fixed_global = fixed_mask & ~local_mask;
    P.free_local = local_mask & ~P.fixed_local;
    uint64_t leaf_local = 0;
    for (int k = 0; k < n_bits; ++k) {
        P.pos[k] = (uint8_t)leaf_pos[k];
        if (leaf_pos[k] < n_local) leaf_local |= 1ull << leaf_pos[k];
    }
    QSV_CHECK_ARG(leaf_local == P.free_local, "qsv_sample_leaf: the leaf bits must be exactly the positions not fixed yet");
    QSV_CUDA(cudaMemsetAsync(buckets_dev, 0, (size_t)32 << n_bits, (cudaStream_t)stream));
    const uint64_t n = 1ull << popc64(P.free_local);
    k_sample_leaf<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const double2 *)state_dev, P, (const SampleState *)sel_dev, (unsigned long long *)buckets_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_sample_select(const void *buckets_dev, int n_bits, const int *bucket_pos, const int *out_bits,
                                 int first, uint64_t u_mant, int u_shift, void *sel_dev, void *stream) {
    QSV_CHECK_ARG(buckets_dev && sel_dev, "qsv_sample_select: NULL pointer");
    if (int rc = check_positions("qsv_sample_select", n_bits, bucket_pos, 40)) return rc;
    QSV_CHECK_ARG(n_bits == 0 || out_bits, "qsv_sample_select: NULL out_bits");
    QSV_CHECK_ARG(!first || (u_mant < (1ull << 53) && u_shift >= 0 && u_shift < 100000), "qsv_sample_select: bad u");
    SelectParams P;
    memset(&P, 0, sizeof(P));
    P.n_bits = n_bits;
    P.first = first ? 1 : 0;
    P.u_mant = u_mant;
    P.u_shift = u_shift > 192 ? 192 : u_shift;
    for (int k = 0; k < n_bits; ++k) {
        QSV_CHECK_ARG(out_bits[k] >= 0 && out_bits[k] < 64, "qsv_sample_select: output bit %d out of range", out_bits[k]);
        P.pos[k] = (uint8_t)bucket_pos[k];
        P.out_bit[k] = (uint8_t)out_bits[k];
    }
    k_sample_select<<<1, kSelThreads, 0, (cudaStream_t)stream>>>((const unsigned long long *)buckets_dev, P,
                                                                 (SampleState *)sel_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_sample_sequential(const void *state_dev, int n_local, const int *src_pos, double u, void *result_dev,
                                     void *stream) {
    QSV_CHECK_ARG(state_dev && result_dev, "qsv_sample_sequential: NULL pointer");
    QSV_CHECK_ARG(n_local >= 0 && n_local <= 40, "qsv_sample_sequential: n_local=%d out of range", n_local);
    SeqMap M;
    memset(&M, 0, sizeof(M));
    M.n = n_local;
    M.identity = 1;
    uint64_t seen = 0;
    for (int b = 0; b < n_local; ++b) {
        const int s = src_pos ? src_pos[b] : b;
        QSV_CHECK_ARG(s >= 0 && s < n_local && !((seen >> s) & 1ull), "qsv_sample_sequential: src_pos is not a permutation");
        seen |= 1ull << s;
        M.src[b] = (uint8_t)s;
        if (s != b) M.identity = 0;
    }
    k_sample_sequential<<<1, 32, 0, (cudaStream_t)stream>>>((const double2 *)state_dev, M, u,
                                                            (unsigned long long *)result_dev);
    QSV_LAUNCH_CHECK();
    return 0;
}
