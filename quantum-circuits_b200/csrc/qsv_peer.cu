// Global<->local qubit swap as ONE kernel over NVLink peer memory (SURVEY 8e's exchange step without
// NCCL, staging buffers or pack/unpack passes).  Each rank swaps, for every peer j, its block j with the
// peer's block r; the two ranks of a pair alternate granules so that both NVLink directions and both
// HBMs carry the same load.  16-byte vector accesses, fully coalesced on both sides.
//
// Partial swaps ("pairwise with rank ^ 2^j for one qubit, half a shard each way"): any subset of
// the global bits may be exchanged; the peers of a rank are then the 2^j - 1 ranks that differ from it in
// those bits only, and (1 - 2^-j) of the shard moves.
// Support hint: a register that started in a basis state is still zero wherever an untouched index bit
// disagrees with its initial value (schedule.Support); pairs of which NEITHER side can be non-zero are
// skipped chunk by chunk, so a swap of a still-sparse state costs next to nothing.
#include "qsv_common.cuh"

namespace qsv {

constexpr int kMaxPeers = 16;
struct PeerPtrs { double2 *p[kMaxPeers]; };
struct SwapPos { int g; int q[4]; int gb[4]; };   // local position q[k] (ascending) trades places with global bit gb[k]

__device__ __forceinline__ uint64_t swap_expand(uint64_t e, const SwapPos &pos) {
    for (int k = 0; k < pos.g; ++k) e = insert_zero(e, (unsigned)pos.q[k]);
    return e;
}

__global__ void __launch_bounds__(256)
k_swap_peer(double2 *__restrict__ local, const __grid_constant__ PeerPtrs peers, const __grid_constant__ SwapPos pos,
            int n_local, int r, uint64_t S, uint64_t G, uint64_t per_pair, uint64_t CH, uint64_t sp_mask,
            uint64_t sp_val) {
    // The shard is indexed by (field, rest): field = the j bits at positions pos.q[] (the qubits that become
    // global), rest = the other n_local - j bits (S = 2^(n_local-j) values).  Element (field = F(peer), rest = e)
    // of rank r trades places with element (field = F(r), rest = e) of rank `peer`, F(x) = the swapped global
    // bits of x.  With the top positions this is the exchange of contiguous blocks; with lower positions the same
    // pairs sit in runs of 2^q[0] amplitudes, so no qubit has to be moved to the top of the shard first.
    // Chunk c (CH contiguous elements) -> peer slot s = 1 + c mod (2^j - 1): consecutive chunks go to DIFFERENT
    // peers, at every moment each rank talks to all its peers at once and the traffic through the NVSwitch is a
    // uniform all-to-all.  (Walking peer by peer made all ranks hit the same GPU at once: 305 GB/s at 8 GPUs
    // against 680 GB/s at 2.)
    const uint64_t n_peers = (1ull << pos.g) - 1ull;
    const uint64_t n_chunks = n_peers * (per_pair / CH);
    // index bits that vary inside a chunk (the low bits of `rest`, moved up past the swapped positions below them)
    const uint64_t var_mask = swap_expand(CH - 1, pos);
    const uint64_t cm = sp_mask & ~var_mask, cv = sp_val & ~var_mask;
    uint64_t fr = 0;
    for (int k = 0; k < pos.g; ++k) fr |= (uint64_t)((r >> pos.gb[k]) & 1) << pos.q[k];
    for (uint64_t c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const unsigned s_ = 1u + (unsigned)(c % n_peers);
        unsigned dj = 0;
        for (int k = 0; k < pos.g; ++k) dj |= ((s_ >> k) & 1u) << pos.gb[k];
        const int j = r ^ (int)dj;
        uint64_t fj = 0;
        for (int k = 0; k < pos.g; ++k) fj |= (uint64_t)((j >> pos.gb[k]) & 1) << pos.q[k];
        const uint64_t hb = (c / n_peers) * CH;
        // granules alternate between the two ranks of the pair: the lower rank owns the even ones
        const uint64_t e0 = S >= 2 ? ((hb / G) * 2 + (r < j ? 0 : 1)) * G + (hb % G) : 0;
        if (S < 2 && r > j) continue;
        const uint64_t x0 = swap_expand(e0, pos);
        const uint64_t mine_g = ((uint64_t)r << n_local) | x0 | fj, theirs_g = ((uint64_t)j << n_local) | x0 | fr;
        if ((mine_g & cm) != cv && (theirs_g & cm) != cv) continue;          // zeros on both sides
        if (sp_mask == 0 && CH == 2 * (uint64_t)blockDim.x) {
            // dense state, full chunk: both elements of a thread are loaded before either is stored
            const uint64_t xa = swap_expand(e0 + threadIdx.x, pos), xb = swap_expand(e0 + threadIdx.x + blockDim.x, pos);
            double2 *m0 = local + (xa | fj), *m1 = local + (xb | fj);
            double2 *t0 = peers.p[j] + (xa | fr), *t1 = peers.p[j] + (xb | fr);
            const double2 b0 = *t0, b1 = *t1;   // remote loads over NVLink
            const double2 a0 = *m0, a1 = *m1;
            *m0 = b0; *m1 = b1;
            *t0 = a0; *t1 = a1;                 // remote stores over NVLink
            continue;
        }
        for (uint64_t l = threadIdx.x; l < CH; l += blockDim.x) {
            const uint64_t x = swap_expand(e0 + l, pos);
            if (sp_mask) {
                const uint64_t mg = ((uint64_t)r << n_local) | x | fj, tg = ((uint64_t)j << n_local) | x | fr;
                if ((mg & sp_mask) != sp_val && (tg & sp_mask) != sp_val) continue;
            }
            double2 *mine = local + (x | fj);
            double2 *theirs = peers.p[j] + (x | fr);
            const double2 a = *mine;
            const double2 b = *theirs;          // remote load over NVLink
            *mine = b;
            *theirs = a;                        // remote store over NVLink
        }
    }
}

}  // namespace qsv

using namespace qsv;

extern "C" int qsv_swap_peer_ex(void *state_dev, const uint64_t *peer_state_ptrs, int n_local, int n_ranks, int rank,
                                int n_swap, const int *positions, const int *gbits, uint64_t sp_mask, uint64_t sp_val,
                                void *stream) {
    QSV_CHECK_ARG(state_dev && peer_state_ptrs, "qsv_swap_peer: NULL pointer");
    QSV_CHECK_ARG(n_ranks >= 2 && n_ranks <= kMaxPeers && !(n_ranks & (n_ranks - 1)), "qsv_swap_peer: %d ranks", n_ranks);
    QSV_CHECK_ARG(n_swap >= 1 && n_swap <= 4 && n_swap <= n_local && (1 << n_swap) <= n_ranks,
                  "qsv_swap_peer: n_swap=%d out of range", n_swap);
    QSV_CHECK_ARG(rank >= 0 && rank < n_ranks, "qsv_swap_peer: rank %d out of range", rank);
    PeerPtrs pp;
    for (int j = 0; j < kMaxPeers; ++j) pp.p[j] = nullptr;
    SwapPos sp;
    sp.g = n_swap;
    for (int k = 0; k < 4; ++k) { sp.q[k] = 0; sp.gb[k] = 0; }
    for (int k = 0; k < n_swap; ++k) {
        sp.q[k] = positions ? positions[k] : n_local - n_swap + k;
        sp.gb[k] = gbits ? gbits[k] : k;
        QSV_CHECK_ARG(sp.q[k] >= 0 && sp.q[k] < n_local && (k == 0 || sp.q[k] > sp.q[k - 1]),
                      "qsv_swap_peer: positions must be ascending local positions (positions[%d]=%d)", k, sp.q[k]);
        QSV_CHECK_ARG(sp.gb[k] >= 0 && (1 << sp.gb[k]) < n_ranks && (k == 0 || sp.gb[k] > sp.gb[k - 1]),
                      "qsv_swap_peer: global bits must be ascending and below log2(ranks) (gbits[%d]=%d)", k, sp.gb[k]);
    }
    for (int s_ = 1; s_ < (1 << n_swap); ++s_) {
        int dj = 0;
        for (int k = 0; k < n_swap; ++k) dj |= ((s_ >> k) & 1) << sp.gb[k];
        const int j = rank ^ dj;
        QSV_CHECK_ARG(peer_state_ptrs[j] != 0, "qsv_swap_peer: peer %d has no mapped pointer", j);
        pp.p[j] = reinterpret_cast<double2 *>(peer_state_ptrs[j]);
    }
    const uint64_t S = 1ull << (n_local - n_swap);
    const uint64_t G = S >= 2 ? (S / 2 < 2048 ? S / 2 : 2048) : 1;
    const uint64_t per_pair = S >= 2 ? S / 2 : 1;
    const uint64_t CH = per_pair < 512 ? per_pair : 512;      // contiguous elements per peer access (8 KiB)
    const uint64_t n_chunks = (uint64_t)((1 << n_swap) - 1) * (per_pair / CH);
    uint64_t grid = n_chunks;
    const uint64_t cap = 148ull * 16ull;
    if (grid > cap) grid = cap;
    k_swap_peer<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((double2 *)state_dev, pp, sp, n_local, rank, S, G,
                                                                 per_pair, CH, sp_mask, sp_val);
    QSV_LAUNCH_CHECK();
    return 0;
}

extern "C" int qsv_swap_peer(void *state_dev, const uint64_t *peer_state_ptrs, int n_local, int g, int rank,
                             const int *positions, void *stream) {
    QSV_CHECK_ARG(g >= 1 && g <= 4, "qsv_swap_peer: g=%d out of range", g);
    return qsv_swap_peer_ex(state_dev, peer_state_ptrs, n_local, 1 << g, rank, g, positions, nullptr, 0, 0, stream);
}
