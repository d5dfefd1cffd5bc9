// Global<->local qubit swap as ONE kernel over NVLink peer memory (SURVEY 8e's exchange step without
// NCCL, staging buffers or pack/unpack passes).  Each rank swaps, for every peer j, its block j with the
// peer's block r; the two ranks of a pair alternate granules so that both NVLink directions and both
// HBMs carry the same load.  16-byte vector accesses, fully coalesced on both sides.
#include "qsv_common.cuh"

namespace qsv {

constexpr int kMaxPeers = 16;
struct PeerPtrs { double2 *p[kMaxPeers]; };
struct SwapPos { int g; int q[4]; };       // local positions (ascending) trading places with global bits 0..g-1

__global__ void __launch_bounds__(256)
k_swap_peer(double2 *__restrict__ local, const __grid_constant__ PeerPtrs peers, const __grid_constant__ SwapPos pos,
            int P, int r, uint64_t S, uint64_t G, uint64_t per_pair, uint64_t CH) {
    // The shard is indexed by (field, rest): field = the g bits at positions pos.q[] (the qubits that become
    // global), rest = the other n_local - g bits (S = 2^(n_local-g) values).  Element (field = j, rest = e) of
    // rank r trades places with element (field = r, rest = e) of rank j.  With the top g positions this is the
    // exchange of contiguous blocks; with lower positions the same pairs sit in runs of 2^q[0] amplitudes, so no
    // qubit has to be moved to the top of the shard first (the re-layout CNots and the passes they cost).
    // Work item w -> (chunk c, lane-in-chunk).  Consecutive chunks go to DIFFERENT peers (j = r ^ s,
    // s = 1 + c mod (P-1)): at every moment each rank talks to all its peers at once and the traffic
    // through the NVSwitch is a uniform all-to-all.  (Walking peer by peer made all ranks hit the same
    // GPU at the same time: 305 GB/s at 8 GPUs against 680 GB/s at 2.)
    const uint64_t total = (uint64_t)(P - 1) * per_pair;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < total;
         w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t c = w / CH;
        const int s_ = 1 + (int)(c % (uint64_t)(P - 1));
        const int j = r ^ s_;
        const uint64_t h = (c / (uint64_t)(P - 1)) * CH + (w % CH);
        // granules alternate between the two ranks of the pair: the lower rank owns the even ones
        uint64_t e = S >= 2 ? ((h / G) * 2 + (r < j ? 0 : 1)) * G + (h % G) : 0;
        if (S < 2 && r > j) continue;
        for (int k = 0; k < pos.g; ++k) e = insert_zero(e, (unsigned)pos.q[k]);
        uint64_t fj = 0, fr = 0;
        for (int k = 0; k < pos.g; ++k) {
            fj |= (uint64_t)((j >> k) & 1) << pos.q[k];
            fr |= (uint64_t)((r >> k) & 1) << pos.q[k];
        }
        double2 *mine = local + (e | fj);
        double2 *theirs = peers.p[j] + (e | fr);
        const double2 a = *mine;
        const double2 b = *theirs;          // remote load over NVLink
        *mine = b;
        *theirs = a;                        // remote store over NVLink
    }
}

}  // namespace qsv

using namespace qsv;

extern "C" int qsv_swap_peer(void *state_dev, const uint64_t *peer_state_ptrs, int n_local, int g, int rank,
                             const int *positions, void *stream) {
    QSV_CHECK_ARG(state_dev && peer_state_ptrs, "qsv_swap_peer: NULL pointer");
    QSV_CHECK_ARG(g >= 1 && g <= 4 && g <= n_local, "qsv_swap_peer: g=%d out of range", g);
    const int P = 1 << g;
    QSV_CHECK_ARG(rank >= 0 && rank < P, "qsv_swap_peer: rank %d out of range", rank);
    PeerPtrs pp;
    for (int j = 0; j < kMaxPeers; ++j) pp.p[j] = nullptr;
    for (int j = 0; j < P; ++j) {
        QSV_CHECK_ARG(j == rank || peer_state_ptrs[j] != 0, "qsv_swap_peer: peer %d has no mapped pointer", j);
        pp.p[j] = reinterpret_cast<double2 *>(peer_state_ptrs[j]);
    }
    SwapPos sp;
    sp.g = g;
    for (int k = 0; k < 4; ++k) sp.q[k] = 0;
    for (int k = 0; k < g; ++k) {
        sp.q[k] = positions ? positions[k] : n_local - g + k;
        QSV_CHECK_ARG(sp.q[k] >= 0 && sp.q[k] < n_local && (k == 0 || sp.q[k] > sp.q[k - 1]),
                      "qsv_swap_peer: positions must be ascending local positions (positions[%d]=%d)", k, sp.q[k]);
    }
    const uint64_t S = 1ull << (n_local - g);
    const uint64_t G = S >= 2 ? (S / 2 < 2048 ? S / 2 : 2048) : 1;
    const uint64_t per_pair = S >= 2 ? S / 2 : 1;
    const uint64_t CH = per_pair < 512 ? per_pair : 512;      // contiguous elements per peer access (8 KiB)
    const uint64_t total = (uint64_t)(P - 1) * per_pair;
    uint64_t grid = (total + 255) / 256;
    const uint64_t cap = 148ull * 16ull;
    if (grid > cap) grid = cap;
    k_swap_peer<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((double2 *)state_dev, pp, sp, P, rank, S, G, per_pair,
                                                                 CH);
    QSV_LAUNCH_CHECK();
    return 0;
}
