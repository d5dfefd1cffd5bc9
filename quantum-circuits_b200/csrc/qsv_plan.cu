// Host-side pass planner kernel: the tile search of qsv.schedule._improve_tile in native code.
//
// Which 2^T amplitudes share a tile decides how many gates one read+write of the state applies (the reference
// applies ONE gate per full-state product, main/circuit.py:315-327; here a pass absorbs every op whose non-diagonal
// targets lie inside the tile and that commutes past what had to stay behind).  The planner evaluates "how many of
// the next ops would a pass with THESE tile positions absorb" a few thousand times per pass; in Python that loop is
// 80 % of the host time of a cold Circuit.run.  Native, one evaluation is ~1 us, which also pays for a randomised
// search over whole schedules (qsv.schedule.plan_schedule_search): several local searches with shuffled candidate
// orders, the cheapest schedule under the measured cost model wins.
//
// No device code: plain C++ compiled into libqsv.so with the rest of the C-ABI.
#include "qsv_common.cuh"
#include <stdint.h>
#include <algorithm>
#include <vector>

namespace {

struct Rng {
    uint64_t s;
    explicit Rng(uint64_t seed) : s(seed * 0x9e3779b97f4a7c15ull + 0xd1b54a32d192ed03ull) {}
    uint64_t next() {
        s ^= s >> 12; s ^= s << 25; s ^= s >> 27;
        return s * 0x2545f4914f6cdd1dull;
    }
    template <class T> void shuffle(std::vector<T> &v) {
        for (size_t i = v.size(); i > 1; --i) std::swap(v[i - 1], v[next() % i]);
    }
};

// ops [0, limit) in program order: an op is taken if it may join a pass at all (never[i] == 0), its non-diagonal
// targets lie inside `allowed`, and it commutes past everything left behind (non-diagonal use of a position blocks
// every later use, diagonal / control use blocks later non-diagonal use only)
inline int absorbed(const uint64_t *nd, const uint64_t *dg, const uint8_t *never, int limit, uint64_t allowed) {
    uint64_t bnd = 0, bd = 0;
    int count = 0;
    for (int i = 0; i < limit; ++i) {
        const uint64_t a = nd[i], d = dg[i];
        if (!never[i] && !(a & (bnd | bd)) && !(d & bnd) && !(a & ~allowed)) {
            ++count;
        } else {
            bnd |= a;
            bd |= d;
        }
    }
    return count;
}

inline uint64_t mask_of(const std::vector<int> &bits) {
    uint64_t m = 0;
    for (int b : bits) m |= 1ull << b;
    return m;
}

}  // namespace

extern "C" int qsv_plan_absorbed(const uint64_t *nd, const uint64_t *dg, const uint8_t *never, int n_ops,
                                 uint64_t allowed) {
    if (!nd || !dg || !never || n_ops < 0) return -1;
    return absorbed(nd, dg, never, n_ops, allowed);
}

extern "C" int qsv_plan_improve_tile(const uint64_t *nd, const uint64_t *dg, const uint8_t *never, int n_ops,
                                     int window, uint64_t low_mask, const int *hi_in, int n_hi_in, int n_free_slots,
                                     int n_local, int low_bits, int max_rounds, uint64_t seed, int *hi_out,
                                     int *n_hi_out, int *absorbed_out) {
    QSV_CHECK_ARG(nd && dg && never && hi_out && n_hi_out, "qsv_plan_improve_tile: NULL pointer");
    QSV_CHECK_ARG(n_ops >= 0 && n_hi_in >= 0 && n_hi_in <= 16 && n_free_slots >= 0 && n_free_slots <= 16 &&
                  n_local >= 1 && n_local <= 63 && low_bits >= 0 && low_bits <= n_local,
                  "qsv_plan_improve_tile: argument out of range");
    const int limit = std::min(n_ops, window);
    // candidates: every position >= low_bits a joinable op of the first half of the window touches non-diagonally
    uint64_t cand_mask = 0;
    for (int i = 0; i < std::min(n_ops, window / 2); ++i)
        if (!never[i]) cand_mask |= nd[i];
    cand_mask &= ~((1ull << low_bits) - 1ull);
    if (n_local < 64) cand_mask &= (1ull << n_local) - 1ull;
    std::vector<int> hi(hi_in, hi_in + n_hi_in);
    std::sort(hi.begin(), hi.end());
    Rng rng(seed);
    int best = absorbed(nd, dg, never, limit, low_mask | mask_of(hi));
    for (int round = 0; round < max_rounds; ++round) {
        bool improved = false;
        const uint64_t hm = mask_of(hi);
        std::vector<int> outside;
        for (int b = low_bits; b < n_local; ++b)
            if (((cand_mask & ~hm) >> b) & 1ull) outside.push_back(b);
        if (seed) rng.shuffle(outside);
        if ((int)hi.size() < n_free_slots) {
            // unused slots first: adding a position never hurts
            for (int inn : outside) {
                const int got = absorbed(nd, dg, never, limit, low_mask | hm | (1ull << inn));
                if (got > best) {
                    best = got;
                    hi.push_back(inn);
                    std::sort(hi.begin(), hi.end());
                    improved = true;
                    break;
                }
            }
            if (improved) continue;
        }
        std::vector<int> order(hi);
        if (seed) rng.shuffle(order);
        for (int out : order) {
            for (int inn : outside) {
                const int got = absorbed(nd, dg, never, limit, low_mask | (hm & ~(1ull << out)) | (1ull << inn));
                if (got > best) {
                    best = got;
                    hi.erase(std::find(hi.begin(), hi.end(), out));
                    hi.push_back(inn);
                    std::sort(hi.begin(), hi.end());
                    improved = true;
                    break;
                }
            }
            if (improved) break;
        }
        if (!improved) break;
    }
    for (size_t i = 0; i < hi.size(); ++i) hi_out[i] = hi[i];
    *n_hi_out = (int)hi.size();
    if (absorbed_out) *absorbed_out = best;
    return 0;
}
