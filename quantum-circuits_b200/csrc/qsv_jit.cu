// Pass-specialised register kernel: the round program of one pass (qsv_rop_t list, include/qsv.h) is
// turned into straight-line sm_100a CUDA for exactly that pass, compiled once with NVRTC, cached by the
// program bytes and launched through the driver API.
//
// Why: k_pass_reg (qsv_tile_reg.cu) INTERPRETS the round program.  ncu showed it instruction-issue bound:
// ~4400 instructions per thread and tile, of which only ~1100 are the FP64 arithmetic of the gates; the
// rest is op decode (constant-bank loads, kind switch, predicates) and address bookkeeping that is the
// same for every tile.  Here all of that is resolved when the kernel is generated:
//   * op kinds, register bits, masks and matrix entries are literals; gates with entries 0, +-1, +-i
//     lose their multiplications; negations are integer sign flips (ALU pipe, not FP64);
//   * thread-level phases that depend only on the thread index are evaluated ONCE per CTA (before the
//     persistent tile loop) instead of once per tile;
//   * diagonal ops on register bits accumulate in a pending per-slot constant that is folded at
//     generation time and applied once (together with the thread scalar) instead of once per op;
//   * H-like real matrices (all entries +-h) are butterflies (1 add per component); their common factor
//     h is deferred and folded into the next multiplication that touches every slot;
//   * X / CNot / Toffoli stay address permutations of the shared-memory exchange (no data movement), with
//     their conditions reduced to one LOP3+compare on literals.
// Tile geometry, TMA bulk load/store and the round structure are those of k_pass_reg; the interpreter
// remains the path for small states (where a ~1 s compile cannot pay off) and for QSV_JIT=0.
//
// The same generated source compiles as plain C++ (-DQSV_JIT_HOST) with the round functions driven by a
// sequential loop over thread ids: that is how tests/ check the generator in a container without a GPU.
#include "qsv_common.cuh"
#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

namespace qsv {
namespace jit {

// ------------------------------------------------------------------------------------------------
// source generation
// ------------------------------------------------------------------------------------------------

struct Out {
    std::string s;
    void f(const char *fmt, ...) __attribute__((format(printf, 2, 3))) {
        char buf[2048];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof(buf), fmt, ap);
        va_end(ap);
        s += buf;
        s += '\n';
    }
};

struct Cx {
    double r, i;
    bool is_one() const { return r == 1.0 && i == 0.0; }
    bool operator==(const Cx &o) const { return r == o.r && i == o.i; }
};
static Cx cx_mul(Cx a, Cx b) { return {a.r * b.r - a.i * b.i, a.r * b.i + a.i * b.r}; }

static std::string dlit(double v) {
    if (v == 0.0) return "0.0";
    if (v == 1.0) return "1.0";
    if (v == -1.0) return "(-1.0)";
    uint64_t b;
    memcpy(&b, &v, 8);
    char buf[48];
    snprintf(buf, sizeof(buf), "QD(0x%016llxULL)", (unsigned long long)b);
    return buf;
}

// Slot r of the current round lives in the variables x<v>r / x<v>i with v = t_slot_var[r]: an unconditional
// (or register-bit controlled) X / CNot on a register bit is a renaming of slots, not an instruction.
static thread_local int t_slot_var[16];
static void reset_slot_vars() { for (int r = 0; r < 16; ++r) t_slot_var[r] = r; }
static std::string xr(int r) { return "x" + std::to_string(t_slot_var[r]) + "r"; }
static std::string xi(int r) { return "x" + std::to_string(t_slot_var[r]) + "i"; }

// x[r] *= c for a literal c
static void emit_cmul_const(Out &o, int r, Cx c) {
    const std::string R = xr(r), I = xi(r);
    if (c.i == 0.0) {
        if (c.r == 1.0) return;
        if (c.r == -1.0) {
            o.f("    %s = QNEG(%s); %s = QNEG(%s);", R.c_str(), R.c_str(), I.c_str(), I.c_str());
            return;
        }
        o.f("    %s *= %s; %s *= %s;", R.c_str(), dlit(c.r).c_str(), I.c_str(), dlit(c.r).c_str());
        return;
    }
    if (c.r == 0.0) {
        if (c.i == 1.0) {
            o.f("    { const double t = %s; %s = QNEG(%s); %s = t; }", R.c_str(), R.c_str(), I.c_str(), I.c_str());
        } else if (c.i == -1.0) {
            o.f("    { const double t = %s; %s = %s; %s = QNEG(t); }", R.c_str(), R.c_str(), I.c_str(), I.c_str());
        } else {
            o.f("    { const double t = %s; %s = %s * %s; %s = %s * t; }", R.c_str(), R.c_str(), dlit(-c.i).c_str(),
                I.c_str(), I.c_str(), dlit(c.i).c_str());
        }
        return;
    }
    o.f("    { const double t = %s * %s - %s * %s; %s = %s * %s + %s * %s; %s = t; }", R.c_str(), dlit(c.r).c_str(),
        I.c_str(), dlit(c.i).c_str(), I.c_str(), R.c_str(), dlit(c.i).c_str(), I.c_str(), dlit(c.r).c_str(), R.c_str());
}

// x[r] *= (mr, mi) for runtime variables
static void emit_cmul_var(Out &o, int r, const char *mr, const char *mi) {
    const std::string R = xr(r), I = xi(r);
    o.f("    { const double t = %s * %s - %s * %s; %s = %s * %s + %s * %s; %s = t; }", R.c_str(), mr, I.c_str(), mi,
        I.c_str(), R.c_str(), mi, I.c_str(), mr, R.c_str());
}

struct Term {
    double c;
    std::string v;
};
static std::string lin(const std::vector<Term> &t) {
    std::string e;
    for (const Term &x : t) {
        if (x.c == 0.0) continue;
        const bool neg = x.c < 0.0;
        const double a = fabs(x.c);
        std::string term = (a == 1.0) ? x.v : dlit(a) + " * " + x.v;
        if (e.empty())
            e = neg ? "-" + term : term;
        else
            e += (neg ? " - " : " + ") + term;
    }
    return e.empty() ? "0.0" : e;
}

static std::string cond_ext(uint64_t em, uint64_t ev) {
    char b[96];
    snprintf(b, sizeof(b), "((gbase & 0x%llxULL) == 0x%llxULL)", (unsigned long long)em, (unsigned long long)ev);
    return b;
}
static std::string cond_thr(uint32_t tm, uint32_t tv) {
    char b[96];
    snprintf(b, sizeof(b), "((tb & 0x%xu) == 0x%xu)", tm, tv);
    return b;
}
// condition of a register op: "" (always), or a C expression
static std::string op_cond(const qsv_rop_t &op) {
    std::string c;
    if (op.ext_mask) c = cond_ext(op.ext_mask, op.ext_val);
    if (op.tmask) c = c.empty() ? cond_thr(op.tmask, op.tval) : c + " && " + cond_thr(op.tmask, op.tval);
    return c;
}

struct RoundInfo {
    int n_hoisted = 0;           // doubles in qsv_H<k> (0: the struct carries a dummy)
    bool barrier = false;        // loads and stores of the round must be separated by a barrier
};

// Pending diagonal of a round: per-slot literal constants (register-bit phases without runtime
// condition), multiplied out at generation time.
struct Pending {
    Cx c[16];
    Pending() { reset(); }
    void reset() { for (auto &v : c) v = {1.0, 0.0}; }
    bool trivial() const {
        for (auto &v : c) if (!v.is_one()) return false;
        return true;
    }
    bool depends_on(int J) const {
        for (int r = 0; r < 16; ++r) if (!(c[r] == c[r ^ (1 << J)])) return true;
        return false;
    }
};

// A per-thread runtime scalar of a round: the product of phase factors whose conditions read thread-held
// or external bits.  Index 0 is S (multiplies every slot); the others belong to one register bit J, value V
// and "window" (the stretch of the op list between two dense gates on J, inside which diagonals commute).
struct Factor {
    std::string tcond;           // condition on the thread index ("" = none)
    uint64_t ext_mask = 0, ext_val = 0;
    Cx s;
};
struct Scalar {
    int J = -1, V = 0, window = 0;
    std::vector<Factor> f;
    int hoist_slot = -1;         // index into qsv_H<k>.v of the part that depends on the thread index only
    bool applied = false;
    std::string re() const { return "q" + std::to_string(id) + "r"; }
    std::string im() const { return "q" + std::to_string(id) + "i"; }
    int id = 0;
};

static void emit_scalar_mul(Out &o, const char *ind, const std::string &re, const std::string &im, const std::string &cond,
                            Cx c) {
    // (re, im) *= c under cond
    std::string body;
    char buf[512];
    if (c.i == 0.0) {
        snprintf(buf, sizeof(buf), "{ %s *= %s; %s *= %s; }", re.c_str(), dlit(c.r).c_str(), im.c_str(), dlit(c.r).c_str());
    } else if (c.r == 0.0) {
        snprintf(buf, sizeof(buf), "{ const double t = %s; %s = %s * %s; %s = %s * t; }", re.c_str(), re.c_str(),
                 dlit(-c.i).c_str(), im.c_str(), im.c_str(), dlit(c.i).c_str());
    } else {
        snprintf(buf, sizeof(buf), "{ const double t = %s * %s - %s * %s; %s = %s * %s + %s * %s; %s = t; }", re.c_str(),
                 dlit(c.r).c_str(), im.c_str(), dlit(c.i).c_str(), im.c_str(), re.c_str(), dlit(c.i).c_str(), im.c_str(),
                 dlit(c.r).c_str(), re.c_str());
    }
    if (cond.empty())
        o.f("%s%s", ind, buf);
    else
        o.f("%sif (%s) %s", ind, cond.c_str(), buf);
}

// Multiply slots by (literal constant) x (product of runtime scalars).  Products are built once per distinct
// combination, sharing prefixes (S, S.P0, S.P0.P1, ...), then applied with one complex multiply per slot.
struct SlotMul {
    Cx c = {1.0, 0.0};
    std::vector<int> sc;         // runtime scalar ids, ascending
};
static void apply_multipliers(Out &o, const SlotMul (&m)[16], const std::vector<Scalar> &scalars, int &tmp_id) {
    std::vector<std::pair<std::string, std::string>> cache;     // key -> variable stem
    auto find = [&](const std::string &k) -> std::string {
        for (auto &e : cache) if (e.first == k) return e.second;
        return "";
    };
    for (int r = 0; r < 16; ++r) {
        if (m[r].sc.empty()) {
            emit_cmul_const(o, r, m[r].c);
            continue;
        }
        // runtime product, prefix by prefix
        std::string key, stem;
        for (size_t i = 0; i < m[r].sc.size(); ++i) {
            const Scalar &sc = scalars[m[r].sc[i]];
            key += "s" + std::to_string(sc.id) + ".";
            if (i == 0) {
                stem = "q" + std::to_string(sc.id);
                continue;
            }
            std::string have = find(key);
            if (have.empty()) {
                have = "w" + std::to_string(tmp_id++);
                o.f("    const double %sr = %sr * %s - %si * %s, %si = %sr * %s + %si * %s;", have.c_str(), stem.c_str(),
                    sc.re().c_str(), stem.c_str(), sc.im().c_str(), have.c_str(), stem.c_str(), sc.im().c_str(),
                    stem.c_str(), sc.re().c_str());
                cache.push_back({key, have});
            }
            stem = have;
        }
        const Cx c = m[r].c;
        if (!c.is_one()) {
            char kb[64];
            uint64_t br, bi;
            memcpy(&br, &c.r, 8);
            memcpy(&bi, &c.i, 8);
            snprintf(kb, sizeof(kb), "c%llx,%llx", (unsigned long long)br, (unsigned long long)bi);
            key += kb;
            std::string have = find(key);
            if (have.empty()) {
                have = "w" + std::to_string(tmp_id++);
                if (c.i == 0.0)
                    o.f("    const double %sr = %sr * %s, %si = %si * %s;", have.c_str(), stem.c_str(), dlit(c.r).c_str(),
                        have.c_str(), stem.c_str(), dlit(c.r).c_str());
                else if (c.r == 0.0)
                    o.f("    const double %sr = %si * %s, %si = %sr * %s;", have.c_str(), stem.c_str(), dlit(-c.i).c_str(),
                        have.c_str(), stem.c_str(), dlit(c.i).c_str());
                else
                    o.f("    const double %sr = %sr * %s - %si * %s, %si = %sr * %s + %si * %s;", have.c_str(), stem.c_str(),
                        dlit(c.r).c_str(), stem.c_str(), dlit(c.i).c_str(), have.c_str(), stem.c_str(), dlit(c.i).c_str(),
                        stem.c_str(), dlit(c.r).c_str());
                cache.push_back({key, have});
            }
            stem = have;
        }
        emit_cmul_var(o, r, (stem + "r").c_str(), (stem + "i").c_str());
    }
}

// Shared-memory layout of the tile and the thread <-> tile-bit mapping (bank conflicts).
// A 16-byte access is conflict-free when the 8 threads of a quarter-warp hit 8 different 16-byte bank groups, i.e.
// differ in bits 0..2 of the slot index.  With the linear layout and thread bits in ascending order that fails as
// soon as a round's register bits include tile bits 0..2 (the threads then differ only in higher bits: 2-, 4- or
// 8-way conflicts on every exchange; ncu: 309 M conflict cycles in a 7.2 ms pass against 42 M in a 5.4 ms one).
// Padded layout: slot(a) = a + (a >> L) + (a >> (L+3)), i.e. run r (2^L amplitudes, one TMA bulk copy) starts at
// slot r * 2^L + r + (r >> 3), so tile bits L+j and L+3+j add 2^j to the bank group like tile bit j does.  Per round,
// thread-index bits 0..2 are given to one tile bit of each class {j, L+j, L+3+j} that is not a register bit: all 8
// combinations then fall into different groups (qsv.schedule.plan_rounds keeps one bit per class out of the registers).
static thread_local int t_pad_L = -1;       // >= 0: padded layout with runs of 2^L amplitudes; -1: linear layout

// Sums variant (the LAST pass of a run that ends in a sample): while a tile's final amplitudes sit in registers, their
// weights re^2 + im^2 are added up in the sampler's exact fixed point (csrc/qsv_sample.cu: units of 2^-96, 128-bit) and
// the tile's total goes to bucket hb = the index bits at pos[0..n_bits) (physical positions OUTSIDE the tile, pos[0] =
// least significant bucket bit) of a 4-limb bucket array - exactly what the first level of the sampler's radix descent
// (qsv_sample_level over the top output bits) would compute with another full read of the state.
struct SumSpec { int n_bits; int pos[12]; };
constexpr int kSumMaxInside = 3;              // bucket positions that may lie INSIDE the tile (2^3 partial sums per tile)
static thread_local const SumSpec *t_sums = nullptr;     // sums variant being generated (read by gen_round)
static thread_local int t_tile_phys[12];                 // physical position of tile-local bit t

// sums variant, STATIC form.  Register slot r of a thread is stored to the tile-local address b0 ^ XOR_{q in r} d_q
// (linear form of the store permutation).  When no op makes a delta d_q depend on the thread or on external bits -
// plain X / CNot riding on the stores, or no permutation at all - the d_q are known when the code is generated: the
// bucket bits INSIDE the tile then split into a per-thread part (read once from the address of slot 0) and a per-slot
// constant.  Returns false when the general form is needed; d[q] = tile-local delta of register bit q.
// (QSV_SUM_STATIC=0: always the general form, for A/B runs)
static bool sum_static(const qsv_round_t &R, const qsv_rop_t *ops, int f_post, int n_post, bool gen_post, uint32_t d[4]) {
    const char *e = getenv("QSV_SUM_STATIC");
    if (gen_post || (e && !strcmp(e, "0"))) return false;
    for (int q = 0; q < 4; ++q) d[q] = 1u << R.rb[q];
    for (int i = 0; i < n_post; ++i) {          // mirrors the linear branch of emit_addresses
        const qsv_rop_t &p = ops[f_post + i];
        const uint32_t xm = p.table_off, q = p.j, pbit = p.flags, cmu = p.tmask & ~pbit;
        if (q == 0xffu || pbit == 0u) continue;                 // moves the base only
        if (cmu || p.ext_mask) return false;                    // the delta would depend on the thread / the tile
        if (d[q] & pbit) d[q] ^= xm;
    }
    return true;
}

static void thread_bit_order(const qsv_round_t &R, int pos[8]) {
    bool is_rb[12] = {false};
    for (int q = 0; q < 4; ++q) is_rb[R.rb[q]] = true;
    bool used[12] = {false};
    int n = 0;
    if (t_pad_L >= 0) {
        for (int j = 0; j < 3; ++j) {
            const int cand[3] = {j, t_pad_L + j, t_pad_L + 3 + j};
            for (int c : cand)
                if (c < 12 && !is_rb[c] && !used[c]) { pos[n++] = c; used[c] = true; break; }
        }
    }
    for (int b = 0; b < 12 && n < 8; ++b)
        if (!is_rb[b] && !used[b]) { pos[n++] = b; used[b] = true; }
}
// ---- search over the thread-bit mapping (on by default since round 2; QSV_JIT_BANKSEARCH=0 keeps the static rule):
// score the ACTUAL slot addresses.  ncu, 30 qubits (profiles/r02_ncu_jit_kernel.md): store conflicts of two passes
// 160-165 M -> 26-35 M, load conflicts of a third 141 M -> 8 M, 5.78 / 5.84 / 5.42 -> 5.60 / 5.56 / 5.33 ms. ---------
// The class argument above is static: it ignores X / CNot / Toffoli that ride on a round's loads or stores and XOR
// thread-dependent bits into the slot index (ncu after the padded layout: 140 M load conflicts per pass are left in
// such rounds).  With the search on, every triple of non-register tile bits is tried as the low thread-index bits: the
// generator evaluates the 16 slot addresses of 8 lanes for a few quarter-warps exactly as the emitted code will
// (eval_addresses mirrors emit_addresses) and keeps the triple with the fewest bank-group collisions.
static thread_local const int *t_tb_override = nullptr;     // mapping chosen for the round being generated

static bool bank_search_enabled() {
    const char *e = getenv("QSV_JIT_BANKSEARCH");
    return !(e && !strcmp(e, "0"));
}

static uint32_t pad_slot(uint32_t a);
static uint32_t slot_offset(const qsv_round_t &R, int r);

static void eval_addresses(const qsv_round_t &R, const qsv_rop_t *ops, int first, int n, bool reverse, bool general,
                           uint32_t tb, bool ext_true, uint32_t a[16]) {
    auto opat = [&](int i) -> const qsv_rop_t & { return ops[reverse ? first + n - 1 - i : first + i]; };
    if (!general) {
        uint32_t b = tb, d[4] = {1u << R.rb[0], 1u << R.rb[1], 1u << R.rb[2], 1u << R.rb[3]};
        for (int i = 0; i < n; ++i) {
            const qsv_rop_t &p = opat(i);
            const uint32_t cm = p.tmask, cv = p.tval, xm = p.table_off, q = p.j, pbit = p.flags;
            const uint32_t cmu = cm & ~pbit;
            const bool e = p.ext_mask ? ext_true : true;
            const uint32_t f = (e && (!cmu || (b & cmu) == (cv & cmu))) ? xm : 0u;
            if (q == 0xffu || pbit == 0u) {
                b ^= f;
            } else {
                const uint32_t fd = (d[q] & pbit) ? f : 0u;
                b ^= (((b ^ cv) & pbit) == 0u) ? f : 0u;
                d[q] ^= fd;
            }
        }
        a[0] = b;
        for (int r = 1; r < 16; ++r) a[r] = a[r & (r - 1)] ^ d[__builtin_ctz(r)];
    } else {
        for (int r = 0; r < 16; ++r) a[r] = tb | slot_offset(R, r);
        for (int i = 0; i < n; ++i) {
            const qsv_rop_t &p = opat(i);
            const bool e = p.ext_mask ? ext_true : true;
            for (int r = 0; r < 16; ++r)
                if (e && (a[r] & p.tmask) == p.tval) a[r] ^= p.table_off;
        }
    }
}

static uint32_t deposit_tid(uint32_t tid, const int pos[8]) {
    uint32_t tb = 0;
    for (int i = 0; i < 8; ++i) tb |= ((tid >> i) & 1u) << pos[i];
    return tb;
}

// bank-group collisions of one mapping: sum over sampled quarter-warps, slots and both values of the external
// conditions of (largest number of lanes in one 16-byte bank group) - 1, for the round's loads and (unless the round
// stores straight to global memory) its stores
static long conflict_score(const qsv_round_t &R, const qsv_rop_t *ops, const int pos[8], bool score_stores) {
    const int n_pre = R.n_pre & 0x7fff, n_post = R.n_post & 0x7fff;
    const bool gen_pre = (R.n_pre & 0x8000) != 0, gen_post = (R.n_post & 0x8000) != 0;
    const int n_reg = (int)(R.n_ops & 0xffffu), n_pt = (int)((R.n_ops >> 16) & 0x3fffu);
    const int f_pre = (int)R.first_op, f_post = f_pre + n_pre + n_pt + n_reg;
    static const int kQw[] = {0, 3, 6, 9, 13, 18, 22, 27, 31};
    long score = 0;
    for (int side = 0; side < (score_stores ? 2 : 1); ++side)
        for (int ext = 0; ext < 2; ++ext)
            for (int qw : kQw) {
                uint32_t a[8][16];
                for (int lane = 0; lane < 8; ++lane) {
                    const uint32_t tb = deposit_tid((uint32_t)(qw * 8 + lane), pos);
                    if (side == 0) eval_addresses(R, ops, f_pre, n_pre, /*reverse=*/true, gen_pre, tb, ext != 0, a[lane]);
                    else eval_addresses(R, ops, f_post, n_post, /*reverse=*/false, gen_post, tb, ext != 0, a[lane]);
                }
                for (int r = 0; r < 16; ++r) {
                    int cnt[8] = {0}, worst = 0;
                    for (int lane = 0; lane < 8; ++lane) {
                        const int g = (int)(pad_slot(a[lane][r]) & 7u);
                        if (++cnt[g] > worst) worst = cnt[g];
                    }
                    score += worst - 1;
                }
            }
    return score;
}

static void search_thread_bits(const qsv_round_t &R, const qsv_rop_t *ops, bool score_stores, int best[8]) {
    thread_bit_order(R, best);
    long best_score = conflict_score(R, ops, best, score_stores);
    if (best_score == 0) return;
    int free_bits[8], n = 0;
    bool is_rb[12] = {false};
    for (int q = 0; q < 4; ++q) is_rb[R.rb[q]] = true;
    for (int b = 0; b < 12; ++b) if (!is_rb[b]) free_bits[n++] = b;
    for (int i = 0; i < 8; ++i)
        for (int j = i + 1; j < 8; ++j)
            for (int k = j + 1; k < 8; ++k) {
                int cand[8], m = 0;
                cand[m++] = free_bits[i]; cand[m++] = free_bits[j]; cand[m++] = free_bits[k];
                for (int t = 0; t < 8; ++t) if (t != i && t != j && t != k) cand[m++] = free_bits[t];
                const long sc = conflict_score(R, ops, cand, score_stores);
                if (sc < best_score) { best_score = sc; memcpy(best, cand, sizeof(cand)); }
            }
}

// best layout (padded or linear) of a pass under the search, with the mapping of every round; returns "use padding"
static bool search_layout(const std::vector<qsv_round_t> &rounds, const qsv_rop_t *ops, int L, bool direct, bool may_pad,
                          std::vector<std::array<int, 8>> &maps, long *score_out = nullptr) {
    long best = -1;
    bool best_pad = false;
    for (int padded = 0; padded < (may_pad ? 2 : 1); ++padded) {
        t_pad_L = padded ? L : -1;
        std::vector<std::array<int, 8>> m(rounds.size());
        long total = 0;
        for (size_t k = 0; k < rounds.size(); ++k) {
            const bool stores = !(k + 1 == rounds.size() && direct);
            search_thread_bits(rounds[k], ops, stores, m[k].data());
            total += conflict_score(rounds[k], ops, m[k].data(), stores);
        }
        if (best < 0 || total < best) { best = total; best_pad = padded != 0; maps.swap(m); }
    }
    t_pad_L = -1;
    if (score_out) *score_out = best;
    return best_pad;
}

static void emit_tb(Out &o, const qsv_round_t &R) {
    int pos[8];
    if (t_tb_override) memcpy(pos, t_tb_override, sizeof(pos));
    else thread_bit_order(R, pos);
    std::string e;
    for (int i = 0; i < 8;) {
        int len = 1;
        while (i + len < 8 && pos[i + len] == pos[i] + len) ++len;
        char b[96];
        snprintf(b, sizeof(b), "(((tid >> %d) & 0x%xu) << %d)", i, (1u << len) - 1u, pos[i]);
        e += (e.empty() ? "" : " | ") + std::string(b);
        i += len;
    }
    o.f("    const u32 tb = %s;", e.c_str());
}

// address permutation list in application order [first, first+n) step dir; result: a0..a15 declared
static void emit_addresses(Out &o, const qsv_round_t &R, const qsv_rop_t *ops, int first, int n, bool reverse,
                           bool general, const char *pfx) {
    auto opat = [&](int i) -> const qsv_rop_t & { return ops[reverse ? first + n - 1 - i : first + i]; };
    if (!general) {
        o.f("    u32 %sb = tb, %sd0 = 0x%xu, %sd1 = 0x%xu, %sd2 = 0x%xu, %sd3 = 0x%xu;", pfx, pfx, 1u << R.rb[0], pfx,
            1u << R.rb[1], pfx, 1u << R.rb[2], pfx, 1u << R.rb[3]);
        for (int i = 0; i < n; ++i) {
            const qsv_rop_t &p = opat(i);
            const uint32_t cm = p.tmask, cv = p.tval, xm = p.table_off, q = p.j, pbit = p.flags;
            const uint32_t cmu = cm & ~pbit;
            std::string e = p.ext_mask ? cond_ext(p.ext_mask, p.ext_val) + " && " : "";
            if (cmu)
                o.f("    { const u32 f = (%s((%sb & 0x%xu) == 0x%xu)) ? 0x%xu : 0u;", e.c_str(), pfx, cmu, cv & cmu, xm);
            else if (p.ext_mask)
                o.f("    { const u32 f = %s ? 0x%xu : 0u;", cond_ext(p.ext_mask, p.ext_val).c_str(), xm);
            else
                o.f("    { const u32 f = 0x%xu;", xm);
            if (q == 0xffu || pbit == 0u) {
                o.f("      %sb ^= f; }", pfx);
            } else {
                o.f("      const u32 fd = (%sd%u & 0x%xu) ? f : 0u;", pfx, q, pbit);
                o.f("      %sb ^= (((%sb ^ 0x%xu) & 0x%xu) == 0u) ? f : 0u;", pfx, pfx, cv, pbit);
                o.f("      %sd%u ^= fd; }", pfx, q);
            }
        }
        o.f("    const u32 %s0 = %sb;", pfx, pfx);
        for (int r = 1; r < 16; ++r) {
            const int low = __builtin_ctz(r);
            o.f("    const u32 %s%d = %s%d ^ %sd%d;", pfx, r, pfx, r & (r - 1), pfx, low);
        }
    } else {
        for (int r = 0; r < 16; ++r) {
            uint32_t off = 0;
            for (int q = 0; q < 4; ++q) if ((r >> q) & 1) off |= 1u << R.rb[q];
            o.f("    u32 %s%d = tb | 0x%xu;", pfx, r, off);
        }
        for (int i = 0; i < n; ++i) {
            const qsv_rop_t &p = opat(i);
            std::string e = p.ext_mask ? cond_ext(p.ext_mask, p.ext_val) + " && " : "";
            for (int r = 0; r < 16; ++r)
                o.f("    %s%d ^= (%s((%s%d & 0x%xu) == 0x%xu)) ? 0x%xu : 0u;", pfx, r, e.c_str(), pfx, r, p.tmask, p.tval,
                    p.table_off);
        }
    }
}

// padded slot of a tile-local address known at generation time
static uint32_t pad_slot(uint32_t a) { return t_pad_L >= 0 ? a + (a >> t_pad_L) + (a >> (t_pad_L + 3)) : a; }
// tile-local offset of register slot r (the round's register bits set according to r)
static uint32_t slot_offset(const qsv_round_t &R, int r) {
    uint32_t off = 0;
    for (int q = 0; q < 4; ++q) if ((r >> q) & 1) off |= 1u << R.rb[q];
    return off;
}

static bool butterfly(const double *m, double &h, int sg[4]) {
    // real matrix with all four entries of the same magnitude h
    if (m[1] != 0.0 || m[3] != 0.0 || m[5] != 0.0 || m[7] != 0.0) return false;
    h = fabs(m[0]);
    if (h == 0.0) return false;
    for (int k = 0; k < 4; ++k) {
        if (fabs(m[2 * k]) != h) return false;
        sg[k] = m[2 * k] < 0.0 ? -1 : 1;
    }
    return true;
}

static bool is_phase_kind(int kind) { return kind >= QSV_ROP_PHASE1 && kind <= QSV_ROP_GEN_PHASE; }
static void phase_rmask(const qsv_rop_t &p, uint32_t &rmask, uint32_t &rval) {
    rmask = p.rmask;
    rval = p.rval;
    if (p.kind != QSV_ROP_GEN_PHASE) {
        const int J = (p.kind - QSV_ROP_PHASE1) >> 1, V = (p.kind - QSV_ROP_PHASE1) & 1;
        rmask = 1u << J;
        rval = (uint32_t)V << J;
    }
}
// a conditional phase on ONE register bit whose factor is not -1 joins a runtime scalar (a conditional -1
// is 16 integer sign flips under a branch: cheaper than any multiplication)
static bool joins_scalar(const qsv_rop_t &p) {
    if (!is_phase_kind(p.kind) || (p.tmask == 0 && p.ext_mask == 0)) return false;
    uint32_t rmask, rval;
    phase_rmask(p, rmask, rval);
    if (rmask == 0 || (rmask & (rmask - 1))) return false;
    return !(p.m[0] == -1.0 && p.m[1] == 0.0);
}

// direct: 0 = the round stores to the shared-memory tile; 1 = LAST round of the direct-store kernel: once every
// thread holds its registers the tile buffer is handed back (QSV_TILE_CONSUMED: barrier + TMA load of the CTA's
// next tile) and the results go from registers straight to global memory.
static int gen_round(Out &o, int k, const qsv_round_t &R, const qsv_rop_t *ops, RoundInfo &info, bool last_round,
                     double &scale, bool defer_scale, std::string &err, int direct = 0, const char *suffix = "",
                     bool emit_hoist = true) {
    const int n_pre = R.n_pre & 0x7fff, n_post = R.n_post & 0x7fff;
    const bool gen_pre = (R.n_pre & 0x8000) != 0, gen_post = (R.n_post & 0x8000) != 0;
    const int n_reg = (int)(R.n_ops & 0xffffu), n_pt = (int)((R.n_ops >> 16) & 0x3fffu);
    const int f_pre = (int)R.first_op, f_pt = f_pre + n_pre, f_reg = f_pt + n_pt, f_post = f_reg + n_reg;
    info.barrier = (n_pre + n_post) > 0;

    // ---- runtime scalars of the round
    std::vector<Scalar> scalars;
    std::vector<int> op_scalar(n_reg, -1);
    if (n_pt > 0) {
        Scalar S;
        for (int i = 0; i < n_pt; ++i) {
            const qsv_rop_t &p = ops[f_pt + i];
            Factor f;
            f.tcond = p.tmask ? cond_thr(p.tmask, p.tval) : "";
            f.ext_mask = p.ext_mask;
            f.ext_val = p.ext_val;
            f.s = {p.m[0], p.m[1]};
            S.f.push_back(f);
        }
        scalars.push_back(S);
    }
    const bool have_S = n_pt > 0;
    {
        int window[4] = {0, 0, 0, 0};
        for (int i = 0; i < n_reg; ++i) {
            const qsv_rop_t &p = ops[f_reg + i];
            if (p.kind >= QSV_ROP_DENSE1 && p.kind < QSV_ROP_PHASE1) {
                window[p.kind & 3]++;
                continue;
            }
            if (p.kind == QSV_ROP_MCX_REG) {
                window[p.j & 3]++;
                continue;
            }
            if (!joins_scalar(p)) continue;
            uint32_t rmask, rval;
            phase_rmask(p, rmask, rval);
            const int J = __builtin_ctz(rmask), V = rval ? 1 : 0;
            int found = -1;
            for (size_t j = 0; j < scalars.size(); ++j)
                if (scalars[j].J == J && scalars[j].V == V && scalars[j].window == window[J]) found = (int)j;
            if (found < 0) {
                Scalar sc;
                sc.J = J;
                sc.V = V;
                sc.window = window[J];
                scalars.push_back(sc);
                found = (int)scalars.size() - 1;
            }
            Factor f;
            f.tcond = p.tmask ? cond_thr(p.tmask, p.tval) : "";
            f.ext_mask = p.ext_mask;
            f.ext_val = p.ext_val;
            f.s = {p.m[0], p.m[1]};
            scalars[found].f.push_back(f);
            op_scalar[i] = found;
        }
    }
    int n_hoist = 0;
    for (size_t j = 0; j < scalars.size(); ++j) {
        scalars[j].id = (int)j;
        bool any = false;
        for (const Factor &f : scalars[j].f) any = any || f.ext_mask == 0;
        if (any) { scalars[j].hoist_slot = n_hoist; n_hoist += 2; }
    }
    info.n_hoisted = n_hoist;

    // ---- the part of the scalars that depends on the thread index only: once per CTA, not once per tile
    if (emit_hoist) {
    o.f("struct qsv_H%d { double v[%d]; };", k, n_hoist ? n_hoist : 1);
    o.f("QSV_FN qsv_H%d qsv_hoist%d(const u32 tid) {", k, k);
    o.f("    qsv_H%d H;", k);
    if (n_hoist) {
        emit_tb(o, R);
        for (const Scalar &sc : scalars) {
            if (sc.hoist_slot < 0) continue;
            o.f("    double %s = 1.0, %s = 0.0;", sc.re().c_str(), sc.im().c_str());
            for (const Factor &f : sc.f)
                if (f.ext_mask == 0) emit_scalar_mul(o, "    ", sc.re(), sc.im(), f.tcond, f.s);
            o.f("    H.v[%d] = %s; H.v[%d] = %s;", sc.hoist_slot, sc.re().c_str(), sc.hoist_slot + 1, sc.im().c_str());
        }
    } else {
        o.f("    H.v[0] = 0.0; (void)tid;");
    }
    o.f("    return H;");
    o.f("}");
    }

    if (direct && t_sums) {
        // Tile sums of the sums variant.  Bucket bits on positions outside the tile are the same for the whole tile.  A
        // bucket bit INSIDE the tile is read from the amplitude's FINAL global index (the address it is stored to:
        // X / CNot riding on the stores move amplitudes between register slots and threads), so every thread keeps one
        // partial sum per combination of the inside bits (acc[]); the warps reduce them with __reduce_add_sync (16-bit
        // limbs: 32 lanes cannot overflow a word), lane 0 adds to a small shared-memory table, and after the barrier one
        // thread per combination adds its 128-bit total to the bucket array in global memory.
        int in_kb[4], n_in = 0;
        for (int kb = 0; kb < t_sums->n_bits; ++kb)
            for (int t = 0; t < 12; ++t)
                if (t_tile_phys[t] == t_sums->pos[kb]) in_kb[n_in++] = kb;
        const int n_ent = 1 << n_in;
        o.f("#ifndef QSV_JIT_HOST");
        o.f("__device__ __forceinline__ void qsv_tile_sum(const qsv_ctx &c, const u64 gbase, const u32 tid, const qsv_u128 *acc) {");
        o.f("    u32 *tab = c.s_tab + c.par * %du;", n_ent * 8);
        o.f("    #pragma unroll");
        o.f("    for (int a_ = 0; a_ < %d; ++a_) {", n_ent);
        o.f("        #pragma unroll");
        o.f("        for (int i_ = 0; i_ < 8; ++i_) {");
        o.f("            const u32 limb = (u32)(((i_ < 4 ? acc[a_].lo : acc[a_].hi) >> (16 * (i_ & 3))) & 0xffffULL);");
        o.f("            const u32 s_ = __reduce_add_sync(0xffffffffu, limb);");
        o.f("            if ((tid & 31u) == 0u && s_) atomicAdd(tab + (u32)a_ * 8u + (u32)i_, s_);");
        o.f("        }");
        o.f("    }");
        o.f("    __syncthreads();");
        o.f("    if (tid < %du) {", n_ent);
        o.f("        u32 *t_ = tab + tid * 8u;");
        o.f("        qsv_u128 v = {0ULL, 0ULL};");
        o.f("        #pragma unroll");
        o.f("        for (int i_ = 0; i_ < 8; ++i_) {");
        o.f("            const u64 s_ = t_[i_];");
        o.f("            t_[i_] = 0u;");
        o.f("            if (i_ < 4) {");
        o.f("                const u64 add = s_ << (16 * i_);");
        o.f("                v.lo += add;");
        o.f("                v.hi += (v.lo < add ? 1ULL : 0ULL) + (i_ ? (s_ >> (64 - 16 * (i_ ? i_ : 1))) : 0ULL);");
        o.f("            } else {");
        o.f("                v.hi += s_ << (16 * (i_ - 4));");
        o.f("            }");
        o.f("        }");
        {
            std::string e = "qsv_bucket_ext(gbase)";
            for (int j = 0; j < n_in; ++j)
                e += " | (((tid >> " + std::to_string(j) + ") & 1u) << " + std::to_string(in_kb[j]) + ")";
            o.f("        if (v.lo | v.hi) {");
            o.f("            u64 *b = c.bk + 4ULL * (%s);", e.c_str());
        }
        o.f("            atomicAdd(b + 0, v.lo & 0xffffffffULL); atomicAdd(b + 1, v.lo >> 32);");
        o.f("            atomicAdd(b + 2, v.hi & 0xffffffffULL); atomicAdd(b + 3, v.hi >> 32);");
        o.f("        }");
        o.f("    }");
        o.f("}");
        uint32_t sd[4];
        if (sum_static(R, ops, f_post, n_post, gen_post, sd)) {
            // per-slot constants c_r (inside bucket bits of XOR_{q in r} d_q) -> one partial sum per distinct value;
            // the per-thread part s0 comes in as an argument.  Lanes of a warp that differ in s0 reduce in separate
            // groups (__match_any_sync), the first lane of a group adds to the shared-memory table.
            int tl_j[4];
            for (int j = 0; j < n_in; ++j)
                for (int t = 0; t < 12; ++t) if (t_tile_phys[t] == t_sums->pos[in_kb[j]]) tl_j[j] = t;
            std::vector<uint32_t> cs;
            for (int r = 0; r < 16; ++r) {
                uint32_t x = 0, c = 0;
                for (int q = 0; q < 4; ++q) if ((r >> q) & 1) x ^= sd[q];
                for (int j = 0; j < n_in; ++j) c |= ((x >> tl_j[j]) & 1u) << j;
                if (std::find(cs.begin(), cs.end(), c) == cs.end()) cs.push_back(c);
            }
            o.f("__device__ __forceinline__ void qsv_sum_one(u32 *tab, const u32 ent, const bool lane0, const qsv_u128 &a) {");
            o.f("    #pragma unroll");
            o.f("    for (int i_ = 0; i_ < 8; ++i_) {");
            o.f("        const u32 limb = (u32)(((i_ < 4 ? a.lo : a.hi) >> (16 * (i_ & 3))) & 0xffffULL);");
            o.f("        const u32 s_ = __reduce_add_sync(0xffffffffu, limb);");
            o.f("        if (lane0 && s_) atomicAdd(tab + ent * 8u + (u32)i_, s_);");
            o.f("    }");
            o.f("}");
            o.f("__device__ __forceinline__ void qsv_tile_sum_s(const qsv_ctx &c, const u64 gbase, const u32 tid, const qsv_u128 *acc,"
                " const u32 s0) {");
            o.f("    u32 *tab = c.s_tab + c.par * %du;", n_ent * 8);
            o.f("    const bool lane0 = (tid & 31u) == 0u;");
            if (n_in == 0) {
                o.f("    qsv_sum_one(tab, 0u, lane0, acc[0]); (void)s0;");
            } else {
                // REDUX needs one mask for the whole warp.  Warps whose lanes agree on s0 (the inside bucket bits are held
                // by warp-index bits of the thread number) reduce every partial sum once; the others offer every table
                // entry the partial sum that belongs to it (or zero).  The thread mapping is fixed per kernel: a warp
                // always takes the same side of this warp-uniform branch.
                o.f("    if (__all_sync(0xffffffffu, s0 == __shfl_sync(0xffffffffu, s0, 0))) {");
                for (size_t ci = 0; ci < cs.size(); ++ci)
                    o.f("        qsv_sum_one(tab, s0 ^ 0x%xu, lane0, acc[%d]);", cs[ci], (int)ci);
                o.f("    } else {");
                for (int e_ = 0; e_ < n_ent; ++e_) {
                    o.f("        { qsv_u128 v_ = {0ULL, 0ULL};");
                    for (size_t ci = 0; ci < cs.size(); ++ci)
                        o.f("          if ((s0 ^ 0x%xu) == %du) v_ = acc[%d];", cs[ci], e_, (int)ci);
                    o.f("          qsv_sum_one(tab, %du, lane0, v_); }", e_);
                }
                o.f("    }");
            }
        o.f("    __syncthreads();");
        o.f("    if (tid < %du) {", n_ent);
        o.f("        u32 *t_ = tab + tid * 8u;");
        o.f("        qsv_u128 v = {0ULL, 0ULL};");
        o.f("        #pragma unroll");
        o.f("        for (int i_ = 0; i_ < 8; ++i_) {");
        o.f("            const u64 s_ = t_[i_];");
        o.f("            t_[i_] = 0u;");
        o.f("            if (i_ < 4) {");
        o.f("                const u64 add = s_ << (16 * i_);");
        o.f("                v.lo += add;");
        o.f("                v.hi += (v.lo < add ? 1ULL : 0ULL) + (i_ ? (s_ >> (64 - 16 * (i_ ? i_ : 1))) : 0ULL);");
        o.f("            } else {");
        o.f("                v.hi += s_ << (16 * (i_ - 4));");
        o.f("            }");
        o.f("        }");
        {
            std::string e = "qsv_bucket_ext(gbase)";
            for (int j = 0; j < n_in; ++j)
                e += " | (((tid >> " + std::to_string(j) + ") & 1u) << " + std::to_string(in_kb[j]) + ")";
            o.f("        if (v.lo | v.hi) {");
            o.f("            u64 *b = c.bk + 4ULL * (%s);", e.c_str());
        }
        o.f("            atomicAdd(b + 0, v.lo & 0xffffffffULL); atomicAdd(b + 1, v.lo >> 32);");
        o.f("            atomicAdd(b + 2, v.hi & 0xffffffffULL); atomicAdd(b + 3, v.hi >> 32);");
        o.f("        }");
        o.f("    }");
        o.f("}");
        }
        o.f("#endif");
    }
    // round 0 reads the tile as it came from global memory: in a FRESH state (no memset behind qsv_init_basis_sparse)
    // everything outside the support is uninitialised, and the amplitudes whose tile-local index disagrees with the
    // support's fixed bits (zf_mask / zf_val, tile-local) are taken as zero instead of what was loaded
    const char *zf_params = k == 0 ? ", const u32 zf_mask, const u32 zf_val" : "";
    if (direct)
        o.f("QSV_FN void qsv_round%d%s(const double2 *tin, double2 *gout, const u64 tile_base, const u32 tid, const u64 gbase,"
            " const qsv_H%d &H, const double2 *tables, const qsv_ctx &ctx%s) {", k, suffix, k, zf_params);
    else
        o.f("QSV_FN void qsv_round%d%s(const double2 *tin, double2 *tout, const u32 tid, const u64 gbase, const qsv_H%d &H,"
            " const double2 *tables%s) {", k, suffix, k, zf_params);
    reset_slot_vars();
    emit_tb(o, R);
    o.f("    double x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i, x4r, x4i, x5r, x5i, x6r, x6i, x7r, x7i;");
    o.f("    double x8r, x8i, x9r, x9i, x10r, x10i, x11r, x11i, x12r, x12i, x13r, x13i, x14r, x14i, x15r, x15i;");
    if (n_pre == 0) {
        // no permutation rides on the loads: slot r sits at tb | offset(r); thread bits and register bits are
        // disjoint, so the padded slot splits into QP(tb) + a literal (an immediate offset of the load)
        o.f("    const u32 pa = QP(tb);");
        for (int r = 0; r < 16; ++r)
            o.f("    { const double2 v = tin[pa + %uu]; x%dr = v.x; x%di = v.y; }", pad_slot(slot_offset(R, r)), r, r);
        if (k == 0) {
            o.f("    if (zf_mask) {");
            o.f("        const u32 zt = (tb ^ zf_val) & zf_mask;");
            for (int r = 0; r < 16; ++r)
                o.f("        if (zt ^ (0x%xu & zf_mask)) { x%dr = 0.0; x%di = 0.0; }", slot_offset(R, r), r, r);
            o.f("    }");
        }
    } else {
        emit_addresses(o, R, ops, f_pre, n_pre, /*reverse=*/true, gen_pre, "a");
        for (int r = 0; r < 16; ++r) o.f("    { const double2 v = tin[QP(a%d)]; x%dr = v.x; x%di = v.y; }", r, r, r);
        if (k == 0) {
            o.f("    if (zf_mask) {");
            for (int r = 0; r < 16; ++r)
                o.f("        if ((a%d ^ zf_val) & zf_mask) { x%dr = 0.0; x%di = 0.0; }", r, r, r);
            o.f("    }");
        }
    }
    o.f("    (void)H; (void)tables; (void)gbase;");
    if (direct) o.f("    QSV_TILE_CONSUMED(ctx);");

    // runtime scalars: hoisted part x the factors that read external bits (uniform per tile)
    for (const Scalar &sc : scalars) {
        if (sc.hoist_slot >= 0)
            o.f("    double %s = H.v[%d], %s = H.v[%d];", sc.re().c_str(), sc.hoist_slot, sc.im().c_str(), sc.hoist_slot + 1);
        else
            o.f("    double %s = 1.0, %s = 0.0;", sc.re().c_str(), sc.im().c_str());
        for (const Factor &f : sc.f) {
            if (f.ext_mask == 0) continue;
            std::string c = cond_ext(f.ext_mask, f.ext_val);
            if (!f.tcond.empty()) c += " && " + f.tcond;
            emit_scalar_mul(o, "    ", sc.re(), sc.im(), c, f.s);
        }
    }

    Pending pd;
    int tmp_id = 0;
    int window[4] = {0, 0, 0, 0};
    // flush: constants (all of them) and the runtime scalars of register bit J's current window
    // (J < 0: everything that is left, S included, and the deferred butterfly factor in the last round)
    auto flush = [&](int J) {
        SlotMul m[16];
        bool any = false;
        const bool final_flush = J < 0;
        double g = 1.0;
        if (final_flush && last_round && scale != 1.0) { g = scale; scale = 1.0; }
        if (final_flush || pd.depends_on(J)) {
            for (int r = 0; r < 16; ++r) {
                m[r].c = {pd.c[r].r * g, pd.c[r].i * g};
                any = any || !m[r].c.is_one();
            }
            pd.reset();
        }
        for (Scalar &sc : scalars) {
            if (sc.applied) continue;
            if (sc.J < 0) {
                if (!final_flush) continue;
                for (int r = 0; r < 16; ++r) m[r].sc.push_back(sc.id);
            } else {
                if (!final_flush && !(sc.J == J && sc.window == window[J])) continue;
                for (int r = 0; r < 16; ++r) if (((r >> sc.J) & 1) == sc.V) m[r].sc.push_back(sc.id);
            }
            sc.applied = true;
            any = true;
        }
        if (any) apply_multipliers(o, m, scalars, tmp_id);
    };

    for (int i = 0; i < n_reg; ++i) {
        const qsv_rop_t &p = ops[f_reg + i];
        const std::string cnd = op_cond(p);
        const int kind = p.kind;
        if (kind >= QSV_ROP_DENSE1 && kind < QSV_ROP_PHASE1) {
            const int J = kind & 3;
            flush(J);
            window[J]++;
            double h;
            int sg[4];
            const bool bf = butterfly(p.m, h, sg);
            const bool defer = bf && defer_scale && cnd.empty();
            if (!cnd.empty()) o.f("    if (%s) {", cnd.c_str());
            for (int r = 0; r < 16; ++r) {
                if ((r >> J) & 1) continue;
                const int r1 = r | (1 << J);
                o.f("    { const double p0r = %s, p0i = %s, p1r = %s, p1i = %s;", xr(r).c_str(), xi(r).c_str(),
                    xr(r1).c_str(), xi(r1).c_str());
                if (bf) {
                    const std::string hs = defer ? "" : dlit(h) + " * ";
                    auto comb = [&](int s0, int s1, const char *a, const char *b) {
                        std::string e = "(";
                        e += (s0 < 0 ? "-" : "");
                        e += a;
                        e += (s1 < 0 ? " - " : " + ");
                        e += b;
                        return e + ")";
                    };
                    o.f("      %s = %s%s; %s = %s%s;", xr(r).c_str(), hs.c_str(), comb(sg[0], sg[1], "p0r", "p1r").c_str(),
                        xi(r).c_str(), hs.c_str(), comb(sg[0], sg[1], "p0i", "p1i").c_str());
                    o.f("      %s = %s%s; %s = %s%s; }", xr(r1).c_str(), hs.c_str(), comb(sg[2], sg[3], "p0r", "p1r").c_str(),
                        xi(r1).c_str(), hs.c_str(), comb(sg[2], sg[3], "p0i", "p1i").c_str());
                } else {
                    const double *m = p.m;      // a = m[0..1], b = m[2..3], c = m[4..5], d = m[6..7]
                    o.f("      %s = %s;", xr(r).c_str(),
                        lin({{m[0], "p0r"}, {-m[1], "p0i"}, {m[2], "p1r"}, {-m[3], "p1i"}}).c_str());
                    o.f("      %s = %s;", xi(r).c_str(),
                        lin({{m[0], "p0i"}, {m[1], "p0r"}, {m[2], "p1i"}, {m[3], "p1r"}}).c_str());
                    o.f("      %s = %s;", xr(r1).c_str(),
                        lin({{m[4], "p0r"}, {-m[5], "p0i"}, {m[6], "p1r"}, {-m[7], "p1i"}}).c_str());
                    o.f("      %s = %s; }", xi(r1).c_str(),
                        lin({{m[4], "p0i"}, {m[5], "p0r"}, {m[6], "p1i"}, {m[7], "p1r"}}).c_str());
                }
            }
            if (!cnd.empty()) o.f("    }");
            if (defer) scale *= h;
        } else if (kind == QSV_ROP_MCX_REG) {
            const int J = p.j & 3;
            flush(J);                                  // diagonals that depend on bit J do not commute with it
            window[J]++;
            if (cnd.empty()) {
                // slot renaming at generation time: zero instructions
                for (int r = 0; r < 16; ++r) {
                    if ((r >> J) & 1) continue;
                    if (((uint32_t)r & p.rmask) != p.rval) continue;
                    std::swap(t_slot_var[r], t_slot_var[r | (1 << J)]);
                }
            } else {
                o.f("    if (%s) {", cnd.c_str());
                for (int r = 0; r < 16; ++r) {
                    if ((r >> J) & 1) continue;
                    if (((uint32_t)r & p.rmask) != p.rval) continue;
                    const int r1 = r | (1 << J);
                    o.f("      { const double tr = %s, ti = %s; %s = %s; %s = %s; %s = tr; %s = ti; }", xr(r).c_str(),
                        xi(r).c_str(), xr(r).c_str(), xr(r1).c_str(), xi(r).c_str(), xi(r1).c_str(), xr(r1).c_str(),
                        xi(r1).c_str());
                }
                o.f("    }");
            }
        } else if (is_phase_kind(kind)) {
            if (op_scalar[i] >= 0) continue;          // part of a runtime scalar, applied at its flush
            uint32_t rmask, rval;
            phase_rmask(p, rmask, rval);
            const Cx s = {p.m[0], p.m[1]};
            if (cnd.empty()) {
                for (int r = 0; r < 16; ++r) if (((uint32_t)r & rmask) == rval) pd.c[r] = cx_mul(pd.c[r], s);
            } else {
                o.f("    if (%s) {", cnd.c_str());
                for (int r = 0; r < 16; ++r) if (((uint32_t)r & rmask) == rval) emit_cmul_const(o, r, s);
                o.f("    }");
            }
        } else if (kind == QSV_ROP_DIAG) {
            // table lookup: index bits from thread-held positions, external positions and register bits
            if (!cnd.empty()) o.f("    if (%s) {", cnd.c_str());
            o.f("    { u32 di = 0u;");
            const int kk = (int)p.k;
            for (int q = 0; q < kk; ++q) {
                const uint32_t sb = p.dsrc[q];
                if (sb == 0xffu) continue;
                if (sb & 0x80u)
                    o.f("      di |= (u32)((gbase >> %u) & 1ULL) << %d;", sb & 0x7fu, kk - 1 - q);
                else
                    o.f("      di |= ((tb >> %u) & 1u) << %d;", sb, kk - 1 - q);
            }
            o.f("      const double2 *tab = tables + %u;", p.table_off / 16u);
            for (int r = 0; r < 16; ++r) {
                if (((uint32_t)r & p.rmask) != p.rval) continue;
                uint32_t d = 0;
                for (int q = 0; q < 4; ++q) if ((r >> q) & 1) d |= p.rc[q];
                o.f("      { const double2 m = QSV_LDG(tab + (di | 0x%xu));", d);
                o.f("        const double t = %s * m.x - %s * m.y; %s = %s * m.y + %s * m.x; %s = t; }", xr(r).c_str(),
                    xi(r).c_str(), xi(r).c_str(), xr(r).c_str(), xi(r).c_str(), xr(r).c_str());
            }
            o.f("    }");
            if (!cnd.empty()) o.f("    }");
        } else {
            err = "register op kind " + std::to_string(kind) + " is not supported by the generator";
            return 1;
        }
    }
    flush(-1);
    (void)have_S;
    if (direct) {
        // registers -> global memory.  Tile-local address -> global index is a bit scatter (linear over XOR):
        // index(r) = tile_base | S(base) ^ XOR_{q in r} S(d_q)
        if (!gen_post) {
            emit_addresses(o, R, ops, f_post, n_post, /*reverse=*/false, false, "b");
            o.f("    const u64 g0 = tile_base | qsv_scatter(bb);");
            o.f("    const u64 gd0 = qsv_scatter(bd0), gd1 = qsv_scatter(bd1), gd2 = qsv_scatter(bd2), gd3 = qsv_scatter(bd3);");
            for (int r = 1; r < 16; ++r)
                o.f("    const u64 g%d = g%d ^ gd%d;", r, r & (r - 1), __builtin_ctz(r));
        } else {
            emit_addresses(o, R, ops, f_post, n_post, /*reverse=*/false, true, "b");
            for (int r = 0; r < 16; ++r) o.f("    const u64 g%d = tile_base | qsv_scatter(b%d);", r, r);
        }
        o.f("    QSV_WAIT_PEER(ctx);");
        for (int r = 0; r < 16; ++r) o.f("    QSV_STG(gout + g%d, %s, %s);", r, xr(r).c_str(), xi(r).c_str());
        if (t_sums) {
            o.f("#ifdef QSV_JIT_HOST");
            uint32_t hd[4];
            if (sum_static(R, ops, f_post, n_post, gen_post, hd)) {
                // host form of the STATIC classification: the bucket is composed the way the device code composes it
                // (bits outside the tile from gbase, the per-thread part from slot 0's final address, the per-slot
                // constant from the deltas known at generation time), not read from the amplitude's full index - the
                // CPU tests check the classification itself
                for (int r = 0; r < 16; ++r) {
                    uint32_t x = 0;
                    for (int q = 0; q < 4; ++q) if ((r >> q) & 1) x ^= hd[q];
                    std::string e = "qsv_bucket_ext(gbase)";
                    for (int kb = 0; kb < t_sums->n_bits; ++kb)
                        for (int t = 0; t < 12; ++t) {
                            if (t_tile_phys[t] != t_sums->pos[kb]) continue;
                            e += " | ((((u32)(g0 >> " + std::to_string(t_sums->pos[kb]) + ") & 1u) ^ " +
                                 std::to_string((x >> t) & 1u) + "u) << " + std::to_string(kb) + ")";
                        }
                    o.f("    qsv_host_add_b(%s, %s, %s);", e.c_str(), xr(r).c_str(), xi(r).c_str());
                }
            } else {
                for (int r = 0; r < 16; ++r) o.f("    qsv_host_add(g%d | gbase, %s, %s);", r, xr(r).c_str(), xi(r).c_str());
            }
            o.f("#else");
            // per-thread partial sums, one per combination of the bucket bits INSIDE the tile, selected by the final
            // global index g<r> of every amplitude
            int in_pos[4], n_in = 0;
            for (int kb = 0; kb < t_sums->n_bits; ++kb)
                for (int t = 0; t < 12; ++t)
                    if (t_tile_phys[t] == t_sums->pos[kb]) in_pos[n_in++] = t_sums->pos[kb];
            uint32_t sd[4];
            if (sum_static(R, ops, f_post, n_post, gen_post, sd)) {
                int tl_j[4];
                for (int j = 0; j < n_in; ++j)
                    for (int t = 0; t < 12; ++t) if (t_tile_phys[t] == in_pos[j]) tl_j[j] = t;
                std::vector<uint32_t> cs;
                int cidx[16];
                for (int r = 0; r < 16; ++r) {
                    uint32_t x = 0, c = 0;
                    for (int q = 0; q < 4; ++q) if ((r >> q) & 1) x ^= sd[q];
                    for (int j = 0; j < n_in; ++j) c |= ((x >> tl_j[j]) & 1u) << j;
                    auto it = std::find(cs.begin(), cs.end(), c);
                    if (it == cs.end()) { cs.push_back(c); cidx[r] = (int)cs.size() - 1; }
                    else cidx[r] = (int)(it - cs.begin());
                }
                std::string s0 = "0u";          // inside bucket bits of slot 0's final address
                for (int j = 0; j < n_in; ++j)
                    s0 += " | ((u32)((g0 >> " + std::to_string(in_pos[j]) + ") & 1ULL) << " + std::to_string(j) + ")";
                o.f("    { qsv_u128 acc[%d];", (int)cs.size());
                o.f("      for (int a_ = 0; a_ < %d; ++a_) acc[a_].lo = acc[a_].hi = 0ULL;", (int)cs.size());
                for (int r = 0; r < 16; ++r) o.f("      qsv_wacc(acc[%d], %s, %s);", cidx[r], xr(r).c_str(), xi(r).c_str());
                o.f("      qsv_tile_sum_s(ctx, gbase, tid, acc, %s); }", s0.c_str());
                o.f("#endif");
                o.f("}");
                return 0;
            }
            o.f("    { qsv_u128 acc[%d];", 1 << n_in);
            o.f("      for (int a_ = 0; a_ < %d; ++a_) acc[a_].lo = acc[a_].hi = 0ULL;", 1 << n_in);
            for (int r = 0; r < 16; ++r) {
                if (n_in == 0) {
                    o.f("      qsv_wacc(acc[0], %s, %s);", xr(r).c_str(), xi(r).c_str());
                    continue;
                }
                std::string sel = "0u";
                for (int j = 0; j < n_in; ++j)
                    sel += " | ((u32)((g" + std::to_string(r) + " >> " + std::to_string(in_pos[j]) + ") & 1ULL) << " +
                           std::to_string(j) + ")";
                o.f("      { qsv_u128 q_ = {0ULL, 0ULL}; qsv_wacc(q_, %s, %s); const u32 sel_ = %s;", xr(r).c_str(),
                    xi(r).c_str(), sel.c_str());
                o.f("        #pragma unroll");
                o.f("        for (int a_ = 0; a_ < %d; ++a_) {", 1 << n_in);
                o.f("          const u64 l_ = sel_ == (u32)a_ ? q_.lo : 0ULL, h_ = sel_ == (u32)a_ ? q_.hi : 0ULL;");
                o.f("          acc[a_].lo += l_; acc[a_].hi += h_ + (acc[a_].lo < l_ ? 1ULL : 0ULL); } }");
            }
            o.f("      qsv_tile_sum(ctx, gbase, tid, acc); }");
            o.f("#endif");
        }
        o.f("}");
        return 0;
    }
    if (info.barrier) o.f("    QSV_BARRIER();");
    if (n_post == 0) {
        o.f("    const u32 pb = QP(tb);");
        for (int r = 0; r < 16; ++r)
            o.f("    tout[pb + %uu] = make_double2(%s, %s);", pad_slot(slot_offset(R, r)), xr(r).c_str(), xi(r).c_str());
    } else {
        emit_addresses(o, R, ops, f_post, n_post, /*reverse=*/false, gen_post, "b");
        for (int r = 0; r < 16; ++r) o.f("    tout[QP(b%d)] = make_double2(%s, %s);", r, xr(r).c_str(), xi(r).c_str());
    }
    o.f("}");
    return 0;
}

static const char *kPrelude = R"SRC(// generated by libqsv (csrc/qsv_jit.cu) for ONE pass program - do not edit
typedef unsigned int u32;
typedef unsigned long long u64;
#ifdef QSV_JIT_HOST
#include <string.h>
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { double2 v; v.x = x; v.y = y; return v; }
static inline double QD(u64 h) { double d; memcpy(&d, &h, 8); return d; }
#define QNEG(v) (-(v))
#define QSV_FN static inline
#define QSV_BARRIER()
#define QSV_LDG(p) (*(p))
#define QSV_STG(p, re, im) (*(p) = make_double2(re, im))
#define QSV_TILE_CONSUMED(c)
#define QSV_WAIT_PEER(c)
struct qsv_ctx { int unused; };
#else
#define QD(h) __longlong_as_double((long long)(h))
#define QNEG(v) __hiloint2double(__double2hiint(v) ^ (int)0x80000000, __double2loint(v))
#define QSV_FN __device__ __forceinline__
#define QSV_BARRIER() __syncthreads()
#define QSV_LDG(p) __ldg(p)
#define QSV_STG(p, re, im) (*(p) = make_double2(re, im))
#define QSV_TILE_CONSUMED(c) qsv_tile_consumed(c)
#ifdef QSV_FUSED_SWAP
#define QSV_WAIT_PEER(c) qsv_wait_peer(c)
#else
#define QSV_WAIT_PEER(c)
#endif
#endif
)SRC";

static const char *kDeviceHelpers = R"SRC(
#ifndef QSV_JIT_HOST
__device__ __forceinline__ u32 smem_u32(const void *p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, u32 parity) {
    u32 done;
    do {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, u32 bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, u32 bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, const void *src, u32 bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)),
                 "r"(bytes) : "memory");
}
#endif
)SRC";

// exact weight sums of the sums variant: the arithmetic of csrc/qsv_sample.cu (weight_fix, u128_add, 4 x 32-bit limbs)
static const char *kSumHelpers = R"SRC(
struct qsv_u128 { u64 lo, hi; };
#ifdef QSV_JIT_HOST
static inline u64 qsv_dbits(double w) { u64 b; memcpy(&b, &w, 8); return b; }
static inline double qsv_weight(double re, double im) { return __builtin_fma(re, re, im * im); }
#else
#define qsv_dbits(w) ((u64)__double_as_longlong(w))
#define qsv_weight(re, im) __fma_rn(re, re, __dmul_rn(im, im))
#endif
QSV_FN void qsv_wacc(qsv_u128 &a, const double re, const double im) {
    const u64 b = qsv_dbits(qsv_weight(re, im));
    const int e = (int)((b >> 52) & 0x7ffULL);
    if (e == 0) return;
    const u64 m = (b & ((1ULL << 52) - 1ULL)) | (1ULL << 52);
    const int sh = e - 1075 + 96;
    u64 lo = 0ULL, hi = 0ULL;
    if (sh >= 75) { hi = ~0ULL; lo = ~0ULL; }
    else if (sh >= 64) hi = m << (sh - 64);
    else if (sh > 0) { lo = m << sh; hi = m >> (64 - sh); }
    else if (sh == 0) lo = m;
    else if (sh > -64) lo = m >> (-sh);
    a.lo += lo;
    a.hi += hi + (a.lo < lo ? 1ULL : 0ULL);
}
)SRC";

// fused swap + pass: two tile buffers and one CTA per SM (QSV_FUSE_BUFFERS=2), or the single buffer and two CTAs per SM
static bool fuse_two_buffers() {
    const char *e = getenv("QSV_FUSE_BUFFERS");
    return e && !strcmp(e, "2");
}
static unsigned long long watchdog_ns() {
    const char *e = getenv("QSV_FUSE_WATCHDOG_MS");
    const long long ms = e ? atoll(e) : 30000;
    return (unsigned long long)(ms > 0 ? ms : 30000) * 1000000ull;
}
static bool pad_enabled() {
    const char *e = getenv("QSV_JIT_PAD");
    return !(e && !strcmp(e, "0"));
}
// shared-memory slots (16 bytes each) of one tile: 2^12 amplitudes plus one pad slot per run
static unsigned tile_slots(int n_hi) { return 4096u + (pad_enabled() ? (1u << n_hi) + ((1u << n_hi) >> 3) : 0u); }

// Fused variant: the global<->local qubit swap that precedes the pass runs INSIDE the pass kernel.  `q[k]` (ascending
// local positions outside the tile, >= low_bits) trades places with global bit k.
// pre = 1: the pass is the one BEFORE the swap - its ops read the coordinates (rank, swapped field) the tile has on the
// rank it is pulled from; pre = 0: the pass AFTER the swap (coordinates of the destination).
// `q[k]` trades places with global bit `gb[k]`; g may be smaller than the number of global qubits (partial swap).
struct SwapSpec { int g; int q[4]; int gb[4]; int pre; int n_local; };
constexpr unsigned kFlagCtas = 1024;      // flag slots per source rank in the peers' flag arrays

int generate(const qsv_pass_t &P, const void *prog_host, uint32_t prog_bytes, std::string &src, std::string &err,
             const SwapSpec *sw = nullptr, const SumSpec *sums = nullptr) {
    const qsv_rprog_hdr_t *hdr = static_cast<const qsv_rprog_hdr_t *>(prog_host);
    const uint8_t *body = static_cast<const uint8_t *>(prog_host) + sizeof(qsv_rprog_hdr_t);
    std::vector<qsv_round_t> rounds(hdr->n_rounds);
    std::vector<qsv_rop_t> ops(hdr->n_ops);
    memcpy(rounds.data(), body, sizeof(qsv_round_t) * hdr->n_rounds);
    memcpy(ops.data(), body + sizeof(qsv_round_t) * hdr->n_rounds, sizeof(qsv_rop_t) * hdr->n_ops);
    (void)prog_bytes;
    if (P.tile_bits != 12) { err = "generator needs 12 tile bits"; return 1; }
    const int L = P.low_bits, n_hi = P.n_hi;
    const bool defer_scale = getenv("QSV_JIT_NO_DEFER") == nullptr;
    const char *pf = getenv("QSV_JIT_PREFETCH");
    const bool prefetch_env = !(pf && !strcmp(pf, "0"));

    const char *ds = getenv("QSV_JIT_DIRECT_STORE");
    const bool direct = sw != nullptr || !(ds && !strcmp(ds, "0"));     // the fused variant stores from registers
    if (sw) {
        if (sw->g < 1 || sw->g > 4) { err = "fused swap: g out of range"; return 1; }
        for (int k = 0; k < sw->g; ++k) {
            if (sw->q[k] < L || sw->q[k] >= P.n_local || (k && sw->q[k] <= sw->q[k - 1])) {
                err = "fused swap: positions must be ascending local positions >= low_bits";
                return 1;
            }
            for (int j = 0; j < n_hi; ++j)
                if (P.hi_pos[j] == sw->q[k]) { err = "fused swap: a swapped position is inside the tile"; return 1; }
        }
    }
    const int n_runs = 1 << n_hi;
    const uint32_t nr = hdr->n_rounds;
    // padding only where it pays: a pass none of whose rounds holds tile bits 0..2 in registers is conflict-free
    // with the linear layout, and the padded slot arithmetic (shift + add per permuted address) is pure overhead
    bool low_rb = false;
    for (const qsv_round_t &R : rounds)
        for (int q = 0; q < 4; ++q) low_rb = low_rb || R.rb[q] < 3;
    const char *pe = getenv("QSV_JIT_PAD");
    const bool bank_search = bank_search_enabled();
    bool pad = pad_enabled() && (low_rb || (pe && !strcmp(pe, "all")));
    std::vector<std::array<int, 8>> tb_maps;
    if (bank_search) {
        // both layouts are scored with their best per-round mappings: padding turns XORs of run bits (tile bits >= L)
        // into bank changes, which helps rounds with low register bits and can hurt permuted rounds without them
        pad = search_layout(rounds, ops.data(), L, direct, pad_enabled(), tb_maps);
    }
    t_pad_L = pad ? L : -1;

    Out o;
    if (sw) { o.f("#define QSV_FUSED_SWAP 1"); o.f("#define QSV_NLOCAL %d", P.n_local); }
    if (sums) {
        if (sw) { err = "sums variant of a fused swap kernel is not supported"; return 1; }
        if (!direct) { err = "sums variant needs the direct-store kernel"; return 1; }
        if (sums->n_bits < 1 || sums->n_bits > 12) { err = "sums: n_bits out of range"; return 1; }
        int n_inside = 0;
        for (int k = 0; k < sums->n_bits; ++k) {
            if (sums->pos[k] < 0 || sums->pos[k] >= 48) { err = "sums: bucket position out of range"; return 1; }
            bool inside = sums->pos[k] < L;
            for (int j = 0; j < n_hi; ++j) inside = inside || P.hi_pos[j] == sums->pos[k];
            n_inside += inside;
        }
        if (n_inside > kSumMaxInside) { err = "sums: more than 3 bucket positions inside the tile"; return 1; }
        o.f("#define QSV_SUMS 1");
    }
    o.s += kPrelude;
    o.s += kDeviceHelpers;
    if (sums) {
        o.s += kSumHelpers;
        // bucket of a FULL global index (host form: every amplitude finds its own bucket)
        std::string e = "0u";
        for (int k = 0; k < sums->n_bits; ++k)
            e += " | ((u32)((full >> " + std::to_string(sums->pos[k]) + ") & 1ULL) << " + std::to_string(k) + ")";
        o.f("QSV_FN u32 qsv_bucket_of(const u64 full) { return %s; }", e.c_str());
        {
            std::string x = "0u";       // the bucket bits on positions OUTSIDE the tile (the same for a whole tile)
            for (int k = 0; k < sums->n_bits; ++k) {
                bool inside = sums->pos[k] < L;
                for (int j = 0; j < n_hi; ++j) inside = inside || P.hi_pos[j] == sums->pos[k];
                if (!inside)
                    x += " | ((u32)((gbase >> " + std::to_string(sums->pos[k]) + ") & 1ULL) << " + std::to_string(k) + ")";
            }
            o.f("QSV_FN u32 qsv_bucket_ext(const u64 gbase) { return %s; }", x.c_str());
        }
        o.f("#ifdef QSV_JIT_HOST");
        o.f("static u64 *qsv_host_bk = 0;");
        o.f("extern \"C\" void qsv_jit_host_set_buckets(u64 *bk) { qsv_host_bk = bk; }");
        o.f("static inline void qsv_host_add_b(const u32 bucket, const double re, const double im) {");
        o.f("    qsv_u128 v = {0ULL, 0ULL};");
        o.f("    qsv_wacc(v, re, im);");
        o.f("    u64 *b = qsv_host_bk + 4ULL * bucket;");
        o.f("    b[0] += v.lo & 0xffffffffULL; b[1] += v.lo >> 32; b[2] += v.hi & 0xffffffffULL; b[3] += v.hi >> 32;");
        o.f("}");
        o.f("static inline void qsv_host_add(const u64 full, const double re, const double im) {");
        o.f("    qsv_u128 v = {0ULL, 0ULL};");
        o.f("    qsv_wacc(v, re, im);");
        o.f("    u64 *b = qsv_host_bk + 4ULL * qsv_bucket_of(full);");
        o.f("    b[0] += v.lo & 0xffffffffULL; b[1] += v.lo >> 32; b[2] += v.hi & 0xffffffffULL; b[3] += v.hi >> 32;");
        o.f("}");
        o.f("#endif");
    }
    if (pad)
        o.f("#define QP(a) ((a) + ((a) >> %d) + ((a) >> %d))", L, L + 3);
    else
        o.f("#define QP(a) (a)");
    o.f("#define QSV_TILE_SLOTS %u", 4096u + (pad ? (unsigned)(n_runs + (n_runs >> 3)) : 0u));
    // tile-local address -> offset inside the state (bit scatter of the high tile bits)
    o.f("QSV_FN u64 qsv_scatter(const u32 a) {");
    {
        std::string e = "(u64)(a & 0x" + [&] { char b_[16]; snprintf(b_, sizeof(b_), "%x", (1u << L) - 1u); return std::string(b_); }() + "u)";
        for (int j = 0; j < n_hi; ++j)
            e += " | ((u64)((a >> " + std::to_string(L + j) + ") & 1u) << " + std::to_string(P.hi_pos[j]) + ")";
        o.f("    return %s;", e.c_str());
    }
    o.f("}");
    if (direct) {
        // hand-back of the tile buffer inside the last round: every thread holds its amplitudes in registers,
        // so the CTA's NEXT tile can stream into shared memory while this one is computed and stored
        o.f("#ifndef QSV_JIT_HOST");
        if (sw)
            o.f("struct qsv_ctx { const double2 *next_src; double2 *tile; u64 *mbar; const u64 *s_off; const u64 *wait_flag;"
                " u64 wait_val; u32 warp; u32 flags; };");
        else
        o.f("struct qsv_ctx { double2 *state; double2 *tile; u64 *mbar; const u64 *s_off; u64 next_base; u32 warp; u32 flags;"
            " u64 *bk; u32 *s_tab; u32 par; };");
        o.f("__device__ __forceinline__ void qsv_issue_load(const qsv_ctx &c) {");
        o.f("    if (c.flags & 2u) mbar_expect_tx(c.mbar, %uu);", 16u << 12);
        o.f("    u32 elected;");
        o.f("    asm volatile(\"{\\n.reg .pred P;\\nelect.sync _|P, 0xffffffff;\\nselp.u32 %%0, 1, 0, P;\\n}\\n\" : \"=r\"(elected));");
        o.f("    if (elected)");
        o.f("        for (u32 r = c.warp; r < %du; r += 8u)", n_runs);
        if (sw)
            o.f("            bulk_g2s(c.tile + (u64)QP(r << %d), c.next_src + c.s_off[r], %uu, c.mbar);", L, 16u << L);
        else
        o.f("            bulk_g2s(c.tile + (u64)QP(r << %d), c.state + c.next_base + c.s_off[r], %uu, c.mbar);", L, 16u << L);
        o.f("}");
        if (sw) {
            // per-tile handshake of a rank pair: "I have read your tile" (release store into the peer's flag array)
            // / "has my tile been read?" (acquire loads of my own flag array) - before the tile is overwritten
            o.f("__device__ __forceinline__ void qsv_signal_peer(u64 *flag, const u64 v) {");
            // relaxed: what the flag orders is MY completed TMA read of the peer's tile (observed through the mbarrier)
            // before the PEER's later stores - a release here would drain this CTA's own stores of the previous tile
            o.f("    asm volatile(\"st.relaxed.sys.global.u64 [%%0], %%1;\" ::\"l\"(flag), \"l\"(v) : \"memory\");");
            o.f("}");
            o.f("__device__ __forceinline__ void qsv_wait_peer(const qsv_ctx &c) {");
            o.f("    if (c.wait_flag) {");
            // bounded spin: a partner that never shows up (failed launch, dead rank) must not hang the node - after
            // QSV_FUSE_WATCHDOG_MS (default 30 s) of polling the CTA traps and the launch fails on the host
            o.f("        if (c.flags & 2u) {");
            o.f("            u64 v, t0 = 0; u32 polls = 0;");
            o.f("            for (;;) {");
            o.f("                asm volatile(\"ld.relaxed.sys.global.u64 %%0, [%%1];\" : \"=l\"(v) : \"l\"(c.wait_flag) : \"memory\");");
            o.f("                if (v >= c.wait_val) break;");
            o.f("                if ((++polls & 0x3ffu) == 0u) {");
            o.f("                    u64 now; asm volatile(\"mov.u64 %%0, %%globaltimer;\" : \"=l\"(now));");
            o.f("                    if (t0 == 0) t0 = now;");
            o.f("                    else if (now - t0 > %lluULL) { asm volatile(\"trap;\"); }", watchdog_ns());
            o.f("                }");
            o.f("            }");
            o.f("        }");
            o.f("        __syncthreads();");
            o.f("    }");
            o.f("}");
        }
        o.f("__device__ __forceinline__ void qsv_tile_consumed(const qsv_ctx &c) {");
        o.f("    asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");");
        o.f("    __syncthreads();");
        o.f("    if (c.flags & 4u) qsv_issue_load(c);");
        o.f("}");
        o.f("#endif");
    }
    std::vector<RoundInfo> info(nr);
    double scale = 1.0;
    t_sums = sums;
    for (int t = 0; t < 12; ++t) t_tile_phys[t] = t < L ? t : (int)P.hi_pos[t - L];
    for (uint32_t k = 0; k < nr; ++k) {
        const bool last = k + 1 == nr;
        const double scale_before = scale;
        // one mapping per round, shared by the hoisted scalars and both forms of the last round
        t_tb_override = bank_search ? tb_maps[k].data() : nullptr;
        if (gen_round(o, (int)k, rounds[k], ops.data(), info[k], last, scale, defer_scale, err, (last && direct) ? 1 : 0))
            return 1;
        if (last && direct) {
            // tile-local form of the last round for the single-tile host entry (tests)
            o.f("#ifdef QSV_JIT_HOST");
            double sc2 = scale_before;
            RoundInfo dummy;
            if (gen_round(o, (int)k, rounds[k], ops.data(), dummy, true, sc2, defer_scale, err, 0, "_local", false)) return 1;
            o.f("#endif");
        }
    }
    t_tb_override = nullptr;
    t_sums = nullptr;

    // is a tile ever skipped?  (every op has an external condition)
    bool always_active = false;
    for (const qsv_rop_t &p : ops) always_active = always_active || p.ext_mask == 0;
    std::string active = "false";
    if (!always_active) {
        std::vector<std::string> seen;
        for (const qsv_rop_t &p : ops) {
            std::string c = cond_ext(p.ext_mask, p.ext_val);
            bool dup = false;
            for (auto &s_ : seen) dup = dup || s_ == c;
            if (!dup) { seen.push_back(c); active += " || " + c; }
        }
    } else {
        active = "true";
    }
    if (sw) active = "true";       // every tile moves, whatever its ops' conditions say
    if (sums) active = "true";     // every tile of the support is weighed, whether an op acts on it or not
    // sp_mask / sp_val: index bits known to be FIXED in the state's support (a register that started in a basis state
    // and has not been touched non-diagonally on those positions): a tile that disagrees holds only zeros and, the
    // pass being linear, stays zero - it is skipped like a tile whose ops are all switched off
    // zf_mask != 0 (zero-fill form): a tile of the support holds uninitialised memory next to its amplitudes and must
    // be written in full even if no op acts on it
    o.f("QSV_FN bool qsv_active(const u64 gbase, const u64 sp_mask, const u64 sp_val, const u32 zf_mask) {");
    o.f("    return ((gbase & sp_mask) == sp_val) && (zf_mask != 0u || (%s));", active.c_str());
    o.f("}");

    // ---- tile base: insert a zero bit at every hi position (ascending)
    o.f("QSV_FN u64 qsv_tile_base(const u64 t) {");
    o.f("    u64 base = t << %d;", L);
    for (int j = 0; j < n_hi; ++j)
        o.f("    base = ((base >> %u) << %u) | (base & 0x%llxULL);", P.hi_pos[j], P.hi_pos[j] + 1u,
            (unsigned long long)((1ull << P.hi_pos[j]) - 1ull));
    o.f("    return base;");
    o.f("}");
    const bool prefetch = prefetch_env && always_active && !direct;      // skipped tiles must not be fetched

    // rounds of one tile.  mode 0: device, 1: host whole-state driver, 2: host single tile (tile-local stores)
    auto rounds_body = [&](Out &w, int mode) {
        for (uint32_t k = 0; k < nr; ++k) {
            const bool last = k + 1 == nr;
            const char *zf = k == 0 ? ", zf_mask, zf_val" : "";
            if (mode == 0) {
                if (k > 0) w.f("        __syncthreads();");
                if (last && direct)
                    w.f("        qsv_round%u(tile, state, base, tid, gbase, H%u, tables, ctx%s);", k, k, zf);
                else
                    w.f("        qsv_round%u(tile, tile, tid, gbase, H%u, tables%s);", k, k, zf);
                continue;
            }
            const char *src_buf = (k & 1) ? "B" : "A", *dst_buf = (k & 1) ? "A" : "B";
            w.f("        for (u32 tid = 0; tid < 256u; ++tid) {");
            w.f("            const qsv_H%u H = qsv_hoist%u(tid);", k, k);
            if (last && direct && mode == 1)
                w.f("            qsv_round%u(%s, state, base, tid, gbase, H, tables, ctx%s);", k, src_buf, zf);
            else if (last && direct)
                w.f("            qsv_round%u_local(%s, %s, tid, gbase, H, tables%s);", k, src_buf, dst_buf, zf);
            else
                w.f("            qsv_round%u(%s, %s, tid, gbase, H, tables%s);", k, src_buf, dst_buf, zf);
            w.f("        }");
        }
    };

    if (sw) {
        // work item u -> (pair slot s = low g bits, rest): the tile whose swapped field equals j = rank ^ s.  The data
        // it needs sits on rank j under field = rank; both ranks of a pair walk the same u sequence with the same CTA,
        // so the per-tile handshake always finds its partner.  Consecutive u go to different peers: the traffic through
        // the NVSwitch is a uniform all-to-all at every moment.
        std::vector<int> ins;
        for (int j = 0; j < n_hi; ++j) ins.push_back(P.hi_pos[j]);
        for (int k = 0; k < sw->g; ++k) ins.push_back(sw->q[k]);
        std::sort(ins.begin(), ins.end());
        o.f("QSV_FN void qsv_fused_tile(const u64 u, const u32 rank, u64 &base, u64 &src_base, u32 &peer) {");
        {
            std::string dep = "0u";
            for (int k = 0; k < sw->g; ++k)
                dep += " | ((u32)((u >> " + std::to_string(k) + ") & 1ULL) << " + std::to_string(sw->gb[k]) + ")";
            o.f("    peer = rank ^ (%s);", dep.c_str());
        }
        o.f("    u64 rest = (u >> %d) << %d;", sw->g, L);
        for (int pos : ins)
            o.f("    rest = ((rest >> %d) << %d) | (rest & 0x%llxULL);", pos, pos + 1, (unsigned long long)((1ull << pos) - 1ull));
        std::string fj = "0ULL", fr = "0ULL";
        for (int k = 0; k < sw->g; ++k) {
            fj += " | ((u64)((peer >> " + std::to_string(sw->gb[k]) + ") & 1u) << " + std::to_string(sw->q[k]) + ")";
            fr += " | ((u64)((rank >> " + std::to_string(sw->gb[k]) + ") & 1u) << " + std::to_string(sw->q[k]) + ")";
        }
        o.f("    base = rest | %s;", fj.c_str());
        o.f("    src_base = rest | %s;", fr.c_str());
        o.f("}");
        // a tile pair of which neither side intersects the support of the state (sp_mask / sp_val over the index bits
        // BEFORE the swap, tile bits excluded) holds only zeros on both ranks: it is neither moved nor computed.  The
        // test is symmetric in the two ranks of a pair, so both skip the same work items and their per-tile
        // handshake counters stay in step.
        o.f("QSV_FN bool qsv_pair_active(const u64 base, const u64 src_base, const u32 peer, const u32 rank, const u64 sp_mask,"
            " const u64 sp_val) {");
        o.f("    return (((base | ((u64)rank << QSV_NLOCAL)) & sp_mask) == sp_val) ||"
            " (((src_base | ((u64)peer << QSV_NLOCAL)) & sp_mask) == sp_val);");
        o.f("}");
    }

    // ---- device kernel
    o.f("#ifndef QSV_JIT_HOST");
    if (sw) o.f("struct qsv_peers { double2 *state[16]; u64 *flags[16]; };");
    const bool two_buf = sw && fuse_two_buffers();
    o.f("extern \"C\" __global__ void __launch_bounds__(256, %d)", two_buf ? 1 : 2);
    if (sw) {
        o.f("qsv_jit_pass(double2 *__restrict__ state, const u64 rank_bits, const u64 n_tiles, const double2 *__restrict__ tables,"
            " const __grid_constant__ qsv_peers peers, const u32 rank, const u64 epoch, const u64 sp_mask, const u64 sp_val,"
            " const u32 zf_mask_in, const u32 zf_val_in, const u32 zf_on) {");
    } else {
    o.f("qsv_jit_pass(double2 *__restrict__ state, const u64 rank_bits, const u64 n_tiles, const double2 *__restrict__ tables,"
        " const u64 sp_mask, const u64 sp_val, const u32 zf_mask, const u32 zf_val%s) {", sums ? ", u64 *__restrict__ bk" : "");
    }
    o.f("    extern __shared__ __align__(128) unsigned char smem_raw[];");
    o.f("    double2 *tile = reinterpret_cast<double2 *>(smem_raw);");
    o.f("    __shared__ __align__(8) u64 mbar;");
    o.f("    const u32 tid = threadIdx.x;");
    o.f("    if (tid == 0) { mbar_init(&mbar, 1); asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\"); }");
    o.f("    __syncthreads();");
    for (uint32_t k = 0; k < nr; ++k) o.f("    const qsv_H%u H%u = qsv_hoist%u(tid);", k, k, k);
    // run r of the tile: 2^L amplitudes at state[base + s_off[r]], s_off[r] = bits of r scattered to hi_pos.
    // The offsets are read from a shared-memory table like in k_pass_reg.  (Literal 64-bit offsets in an
    // unrolled issue sequence made ptxas 12.9 schedule a register copy AFTER its use across the ELECT
    // waterfall loops of UBLKCP - the first tile of a CTA then copied from a garbage address whenever an
    // offset needed a carry into the high word, i.e. tile bits at positions >= 28.)
    o.f("    __shared__ u64 s_off[%d];", n_runs);
    o.f("    if (tid < %du) s_off[tid] = qsv_scatter(tid << %d);", n_runs, L);
    o.f("    __syncthreads();");
    o.f("    const u32 warp = tid >> 5;");
    o.f("    u32 parity = 0;");
    if (sw) {
        // Fused swap + pass (pull model): every tile is TMA-loaded from the rank that holds it BEFORE the swap, run
        // through the rounds, and stored from registers to its place AFTER the swap in the local shard.  That place
        // holds the tile the partner pulls: it is overwritten only once the partner has flagged "read" (it flags right
        // after its TMA load of the tile completed).  All CTAs are co-resident (grid <= 2 per SM), so the partner CTA
        // always makes progress; nobody touches a peer's memory except to read it and to set flags.
        // Two tile buffers (one CTA per SM): the NVLink load of tile k+1 is in flight during ALL rounds of tile k.
        // With the single buffer of the local kernel the link idles while a tile is being computed (measured at 2
        // GPUs: 496 GB/s against the 677 GB/s of the plain swap kernel).
        o.f("    __shared__ __align__(8) u64 mbar2;");
        o.f("    if (tid == 0) { mbar_init(&mbar2, 1); asm volatile(\"fence.mbarrier_init.release.cluster;\" ::: \"memory\"); }");
        o.f("    __syncthreads();");
        o.f("    qsv_ctx ctx;");
        o.f("    ctx.tile = tile; ctx.mbar = &mbar; ctx.s_off = s_off; ctx.warp = warp;");
        o.f("    ctx.flags = (tid == 0 ? 2u : 0u)%s;", two_buf ? "" : " | 4u");
        o.f("    u64 u = blockIdx.x;");
        o.f("    u64 base = 0, src_base = 0; u32 peer = 0;");
        o.f("    for (;; u += gridDim.x) {");
        o.f("        if (u >= n_tiles) return;");
        o.f("        qsv_fused_tile(u, rank, base, src_base, peer);");
        o.f("        if (qsv_pair_active(base, src_base, peer, rank, sp_mask, sp_val)) break;");
        o.f("    }");
        o.f("    ctx.next_src = peers.state[peer] + src_base;");
        o.f("    qsv_issue_load(ctx);");
        o.f("    u64 it = 0;");
        o.f("    u32 buf = 0, parity2 = 0;");
        o.f("    for (;;) {");
        if (sw->pre)
            o.f("        const u64 gbase = src_base | ((u64)peer << %d);", sw->n_local);
        else
            o.f("        const u64 gbase = base | rank_bits;");
        // zero-fill form (zf_on: the register was NOT cleared before the run): a tile pulled from outside the support
        // of the state holds garbage and counts as zeros; inside the support only the amplitudes that agree with the
        // fixed bits of the tile (zf_mask_in / zf_val_in, tile-local) are data.  Bit 31 is set in no tile-local
        // address: (mask, val) = (2^31, 2^31) zeroes every amplitude of the tile in round 0.
        o.f("        const bool src_in = (((src_base | ((u64)peer << %d)) & sp_mask) == sp_val);", sw->n_local);
        o.f("        const u32 zf_mask = zf_on ? (src_in ? zf_mask_in : 0x80000000u) : 0u;");
        o.f("        const u32 zf_val = src_in ? zf_val_in : 0x80000000u;");
        o.f("        const u32 cur_peer = peer;");
        o.f("        const u64 cur_base = base;");
        o.f("        u64 un = u + gridDim.x;");
        o.f("        for (; un < n_tiles; un += gridDim.x) {");
        o.f("            qsv_fused_tile(un, rank, base, src_base, peer);");
        o.f("            if (qsv_pair_active(base, src_base, peer, rank, sp_mask, sp_val)) break;");
        o.f("        }");
        o.f("        double2 *tbuf = tile + (u64)buf * QSV_TILE_SLOTS;");
        o.f("        if (un < n_tiles) {");
        o.f("            ctx.next_src = peers.state[peer] + src_base;");
        if (two_buf) {
            o.f("            // the other buffer was emptied into registers by the previous tile's last round (barrier inside)");
            o.f("            ctx.tile = tile + (u64)(buf ^ 1u) * QSV_TILE_SLOTS;");
            o.f("            ctx.mbar = buf ? &mbar : &mbar2;");
            o.f("            qsv_issue_load(ctx);");
        } else {
            o.f("            ctx.flags |= 4u;        // issued from inside the last round, when the single buffer is free");
        }
        o.f("        } else {");
        o.f("            ctx.flags &= 3u;");
        o.f("        }");
        o.f("        if (buf) { mbar_wait(&mbar2, parity2); parity2 ^= 1u; } else { mbar_wait(&mbar, parity); parity ^= 1u; }");
        o.f("        ++it;");
        o.f("        if (cur_peer != rank) {");
        o.f("            if (tid == 0) qsv_signal_peer(peers.flags[cur_peer] + (u64)rank * %uu + blockIdx.x, epoch + it);", kFlagCtas);
        o.f("            ctx.wait_flag = peers.flags[rank] + (u64)cur_peer * %uu + blockIdx.x;", kFlagCtas);
        o.f("            ctx.wait_val = epoch + it;");
        o.f("        } else {");
        o.f("            ctx.wait_flag = nullptr; ctx.wait_val = 0;");
        o.f("        }");
        {
            Out w;
            rounds_body(w, 0);
            // the rounds refer to the tile's base as `base`: give them the current tile's
            std::string body = w.s;
            size_t pos_ = 0;
            while ((pos_ = body.find("state, base, tid", pos_)) != std::string::npos) body.replace(pos_, 16, "state, cur_base, tid");
            pos_ = 0;
            while ((pos_ = body.find("(tile, tile, ", pos_)) != std::string::npos) body.replace(pos_, 13, "(tbuf, tbuf, ");
            pos_ = 0;
            while ((pos_ = body.find("(tile, state, ", pos_)) != std::string::npos) body.replace(pos_, 14, "(tbuf, state, ");
            o.s += body;
        }
        o.f("        if (un >= n_tiles) break;");
        o.f("        u = un;");
        if (two_buf) o.f("        buf ^= 1u;");
        o.f("    }");
    } else if (direct) {
        // Persistent loop with the load of tile k+1 issued from inside tile k's last round.
        o.f("    qsv_ctx ctx;");
        o.f("    ctx.state = state; ctx.tile = tile; ctx.mbar = &mbar; ctx.s_off = s_off; ctx.warp = warp;");
        o.f("    ctx.flags = (tid == 0 ? 2u : 0u) | 4u;");
        if (sums) {
            o.f("    __shared__ u32 s_tab[2 * 8 * 8];");
            o.f("    if (tid < 128u) s_tab[tid] = 0u;");
            o.f("    __syncthreads();");
            o.f("    ctx.bk = bk; ctx.s_tab = s_tab; ctx.par = 0u;");
        } else {
            o.f("    ctx.bk = nullptr; ctx.s_tab = nullptr; ctx.par = 0u;");
        }
        o.f("    u64 t = blockIdx.x;");
        o.f("    while (t < n_tiles && !qsv_active(qsv_tile_base(t) | rank_bits, sp_mask, sp_val, zf_mask)) t += gridDim.x;");
        o.f("    if (t >= n_tiles) return;");
        o.f("    ctx.next_base = qsv_tile_base(t);");
        o.f("    qsv_issue_load(ctx);");
        o.f("    for (;;) {");
        o.f("        const u64 base = qsv_tile_base(t);");
        o.f("        const u64 gbase = base | rank_bits;");
        o.f("        u64 tn = t + gridDim.x;");
        o.f("        while (tn < n_tiles && !qsv_active(qsv_tile_base(tn) | rank_bits, sp_mask, sp_val, zf_mask)) tn += gridDim.x;");
        o.f("        ctx.next_base = tn < n_tiles ? qsv_tile_base(tn) : 0ULL;");
        o.f("        ctx.flags = (ctx.flags & 3u) | (tn < n_tiles ? 4u : 0u);");
        o.f("        mbar_wait(&mbar, parity);");
        o.f("        parity ^= 1u;");
        o.f("        ctx.par ^= 1u;");
        rounds_body(o, 0);
        o.f("        if (tn >= n_tiles) break;");
        o.f("        t = tn;");
        o.f("    }");
    } else {
        // one elected lane per warp issues the bulk copies.  elect.sync (instead of lane == 0) tells ptxas that
        // exactly one lane is active: addresses go through R2UR + uniform arithmetic straight into UBLKCP, without
        // the ELECT/PLOP3/BRA.U.ANY waterfall loop it otherwise wraps around every copy (16 -> 8 instructions each)
        o.f("    u32 elected;");
        o.f("    asm volatile(\"{\\n.reg .pred P;\\nelect.sync _|P, 0xffffffff;\\nselp.u32 %%0, 1, 0, P;\\n}\\n\" : \"=r\"(elected));");
        o.f("    const bool issuer = elected != 0u;");
        o.f("    for (u64 t = blockIdx.x; t < n_tiles; t += gridDim.x) {");
        o.f("        const u64 base = qsv_tile_base(t);");
        o.f("        const u64 gbase = base | rank_bits;");
        o.f("        if (!qsv_active(gbase, sp_mask, sp_val, zf_mask)) continue;");
        o.f("        if (tid == 0) mbar_expect_tx(&mbar, %uu);", 16u << 12);
        o.f("        if (issuer)");
        o.f("            for (u32 r = warp; r < %du; r += 8u)", n_runs);
        o.f("                bulk_g2s(tile + (u64)QP(r << %d), state + base + s_off[r], %uu, &mbar);", L, 16u << L);
        if (prefetch) {
            // the CTA's NEXT tile is pulled into L2 while this one is computed
            o.f("        if (issuer && t + gridDim.x < n_tiles) {");
            o.f("            const u64 nbase = qsv_tile_base(t + gridDim.x);");
            o.f("            for (u32 r = warp; r < %du; r += 8u) bulk_prefetch_l2(state + nbase + s_off[r], %uu);", n_runs,
                16u << L);
            o.f("        }");
        }
        o.f("        mbar_wait(&mbar, parity);");
        o.f("        parity ^= 1u;");
        rounds_body(o, 0);
        o.f("        asm volatile(\"fence.proxy.async.shared::cta;\" ::: \"memory\");");
        o.f("        __syncthreads();");
        o.f("        if (issuer)");
        o.f("            for (u32 r = warp; r < %du; r += 8u)", n_runs);
        o.f("                bulk_s2g(state + base + s_off[r], tile + (u64)QP(r << %d), %uu);", L, 16u << L);
        o.f("        asm volatile(\"cp.async.bulk.commit_group;\" ::: \"memory\");");
        o.f("        asm volatile(\"cp.async.bulk.wait_group.read 0;\" ::: \"memory\");");
        o.f("    }");
    }
    o.f("}");
    o.f("#else");
    // ---- host emulation driver (tests only): same round functions, threads run sequentially
    o.f("extern \"C\" void qsv_jit_host_run(double2 *state, const u64 rank_bits, const u64 n_tiles, const double2 *tables,"
        " const u64 sp_mask, const u64 sp_val, const u32 zf_mask, const u32 zf_val) {");
    o.f("    static double2 A[QSV_TILE_SLOTS], B[QSV_TILE_SLOTS];");
    o.f("    qsv_ctx ctx; (void)ctx;");
    o.f("    for (u64 t = 0; t < n_tiles; ++t) {");
    o.f("        const u64 base = qsv_tile_base(t);");
    o.f("        const u64 gbase = base | rank_bits;");
    o.f("        if (!qsv_active(gbase, sp_mask, sp_val, zf_mask)) continue;");
    o.f("        for (u32 r = 0; r < %du; ++r)", n_runs);
    o.f("            memcpy(A + (u64)QP(r << %d), state + base + qsv_scatter(r << %d), %uu);", L, L, 16u << L);
    rounds_body(o, 1);
    if (!direct) {
        o.f("        const double2 *fin = %s;", (nr & 1) ? "B" : "A");
        o.f("        for (u32 r = 0; r < %du; ++r)", n_runs);
        o.f("            memcpy(state + base + qsv_scatter(r << %d), fin + (u64)QP(r << %d), %uu);", L, L, 16u << L);
    }
    o.f("    }");
    o.f("}");
    if (sw) {
        // fused variant on the host: `state` is this rank's shard AFTER the swap (output), src_states[j] the shards
        // BEFORE the swap (inputs, never written)
        o.f("extern \"C\" void qsv_jit_host_run_swap(double2 *state, const double2 *const *src_states, const u32 rank,"
            " const u64 rank_bits, const u64 n_tiles, const double2 *tables, const u64 sp_mask, const u64 sp_val,"
            " const u32 zf_mask_in, const u32 zf_val_in, const u32 zf_on) {");
        o.f("    static double2 A[QSV_TILE_SLOTS], B[QSV_TILE_SLOTS];");
        o.f("    qsv_ctx ctx; (void)ctx;");
        o.f("    for (u64 u = 0; u < n_tiles; ++u) {");
        o.f("        u64 base, src_base; u32 peer;");
        o.f("        qsv_fused_tile(u, rank, base, src_base, peer);");
        o.f("        if (!qsv_pair_active(base, src_base, peer, rank, sp_mask, sp_val)) continue;");
        o.f("        const bool src_in = (((src_base | ((u64)peer << %d)) & sp_mask) == sp_val);", sw->n_local);
        o.f("        const u32 zf_mask = zf_on ? (src_in ? zf_mask_in : 0x80000000u) : 0u;");
        o.f("        const u32 zf_val = src_in ? zf_val_in : 0x80000000u;");
        if (sw->pre)
            o.f("        const u64 gbase = src_base | ((u64)peer << %d); (void)rank_bits;", sw->n_local);
        else
            o.f("        const u64 gbase = base | rank_bits;");
        o.f("        for (u32 r = 0; r < %du; ++r)", n_runs);
        o.f("            memcpy(A + (u64)QP(r << %d), src_states[peer] + src_base + qsv_scatter(r << %d), %uu);", L, L, 16u << L);
        rounds_body(o, 1);
        o.f("    }");
        o.f("}");
    }
    // one tile with a caller-chosen global base index (programs of states too large for a host array)
    o.f("extern \"C\" void qsv_jit_host_tile(double2 *tile, const u64 gbase, const double2 *tables) {");
    o.f("    static double2 A[QSV_TILE_SLOTS], B[QSV_TILE_SLOTS];");
    o.f("    const u32 zf_mask = 0u, zf_val = 0u;");
    o.f("    for (u32 r = 0; r < %du; ++r) memcpy(A + (u64)QP(r << %d), tile + ((u64)r << %d), %uu);", n_runs, L, L,
        16u << L);
    o.f("    {");
    rounds_body(o, 2);
    o.f("    }");
    o.f("    for (u32 r = 0; r < %du; ++r) memcpy(tile + ((u64)r << %d), %s + (u64)QP(r << %d), %uu);", n_runs, L,
        (nr & 1) ? "B" : "A", L, 16u << L);
    o.f("}");
    o.f("#endif");
    src.swap(o.s);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// NVRTC (dlopen'ed: the library must load in containers without the toolkit on the loader path)
// ------------------------------------------------------------------------------------------------

struct Nvrtc {
    void *h = nullptr;
    int (*CreateProgram)(void **, const char *, const char *, int, const char *const *, const char *const *) = nullptr;
    int (*CompileProgram)(void *, int, const char *const *) = nullptr;
    int (*GetCUBINSize)(void *, size_t *) = nullptr;
    int (*GetCUBIN)(void *, char *) = nullptr;
    int (*GetProgramLogSize)(void *, size_t *) = nullptr;
    int (*GetProgramLog)(void *, char *) = nullptr;
    int (*DestroyProgram)(void **) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};

static Nvrtc *nvrtc(std::string &err) {
    static Nvrtc lib;
    static std::once_flag once;
    static std::string load_err;
    std::call_once(once, [] {
        const char *names[] = {getenv("QSV_NVRTC_LIB"), "libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12",
                               "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so"};
        for (const char *n : names) {
            if (!n || !*n) continue;
            lib.h = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (lib.h) break;
        }
        if (!lib.h) {
            load_err = "cannot load libnvrtc (set QSV_NVRTC_LIB or QSV_JIT=0)";
            return;
        }
#define QSV_SYM(field, name)                                                        \
    *reinterpret_cast<void **>(&lib.field) = dlsym(lib.h, name);                    \
    if (!lib.field) load_err = std::string("libnvrtc lacks ") + name;
        QSV_SYM(CreateProgram, "nvrtcCreateProgram")
        QSV_SYM(CompileProgram, "nvrtcCompileProgram")
        QSV_SYM(GetCUBINSize, "nvrtcGetCUBINSize")
        QSV_SYM(GetCUBIN, "nvrtcGetCUBIN")
        QSV_SYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
        QSV_SYM(GetProgramLog, "nvrtcGetProgramLog")
        QSV_SYM(DestroyProgram, "nvrtcDestroyProgram")
        QSV_SYM(GetErrorString, "nvrtcGetErrorString")
#undef QSV_SYM
    });
    if (!load_err.empty()) { err = load_err; return nullptr; }
    return &lib;
}

static int compile_cubin(const std::string &src, std::string &cubin, std::string &err) {
    Nvrtc *rt = nvrtc(err);
    if (!rt) return 1;
    void *prog = nullptr;
    int rc = rt->CreateProgram(&prog, src.c_str(), "qsv_jit_pass.cu", 0, nullptr, nullptr);
    if (rc != 0) { err = std::string("nvrtcCreateProgram: ") + rt->GetErrorString(rc); return 1; }
    const char *opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "-lineinfo"};
    rc = rt->CompileProgram(prog, 3, opts);
    if (rc != 0) {
        size_t n = 0;
        rt->GetProgramLogSize(prog, &n);
        std::string log(n, '\0');
        if (n) rt->GetProgramLog(prog, &log[0]);
        err = std::string("nvrtcCompileProgram: ") + rt->GetErrorString(rc) + "\n" + log.substr(0, 1500);
        rt->DestroyProgram(&prog);
        return 1;
    }
    size_t n = 0;
    rt->GetCUBINSize(prog, &n);
    cubin.assign(n, '\0');
    rt->GetCUBIN(prog, &cubin[0]);
    rt->DestroyProgram(&prog);
    return 0;
}

// ------------------------------------------------------------------------------------------------
// driver API (dlopen'ed: libcuda.so.1 does not exist in the build container)
// ------------------------------------------------------------------------------------------------

struct Driver {
    void *h = nullptr;
    int (*ModuleLoadData)(void **, const void *) = nullptr;
    int (*ModuleGetFunction)(void **, void *, const char *) = nullptr;
    int (*FuncSetAttribute)(void *, int, int) = nullptr;
    int (*LaunchKernel)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *, void **,
                        void **) = nullptr;
    int (*GetErrorString)(int, const char **) = nullptr;
    int (*OccupancyMaxActiveBlocks)(int *, void *, int, size_t) = nullptr;
    int (*LaunchCooperativeKernel)(void *, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void *,
                                   void **) = nullptr;
};

static Driver *driver(std::string &err) {
    static Driver lib;
    static std::once_flag once;
    static std::string load_err;
    std::call_once(once, [] {
        lib.h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!lib.h) { load_err = "cannot load libcuda.so.1"; return; }
#define QSV_SYM(field, name)                                                        \
    *reinterpret_cast<void **>(&lib.field) = dlsym(lib.h, name);                    \
    if (!lib.field) load_err = std::string("libcuda lacks ") + name;
        QSV_SYM(ModuleLoadData, "cuModuleLoadData")
        QSV_SYM(ModuleGetFunction, "cuModuleGetFunction")
        QSV_SYM(FuncSetAttribute, "cuFuncSetAttribute")
        QSV_SYM(LaunchKernel, "cuLaunchKernel")
        QSV_SYM(GetErrorString, "cuGetErrorString")
        QSV_SYM(OccupancyMaxActiveBlocks, "cuOccupancyMaxActiveBlocksPerMultiprocessor")
        QSV_SYM(LaunchCooperativeKernel, "cuLaunchCooperativeKernel")
#undef QSV_SYM
    });
    if (!load_err.empty()) { err = load_err; return nullptr; }
    return &lib;
}

// ------------------------------------------------------------------------------------------------
// cache: key = tile geometry + program bytes (n_local and rank_bits are kernel arguments)
// ------------------------------------------------------------------------------------------------

struct Entry {
    std::mutex mu;
    bool compiled = false, failed = false;
    std::string cubin, err;
    std::unordered_map<int, void *> fn;     // device ordinal -> CUfunction
};

static std::mutex g_mu;
static std::unordered_map<std::string, Entry *> g_cache;
static std::atomic<uint64_t> g_compiles{0}, g_hits{0}, g_compile_us{0};

static std::string cache_key(const qsv_pass_t &P, const void *prog, uint32_t bytes, const SwapSpec *sw = nullptr,
                             const SumSpec *sums = nullptr) {
    std::string k;
    if (sums) {
        k.append("SUM");
        k.append(reinterpret_cast<const char *>(sums), sizeof(*sums));
    }
    if (sw) {
        k.append(fuse_two_buffers() ? "F2" : "F1");
        k.append(reinterpret_cast<const char *>(sw), sizeof(*sw));
    }
    k.append(reinterpret_cast<const char *>(&P.tile_bits), sizeof(P.tile_bits));
    k.append(reinterpret_cast<const char *>(&P.low_bits), sizeof(P.low_bits));
    k.append(reinterpret_cast<const char *>(P.hi_pos), sizeof(P.hi_pos));
    k.append(getenv("QSV_JIT_NO_DEFER") ? "N" : "D");
    { const char *ds = getenv("QSV_JIT_DIRECT_STORE"); k.append(ds && !strcmp(ds, "0") ? "s" : "S"); }
    { const char *pf = getenv("QSV_JIT_PREFETCH"); k.append(pf && !strcmp(pf, "0") ? "p" : "P"); }
    { const char *pe = getenv("QSV_JIT_PAD"); k.append(pe ? pe : "-"); k.append("|"); }
    k.append(bank_search_enabled() ? "S" : "s");
    k.append(static_cast<const char *>(prog), bytes);
    return k;
}

static Entry *lookup(const qsv_pass_t &P, const void *prog, uint32_t bytes, const SwapSpec *sw = nullptr,
                     const SumSpec *sums = nullptr) {
    const std::string k = cache_key(P, prog, bytes, sw, sums);
    std::lock_guard<std::mutex> g(g_mu);
    auto it = g_cache.find(k);
    if (it != g_cache.end()) return it->second;
    Entry *e = new Entry();
    g_cache.emplace(k, e);
    return e;
}

// ---- on-disk cubin cache: a fresh process (Shor builds a new circuit per base; benchmarks and servers restart) finds
// the kernels of circuits it has seen before instead of paying NVRTC again (~0.3-0.4 s of CPU per kernel).  Keyed by
// the GENERATED SOURCE (so a change of the generator can never serve a stale kernel), the NVRTC options and the
// compiler version.  QSV_CACHE_DIR (default ~/.cache/qsv), QSV_DISK_CACHE=0 disables.  Atomic writes (temp + rename).
static bool disk_cache_dir(std::string &dir) {
    const char *off = getenv("QSV_DISK_CACHE");
    if (off && !strcmp(off, "0")) return false;
    const char *d = getenv("QSV_CACHE_DIR");
    if (d && *d) { dir = d; return true; }
    const char *home = getenv("HOME");
    if (!home || !*home) return false;
    dir = std::string(home) + "/.cache/qsv";
    return true;
}

static void hash128(const std::string &s_, uint64_t &h1, uint64_t &h2) {
    h1 = 0xcbf29ce484222325ull; h2 = 0x9e3779b97f4a7c15ull;
    for (unsigned char c : s_) {
        h1 = (h1 ^ c) * 0x100000001b3ull;                         // FNV-1a
        h2 = (h2 + c) * 0xff51afd7ed558ccdull; h2 ^= h2 >> 29;    // an independent multiply-xorshift chain
    }
    h2 ^= (uint64_t)s_.size() * 0xc4ceb9fe1a85ec53ull;
}

static std::string disk_cache_path(const std::string &src) {
    std::string dir;
    if (!disk_cache_dir(dir)) return std::string();
    static std::string ver;
    static std::once_flag once;
    std::call_once(once, [] {
        std::string err;
        int major = 0, minor = 0;
        Nvrtc *rt = nvrtc(err);
        if (rt) {
            auto fn = reinterpret_cast<int (*)(int *, int *)>(dlsym(rt->h, "nvrtcVersion"));
            if (fn) fn(&major, &minor);
        }
        ver = "nvrtc" + std::to_string(major) + "." + std::to_string(minor) + "|sm_100a|c++17|lineinfo|";
    });
    uint64_t h1, h2;
    hash128(ver + src, h1, h2);
    char name[80];
    snprintf(name, sizeof(name), "/jit-%016llx%016llx.cubin", (unsigned long long)h1, (unsigned long long)h2);
    return dir + name;
}

static bool disk_cache_load(const std::string &path, std::string &cubin) {
    if (path.empty()) return false;
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) return false;
    std::string data;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, n);
    fclose(f);
    if (data.size() < 64 || memcmp(data.data(), "\x7f" "ELF", 4) != 0) return false;     // truncated / foreign file
    cubin.swap(data);
    return true;
}

static void disk_cache_store(const std::string &path, const std::string &cubin) {
    if (path.empty()) return;
    const std::string dir = path.substr(0, path.rfind('/'));
    for (size_t i = 1; i <= dir.size(); ++i)          // mkdir -p
        if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0755);
    char tmp[32];
    snprintf(tmp, sizeof(tmp), ".tmp-%d-%p", (int)getpid(), (const void *)&cubin);
    const std::string tpath = dir + "/" + tmp;
    FILE *f = fopen(tpath.c_str(), "wb");
    if (!f) return;
    const bool ok = fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
    fclose(f);
    if (!ok || rename(tpath.c_str(), path.c_str()) != 0) remove(tpath.c_str());
}

static std::atomic<uint64_t> g_disk_hits{0};

static int ensure_compiled(Entry *e, const qsv_pass_t &P, const void *prog, uint32_t bytes, const SwapSpec *sw = nullptr,
                           const SumSpec *sums = nullptr) {
    std::lock_guard<std::mutex> g(e->mu);
    if (e->compiled) { g_hits.fetch_add(1); return 0; }
    if (e->failed) { set_error("qsv jit: %s", e->err.c_str()); return 1; }
    const auto t0 = std::chrono::steady_clock::now();
    std::string src;
    if (generate(P, prog, bytes, src, e->err, sw, sums)) {
        e->failed = true;
        set_error("qsv jit: %s", e->err.c_str());
        return 1;
    }
    const std::string cpath = disk_cache_path(src);
    if (disk_cache_load(cpath, e->cubin)) {
        e->compiled = true;
        g_disk_hits.fetch_add(1);
        return 0;
    }
    if (compile_cubin(src, e->cubin, e->err)) {
        e->failed = true;
        set_error("qsv jit: %s", e->err.c_str());
        return 1;
    }
    disk_cache_store(cpath, e->cubin);
    e->compiled = true;
    g_compiles.fetch_add(1);
    g_compile_us.fetch_add((uint64_t)std::chrono::duration_cast<std::chrono::microseconds>(
                               std::chrono::steady_clock::now() - t0).count());
    return 0;
}

static int validate(const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes) {
    QSV_CHECK_ARG(P && prog_host, "qsv jit: NULL pointer");
    QSV_CHECK_ARG(prog_bytes >= sizeof(qsv_rprog_hdr_t), "qsv jit: program too short");
    const qsv_rprog_hdr_t *hdr = static_cast<const qsv_rprog_hdr_t *>(prog_host);
    QSV_CHECK_ARG(hdr->n_rounds >= 1 && hdr->n_rounds <= QSV_JIT_MAX_ROUNDS && hdr->n_ops <= QSV_JIT_MAX_OPS,
                  "qsv jit: %u rounds / %u ops exceed the limits (%d / %d)", hdr->n_rounds, hdr->n_ops,
                  QSV_JIT_MAX_ROUNDS, QSV_JIT_MAX_OPS);
    const size_t need = sizeof(qsv_rprog_hdr_t) + sizeof(qsv_round_t) * hdr->n_rounds + sizeof(qsv_rop_t) * hdr->n_ops;
    QSV_CHECK_ARG(prog_bytes >= need, "qsv jit: prog_bytes=%u smaller than the %zu bytes the header implies", prog_bytes,
                  need);
    QSV_CHECK_ARG(P->tile_bits == 12 && P->n_hi >= 0 && P->n_hi <= 8 && P->low_bits == 12 - P->n_hi,
                  "qsv jit: needs a 12-bit tile with at most 8 high positions");
    QSV_CHECK_ARG(hdr->reserved == 0 || hdr->reserved == 4, "qsv jit: needs 4 register bits per round");
    const qsv_round_t *rounds = reinterpret_cast<const qsv_round_t *>(static_cast<const uint8_t *>(prog_host) + 16);
    for (uint32_t r = 0; r < hdr->n_rounds; ++r) {
        const qsv_round_t &R = rounds[r];
        QSV_CHECK_ARG(R.rb[0] < R.rb[1] && R.rb[1] < R.rb[2] && R.rb[2] < R.rb[3] && R.rb[3] < 12,
                      "qsv jit: round %u register bits not ascending inside the tile", r);
        QSV_CHECK_ARG((uint64_t)R.first_op + (R.n_pre & 0x7fffu) + (R.n_ops & 0xffffu) + ((R.n_ops >> 16) & 0x3fffu) +
                              (R.n_post & 0x7fffu) <= hdr->n_ops, "qsv jit: round %u op range", r);
    }
    return 0;
}

// policy: QSV_JIT=0 never, =force always (tile_bits 12), default: states of >= 2^QSV_JIT_MIN_QUBITS amplitudes
bool wanted(const qsv_pass_t *P, const void *prog_host) {
    const char *m = getenv("QSV_JIT");
    if (m && !strcmp(m, "0")) return false;
    const qsv_rprog_hdr_t *hdr = static_cast<const qsv_rprog_hdr_t *>(prog_host);
    if (P->tile_bits != 12 || !(hdr->reserved == 0 || hdr->reserved == 4)) return false;
    if (m && !strcmp(m, "force")) return true;
    const char *q = getenv("QSV_JIT_MIN_QUBITS");
    const int minq = q ? atoi(q) : 28;
    return P->n_local >= minq;
}

static int get_function(const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes, const SwapSpec *sw, void **fn_out,
                        const SumSpec *sums = nullptr);

static int launch_any(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes,
                      const void *tables_dev, cudaStream_t st, unsigned grid, unsigned long long sp_mask,
                      unsigned long long sp_val, unsigned zf_mask, unsigned zf_val, const SumSpec *sums, void *bk_dev);

int launch(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes, const void *tables_dev,
           cudaStream_t st, unsigned grid, unsigned long long sp_mask, unsigned long long sp_val, unsigned zf_mask,
           unsigned zf_val) {
    return launch_any(state_dev, P, prog_host, prog_bytes, tables_dev, st, grid, sp_mask, sp_val, zf_mask, zf_val, nullptr,
                      nullptr);
}

int launch_sums(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes, const void *tables_dev,
                cudaStream_t st, unsigned grid, unsigned long long sp_mask, unsigned long long sp_val, unsigned zf_mask,
                unsigned zf_val, const SumSpec &sums, void *bk_dev) {
    return launch_any(state_dev, P, prog_host, prog_bytes, tables_dev, st, grid, sp_mask, sp_val, zf_mask, zf_val, &sums,
                      bk_dev);
}

static int launch_any(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes,
                      const void *tables_dev, cudaStream_t st, unsigned grid, unsigned long long sp_mask,
                      unsigned long long sp_val, unsigned zf_mask, unsigned zf_val, const SumSpec *sums, void *bk_dev) {
    void *fn = nullptr;
    if (int rc = get_function(P, prog_host, prog_bytes, nullptr, &fn, sums)) return rc;
    std::string err;
    Driver *d = driver(err);
    QSV_CHECK_ARG(d, "qsv jit: %s", err.c_str());
    unsigned long long rank_bits = P->rank_bits, n_tiles = 1ull << (P->n_local - P->tile_bits);
    void *args[] = {&state_dev, &rank_bits, &n_tiles, const_cast<void **>(&tables_dev), &sp_mask, &sp_val, &zf_mask,
                    &zf_val, &bk_dev};      // the sums variant takes the bucket array as its ninth parameter
    const int rc = d->LaunchKernel(fn, grid, 1, 1, 256, 1, 1, 16u * tile_slots(P->n_hi), (void *)st, args, nullptr);
    if (rc) {
        const char *es = nullptr;
        d->GetErrorString(rc, &es);
        set_error("qsv jit: cuLaunchKernel: %s", es ? es : "?");
        return 2;
    }
    count_launch();
    return 0;
}

struct PeerTable { void *state[16]; void *flags[16]; };

static int sm_count() {
    static std::mutex mu;
    static std::unordered_map<int, int> per_device;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    std::lock_guard<std::mutex> g(mu);
    auto it = per_device.find(dev);
    if (it != per_device.end()) return it->second;
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    per_device[dev] = n;
    return n;
}

// fused swap + pass: all CTAs must be co-resident (they hand-shake with the partner rank's CTA of the same index)
// -> a COOPERATIVE launch: the driver starts the grid only when every CTA fits on the device at once (and refuses a
// grid that never could), whatever else is resident; the flag spin inside is bounded by a watchdog.
int launch_swap(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes, const void *tables_dev,
                const SwapSpec &sw, const PeerTable &peers, unsigned rank, unsigned long long epoch,
                unsigned long long sp_mask, unsigned long long sp_val, unsigned zf_mask, unsigned zf_val, unsigned zf_on,
                cudaStream_t st) {
    void *fn = nullptr;
    if (int rc = get_function(P, prog_host, prog_bytes, &sw, &fn)) return rc;
    std::string err;
    Driver *d = driver(err);
    QSV_CHECK_ARG(d, "qsv jit: %s", err.c_str());
    const unsigned smem = (fuse_two_buffers() ? 2u : 1u) * 16u * tile_slots(P->n_hi);
    {
        std::string e2;
        Driver *d2 = driver(e2);
        if (d2) d2->FuncSetAttribute(fn, /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/ 8, (int)smem);
    }
    int per_sm = 0;
    int rc = d->OccupancyMaxActiveBlocks(&per_sm, fn, 256, smem);
    QSV_CHECK_ARG(rc == 0 && per_sm >= 1, "qsv jit: occupancy query failed for the fused swap kernel (rc=%d, %d CTAs/SM)", rc, per_sm);
    if (per_sm > 2) per_sm = 2;
    unsigned long long rank_bits = P->rank_bits, n_tiles = 1ull << (P->n_local - P->tile_bits);
    unsigned long long grid = (unsigned long long)sm_count() * (unsigned)per_sm;
    if (grid > kFlagCtas) grid = kFlagCtas;
    if (grid > n_tiles) grid = n_tiles;
    PeerTable pt = peers;
    void *args[] = {&state_dev, &rank_bits, &n_tiles, const_cast<void **>(&tables_dev), &pt, &rank, &epoch, &sp_mask,
                    &sp_val, &zf_mask, &zf_val, &zf_on};
    const char *coop = getenv("QSV_FUSE_COOPERATIVE");
    if (coop && !strcmp(coop, "0"))
        rc = d->LaunchKernel(fn, (unsigned)grid, 1, 1, 256, 1, 1, smem, (void *)st, args, nullptr);
    else
        rc = d->LaunchCooperativeKernel(fn, (unsigned)grid, 1, 1, 256, 1, 1, smem, (void *)st, args);
    if (rc) {
        const char *es = nullptr;
        d->GetErrorString(rc, &es);
        set_error("qsv jit: cuLaunchCooperativeKernel (fused swap, grid %llu): %s", grid, es ? es : "?");
        return 2;
    }
    count_launch();
    return 0;
}

static int get_function(const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes, const SwapSpec *sw, void **fn_out,
                        const SumSpec *sums) {
    if (validate(P, prog_host, prog_bytes)) return 1;
    Entry *e = lookup(*P, prog_host, prog_bytes, sw, sums);
    if (ensure_compiled(e, *P, prog_host, prog_bytes, sw, sums)) return 1;
    int dev = 0;
    QSV_CUDA(cudaGetDevice(&dev));
    void *fn = nullptr;
    {
        std::lock_guard<std::mutex> g(e->mu);
        auto it = e->fn.find(dev);
        if (it != e->fn.end()) fn = it->second;
        if (!fn) {
            std::string err;
            Driver *d = driver(err);
            QSV_CHECK_ARG(d, "qsv jit: %s", err.c_str());
            QSV_CUDA(cudaFree(0));        // the primary context is current before the first driver call
            void *mod = nullptr;
            int rc = d->ModuleLoadData(&mod, e->cubin.data());
            const char *es = nullptr;
            if (rc) { d->GetErrorString(rc, &es); set_error("qsv jit: cuModuleLoadData: %s", es ? es : "?"); return 2; }
            rc = d->ModuleGetFunction(&fn, mod, "qsv_jit_pass");
            if (rc) { d->GetErrorString(rc, &es); set_error("qsv jit: cuModuleGetFunction: %s", es ? es : "?"); return 2; }
            rc = d->FuncSetAttribute(fn, /*CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES*/ 8,
                                     (int)(16u * tile_slots(P->n_hi)));
            if (rc) { d->GetErrorString(rc, &es); set_error("qsv jit: cuFuncSetAttribute: %s", es ? es : "?"); return 2; }
            e->fn[dev] = fn;
        }
    }
    *fn_out = fn;
    return 0;
}

}  // namespace jit
}  // namespace qsv

using namespace qsv;

static int copy_out(const std::string &s, char *buf, size_t buf_len, size_t *needed) {
    if (needed) *needed = s.size();
    if (buf && buf_len) {
        const size_t n = s.size() < buf_len ? s.size() : buf_len;
        memcpy(buf, s.data(), n);
    }
    return 0;
}

extern "C" int qsv_jit_source(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, char *buf,
                              size_t buf_len, size_t *needed) {
    if (jit::validate(pass, prog_host, prog_bytes)) return 1;
    std::string src, err;
    if (jit::generate(*pass, prog_host, prog_bytes, src, err)) { set_error("qsv jit: %s", err.c_str()); return 1; }
    return copy_out(src, buf, buf_len, needed);
}

extern "C" int qsv_jit_compile(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, char *cubin_buf,
                               size_t buf_len, size_t *needed) {
    if (jit::validate(pass, prog_host, prog_bytes)) return 1;
    jit::Entry *e = jit::lookup(*pass, prog_host, prog_bytes);
    if (jit::ensure_compiled(e, *pass, prog_host, prog_bytes)) return 1;
    return copy_out(e->cubin, cubin_buf, buf_len, needed);
}

static int make_sum_spec(jit::SumSpec &sm, int n_bits, const int *positions) {
    QSV_CHECK_ARG(n_bits >= 1 && n_bits <= 12 && positions, "qsv sums: n_bits=%d out of range or NULL positions", n_bits);
    memset(&sm, 0, sizeof(sm));
    sm.n_bits = n_bits;
    for (int k = 0; k < n_bits; ++k) {
        QSV_CHECK_ARG(positions[k] >= 0 && positions[k] < 48, "qsv sums: position %d out of range", positions[k]);
        for (int j = 0; j < k; ++j) QSV_CHECK_ARG(positions[j] != positions[k], "qsv sums: duplicate position %d", positions[k]);
        sm.pos[k] = positions[k];
    }
    return 0;
}

extern "C" int qsv_jit_source_sums(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int n_bits,
                                   const int *positions, char *buf, size_t buf_len, size_t *needed) {
    if (jit::validate(pass, prog_host, prog_bytes)) return 1;
    jit::SumSpec sm;
    if (make_sum_spec(sm, n_bits, positions)) return 1;
    std::string src, err;
    if (jit::generate(*pass, prog_host, prog_bytes, src, err, nullptr, &sm)) { set_error("qsv jit: %s", err.c_str()); return 1; }
    return copy_out(src, buf, buf_len, needed);
}

extern "C" int qsv_jit_compile_sums(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int n_bits,
                                    const int *positions, char *cubin_buf, size_t buf_len, size_t *needed) {
    if (jit::validate(pass, prog_host, prog_bytes)) return 1;
    jit::SumSpec sm;
    if (make_sum_spec(sm, n_bits, positions)) return 1;
    jit::Entry *e = jit::lookup(*pass, prog_host, prog_bytes, nullptr, &sm);
    if (jit::ensure_compiled(e, *pass, prog_host, prog_bytes, nullptr, &sm)) return 1;
    return copy_out(e->cubin, cubin_buf, buf_len, needed);
}

namespace qsv { namespace jit {
int run_sums(void *state_dev, const qsv_pass_t *P, const void *prog_host, uint32_t prog_bytes, const void *tables_dev,
             cudaStream_t st, unsigned grid, unsigned long long sp_mask, unsigned long long sp_val, unsigned zf_mask,
             unsigned zf_val, int n_bits, const int *positions, void *bk_dev) {
    SumSpec sm;
    if (make_sum_spec(sm, n_bits, positions)) return 1;
    return launch_sums(state_dev, P, prog_host, prog_bytes, tables_dev, st, grid, sp_mask, sp_val, zf_mask, zf_val, sm, bk_dev);
}
} }

static int make_swap_spec(jit::SwapSpec &sw, int j, const int *positions, const int *gbits, int pass_first,
                          const qsv_pass_t *pass) {
    QSV_CHECK_ARG(j >= 1 && j <= 4 && positions && pass, "qsv fused swap: n_swap=%d out of range or NULL pointer", j);
    memset(&sw, 0, sizeof(sw));
    sw.g = j;
    for (int k = 0; k < j; ++k) {
        sw.q[k] = positions[k];
        sw.gb[k] = gbits ? gbits[k] : k;
        QSV_CHECK_ARG(sw.gb[k] >= 0 && sw.gb[k] < 4 && (k == 0 || sw.gb[k] > sw.gb[k - 1]),
                      "qsv fused swap: global bits must be ascending in [0, 4)");
    }
    sw.pre = pass_first ? 1 : 0;
    sw.n_local = pass->n_local;       // baked into the code (rank bits of the support test, pass-first coordinates)
    return 0;
}

extern "C" int qsv_jit_source_swap_ex(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int n_swap,
                                      const int *positions, const int *gbits, int pass_first, char *buf, size_t buf_len,
                                      size_t *needed) {
    if (jit::validate(pass, prog_host, prog_bytes)) return 1;
    jit::SwapSpec sw;
    if (make_swap_spec(sw, n_swap, positions, gbits, pass_first, pass)) return 1;
    std::string src, err;
    if (jit::generate(*pass, prog_host, prog_bytes, src, err, &sw)) { set_error("qsv jit: %s", err.c_str()); return 1; }
    return copy_out(src, buf, buf_len, needed);
}

extern "C" int qsv_jit_source_swap(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int g,
                                   const int *positions, int pass_first, char *buf, size_t buf_len, size_t *needed) {
    return qsv_jit_source_swap_ex(pass, prog_host, prog_bytes, g, positions, nullptr, pass_first, buf, buf_len, needed);
}

extern "C" int qsv_jit_compile_swap_ex(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int n_swap,
                                       const int *positions, const int *gbits, int pass_first, char *cubin_buf,
                                       size_t buf_len, size_t *needed) {
    if (jit::validate(pass, prog_host, prog_bytes)) return 1;
    jit::SwapSpec sw;
    if (make_swap_spec(sw, n_swap, positions, gbits, pass_first, pass)) return 1;
    jit::Entry *e = jit::lookup(*pass, prog_host, prog_bytes, &sw);
    if (jit::ensure_compiled(e, *pass, prog_host, prog_bytes, &sw)) return 1;
    return copy_out(e->cubin, cubin_buf, buf_len, needed);
}

extern "C" int qsv_jit_compile_swap(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int g,
                                    const int *positions, int pass_first, char *cubin_buf, size_t buf_len, size_t *needed) {
    return qsv_jit_compile_swap_ex(pass, prog_host, prog_bytes, g, positions, nullptr, pass_first, cubin_buf, buf_len, needed);
}

static int run_swap(void *state_dev, const qsv_pass_t *pass, const void *prog_host,
                    uint32_t prog_bytes, const void *tables_dev, const uint64_t *peer_state_ptrs,
                    const uint64_t *peer_flag_ptrs, int n_ranks, int n_swap, const int *positions,
                    const int *gbits, int pass_first, int rank, uint64_t epoch, uint64_t sp_mask,
                    uint64_t sp_val, uint32_t zf_mask, uint32_t zf_val, uint32_t zf_on, void *stream) {
    QSV_CHECK_ARG(state_dev && pass && prog_host && peer_state_ptrs && peer_flag_ptrs, "qsv_run_pass_rounds_swap: NULL pointer");
    jit::SwapSpec sw;
    if (make_swap_spec(sw, n_swap, positions, gbits, pass_first, pass)) return 1;
    QSV_CHECK_ARG(n_ranks >= 2 && n_ranks <= 16 && !(n_ranks & (n_ranks - 1)), "qsv_run_pass_rounds_swap: %d ranks", n_ranks);
    QSV_CHECK_ARG(rank >= 0 && rank < n_ranks, "qsv_run_pass_rounds_swap: rank %d out of range", rank);
    for (int k = 0; k < n_swap; ++k)
        QSV_CHECK_ARG((1 << sw.gb[k]) < n_ranks, "qsv_run_pass_rounds_swap: global bit %d beyond %d ranks", sw.gb[k], n_ranks);
    jit::PeerTable pt;
    for (int j = 0; j < 16; ++j) { pt.state[j] = nullptr; pt.flags[j] = nullptr; }
    for (int j = 0; j < n_ranks; ++j) {
        pt.state[j] = j == rank ? state_dev : reinterpret_cast<void *>(peer_state_ptrs[j]);
        pt.flags[j] = reinterpret_cast<void *>(peer_flag_ptrs[j]);
        QSV_CHECK_ARG(pt.state[j] && pt.flags[j], "qsv_run_pass_rounds_swap: peer %d has no mapped pointer", j);
    }
    return jit::launch_swap(state_dev, pass, prog_host, prog_bytes, tables_dev, sw, pt, (unsigned)rank, epoch, sp_mask,
                            sp_val, zf_mask, zf_val, zf_on, (cudaStream_t)stream);
}

extern "C" int qsv_run_pass_rounds_swap_ex(void *state_dev, const qsv_pass_t *pass, const void *prog_host,
                                           uint32_t prog_bytes, const void *tables_dev, const uint64_t *peer_state_ptrs,
                                           const uint64_t *peer_flag_ptrs, int n_ranks, int n_swap, const int *positions,
                                           const int *gbits, int pass_first, int rank, uint64_t epoch, uint64_t sp_mask,
                                           uint64_t sp_val, void *stream) {
    return run_swap(state_dev, pass, prog_host, prog_bytes, tables_dev, peer_state_ptrs, peer_flag_ptrs, n_ranks, n_swap,
                    positions, gbits, pass_first, rank, epoch, sp_mask, sp_val, 0, 0, 0, stream);
}

extern "C" int qsv_run_pass_rounds_swap_fresh(void *state_dev, const qsv_pass_t *pass, const void *prog_host,
                                              uint32_t prog_bytes, const void *tables_dev, const uint64_t *peer_state_ptrs,
                                              const uint64_t *peer_flag_ptrs, int n_ranks, int n_swap, const int *positions,
                                              const int *gbits, int pass_first, int rank, uint64_t epoch, uint64_t sp_mask,
                                              uint64_t sp_val, uint32_t zf_mask, uint32_t zf_val, void *stream) {
    QSV_CHECK_ARG(!((zf_mask | zf_val) >> 12), "qsv_run_pass_rounds_swap_fresh: zf_mask / zf_val are tile-local (12 bits)");
    return run_swap(state_dev, pass, prog_host, prog_bytes, tables_dev, peer_state_ptrs, peer_flag_ptrs, n_ranks, n_swap,
                    positions, gbits, pass_first, rank, epoch, sp_mask, sp_val, zf_mask, zf_val, 1, stream);
}

extern "C" int qsv_run_pass_rounds_swap(void *state_dev, const qsv_pass_t *pass, const void *prog_host,
                                        uint32_t prog_bytes, const void *tables_dev, const uint64_t *peer_state_ptrs,
                                        const uint64_t *peer_flag_ptrs, int g, const int *positions, int pass_first,
                                        int rank, uint64_t epoch, void *stream) {
    return qsv_run_pass_rounds_swap_ex(state_dev, pass, prog_host, prog_bytes, tables_dev, peer_state_ptrs, peer_flag_ptrs,
                                       1 << g, g, positions, nullptr, pass_first, rank, epoch, 0, 0, stream);
}

extern "C" int qsv_jit_sums_static(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes) {
    if (jit::validate(pass, prog_host, prog_bytes)) return -1;
    const qsv_rprog_hdr_t *hdr = static_cast<const qsv_rprog_hdr_t *>(prog_host);
    const uint8_t *body = static_cast<const uint8_t *>(prog_host) + sizeof(qsv_rprog_hdr_t);
    if (hdr->n_rounds == 0) return 0;
    std::vector<qsv_round_t> rounds(hdr->n_rounds);
    std::vector<qsv_rop_t> ops(hdr->n_ops);
    memcpy(rounds.data(), body, sizeof(qsv_round_t) * hdr->n_rounds);
    memcpy(ops.data(), body + sizeof(qsv_round_t) * hdr->n_rounds, sizeof(qsv_rop_t) * hdr->n_ops);
    const qsv_round_t &R = rounds.back();
    const int n_pre = R.n_pre & 0x7fff, n_post = R.n_post & 0x7fff;
    const int n_reg = (int)(R.n_ops & 0xffffu), n_pt = (int)((R.n_ops >> 16) & 0x3fffu);
    uint32_t d[4];
    return jit::sum_static(R, ops.data(), (int)R.first_op + n_pre + n_pt + n_reg, n_post, (R.n_post & 0x8000) != 0, d) ? 1 : 0;
}

extern "C" int qsv_jit_bank_score(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                                  uint64_t *default_score, uint64_t *searched_score) {
    if (jit::validate(pass, prog_host, prog_bytes)) return 1;
    const qsv_rprog_hdr_t *hdr = static_cast<const qsv_rprog_hdr_t *>(prog_host);
    const uint8_t *body = static_cast<const uint8_t *>(prog_host) + sizeof(qsv_rprog_hdr_t);
    std::vector<qsv_round_t> rounds(hdr->n_rounds);
    std::vector<qsv_rop_t> ops(hdr->n_ops);
    memcpy(rounds.data(), body, sizeof(qsv_round_t) * hdr->n_rounds);
    memcpy(ops.data(), body + sizeof(qsv_round_t) * hdr->n_rounds, sizeof(qsv_rop_t) * hdr->n_ops);
    bool low_rb = false;
    for (const qsv_round_t &R : rounds)
        for (int q = 0; q < 4; ++q) low_rb = low_rb || R.rb[q] < 3;
    const bool direct = [] { const char *ds = getenv("QSV_JIT_DIRECT_STORE"); return !(ds && !strcmp(ds, "0")); }();
    uint64_t d = 0, sch = 0;
    for (uint32_t k = 0; k < hdr->n_rounds; ++k) {
        const bool stores = !(k + 1 == hdr->n_rounds && direct);
        int pos[8];
        jit::t_pad_L = (jit::pad_enabled() && low_rb) ? pass->low_bits : -1;       // what the default kernel does
        jit::thread_bit_order(rounds[k], pos);
        d += (uint64_t)jit::conflict_score(rounds[k], ops.data(), pos, stores);
    }
    {
        std::vector<std::array<int, 8>> maps;
        long total = 0;
        jit::search_layout(rounds, ops.data(), pass->low_bits, direct, jit::pad_enabled(), maps, &total);
        sch = (uint64_t)total;
    }
    jit::t_pad_L = -1;
    if (default_score) *default_score = d;
    if (searched_score) *searched_score = sch;
    return 0;
}

extern "C" int qsv_jit_wanted(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes) {
    if (!pass || !prog_host || prog_bytes < sizeof(qsv_rprog_hdr_t)) return 0;
    return jit::wanted(pass, prog_host) ? 1 : 0;
}

extern "C" uint64_t qsv_jit_disk_hits(void) { return jit::g_disk_hits.load(); }

extern "C" void qsv_jit_stats(uint64_t *kernels_compiled, uint64_t *cache_hits, double *compile_seconds) {
    if (kernels_compiled) *kernels_compiled = jit::g_compiles.load();
    if (cache_hits) *cache_hits = jit::g_hits.load();
    if (compile_seconds) *compile_seconds = 1e-6 * (double)jit::g_compile_us.load();
}
