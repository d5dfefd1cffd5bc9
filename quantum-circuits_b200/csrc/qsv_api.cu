// C-ABI glue: error reporting, launch counter, qsv_run_pass and the single-gate conveniences
// (which plan a one-op pass on the host, upload it with stream-ordered memory and launch k_pass).
#include "qsv_common.cuh"
#include <atomic>
#include <vector>
#include <string.h>
#include <algorithm>

namespace qsv {

static thread_local char g_err[512] = "";
static std::atomic<uint64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int launch_pass(void *state_dev, const qsv_pass_t *P, const void *prog_dev, const uint32_t *op_off_dev, int n_ops,
                bool wide_k, cudaStream_t st, uint64_t fixed_mask = 0, uint64_t fixed_val = 0);

constexpr int kDefaultT = 12;
constexpr int kDefaultL = 6;

struct Plan {
    qsv_pass_t pass;
    int loc[64];   // physical position -> tile-local position, -1 if outside the tile
};

// Choose a tile containing every bit of need_mask, avoiding avoid_mask where possible.
static int plan_single(int n_local, uint64_t need_mask, uint64_t avoid_mask, uint64_t rank_bits, Plan &pl) {
    QSV_CHECK_ARG(n_local >= 1 && n_local <= 40, "n_local=%d out of range", n_local);
    const int T = std::min(n_local, kDefaultT);
    int L = std::min(T, kDefaultL);
    auto high_needed = [&](int l) { return __builtin_popcountll(need_mask >> l); };
    while (L > 0 && (high_needed(L) > T - L)) --L;
    QSV_CHECK_ARG(high_needed(L) <= T - L && T - L <= 8, "gate does not fit one tile (T=%d)", T);
    memset(&pl.pass, 0, sizeof(pl.pass));
    pl.pass.n_local = n_local;
    pl.pass.tile_bits = T;
    pl.pass.low_bits = L;
    pl.pass.n_hi = T - L;
    pl.pass.rank_bits = rank_bits;
    uint64_t chosen = need_mask >> L << L;
    int have = __builtin_popcountll(chosen);
    for (int round = 0; round < 2 && have < T - L; ++round)          // round 0 avoids control bits
        for (int b = L; b < n_local && have < T - L; ++b) {
            if ((chosen >> b) & 1ull) continue;
            if (round == 0 && ((avoid_mask >> b) & 1ull)) continue;
            chosen |= 1ull << b;
            ++have;
        }
    for (int b = 0; b < 64; ++b) pl.loc[b] = (b < L) ? b : -1;
    int j = 0;
    for (int b = L; b < n_local; ++b)
        if ((chosen >> b) & 1ull) {
            pl.pass.hi_pos[j] = (uint8_t)b;
            pl.loc[b] = L + j;
            ++j;
        }
    return 0;
}

struct OpBuilder {
    std::vector<uint8_t> bytes;
    qsv_op_t *hdr() { return reinterpret_cast<qsv_op_t *>(bytes.data()); }
    void init(uint32_t kind, uint32_t k, size_t payload_bytes) {
        const size_t total = (QSV_OP_HEADER_BYTES + payload_bytes + 15) / 16 * 16;
        bytes.assign(total, 0);
        hdr()->kind = kind;
        hdr()->k = k;
        hdr()->payload_off = QSV_OP_HEADER_BYTES;
        hdr()->op_bytes = (uint32_t)total;
    }
    // split controls into tile-local (ins + loc_ctrl_val) and external (mask/val)
    int add_controls(const Plan &pl, int n_ctrl, const int *controls, std::vector<int> &ins) {
        for (int c = 0; c < n_ctrl; ++c) {
            const int p = controls[c];
            QSV_CHECK_ARG(p >= 0 && p < 64, "control position %d out of range", p);
            if (p < pl.pass.n_local && pl.loc[p] >= 0) {
                ins.push_back(pl.loc[p]);
                hdr()->loc_ctrl_val |= 1u << pl.loc[p];
            } else {
                hdr()->ext_ctrl_mask |= 1ull << p;
                hdr()->ext_ctrl_val |= 1ull << p;
            }
        }
        return 0;
    }
    int finish_ins(std::vector<int> &ins) {
        std::sort(ins.begin(), ins.end());
        QSV_CHECK_ARG(ins.size() <= 16, "too many tile-local bits in one op");
        for (size_t i = 1; i < ins.size(); ++i)
            QSV_CHECK_ARG(ins[i] != ins[i - 1], "a qubit is used twice by one gate");
        hdr()->n_ins = (uint32_t)ins.size();
        for (size_t i = 0; i < ins.size(); ++i) hdr()->ins[i] = (uint8_t)ins[i];
        return 0;
    }
};

static int run_single(void *state_dev, const Plan &pl, OpBuilder &ob, int dense_k, cudaStream_t st) {
    const size_t off_bytes = 16;
    const size_t total = off_bytes + ob.bytes.size();
    std::vector<uint8_t> host(total, 0);
    uint32_t first = (uint32_t)off_bytes;
    memcpy(host.data(), &first, 4);
    memcpy(host.data() + off_bytes, ob.bytes.data(), ob.bytes.size());
    void *dev = nullptr;
    QSV_CUDA(cudaMallocAsync(&dev, total, st));
    QSV_CUDA(cudaMemcpyAsync(dev, host.data(), total, cudaMemcpyHostToDevice, st));
    qsv_pass_t pass = pl.pass;
    pass.max_dense_k = dense_k;
    // the single op's external controls pin those index bits for every tile the launch has to visit
    int rc = launch_pass(state_dev, &pass, dev, reinterpret_cast<const uint32_t *>(dev), 1, dense_k > 3, st,
                         ob.hdr()->ext_ctrl_mask, ob.hdr()->ext_ctrl_val);
    cudaFreeAsync(dev, st);
    return rc;
}

static int masks_of(int n, const int *pos, int limit, uint64_t *mask, const char *who) {
    *mask = 0;
    for (int i = 0; i < n; ++i) {
        QSV_CHECK_ARG(pos[i] >= 0 && pos[i] < limit, "%s: position %d out of range", who, pos[i]);
        *mask |= 1ull << pos[i];
    }
    return 0;
}

}  // namespace qsv

using namespace qsv;

extern "C" int qsv_abi_version(void) { return QSV_ABI_VERSION; }
extern "C" const char *qsv_last_error(void) { return g_err; }
extern "C" uint64_t qsv_launch_count(void) { return g_launches.load(); }
extern "C" void qsv_reset_launch_count(void) { g_launches.store(0); }

extern "C" int qsv_run_pass(void *state_dev, const qsv_pass_t *pass, const void *prog_dev,
                            const uint32_t *op_offsets_dev, int n_ops, void *stream) {
    QSV_CHECK_ARG(pass, "qsv_run_pass: pass is NULL");
    QSV_CHECK_ARG(pass->max_dense_k >= 0 && pass->max_dense_k <= QSV_MAX_TILE_K,
                  "qsv_run_pass: max_dense_k=%d exceeds %d", pass->max_dense_k, QSV_MAX_TILE_K);
    return launch_pass(state_dev, pass, prog_dev, op_offsets_dev, n_ops, pass->max_dense_k > 3,
                       (cudaStream_t)stream);
}

extern "C" int qsv_run_pass_sparse(void *state_dev, const qsv_pass_t *pass, const void *prog_dev,
                                   const uint32_t *op_offsets_dev, int n_ops, uint64_t fixed_mask, uint64_t fixed_val,
                                   void *stream) {
    QSV_CHECK_ARG(pass, "qsv_run_pass_sparse: pass is NULL");
    QSV_CHECK_ARG(pass->max_dense_k >= 0 && pass->max_dense_k <= QSV_MAX_TILE_K,
                  "qsv_run_pass_sparse: max_dense_k=%d exceeds %d", pass->max_dense_k, QSV_MAX_TILE_K);
    QSV_CHECK_ARG((fixed_val & ~fixed_mask) == 0, "qsv_run_pass_sparse: fixed_val has bits outside fixed_mask");
    return launch_pass(state_dev, pass, prog_dev, op_offsets_dev, n_ops, pass->max_dense_k > 3,
                       (cudaStream_t)stream, fixed_mask, fixed_val);
}

extern "C" int qsv_apply_dense(void *state_dev, int n_local, int k, const int *targets, int n_ctrl,
                               const int *controls, const double *matrix_re_im, uint64_t rank_bits, void *stream) {
    QSV_CHECK_ARG(state_dev && targets && matrix_re_im, "qsv_apply_dense: NULL pointer");
    QSV_CHECK_ARG(k >= 1 && k <= QSV_MAX_TILE_K, "qsv_apply_dense: k=%d (1..%d; use qsv_apply_dense_wide)", k,
                  QSV_MAX_TILE_K);
    QSV_CHECK_ARG(n_ctrl >= 0 && (n_ctrl == 0 || controls), "qsv_apply_dense: bad controls");
    uint64_t need, avoid;
    if (int rc = masks_of(k, targets, n_local, &need, "qsv_apply_dense")) return rc;
    if (int rc = masks_of(n_ctrl, controls, 64, &avoid, "qsv_apply_dense")) return rc;
    QSV_CHECK_ARG(!(need & avoid), "qsv_apply_dense: a qubit is both target and control");
    Plan pl;
    if (int rc = plan_single(n_local, need, avoid, rank_bits, pl)) return rc;
    OpBuilder ob;
    const size_t D = (size_t)1 << k;
    ob.init(QSV_OP_DENSE, (uint32_t)k, D * D * 16);
    memcpy(ob.bytes.data() + QSV_OP_HEADER_BYTES, matrix_re_im, D * D * 16);
    bool real = true;
    for (size_t i = 0; i < D * D; ++i) real = real && (matrix_re_im[2 * i + 1] == 0.0);
    if (real) ob.hdr()->flags |= QSV_OPF_REAL_MATRIX;
    std::vector<int> ins;
    for (int p = 0; p < k; ++p) {
        ob.hdr()->tgt[p] = (uint8_t)pl.loc[targets[p]];
        ins.push_back(pl.loc[targets[p]]);
    }
    if (int rc = ob.add_controls(pl, n_ctrl, controls, ins)) return rc;
    if (int rc = ob.finish_ins(ins)) return rc;
    return run_single(state_dev, pl, ob, k, (cudaStream_t)stream);
}

extern "C" int qsv_apply_diag(void *state_dev, int n_local, int k, const int *targets, const double *diag_re_im,
                              uint64_t rank_bits, void *stream) {
    QSV_CHECK_ARG(state_dev && targets && diag_re_im, "qsv_apply_diag: NULL pointer");
    QSV_CHECK_ARG(k >= 1 && k <= 8, "qsv_apply_diag: k=%d (1..8; use qsv_apply_diag_wide)", k);
    uint64_t tm;
    if (int rc = masks_of(k, targets, 64, &tm, "qsv_apply_diag")) return rc;
    QSV_CHECK_ARG(__builtin_popcountll(tm) == k, "qsv_apply_diag: duplicate target");
    Plan pl;
    if (int rc = plan_single(n_local, 0, tm, rank_bits, pl)) return rc;   // diagonal: nothing has to be local
    OpBuilder ob;
    ob.init(QSV_OP_DIAG, (uint32_t)k, ((size_t)1 << k) * 16);
    memcpy(ob.bytes.data() + QSV_OP_HEADER_BYTES, diag_re_im, ((size_t)1 << k) * 16);
    for (int p = 0; p < k; ++p) {
        const int b = targets[p];
        ob.hdr()->tgt[p] = (b < n_local && pl.loc[b] >= 0) ? (uint8_t)pl.loc[b] : (uint8_t)(0x80 | b);
    }
    std::vector<int> ins;
    if (int rc = ob.finish_ins(ins)) return rc;
    return run_single(state_dev, pl, ob, 0, (cudaStream_t)stream);
}

extern "C" int qsv_apply_phase(void *state_dev, int n_local, int n_ctrl, const int *controls, double re, double im,
                               uint64_t rank_bits, void *stream) {
    QSV_CHECK_ARG(state_dev && n_ctrl >= 1 && controls, "qsv_apply_phase: need at least one control bit");
    uint64_t avoid;
    if (int rc = masks_of(n_ctrl, controls, 64, &avoid, "qsv_apply_phase")) return rc;
    Plan pl;
    if (int rc = plan_single(n_local, 0, avoid, rank_bits, pl)) return rc;
    OpBuilder ob;
    ob.init(QSV_OP_PHASE, 0, 0);
    ob.hdr()->scalar[0] = re;
    ob.hdr()->scalar[1] = im;
    std::vector<int> ins;
    if (int rc = ob.add_controls(pl, n_ctrl, controls, ins)) return rc;
    if (int rc = ob.finish_ins(ins)) return rc;
    return run_single(state_dev, pl, ob, 0, (cudaStream_t)stream);
}

extern "C" int qsv_apply_cx(void *state_dev, int n_local, int n_ctrl, const int *controls, int target,
                            uint64_t rank_bits, void *stream) {
    QSV_CHECK_ARG(state_dev, "qsv_apply_cx: NULL state");
    QSV_CHECK_ARG(target >= 0 && target < n_local, "qsv_apply_cx: target %d is not a local qubit", target);
    QSV_CHECK_ARG(n_ctrl >= 0 && (n_ctrl == 0 || controls), "qsv_apply_cx: bad controls");
    uint64_t avoid;
    if (int rc = masks_of(n_ctrl, controls, 64, &avoid, "qsv_apply_cx")) return rc;
    QSV_CHECK_ARG(!((avoid >> target) & 1ull), "qsv_apply_cx: target is also a control");
    Plan pl;
    if (int rc = plan_single(n_local, 1ull << target, avoid, rank_bits, pl)) return rc;
    OpBuilder ob;
    ob.init(QSV_OP_MCX, 1, 0);
    ob.hdr()->tgt[0] = (uint8_t)pl.loc[target];
    std::vector<int> ins{pl.loc[target]};
    if (int rc = ob.add_controls(pl, n_ctrl, controls, ins)) return rc;
    if (int rc = ob.finish_ins(ins)) return rc;
    return run_single(state_dev, pl, ob, 0, (cudaStream_t)stream);
}
