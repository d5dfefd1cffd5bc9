/*
 * qsv.h — C-ABI of the B200-native state-vector engine (libqsv.so).
 *
 * This is the drop-in boundary underneath the reference's Python API
 * (MatejVitek/Quantum-Circuits, main/circuit.py + main/matrix.py).  The reference has no
 * FFI of its own: its hot path is Circuit.run_method2 (main/circuit.py:279-337) plus the
 * sampling in Circuit.run (main/circuit.py:341-366).  Every entry point below replaces one
 * step of that path and cites it.  The Python mirror of the reference API
 * (quantum-circuits_b200/main/circuit.py) binds these symbols with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - All pointers named *_dev are CUDA device pointers; everything else is host memory.
 *   - `state_dev` is an array of 2^n_local complex128 amplitudes (re, im interleaved, 16 B each).
 *   - Bit positions are PHYSICAL index bits: position 0 = least-significant bit of the amplitude
 *     index.  The reference is big-endian (wire 0 = MSB, main/circuit.py:290-299), so for an
 *     unsharded n-qubit state wire w lives at position n-1-w.  Positions >= n_local are "global"
 *     qubits held in `rank_bits` (multi-GPU sharding by the top log2(P) qubits).
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - Every function returns 0 on success, non-zero on error; qsv_last_error() then returns a
 *     thread-local message.  Nothing here ever falls back to the CPU.
 *   - Threading: every entry point may be called from several host threads at once (each on its own
 *     stream and state, or for compilation only); the library keeps no per-call state in shared statics
 *     (the kernel cache is mutex-protected, launch scratch and the error string are thread-local).
 *     Two calls that touch the SAME state buffer must be ordered by the caller (same stream or events).
 */
#ifndef QSV_H
#define QSV_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QSV_ABI_VERSION 1

/* ---- in-tile op program (consumed by qsv_run_pass) -------------------------------------- */

#define QSV_MAX_TILE_K   5   /* widest dense / diagonal gate applied inside a shared-memory tile */
#define QSV_MAX_TILE_T  13   /* largest tile: 2^13 amplitudes = 128 KiB of shared memory        */
#define QSV_OP_HEADER_BYTES 96

enum qsv_op_kind {
    QSV_OP_DENSE = 1,   /* k-qubit dense matrix on tile-local targets, optional controls        */
    QSV_OP_MCX   = 2,   /* (multi-)controlled X: exact pair swap (X, CNot, Toffoli)             */
    QSV_OP_PHASE = 3,   /* multiply amplitudes whose control bits match by one complex scalar   */
    QSV_OP_DIAG  = 4,   /* k-qubit diagonal table; bits may be tile-local or external           */
    QSV_OP_SWAP  = 5    /* exchange two tile-local qubits (exact data movement)                 */
};

/* One op of a pass program.  Laid out for direct device consumption; ops are variable-size:
 * header (96 B) followed by `payload` (dense matrix row-major 2^k x 2^k complex128, or diagonal
 * table 2^k complex128).  All positions in `ins`/`tgt` are TILE-LOCAL bit positions unless the
 * 0x80 flag marks them as external physical positions (DIAG only). */
typedef struct qsv_op {
    uint32_t kind;          /* enum qsv_op_kind                                                   */
    uint32_t k;             /* number of target qubits (DENSE, DIAG); 1 for MCX; 2 for SWAP        */
    uint32_t n_ins;         /* number of entries of ins[]                                          */
    uint32_t flags;         /* QSV_OPF_*                                                           */
    uint64_t ext_ctrl_mask; /* physical-index bits OUTSIDE the tile that gate the whole op         */
    uint64_t ext_ctrl_val;  /* required value of those bits                                        */
    uint32_t loc_ctrl_val;  /* tile-local control bits that must be 1 (OR-ed into the local index) */
    uint32_t payload_off;   /* byte offset of the payload from the start of this op (16-aligned)   */
    uint32_t op_bytes;      /* total bytes of this op including payload (multiple of 16)           */
    uint32_t reserved;
    uint8_t  ins[16];       /* sorted ascending: tile-local bits fixed by the op (targets+controls) */
    uint8_t  tgt[8];        /* tgt[0] = most-significant bit of the gate index (= gate port 0)     */
    double   scalar[2];     /* PHASE: (re, im)                                                     */
} qsv_op_t;

#define QSV_OPF_REAL_MATRIX 1u   /* DENSE: all imaginary parts are exactly zero                   */

/* Geometry of one pass over the state: which physical bits form the shared-memory tile. */
typedef struct qsv_pass {
    int32_t  n_local;       /* log2(number of amplitudes on this device)                           */
    int32_t  tile_bits;     /* T: tile holds 2^T amplitudes (T <= QSV_MAX_TILE_T, T <= n_local)    */
    int32_t  low_bits;      /* L: physical bits 0..L-1 are always in the tile (contiguous runs)    */
    int32_t  n_hi;          /* T - L                                                              */
    uint8_t  hi_pos[16];    /* ascending physical positions (>= L) of the remaining tile bits      */
    uint64_t rank_bits;     /* value of the global index bits, pre-shifted by n_local              */
    int32_t  max_dense_k;   /* widest DENSE op in the program (selects the register-heavy variant)  */
    int32_t  reserved;      /* qsv_run_pass_rounds: register bits per round, 3 or 4 (0 = 4)         */
} qsv_pass_t;

/* ---- register-resident pass program (consumed by qsv_run_pass_rounds) --------------------- */
/* The tile (2^T amplitudes, T >= 4) is held in REGISTERS: 2^(T-4) threads x 16 amplitudes.  A pass is a
 * sequence of ROUNDS; a round fixes 4 tile-local "register bits" rb[0..3] (ascending): thread t owns the
 * 16 amplitudes that differ only in those bits, so every op whose non-diagonal target is a register bit
 * runs without touching shared memory.  Between rounds the tile is re-distributed through shared memory.
 * Host image: [qsv_rprog_hdr_t][qsv_round_t x n_rounds][qsv_rop_t x n_ops]; it is passed to the kernel
 * BY VALUE (constant bank), so op fields and matrix entries are read through the uniform datapath.
 * Diagonal tables (rare) stay in a device buffer; qsv_rop_t.table_off is a byte offset into it. */

#define QSV_RPROG_MAX_ROUNDS 32      /* interpreting kernel: the program travels as a kernel parameter */
#define QSV_RPROG_MAX_OPS    120
#define QSV_JIT_MAX_ROUNDS   256     /* pass-specialised kernel: the program becomes code, not data      */
#define QSV_JIT_MAX_OPS      2048
/* kind byte of qsv_rop_t.  0..17 are contiguous (one jump table in the kernel). */
#define QSV_ROP_DENSE1       0   /* +J      complex 1-qubit matrix on register bit J                    */
#define QSV_ROP_DENSE1_REAL  4   /* +J      real matrix                                                 */
#define QSV_ROP_PHASE1       8   /* +2J+V   phase on the slots with register bit J == V                 */
#define QSV_ROP_GEN_PHASE   16   /*         phase with any (rmask, rval) condition on the register index */
#define QSV_ROP_DIAG        17   /*         diagonal table (k <= 8 bits: thread, register or external)  */
#define QSV_ROP_PHASE_T     18   /*         phase independent of the register bits: kept in a separate
                                            list per round and folded into one scalar per thread       */
#define QSV_ROP_PERM        19   /* (multi-)controlled X as an address permutation at a round boundary:
                                    if ((addr & tmask) == tval) addr ^= table_off   (tile-local bits).
                                    Linear lists: j = 0xff if the condition is slot-independent, else
                                    j = register index q, rval = position and flags = mask of the one
                                    condition bit that varies with r_q (flags = 0 when j = 0xff)        */
#define QSV_ROP_MCX_REG     20   /* pass-specialised kernel only: (multi-)controlled X whose target is register bit j,
                                    executed in the register phase: slots r (bit j clear) with (r & rmask) == rval trade
                                    places with r | 1<<j where the thread / external condition holds */
#define QSV_ROPF_REAL  1u   /* matrix entries are real                                               */
#define QSV_ROPF_NEG   2u   /* PHASE: scalar is exactly -1                                            */

typedef struct qsv_rprog_hdr {
    uint32_t n_rounds, total_bytes, n_ops, reserved;   /* reserved: register bits per round (informational) */
} qsv_rprog_hdr_t;

typedef struct qsv_round {
    uint8_t  rb[4];         /* ascending tile-local positions held in registers this round             */
    uint32_t n_ops;         /* bits 0-15: ordered register ops; bits 16-29: QSV_ROP_PHASE_T ops        */
    uint32_t first_op;      /* index of the round's first op: [PERM pre][PHASE_T][register ops][PERM post] */
    uint16_t n_pre;         /* permutation ops applied to the addresses the round LOADS from;
                               bit 15: the list is not XOR-linear, use the per-slot general form       */
    uint16_t n_post;        /* same for the addresses the round STORES to                              */
} qsv_round_t;

typedef struct qsv_rop {
    uint8_t  kind;          /* QSV_ROP_*                                                               */
    uint8_t  j;             /* DENSE1 / MCX: target register bit 0..3                                  */
    uint8_t  rmask, rval;   /* condition on the register index r (controls that are register bits)     */
    uint32_t tmask, tval;   /* condition on the thread's tile-local base index (register bits are 0)   */
    uint32_t flags2;        /* bit 0: ext_mask != 0 (the external condition has to be evaluated)       */
    uint64_t ext_mask;      /* condition on physical index bits outside the tile                       */
    uint64_t ext_val;
    uint32_t flags;         /* QSV_ROPF_*                                                              */
    uint32_t table_off;     /* DIAG: byte offset of the 2^k complex128 table inside tables_dev         */
    uint8_t  dsrc[8];       /* DIAG table bit p (p=0 most significant): 0x00..0x3f thread-held local
                               position, 0x80|pos external physical position, 0xff = register bit      */
    uint8_t  rc[4];         /* DIAG: table-index bits contributed by register bit 0..3                 */
    uint32_t k;             /* DIAG: number of table bits                                              */
    double   m[8];          /* DENSE1: m00,m01,m10,m11 as (re,im); PHASE: m[0..1] = scalar             */
    uint8_t  pad[8];
} qsv_rop_t;                /* 128 bytes                                                               */

/* ---- library ---------------------------------------------------------------------------- */

int         qsv_abi_version(void);
const char *qsv_last_error(void);
/* Number of kernels this library has launched since load / last reset (bench.py's gpu_launches). */
uint64_t    qsv_launch_count(void);
void        qsv_reset_launch_count(void);

/* ---- state preparation ------------------------------------------------------------------ */

/* |in_v> basis state: replaces the tensor product of e0/e1 columns, main/circuit.py:290-301 and
 * main/matrix.py:6-11.  `index` is the global basis index; only the shard that owns it (top bits
 * == rank) writes the 1. */
int qsv_init_basis(void *state_dev, int n_local, uint64_t index, uint64_t rank, void *stream);

/* |index> WITHOUT the 16 * 2^n_local byte clear (the reference builds its initial column entry by entry,
 * main/circuit.py:290-301; here the clear alone is 2.6 ms at 30 qubits): only the basis amplitude is written, every other
 * amplitude of the register stays UNINITIALISED.  Valid only as the start of a sequence of qsv_run_pass_rounds_fresh
 * passes that ends with one touching every tile; the engine (qsv.engine.StateVector.run_from_basis) checks that. */
int qsv_init_basis_sparse(void *state_dev, int n_local, uint64_t index, uint64_t rank, void *stream);

/* All basis inputs of an n-qubit circuit in ONE register of 2n qubits (Circuit.get_subcircuit,
 * main/circuit.py:122-157: the whole-circuit unitary column by column): amplitude 1 at index (c << n) | (c ^ flip)
 * for every column c, 0 elsewhere.  The circuit's ops act on positions < n; afterwards state[(c << n) | y] is
 * entry (y, c) of the unitary in physical order.  `flip` = the X gates the lowering pushed back to the input. */
int qsv_init_columns(void *state_dev, int n, uint64_t flip, void *stream);

/* ---- gate application: replaces P2*psi; M*psi; P1*psi, main/circuit.py:303-327 ----------- */

/* Apply `n_ops` ops to every tile of one pass.  `op_offsets_dev[i]` is the byte offset of op i
 * from `prog_dev`.  One read + one write of every tile that has at least one active op; tiles whose
 * external controls switch every op off are skipped entirely. */
int qsv_run_pass(void *state_dev, const qsv_pass_t *pass, const void *prog_dev,
                 const uint32_t *op_offsets_dev, int n_ops, void *stream);

/* qsv_run_pass restricted to the tiles whose global base index equals `fixed_val` on the bits of `fixed_mask`
 * (local positions outside the tile and/or global positions as in pass->rank_bits; bits inside the tile are ignored).
 * The caller pins what it knows: the support of a register that started in a basis state (main/circuit.py:290-301,
 * see qsv_run_pass_rounds_sparse) and/or a control value shared by every op of the program.  The grid enumerates the
 * remaining tiles only, so a CZ / Toffoli pass spends no CTA on the 3/4 of the tiles it would skip anyway. */
int qsv_run_pass_sparse(void *state_dev, const qsv_pass_t *pass, const void *prog_dev, const uint32_t *op_offsets_dev,
                        int n_ops, uint64_t fixed_mask, uint64_t fixed_val, void *stream);

/* Same contract as qsv_run_pass for a register-resident program (see qsv_rop_t).  `prog_host` is HOST
 * memory (at most QSV_RPROG_MAX_ROUNDS rounds / QSV_RPROG_MAX_OPS ops); `tables_dev` is the device
 * buffer diagonal ops index with table_off (may be NULL if the program has none).  tile_bits >= 4. */
int qsv_run_pass_rounds(void *state_dev, const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                        const void *tables_dev, void *stream);

/* qsv_run_pass_rounds for a state whose SUPPORT is known: index bits under `fixed_mask` (local positions outside
 * the tile and/or global positions, as in pass->rank_bits) equal `fixed_val` wherever the state is non-zero - true
 * for a register that started in a basis state (main/circuit.py:290-301) on every position no gate has touched
 * non-diagonally yet.  Tiles that disagree hold only zeros and, the pass being linear, stay zero: they are neither
 * read nor written.  The result is identical to qsv_run_pass_rounds; kernels without the skip ignore the hint. */
int qsv_run_pass_rounds_sparse(void *state_dev, const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                               const void *tables_dev, uint64_t fixed_mask, uint64_t fixed_val, void *stream);

/* qsv_run_pass_rounds_sparse on a register that was started with qsv_init_basis_sparse and has not been swept in full
 * yet: memory outside the support is uninitialised.  The pass loads the tiles of the support and takes every amplitude
 * whose TILE-LOCAL index (bit t = position t for t < low_bits, else pass->hi_pos[t - low_bits]) differs from zf_val on
 * the bits of zf_mask - the support's fixed bits inside the tile - as zero, whatever was loaded; every tile it touches
 * is written in full.  Only the pass-specialised kernel implements this form (error otherwise; no fallback). */
int qsv_run_pass_rounds_fresh(void *state_dev, const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                              const void *tables_dev, uint64_t fixed_mask, uint64_t fixed_val, uint32_t zf_mask,
                              uint32_t zf_val, void *stream);

/* The LAST pass of a run that ends in a sample (main/circuit.py:360-364): qsv_run_pass_rounds_sparse / _fresh
 * (zf_mask = 0: nothing to zero-fill) that also adds up, tile by tile while the final amplitudes sit in registers,
 * the weights re^2 + im^2 in the sampler's exact fixed point and leaves them in `buckets_dev` (4 << n_bits uint64 limbs,
 * cleared by this call): bucket = the index bits at bucket_pos[0..n_bits) (physical positions OUTSIDE the pass's tile,
 * bucket_pos[0] = least significant; global positions are read from pass->rank_bits).  With bucket_pos = the positions of
 * the top output bits this is the first level of the radix descent of qsv_sample_level - without its read of the whole
 * state (17 GB = 2.6 ms at 30 qubits); continue with qsv_sample_select.  Pass-specialised kernel only. */
int qsv_run_pass_rounds_sums(void *state_dev, const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                             const void *tables_dev, uint64_t fixed_mask, uint64_t fixed_val, uint32_t zf_mask,
                             uint32_t zf_val, int n_bits, const int *bucket_pos, void *buckets_dev, void *stream);

/* Pass-specialised form of qsv_run_pass_rounds.  For large states (n_local >= 28 by default; environment
 * QSV_JIT=0 never, QSV_JIT=force always, QSV_JIT_MIN_QUBITS=n) qsv_run_pass_rounds does not interpret the
 * round program: libqsv generates straight-line sm_100a CUDA for exactly this program (same tile geometry,
 * TMA bulk load/store and round structure), compiles it once with NVRTC and caches the kernel by the
 * program bytes.  These entry points expose the generator without a GPU:
 *   qsv_jit_source   the generated CUDA source (also valid C++ with -DQSV_JIT_HOST: tests/ run it on the CPU)
 *   qsv_jit_compile  compile (or fetch from the cache) and return the sm_100a cubin; safe to call from
 *                    several host threads at once to prepare the kernels of a circuit in parallel
 *   qsv_jit_wanted   1 if qsv_run_pass_rounds would take the specialised path for this pass
 *   qsv_jit_stats    kernels compiled, cache hits and host seconds spent compiling since load
 * buf may be NULL; *needed receives the full size.  tile_bits must be 12 with 4 register bits per round. */
int  qsv_jit_source(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, char *buf,
                    size_t buf_len, size_t *needed);
int  qsv_jit_compile(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, char *cubin_buf,
                     size_t buf_len, size_t *needed);
int  qsv_jit_wanted(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes);
void qsv_jit_stats(uint64_t *kernels_compiled, uint64_t *cache_hits, double *compile_seconds);
/* Kernels served from the on-disk cubin cache instead of NVRTC since the library was loaded (QSV_CACHE_DIR, default
 * ~/.cache/qsv; keyed by the generated source + compiler version; QSV_DISK_CACHE=0 disables). */
uint64_t qsv_jit_disk_hits(void);
/* Source / cubin of the sums variant of a pass kernel (see qsv_run_pass_rounds_sums; tests compile the source as host
 * C++, where qsv_jit_host_set_buckets(uint64_t*) names the bucket array before qsv_jit_host_run). */
int  qsv_jit_source_sums(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int n_bits,
                         const int *bucket_pos, char *buf, size_t buf_len, size_t *needed);
int  qsv_jit_compile_sums(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int n_bits,
                          const int *bucket_pos, char *cubin_buf, size_t buf_len, size_t *needed);
/* Modelled shared-memory bank-group collisions of one pass program (sampled quarter-warps x slots x rounds; 0 = none):
 * with the default thread-index mapping, and with the per-round mapping the opt-in search (QSV_JIT_BANKSEARCH=1) picks.
 * A planning aid that needs no GPU (tools/plan_report.py --banks). */
/* 1 if the sums variant of this pass classifies amplitudes into buckets with constants fixed at generation time (nothing
 * that depends on the thread or the tile permutes the last round's stores: any number of bucket bits inside the tile is
 * then as cheap as none), 0 if it reads every amplitude's bucket from its final index (~2.4 ms per 2^30 amplitudes for two
 * inside bits, ~6 ms for three: the host then settles for fewer bits), -1 on an invalid program. */
int  qsv_jit_sums_static(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes);
int  qsv_jit_bank_score(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                        uint64_t *default_score, uint64_t *searched_score);

/* ---- Feynman path sum: Circuit.run_method1, main/circuit.py:221-276 ------------------------------------------
 * weights[out] = | sum over the 2^n_internal assignments of the internal wires of prod over the gates of
 * matrix_g[i][j] |^2, i / j = the values of the gate's in- / out-wires (first wire most significant), for every
 * assignment `out` of the n_outputs output wires (index bit n_outputs-1-i = output wire i).  Device arrays:
 * wires [n_wires][2] uint8 = (kind, idx): kind 0 constant bit idx (fed by a circuit input), 1 bit idx of `out`,
 * 2 bit idx of the internal assignment; gate_hdr [n_gates][3] uint32 = (k, offset of the row-major 2^k x 2^k complex128
 * matrix in mats, offset of the gate's 2k wire ids in ports: k in-wires then k out-wires); direct [n_direct][2] uint32 =
 * (input bit, bit position of `out`) for wires that run from a circuit input straight to an output.  The reference's
 * own algorithm and cost (2^(n + internal) * gates), one CTA per output assignment. */
int qsv_path_sum(const void *wires_dev, int n_wires, const void *gate_hdr_dev, const void *ports_dev,
                 const void *mats_dev, int n_gates, int n_outputs, int n_internal, const void *direct_dev,
                 int n_direct, void *weights_dev, void *stream);

/* ---- host-side pass planning (no device work; qsv.schedule) ---------------------------------------------------
 * The reference applies one gate per full-state product (main/circuit.py:315-327); here a PASS (one read+write of the
 * state) applies every op whose non-diagonal targets lie inside its tile and that commutes past what stays behind.
 * Op i of the pending list is described by nd[i] / dg[i] (bit masks of the positions it uses non-diagonally /
 * diagonally or as a control) and never[i] != 0 when it cannot join a pass at all (whole-state ops, swaps).
 * qsv_plan_absorbed: how many of the n_ops ops a pass whose tile holds exactly the positions in `allowed` takes.
 * qsv_plan_improve_tile: local search over the tile's high positions (exchange one position at a time while the pass
 * absorbs more of the first `window` ops); seed 0 = deterministic candidate order, seed > 0 = shuffled orders (the
 * randomised schedule search).  hi_out must hold n_free_slots ints.  Returns 0, or -1 / non-zero on bad arguments. */
int qsv_plan_absorbed(const uint64_t *nd, const uint64_t *dg, const uint8_t *never, int n_ops, uint64_t allowed);
int qsv_plan_improve_tile(const uint64_t *nd, const uint64_t *dg, const uint8_t *never, int n_ops, int window,
                          uint64_t low_mask, const int *hi_in, int n_hi_in, int n_free_slots, int n_local,
                          int low_bits, int max_rounds, uint64_t seed, int *hi_out, int *n_hi_out, int *absorbed_out);

/* Single-gate conveniences (host arguments only; each is one pass with a one-op program built by
 * the library).  `targets[0]` is the gate's port 0 = most-significant bit of the matrix index
 * (main/circuit.py:305-313).  Controls are physical positions that must be 1. */
int qsv_apply_dense(void *state_dev, int n_local, int k, const int *targets,
                    int n_ctrl, const int *controls, const double *matrix_re_im, /* 2*4^k doubles */
                    uint64_t rank_bits, void *stream);
int qsv_apply_diag(void *state_dev, int n_local, int k, const int *targets,
                   const double *diag_re_im, /* 2*2^k doubles */ uint64_t rank_bits, void *stream);
int qsv_apply_phase(void *state_dev, int n_local, int n_ctrl, const int *controls,
                    double re, double im, uint64_t rank_bits, void *stream);
int qsv_apply_cx(void *state_dev, int n_local, int n_ctrl, const int *controls, int target,
                 uint64_t rank_bits, void *stream);

/* Wide gates (k > QSV_MAX_TILE_K), used for the reference's full-width dense operators
 * (main/Grover.py:14-31, main/Shor.py:54-83, main/matrix.py:212-219).  Out-of-place for dense and
 * permutation, in-place for diagonal.  `perm_dev[c]` = row r with M[r][c] == 1. */
int qsv_apply_dense_wide(const void *state_in_dev, void *state_out_dev, int n_local, int k,
                         const int *targets, const void *matrix_dev, void *stream);
int qsv_apply_perm_wide(const void *state_in_dev, void *state_out_dev, int n_local, int k,
                        const int *targets, const uint32_t *perm_dev, void *stream);
int qsv_apply_diag_wide(void *state_dev, int n_local, int k, const int *targets,
                        const void *diag_dev, void *stream);

/* ---- Grover: main/Grover.py:14-22 (oracle F) and :24-31,40-44 (H U0 H) -------------------- */

/* Oracle for one marked item: exchanges the two amplitudes idx0 <-> idx1 (local indices). */
int qsv_grover_oracle(void *state_dev, uint64_t idx0, uint64_t idx1, void *stream);
/* Diffusion 2|s><s| - I on the search register = all physical bits >= anc_bits, independently for
 * each value of the low `anc_bits` ancilla bits.  Three steps so that a multi-GPU caller can
 * all-reduce the sums between them:
 *   qsv_diffusion_sums   : sums_dev[2*a .. 2*a+1] = sum over the search register of psi[., a]
 *   (all-reduce sums_dev across ranks)
 *   qsv_diffusion_update : psi[i, a] = 2 * sums[a] / 2^n_search_total - psi[i, a]
 * `scratch_dev` must hold qsv_diffusion_scratch_bytes(n_local, anc_bits) bytes. */
size_t qsv_diffusion_scratch_bytes(int n_local, int anc_bits);
int qsv_diffusion_sums(const void *state_dev, int n_local, int anc_bits, void *scratch_dev,
                       void *sums_dev, void *stream);
int qsv_diffusion_update(void *state_dev, int n_local, int anc_bits, const void *sums_dev,
                         int n_search_total, void *stream);

/* ---- Shor: modular-exponentiation involution, main/Shor.py:54-83 -------------------------- */

/* For every x (bits at x_pos[0..nx), x_pos[0] = MSB) with f = a^x mod N:
 *   (x, y=0) <-> (x, y=f)   and   (x, y=2^h-1) <-> (x, y=(2^h-1) & ~f),   y bits at y_pos[0..h).
 * x bits at positions >= n_local are taken from rank_bits. */
int qsv_modexp_involution(void *state_dev, int n_local, int nx, const int *x_pos, int h,
                          const int *y_pos, uint64_t a, uint64_t N, uint64_t rank_bits,
                          void *stream);

/* ---- measurement: main/circuit.py:329-337 (relabel + |psi|^2) and :360-364 (sampling) ----- */

/* out_dev[z] = |psi[y(z)]|^2 where bit b of z is taken from physical bit src_pos[b] of y
 * (src_pos == NULL: identity).  Output relabel of main/circuit.py:329-332 folded into the read. */
int qsv_probs(const void *state_dev, int n_local, const int *src_pos, void *out_dev, void *stream);
/* sums_dev[j] = sum of |psi|^2 over output indices z in [j*2^block_bits, (j+1)*2^block_bits). */
int qsv_block_sums(const void *state_dev, int n_local, const int *src_pos, int block_bits,
                   void *sums_dev, void *stream);
/* ---- exact sampling on large and sharded registers (csrc/qsv_sample.cu) -----------------------------------------
 * random.choices (main/circuit.py:360-364) picks index = bisect_right(accumulate(weights), u * total, 0, N - 1).  Above
 * 2^20 outcomes, and on every sharded register, that rule is evaluated in OUTPUT order with exact 128-bit fixed-point
 * sums (2^-96 units; integer addition is associative, so the outcome is the same for 1, 2, 4 or 8 shards) by a radix
 * descent over the output index: per level qsv_sample_level accumulates 2^n_bits bucket sums over the elements still
 * compatible with the bits chosen so far, the caller all-reduces the int64 bucket arrays over the ranks, and
 * qsv_sample_select picks the bucket holding the target; qsv_sample_leaf + qsv_sample_select resolve the last <= 12
 * bits.  Only the 64-byte state below returns to the host. */
#define QSV_SAMPLE_MAX_LEVEL_BITS 12
typedef struct {
    uint64_t fixed_val;     /* physical index bits (global ones included) pinned by the levels so far          */
    uint64_t z;             /* output index bits chosen so far: the outcome once every level has run            */
    uint64_t target[2];     /* remaining target weight, fixed point 2^-96 (lo, hi)                               */
    uint64_t total[2];      /* total weight of the state (set by the first select)                               */
    uint64_t status;        /* 0 = ok, 1 = the total weight is zero (random.choices raises ValueError)           */
    uint64_t levels_done;
} qsv_sample_state_t;
/* Zero the device-side state before a new draw. */
int qsv_sample_begin(void *sel_dev, void *stream);
/* Bucket sums of one level.  `fixed_mask`: physical positions (local and pre-shifted global) pinned by earlier levels;
 * `bucket_pos[k]`: physical position feeding bucket bit k (k = 0 least significant; at most two of them among the ten
 * lowest free local positions - split the level otherwise); `buckets_dev`: int64[4 << n_bits], zeroed here, four
 * 32-bit limbs per bucket held in 64-bit words (sums of up to 2^32 ranks/CTAs cannot overflow). */
int qsv_sample_level(const void *state_dev, int n_local, uint64_t rank_bits, uint64_t fixed_mask, int n_bits,
                     const int *bucket_pos, const void *sel_dev, void *buckets_dev, void *stream);
/* The last bits: every remaining element is stored at its own slot of `buckets_dev` (same layout). */
int qsv_sample_leaf(const void *state_dev, int n_local, uint64_t rank_bits, uint64_t fixed_mask, int n_bits,
                    const int *leaf_pos, const void *sel_dev, void *buckets_dev, void *stream);
/* Pick the bucket that holds the target and update *sel_dev.  first != 0: also computes the total and the target
 * floor(total * u) for u = u_mant * 2^-u_shift (u_mant < 2^53: math.frexp of the Python float, exact). */
int qsv_sample_select(const void *buckets_dev, int n_bits, const int *bucket_pos, const int *out_bits, int first,
                      uint64_t u_mant, int u_shift, void *sel_dev, void *stream);
/* Registers of up to 2^20 outcomes on one GPU: the sequential float scan (bit-exact with CPython's accumulate +
 * bisect) in ONE launch - total, then target = u * total, then the search.  result_dev: {uint64 index, double total}. */
int qsv_sample_sequential(const void *state_dev, int n_local, const int *src_pos, double u, void *result_dev,
                          void *stream);

/* Sequential left-to-right accumulation from `prefix`, starting at output index z_start: writes the
 * first z with cum > target to result_dev[0] (uint64), clamped to z_last, and the running sum at
 * the point where the scan stopped to result_dev[1] (double) — with target = +inf that is the
 * exact sequential total.  Reproduces bisect_right(accumulate(weights), u*total, 0, N-1) of
 * random.choices (main/circuit.py:364).  result_dev holds 16 bytes. */
int qsv_sample_scan(const void *state_dev, int n_local, const int *src_pos, uint64_t z_start,
                    uint64_t z_last, double prefix, double target, void *result_dev, void *stream);

/* ---- multi-GPU re-layout helpers ---------------------------------------------------------- */

/* Pack / unpack between the state and a contiguous staging buffer for the chunked all-to-all
 * qubit swap: copies n_chunks slices of `chunk_elems` amplitudes, slice j starting at
 * state[j*stride_elems + offset_elems], to/from staging[j*chunk_elems]. */
int qsv_pack_slices(const void *state_dev, void *staging_dev, uint64_t n_chunks,
                    uint64_t stride_elems, uint64_t offset_elems, uint64_t chunk_elems, void *stream);
int qsv_unpack_slices(void *state_dev, const void *staging_dev, uint64_t n_chunks,
                      uint64_t stride_elems, uint64_t offset_elems, uint64_t chunk_elems, void *stream);

/* In-place global<->local qubit swap over NVLink PEER MEMORY (no staging, no NCCL on the data path).
 * `peer_state_ptrs[j]` is rank j's state pointer mapped into this process (CUDA VMM / symmetric memory);
 * entry [rank] is ignored.  `positions[k]` (ascending local positions; NULL = the top g) trades places with
 * global bit k.  Writing a local index as (field, rest) with field = the bits at `positions`, element
 * (field = j, rest = e) of rank r trades places with element (field = r, rest = e) of rank j.  Every pair of
 * ranks splits the work: in alternating granules one of the two ranks reads both sides and writes both sides
 * (one remote read + one remote write per amplitude), so each amplitude crosses NVLink exactly once per
 * direction and touches local HBM once.  Positions >= 9 keep every peer access a contiguous 8 KiB run.
 * The caller must barrier all ranks before and after the call. */
int qsv_swap_peer(void *state_dev, const uint64_t *peer_state_ptrs, int n_local, int g, int rank,
                  const int *positions, void *stream);
/* General form.  `n_ranks` = 2^g ranks share the register; `n_swap` <= g qubits are exchanged: positions[k]
 * (ascending local positions) trades places with global bit gbits[k] (ascending, NULL = 0..n_swap-1).  With
 * n_swap < g this is a PARTIAL swap (pairwise exchange with rank ^ 2^j for one qubit): the peers of a
 * rank are the 2^n_swap - 1 ranks that differ from it in those bits only and (1 - 2^-n_swap) of the shard moves.
 * `sp_mask` / `sp_val`: index bits (local positions and, from bit n_local up, the rank) that are FIXED in the
 * support of the state BEFORE the swap - a register that started in a basis state (main/circuit.py:290-301) is
 * zero wherever such a bit disagrees.  Pairs of which neither element can be non-zero are not moved (0 / 0: dense). */
int qsv_swap_peer_ex(void *state_dev, const uint64_t *peer_state_ptrs, int n_local, int n_ranks, int rank,
                     int n_swap, const int *positions, const int *gbits, uint64_t sp_mask, uint64_t sp_val,
                     void *stream);

/* Global<->local qubit swap FUSED with the gate pass that follows it: one kernel over NVLink peer memory
 * (SURVEY 8e's exchange step + one run of main/circuit.py:303-327's gate loop in a single sweep).  Contract of
 * qsv_swap_peer followed by qsv_run_pass_rounds on the same state (pass_first = 0), or of qsv_run_pass_rounds followed
 * by qsv_swap_peer (pass_first = 1: the ops see the rank and field bits the tile has BEFORE the swap, i.e. on the rank
 * it is pulled from; pass->rank_bits is ignored), with these differences: the pass's tile must not
 * hold any of `positions` (which must be >= pass->low_bits); every tile is TMA-loaded from the rank that holds it
 * before the swap (pull), run through the rounds, and stored into the local shard after the rank pair has
 * hand-shaken on that tile through `peer_flag_ptrs` (per rank an array of 16 * 1024 uint64 in peer-mapped memory,
 * zero-initialised once; entry [rank] is the caller's own array).  `epoch` must be the same on all ranks and grow
 * by at least 2^32 from one call to the next on the same flag arrays.  All ranks must have finished their earlier
 * kernels on the state before the call (barrier); no barrier is needed after it.  Specialised kernel only
 * (tile_bits 12); the grid is limited to the CTAs that are resident at once.
 *   qsv_jit_source_swap / qsv_jit_compile_swap: generator access as for qsv_jit_source / qsv_jit_compile; the host
 *   form of the source exports qsv_jit_host_run_swap(out_shard, shards_before_swap[], rank, rank_bits, n_tiles, tables). */
int qsv_run_pass_rounds_swap(void *state_dev, const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                             const void *tables_dev, const uint64_t *peer_state_ptrs, const uint64_t *peer_flag_ptrs,
                             int g, const int *positions, int pass_first, int rank, uint64_t epoch, void *stream);
int qsv_jit_source_swap(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int g,
                        const int *positions, int pass_first, char *buf, size_t buf_len, size_t *needed);
int qsv_jit_compile_swap(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int g,
                         const int *positions, int pass_first, char *cubin_buf, size_t buf_len, size_t *needed);
/* General forms: partial swaps (`n_swap` of the log2(n_ranks) global qubits, positions[k] <-> global bit gbits[k],
 * NULL = 0..n_swap-1) and the support hint of qsv_swap_peer_ex (`sp_mask` / `sp_val` over the index bits before
 * the swap, the pass's tile bits excluded): tile pairs that hold only zeros on both ranks are neither moved nor
 * computed.  The kernel is launched COOPERATIVELY (all CTAs co-resident or the launch fails) and its flag spin is
 * bounded: after QSV_FUSE_WATCHDOG_MS (default 30000) without an answer from the partner the kernel traps and the
 * launch reports an error instead of hanging the node.  The host form's qsv_jit_host_run_swap takes
 * (..., tables, sp_mask, sp_val). */
int qsv_run_pass_rounds_swap_ex(void *state_dev, const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                                const void *tables_dev, const uint64_t *peer_state_ptrs,
                                const uint64_t *peer_flag_ptrs, int n_ranks, int n_swap, const int *positions,
                                const int *gbits, int pass_first, int rank, uint64_t epoch, uint64_t sp_mask,
                                uint64_t sp_val, void *stream);
/* Zero-fill form (sharded counterpart of qsv_run_pass_rounds_fresh): the register was NOT cleared before the run
 * (qsv_init_basis_sparse wrote the basis amplitude only), so whatever lies outside the support of the state is
 * uninitialised.  A tile pulled from outside the support (sp_mask / sp_val) counts as zeros, a tile inside it keeps only
 * the amplitudes whose tile-local index agrees with the fixed bits inside the tile (`zf_mask` / `zf_val`, 12 tile-local
 * bits); every tile of an active pair is written in full.  The host form's qsv_jit_host_run_swap takes
 * (..., sp_mask, sp_val, zf_mask, zf_val, zf_on). */
int qsv_run_pass_rounds_swap_fresh(void *state_dev, const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes,
                                   const void *tables_dev, const uint64_t *peer_state_ptrs,
                                   const uint64_t *peer_flag_ptrs, int n_ranks, int n_swap, const int *positions,
                                   const int *gbits, int pass_first, int rank, uint64_t epoch, uint64_t sp_mask,
                                   uint64_t sp_val, uint32_t zf_mask, uint32_t zf_val, void *stream);
int qsv_jit_source_swap_ex(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int n_swap,
                           const int *positions, const int *gbits, int pass_first, char *buf, size_t buf_len,
                           size_t *needed);
int qsv_jit_compile_swap_ex(const qsv_pass_t *pass, const void *prog_host, uint32_t prog_bytes, int n_swap,
                            const int *positions, const int *gbits, int pass_first, char *cubin_buf, size_t buf_len,
                            size_t *needed);

#ifdef __cplusplus
}
#endif
#endif /* QSV_H */
