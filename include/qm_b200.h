/*
 * qm_b200.h — C ABI of the B200-native state-vector engine that sits beneath
 * Quromorphic.jl's src/Quantum simulator API.
 *
 * Every entry point is `extern "C"`, takes plain pointers and sizes, and returns
 * an int32 status (0 = QM_OK).  Nothing here names a torch/CUDA type: the Julia
 * host binds these with `ccall` (see INTEGRATION.md and
 * quromorphic.jl_b200/julia/B200QSim.jl), the Python harness with ctypes.
 *
 * Conventions at this boundary:
 *   - qubit arguments are 0-based BIT positions of the amplitude index
 *     (reference qubit t  <->  bit t-1, QSim.jl:56-58, :304).  The host shim
 *     converts t -> t-1 and applies the reference's non-adjacent-control index
 *     map (QSim.jl:181-197) before calling; the kernels are textbook.
 *   - a 2x2 matrix is 8 doubles in Julia's Matrix{ComplexF64} memory order
 *     (column-major, interleaved):  re U11, im U11, re U21, im U21,
 *                                   re U12, im U12, re U22, im U22.
 *   - amplitudes are interleaved (re, im) doubles, amplitude i at doubles
 *     [2i, 2i+1] — exactly Vector{ComplexF64} memory (QSim.jl:33-37).
 *   - host buffers belong to the caller and are never retained after return.
 *   - gate calls enqueue (they are fused into HBM passes at the next flush);
 *     readouts, download and qm_sync flush and synchronise.
 *   - a handle is not thread-safe; distinct handles are independent.
 */
#ifndef QM_B200_H
#define QM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QM_ABI_VERSION 4

/* status codes */
enum {
    QM_OK              = 0,
    QM_ERR_ARG         = 1,  /* bad argument (range, null, c == t, ...)        */
    QM_ERR_CUDA        = 2,  /* CUDA runtime/driver error (see qm_last_error) */
    QM_ERR_NOMEM       = 3,  /* device or host allocation failed              */
    QM_ERR_NO_DEVICE   = 4,  /* no CUDA device: there is NO CPU fallback      */
    QM_ERR_STATE       = 5,  /* call not valid for this handle's mode         */
    QM_ERR_COMM        = 6   /* NCCL / peer-memory failure                    */
};

/* gate kinds of the wire format (qm_gate.kind) */
enum {
    QM_GATE_U2      = 0,  /* 1-qubit 2x2 `m` on bit q0            — replaces u2!         QSim.jl:52-63   */
    QM_GATE_CU2     = 1,  /* control bit q0, target bit q1, `m`   — replaces controlled! QSim.jl:175-203 */
    QM_GATE_SWAP    = 2,  /* exchange bits q0 and q1              — replaces swap!       QSim.jl:262-275 */
    QM_GATE_U2_SLOT = 3   /* batched only: 2x2 on bit q0 taken per state from slot q1 of `per_state_mats` */
};

/* measurement bases for qm_measure */
enum { QM_BASIS_Z = 0, QM_BASIS_X = 1, QM_BASIS_Y = 2 };

/* qm_set_option keys */
enum {
    QM_OPT_EAGER      = 0,  /* 1: run every gate call at once with the unfused kernels (default 0: queue + fuse) */
    QM_OPT_TILE_BITS  = 1,  /* fused pass tile size m: 11 or 12 for the TMA kernels; default 0 = automatic (12) */
    QM_OPT_LOW_BITS   = 2,  /* contiguous low bits every tile keeps (default 5: 512-byte rows; sharded registers 4)    */
    QM_OPT_FUSE       = 3,  /* 0: planner emits one pass per gate (debug), 1: fuse (default)                      */
    QM_OPT_TRACE      = 4,  /* 1: CTA 0 of every fused pass records SM-clock stamps per tile (qm_trace_read; diagnostics) */
    QM_OPT_PIPELINE   = 5,  /* 0: one CTA per tile kernel, 1: persistent TMA-pipelined kernel                     */
    QM_OPT_MAX_COST   = 6,  /* planner: budget of a pass = gate costs (FP64 FMAs per amplitude + 1) + 18 per register-block group (planner.hpp gate_cost); default 80, 0 = no cap */
    QM_OPT_TIME_PASSES = 7, /* 1: CUDA events around every fused pass, read back with qm_pass_times (diagnostics) */
    QM_OPT_OVERLAP   = 9,   /* sharded registers: 1 (default) run the global<->local exchange slab by slab beside the passes before and after it, 0: exchange between passes */
    QM_OPT_XSM       = 10,  /* ... SMs the overlapped exchange kernel claims (default 40: ~17 GB/s per SM per direction, the passes beside it use the rest) */
    QM_OPT_XUN       = 11,  /* ... amplitude pairs in flight per exchange thread: 4 (1024 threads per SM) or 8 (512, default) */
    QM_OPT_OVERLAP_DEPTH = 13, /* ... how many of the last passes before an exchange run slab by slab in front of it (1..4, default 2) */
    QM_OPT_SLAB_BITS = 12,  /* ... log2 of the number of slabs (1..3, default 3) */
    QM_OPT_DEFER_PASS = 15, /* sharded registers: 1 (default) the last fused pass planned before an exchange is postponed to after it when it does not target an outgoing qubit and the passes after the exchange do not become more (a reservoir layer: 4 passes per step instead of 5), 0: never */
    QM_OPT_ZFOLD     = 16,  /* 1 (default): qm_apply_circuit_expect_z folds the all-qubit <Z> sums into the last fused pass of the circuit when that pass runs as a compiled kernel over one whole register (the state is not read again by the readout kernels); 0: the circuit, then the ordinary readout */
    QM_OPT_JIT       = 14,  /* compiled pass kernels: a fused pass whose opcode sequence (gate kinds, sign variants, register-block positions) has been seen this many times is compiled into straight-line code and launched instead of the interpreter; default 2 (the second time a circuit of the same structure runs; only on registers of 2^26 amplitudes and more, batch included — below that a pass takes microseconds and a 0.25 s compilation never pays), 1 = every pass at first sight whatever the size, 0 = never.  The planner prices gates for the kernel that will run them. */
    QM_OPT_SHORT_PCT = 8    /* planner: a circuit shorter than the look-ahead window (a reservoir step before a readout) is planned at the multiple of the budget, up to this many per cent, that the pass model says finishes first; default 0 = no limit, 100 = off */
};

typedef struct qm_state qm_state;

/* One gate of a circuit.  80 bytes, no padding surprises: 4 x int32 + 8 x double. */
typedef struct qm_gate {
    int32_t kind;   /* QM_GATE_*                                   */
    int32_t q0;     /* target (U2), control (CU2), first bit (SWAP) */
    int32_t q1;     /* target (CU2), second bit (SWAP), slot (U2_SLOT) */
    int32_t flags;  /* reserved, must be 0                          */
    double  m[8];   /* 2x2 matrix, Julia memory order (unused for SWAP/U2_SLOT) */
} qm_gate;

/* Counters since creation (or the last qm_counters_reset). */
typedef struct qm_counters_t {
    uint64_t gates;          /* gates accepted                                  */
    uint64_t passes;         /* fused HBM passes launched                       */
    uint64_t launches;       /* kernels launched (all kinds)                    */
    uint64_t bytes_hbm;      /* algorithmic HBM bytes of those passes/readouts  */
    uint64_t bytes_link;     /* bytes sent over NVLink by remaps (this rank)    */
    uint64_t bytes_h2d;      /* host->device bytes copied (pass programs, slots, uploads) */
    uint64_t bytes_d2h;      /* device->host bytes copied (readouts, downloads)  */
    double   ms_device;      /* CUDA-event time of the last flushed circuit     */
    double   ms_plan;        /* host planning time of the last flushed circuit  */
    double   ms_remap;       /* CUDA-event time spent in global-qubit remaps (all-to-all) since reset */
    uint64_t overlapped_remaps; /* remaps that ran slab by slab beside the fused passes before and after them */
    uint64_t overlapped_passes; /* fused passes that were cut into slabs to run beside an exchange */
    uint64_t compiled_passes;   /* fused passes that ran as compiled straight-line kernels (QM_OPT_JIT) instead of the interpreter */
    uint64_t deferred_passes;   /* sharded: passes postponed from before an exchange to after it (QM_OPT_DEFER_PASS) */
    uint64_t folded_readouts;   /* qm_apply_circuit_expect_z calls whose <Z> sums were taken inside the last pass (QM_OPT_ZFOLD); ABI 4 */
} qm_counters_t;

/* ---- lifetime -------------------------------------------------------------------- */

/* statevector(n, m) for `batch` independent registers on CUDA device `device`
 * (QSim.jl:33-37: zeros, then amplitude `basis_index` = 1).  batch >= 1. */
int32_t qm_create(int32_t n_qubits, uint64_t basis_index, int32_t batch, int32_t device,
                  qm_state** out);
int32_t qm_destroy(qm_state* st);
/* back to |basis_index> for every register of the batch */
int32_t qm_reset(qm_state* st, uint64_t basis_index);
int32_t qm_set_option(qm_state* st, int32_t key, int64_t value);
int32_t qm_num_qubits(const qm_state* st, int32_t* n_qubits, int32_t* batch);

/* ---- host <-> device state (B200State(v) / Vector{ComplexF64}(st)) ----------------- */
int32_t qm_upload(qm_state* st, const double* host_interleaved, uint64_t n_amps, int32_t batch_idx);
int32_t qm_download(qm_state* st, double* host_interleaved, uint64_t n_amps, int32_t batch_idx);

/* ---- gates ----------------------------------------------------------------------------
 * qm_apply_u2          replaces u2!          QSim.jl:52-63  (and h!,x!,y!,z!,rx!,ry!,rz!  :81-172)
 * qm_apply_controlled  replaces controlled!  QSim.jl:175-203 (and cnot!,crx!,cry!,crz!    :205-260)
 * qm_apply_swap        replaces swap!        QSim.jl:262-275
 * qm_apply_circuit     a gate list, fused into as few HBM passes as the planner finds
 * qm_apply_batched     a gate list applied to every register of the batch; QM_GATE_U2_SLOT
 *                      gates take their matrix from per_state_mats[b][slot][8] (host, doubles)
 */
/* compiled pass kernels ahead of time: plans `gates` as qm_apply_circuit would, compiles the kernels of its passes (QM_OPT_JIT)
 * and keeps the plan, without touching the register — the first application then already runs compiled.  *n_compiled
 * (optional) = kernels compiled by this call.  No-op when compiled kernels are not in use for this register. */
int32_t qm_precompile_circuit(qm_state* st, const qm_gate* gates, int32_t ngates, int32_t* n_compiled);
/* a reservoir step in one call: qm_apply_circuit(gates) followed by qm_expect_z_all(threshold, out) — same results, same
 * layout of `out` ([n][2] doubles; batch x n x 2 for a batch).  When the last fused pass of the circuit runs as a compiled
 * kernel over one whole register (QM_OPT_JIT, batch 1), the thresholded |a|^2 sums of measure_z (QSim.jl:297-318, every qubit)
 * are taken from the tiles of that pass while they are still in shared memory, so the 16 * 2^n byte read of the readout kernels
 * disappears (QM_OPT_ZFOLD, counters.folded_readouts); otherwise the two calls are made one after the other. */
int32_t qm_apply_circuit_expect_z(qm_state* st, const qm_gate* gates, int32_t ngates, double threshold, double* out);
int32_t qm_apply_u2(qm_state* st, int32_t t0, const double U[8]);
int32_t qm_apply_controlled(qm_state* st, int32_t c0, int32_t t0, const double V[8]);
int32_t qm_apply_swap(qm_state* st, int32_t a0, int32_t b0);
int32_t qm_apply_circuit(qm_state* st, const qm_gate* gates, int32_t ngates);
int32_t qm_apply_batched(qm_state* st, const qm_gate* gates, int32_t ngates,
                         const double* per_state_mats, int32_t nslots);

/* ---- readouts -------------------------------------------------------------------------
 * qm_probs          replaces mp               QSim.jl:293  (out: 2^n doubles per register, host)
 * qm_measure        replaces measure_z/x/y    QSim.jl:297-330: (p0,p1) of bit t0 after the basis
 *                   rotation, summing only probabilities > threshold (reference: 1e-10, strict;
 *                   threshold < 0 disables the test).  batch_idx selects the register.
 * qm_expect_z_all   all-qubit measure_z in ONE pass: out[2q] = p0, out[2q+1] = p1 for q = 0..n-1,
 *                   for every register of the batch (out: batch * 2n doubles, host)
 * qm_expect_all     same for basis X or Y (measure_x / measure_y of every qubit, QSim.jl:320-330): the state is read
 *                   once per 12 (then 9) qubits, every pair rotated in registers; qm_expect_all_with takes the
 *                   caller's own rotation matrix like qm_measure_with (R == NULL: the Z basis)
 * qm_reduced_density  statevector_to_density_matrix + partial_trace, QSim.jl:39-41 and
 *                   QInfo.jl:381-405, without ever forming rho: keep_low=1 keeps B = the low k
 *                   qubits (partial_trace(..., 1)), keep_low=0 keeps A = the high k qubits
 *                   (partial_trace(..., 2)).  out: 2^k x 2^k ComplexF64, column-major interleaved.
 * qm_sample         nshots basis-state indices by inverse CDF over abs2(psi) with the
 *                   counter-based generator documented in DESIGN.md (bit-exact vs the oracle)
 * qm_norm2          sum of abs2 (no threshold)
 */
int32_t qm_probs(qm_state* st, double* out_host, int32_t batch_idx);
int32_t qm_measure(qm_state* st, int32_t basis, int32_t t0, double threshold, int32_t batch_idx,
                   double out[2]);
/* qm_measure with the caller's own rotation matrix R (Julia passes H or rx_gate(pi/2) so that the
 * matrix bits are the host's; R == NULL means no rotation, i.e. the Z basis) */
int32_t qm_measure_with(qm_state* st, int32_t t0, const double* R, double threshold, int32_t batch_idx,
                        double out[2]);
int32_t qm_expect_z_all(qm_state* st, double threshold, double* out_host);
/* the same readout in two halves, for loops whose host side prepares the next step while the device still runs this one:
 * _begin enqueues the readout of the state as it is at this point of the stream and returns at once (gate calls may
 * follow immediately), _end waits for the oldest pending one and copies it out.  At most two pending; unsharded registers. */
int32_t qm_expect_z_all_begin(qm_state* st, double threshold);
int32_t qm_expect_z_all_end(qm_state* st, double* out_host);
int32_t qm_expect_all(qm_state* st, int32_t basis, double threshold, double* out_host);
int32_t qm_expect_all_with(qm_state* st, const double* R, double threshold, double* out_host);
int32_t qm_reduced_density(qm_state* st, int32_t keep_low, int32_t k, int32_t batch_idx,
                           double* out_colmajor);
int32_t qm_sample(qm_state* st, uint64_t seed, int32_t nshots, int32_t batch_idx, uint64_t* out);
int32_t qm_norm2(qm_state* st, int32_t batch_idx, double* out);

/* ---- QInfo measures for MANY small matrices, on the device ---------------------------
 * The reference evaluates every measure from a host eigen-decomposition of one 2^k x 2^k matrix (QInfo.jl:24-28);
 * a batched reservoir has thousands per step.  One warp diagonalises one matrix (cyclic complex Jacobi, d <= 16).
 * qm_reduced_density_batch  qm_reduced_density for every register of the batch: out [batch][4^k] ComplexF64, column-major
 * qm_entropy_batch          entropy of the reduced matrix of every register (k <= 4): kind 0 von_neumann_entropy
 *                           (QInfo.jl:47-50, eigenvalues clamped at 1e-10 inside the log as :26), 1 renyi_entropy(alpha)
 *                           (:199-204), 2 min_entropy (:353-357), 3 max_entropy (:332-337); out [batch] doubles
 * qm_herm_measure_batch     the same measures for `count` host matrices (d x d, column-major), plus 4 trace_distance of
 *                           (A, B) (:311-316), 5 fidelity(A, B) (:289-294), 6 ascending eigenvalues (out [count][d])
 */
int32_t qm_reduced_density_batch(qm_state* st, int32_t keep_low, int32_t k, double* out_colmajor);
int32_t qm_entropy_batch(qm_state* st, int32_t keep_low, int32_t k, int32_t kind, double alpha, double* out_host);
int32_t qm_herm_measure_batch(int32_t device, int32_t kind, double alpha, const double* A, const double* B,
                              int32_t d, int32_t count, double* out_host);

/* ---- density matrices (the Matrix{ComplexF64} methods of QSim.jl) ----------------------------
 * rho on q qubits is a register of 2q qubits holding vec(rho) in Julia's column-major Matrix memory:
 * amplitude r + c 2^q = rho[r, c].  rho <- op rho op' (u2!  QSim.jl:67-79, controlled!  :209-232) is then
 * qm_apply_u2(t0, U) + qm_apply_u2(q + t0, conj(U)) (same for qm_apply_controlled): the host shim issues
 * both, the planner fuses them like any other gates.  density_matrix(q, m) = qm_create(2q, m (2^q + 1)).
 * qm_dm_from_state  replaces statevector_to_density_matrix  QSim.jl:39-41  (rho[r,c] = s[r] conj(s[c]))
 * qm_dm_diag        replaces mp(rho) = real.(diag(rho))      QSim.jl:295   (out: 2^q doubles, host)
 * qm_dm_measure     replaces measure_z / measure_x / measure_y of rho  QSim.jl:332-365: diagonal of R rho R'
 *                   summed by bit t0, entries <= threshold dropped; R == NULL is the Z basis
 */
int32_t qm_dm_from_state(qm_state* sv, qm_state* rho);
int32_t qm_dm_diag(qm_state* rho, double* out_host);
int32_t qm_dm_measure(qm_state* rho, int32_t t0, const double* R, double threshold, double out[2]);

/* ---- sharded registers (one process per GPU; high-order qubits global) -----------------
 * qm_dist_unique_id  fills a 128-byte NCCL unique id (rank 0 calls it, the host broadcasts it)
 * qm_create_dist     rank `rank` of `world` (power of two) holds the amplitudes whose top
 *                    log2(world) index bits equal `rank`; n_qubits is the GLOBAL register size.
 * Gates, readouts and counters then work on the global register; global-qubit gates are
 * remapped by an in-place chunked all-to-all over NVLink.
 * Collective calls (every rank makes them, in the same order, and gets the same answer):
 * qm_expect_z_all, qm_norm2, qm_measure*, qm_reduced_density (keep_low = 0 needs k <= n/2),
 * qm_sample (>= 12 local bits; the shots equal those of an unsharded register bit for bit).
 * qm_upload / qm_download move this rank's contiguous slice.
 */
int32_t qm_dist_unique_id(uint8_t id_out[128]);
int32_t qm_create_dist(int32_t n_qubits, uint64_t basis_index, int32_t device, int32_t rank,
                       int32_t world, const uint8_t nccl_id[128], qm_state** out);

/* Optional, after qm_create_dist: map the other ranks' shards into this process (CUDA IPC).  With peer memory the
 * global<->local exchange is a kernel of loads/stores over NVLink (no staging), ranks are ordered by flags in each
 * other's memory that the consumer's STREAM waits on (no NCCL kernel, no SM), the exchange runs slab by slab beside the
 * fused passes before and after it, and the readouts' small collectives go through mailboxes in rank order.  Every rank
 * exports 64 bytes, the host all-gathers them (world x 64 bytes, rank-major) and every rank imports.  If the import
 * fails the engine keeps NCCL (grouped send/recv through a staging ring, ncclAllReduce for the readouts). */
int32_t qm_dist_ipc_export(qm_state* st, uint8_t handle_out[64]);
int32_t qm_dist_ipc_import(qm_state* st, const uint8_t* handles);

/* All ranks of a sharded register inside ONE process (one host thread per rank, one device or several with peer
 * access): no NCCL, no IPC — the ranks see each other through plain pointers.  `key` names the world; every rank calls
 * qm_create_dist_local with the same key, and once all of them have returned (the caller's own barrier) every rank
 * calls qm_dist_local_connect.  This is how the sharded kernels, fences and collectives are tested on a 1-GPU box. */
int32_t qm_create_dist_local(int32_t n_qubits, uint64_t basis_index, int32_t device, int32_t rank,
                             int32_t world, uint64_t key, qm_state** out);
int32_t qm_dist_local_connect(qm_state* st);

/* ---- host-only planner introspection (works without a GPU; used by the CPU tests) ------
 * Plans `gates` for an n-qubit register with the given tile/low-bit sizes and `global_bits`
 * top qubits sharded; writes a flat int32 description of the passes into `out` (layout in
 * DESIGN.md §planner) and the number of int32 written into *out_len.  Returns QM_ERR_ARG
 * if cap is too small (then *out_len = needed). */
/* max_cost: cap on a pass's arithmetic (0 = none); pipeline != 0: the limits of the persistent TMA kernel. */
int32_t qm_plan_describe_ex(int32_t n_qubits, int32_t global_bits, int32_t tile_bits, int32_t low_bits,
                            double max_cost, int32_t pipeline, const qm_gate* gates, int32_t ngates,
                            int32_t* out, int64_t cap, int64_t* out_len);

/* The same plan as a flat gate list in execution order (wire format on physical bits, after
 * merging): applying `out` one gate at a time must equal applying `gates` in their given order.
 * pass_of[i] receives the pass index of out[i] (may be NULL). */
int32_t qm_plan_gates(int32_t n_qubits, int32_t global_bits, int32_t tile_bits, int32_t low_bits,
                      double max_cost, const qm_gate* gates, int32_t ngates, qm_gate* out,
                      int32_t* pass_of, int32_t cap, int32_t* out_len);

/* host-only introspection: record counts of the encoded pass programs by family:
 * hist[0..8] = ROT, ROTX, PHASE, PERMX, THR, SLOT, CTRL records, register-block groups, passes */
int32_t qm_plan_ophist(int32_t n_qubits, int32_t tile_bits, int32_t low_bits, double max_cost,
                       const qm_gate* gates, int32_t ngates, int64_t hist[9]);

/* ---- misc ---------------------------------------------------------------------------- */
int32_t qm_sync(qm_state* st);
int32_t qm_last_error(char* buf, int32_t cap);
int32_t qm_counters(qm_state* st, qm_counters_t* out);
int32_t qm_counters_reset(qm_state* st);
int32_t qm_abi_version(void);
/* compiled pass kernels, process-wide: out = {compiled, failed, kernels in the cache, backend (1 nvJitLink, 2 driver PTX
 * compiler, 0 none yet)}; *compile_ms = wall time spent compiling so far */
int32_t qm_jit_stats(uint64_t out[4], double* compile_ms);
/* optional disk cache of the compiled pass kernels, process-wide: a directory (created if missing; NULL or "" = none, the
 * default unless the environment variable QM_JIT_CACHE_DIR names one) that keeps every assembled kernel as a file keyed by a
 * hash of the complete PTX text and of the assembler's version, so that a later process running a circuit of the same
 * structure loads its kernels instead of assembling them again.  Processes may share the directory (files are renamed into
 * place); a file that does not verify is ignored.  The checks guard against truncation and stale files, not against an
 * adversary: the directory holds code the GPU will run, so give it the permissions of the library itself.  qm_jit_cache_stats: out = {kernels read from the directory, kernels
 * written to it} since the process started. */
int32_t qm_jit_cache_dir(const char* path);
int32_t qm_jit_cache_stats(uint64_t out[2]);
/* host-only check of the compiled pass kernels (no GPU needed): every 2^12-tile pass of the planned circuit is turned into PTX
 * text and, with bit 0 of `compile` set and libnvJitLink present, assembled for sm_100a (bit 1: the ZFOLD flavour of the kernels,
 * qm_apply_circuit_expect_z).  out = {passes, passes with PTX text, passes
 * assembled, distinct opcode sequences, bytes of the largest cubin, hash over all passes' opcode sequences (canonical signs: it
 * depends on the structure of the circuit, not on its angles)}; ptx_out (optional) receives the text of the first pass */
int32_t qm_jit_check(int32_t n_qubits, int32_t low_bits, double max_cost, const qm_gate* gates, int32_t ngates,
                     int32_t compile, int64_t out[6], char* ptx_out, int64_t ptx_cap);
/* diagnostics (QM_OPT_TIME_PASSES): per fused pass since the last call, 8 floats {ms, groups, gates, planner cost,
 * records with full FP64 cost, with a control in the register block (half cost), xor exchanges, thread-constant phases};
 * *n_passes receives the number of passes timed, at most cap_passes are written */
int32_t qm_pass_times(qm_state* st, float* out, int32_t cap_passes, int32_t* n_passes);
/* diagnostics (QM_OPT_TRACE): 8 SM-clock stamps per tile of CTA 0 of the last fused pass: the compute team's wait
 * start, tile landed, after register-block group 0, 1, 2, arrival; the producer's "computed" seen and next load issued */
int32_t qm_trace_read(qm_state* st, uint64_t* out, int64_t cap_u64);
/* device-side view for harnesses that time on the library's stream (no ownership transfer) */
int32_t qm_device_ptr(qm_state* st, void** amps, void** cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* QM_B200_H */
