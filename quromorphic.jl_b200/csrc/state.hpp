// state.hpp — the handle behind qm_state* and the small helpers shared by qm_api.cu and dist.cu
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <utility>
#include <vector>

#include "../../include/qm_b200.h"
#include "planner.hpp"

struct DistCtx;

int set_err(int code, const char* fmt, ...);

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return set_err( \
    (e_ == cudaErrorNoDevice || e_ == cudaErrorInsufficientDriver) ? QM_ERR_NO_DEVICE : \
    (e_ == cudaErrorMemoryAllocation ? QM_ERR_NOMEM : QM_ERR_CUDA), "%s failed: %s (%s:%d)", #x, \
    cudaGetErrorString(e_), __FILE__, __LINE__); } while (0)
#define QM_TRY(x) do { int rc_ = (x); if (rc_ != QM_OK) return rc_; } while (0)

struct qm_state {
    int n = 0;            // global qubits
    int nl = 0;           // local index bits
    int g = 0;            // sharded (rank) bits
    int rank = 0, world = 1;
    int batch = 1;
    int device = 0;
    int sms = 148;
    double2* amp = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_stage = nullptr;
    bool ev_pending = false;
    // options
    bool eager = false, fuse = true;
    int pipeline = 1;           // QM_OPT_PIPELINE: 0 general kernel, 1 persistent TMA kernel
    bool overlap = true;        // sharded: run the exchange beside the passes when the plan allows (QM_OPT_OVERLAP)
    int xsm = 40, xun = 8;      // sharded, overlapped exchange: SMs the exchange kernel claims, loads in flight per thread
    int overlap_depth = 2;      // ... passes in front of an exchange that run slab by slab beside it (1..4)
    int slab_bits = 3;          // ... and log2 of the number of slabs (measured: profiles/r2f_*, r2g_*)
    bool defer_pass = true;     // QM_OPT_DEFER_PASS: sharded registers may postpone the last pass before an exchange to after it (qm_api.cu run_circuit)
    bool zfold = true;          // QM_OPT_ZFOLD: qm_apply_circuit_expect_z may take the <Z> sums inside the last pass
    bool zfold_want = false;    // ... inside that call: run_circuit folds the readout into its last pass if it can
    bool zfold_done = false;    // ... and says here whether it did (the sums are then in d_zfold)
    double zfold_thr = 0.0;
    int dry_run = 0;            // inside qm_precompile_circuit: run_circuit launches nothing; 1 = plan only (into the plan cache), 2 = replay the plan and compile its kernels
    int jit = 2;                // QM_OPT_JIT: a pass program whose opcode sequence has been seen this many times is compiled (0 = never)
    bool time_passes = false;   // QM_OPT_TIME_PASSES: CUDA events around every fused pass (diagnostics, tools/pass_model.py)
    std::vector<cudaEvent_t> pass_events;      // 2 per timed pass, resolved by qm_pass_times
    std::vector<float> pass_desc;              // per timed pass: groups, records, planner cost
    int tile_bits = 0, low_bits = 5;     // tile_bits 0 = automatic (12)
    double short_mult = 1e9;    // QM_OPT_SHORT_PCT / 100 (planner.hpp budget_for_pending); 1e9 = no limit
    double max_cost = 80.0;  // per-pass budget of the planner cost model (gate costs + 18 per register-block group, planner.hpp); 0 = unlimited
    // plans of recent circuits keyed by STRUCTURE (targets, controls, diagonal / slot flags, cost classes — everything
    // the planner looks at; not the matrix values): a reservoir loop applies the same layer with new angles every
    // step, and planning a 50-gate step costs ~1 ms on the host against ~0.1 ms of kernels at 20 qubits
    struct CachedPlan {
        uint64_t key[3] = { 0, 0, 0 };    // two independent hashes of everything the planner reads + the gate count
        std::vector<std::pair<qm::PlanPass, std::vector<int>>> passes;     // pass (stale matrices) + gate indices in group order
    };
    std::vector<CachedPlan> plan_cache;       // most recent first, <= 8 entries
    // Global phase of the register that the fused passes have NOT applied to the amplitudes (encode_v4: uncontrolled diagonal
    // gates run in D1 form, diag(d0, d1) = d0 * diag(1, d1 / d0), and d0 lands here): true amplitudes = gphase * stored ones.
    // Probabilities, expectations, reduced densities and samples do not depend on it; whoever reads amplitudes linearly
    // (download, the density-matrix readouts, qm_device_ptr) calls materialize_phase first.
    double gphase[2] = { 1.0, 0.0 };
    // queue
    std::vector<qm_gate> queue;
    // logical -> physical qubit map (identity unless a sharded remap happened)
    std::vector<int> perm;
    // device/host scratch
    unsigned char* d_blob = nullptr; size_t d_blob_cap = 0;
    unsigned char* h_blob = nullptr; size_t h_blob_cap = 0;
    double* d_scratch = nullptr; size_t d_scratch_cap = 0;   // partial sums
    double* d_zfold = nullptr; size_t d_zfold_cap = 0;       // folded readout: V4Fold | per-CTA partials | [nl][2] pairs | total
    double* h_zfold = nullptr; size_t h_zfold_cap = 0;       // pinned staging of the V4Fold block
    double* h_scratch = nullptr; size_t h_scratch_cap = 0;   // pinned
    double* d_slots = nullptr; size_t d_slots_cap = 0;
    double* h_slots = nullptr; size_t h_slots_cap = 0;       // pinned staging of the per-state coefficients, two halves used alternately
    cudaEvent_t ev_slots[2] = { nullptr, nullptr }; int slots_half = 0;
    // split readout (qm_expect_z_all_begin / _end): two pinned result buffers used as a FIFO
    double* h_zq[2] = { nullptr, nullptr }; size_t h_zq_cap[2] = { 0, 0 }; cudaEvent_t ev_zq[2] = { nullptr, nullptr };
    int zq_head = 0, zq_count = 0;
    std::vector<void*> retired_dev, retired_pinned;          // outgrown scratch buffers, released by qm_destroy (ensure_dev)
    qm_counters_t ctr{};
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> remap_events;   // resolved into ctr.ms_remap at the next sync
    DistCtx* dist = nullptr;
    unsigned long long* trace_buf = nullptr;     // QM_OPT_TRACE: phase trace of CTA 0 of the next fused pass (diagnostics)
};

// one slab of a sharded register: the amplitudes whose index bits bit[0..nbits) have the values in value_or
struct PassSlab {
    int nbits = 0;
    int bit[3] = { 0, 0, 0 };
    uint64_t value_or = 0;
    int index = 0;          // which slab (counters: a pass is counted once, with its slab 0)
    int sm_cap = 0;         // CTAs of the persistent grid (0 = all SMs)
};

int ensure_dev(void** p, size_t* cap, size_t need, std::vector<void*>* retired);
int ensure_pinned(void** p, size_t* cap, size_t need, std::vector<void*>* retired);
#define ENSURE_DEV(st, field, need) ensure_dev((void**)&(st)->field, &(st)->field##_cap, (need), &(st)->retired_dev)
#define ENSURE_PIN(st, field, need) ensure_pinned((void**)&(st)->field, &(st)->field##_cap, (need), &(st)->retired_pinned)

static inline uint64_t local_amps(const qm_state* st) { return 1ull << st->nl; }
static inline uint64_t total_amps(const qm_state* st) { return (uint64_t)st->batch << st->nl; }

// implemented in qm_api.cu, used by dist.cu
int qm_create_common(int n, uint64_t basis, int batch, int device, int rank, int world, qm_state** out);
int qm_zstats(qm_state* st, double thr, double** h_pairs, double** h_tot);
int qm_measure_phys(qm_state* st, int phys_bit, double thr, int batch_idx, const double* R, double out[2]);
// reduced density of this device's 2^nl amplitudes (kernels enqueued on st->stream, result [4^k] double2 on the device)
int qm_reduced_local(qm_state* st, int keep_low, int k, int batch_idx, double2** d_out);
// sampler levels: leaf sums of this device's amplitudes into s1_local[0 .. 2^nl/4096), and the descent over the full
// level-1 array (n1 leaves of the WHOLE register); a shot whose leaf is not in [leaf_lo, leaf_lo + 2^nl/4096) yields 0
int qm_sample_leaves(qm_state* st, int batch_idx, double* d_s1_local);
int qm_sample_descend(qm_state* st, int batch_idx, const double* d_s1, uint64_t n1, uint64_t leaf_lo, uint64_t seed, int nshots,
                      double* d_levels, uint64_t* d_out);
int qm_flush(qm_state* st);
int materialize_phase(qm_state* st);       // multiplies the pending global phase into the amplitudes (one pass; no-op when it is 1)
