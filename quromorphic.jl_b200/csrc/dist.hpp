// dist.hpp — sharded registers: one process per GPU, the top log2(world) qubits global.
#pragma once
#include <vector>

#include "planner.hpp"
#include "state.hpp"

int  dist_destroy(DistCtx* d);
bool dist_owns_stream(const qm_state* st);      // the state's stream belongs to a same-process world
// exchange the g global index bits with the top g local ones (in-place chunked all-to-all over NVLink),
// update st->perm, and relabel the not-yet-executed lowered gates (may be null)
int  dist_remap(qm_state* st, std::vector<qm::LGate>* gates, std::vector<char>* done);
int  dist_exchange_now(qm_state* st);
int  dist_canonicalize(qm_state* st);
void dist_reset(qm_state* st);
size_t dist_window_bytes();
// bookkeeping of an exchange without running one (the overlapped path runs it slab by slab): perm, pending gates, counters
void dist_relabel(qm_state* st, std::vector<qm::LGate>* gates, std::vector<char>* done);
// overlapped exchange (dist.cu): can it be used, the epoch pair of the next one, and the three per-slab steps
bool dist_can_overlap(const qm_state* st);
void dist_next_epochs(qm_state* st, uint32_t* ep_pass, uint32_t* ep_exch);
int  dist_slab_pass_done(qm_state* st, const PassSlab& slab, uint32_t ep_pass);
int  dist_slab_exchange(qm_state* st, const PassSlab& slab, uint32_t ep_pass, uint32_t ep_exch, bool first, bool last);
int  dist_slab_wait_exchanged(qm_state* st, const PassSlab& slab, uint32_t ep_exch);
int  dist_finish_z(qm_state* st, const double* pairs, const double* tot, double* out_host);
int  dist_finish_norm(qm_state* st, double local_tot, double* out);
int  dist_measure(qm_state* st, int t0, double thr, const double* R, double out[2]);
int  dist_reduced_density(qm_state* st, int keep_low, int k, double* out);
int  dist_sample(qm_state* st, uint64_t seed, int nshots, uint64_t* out);

#include "dist_bits.hpp"
