// dist.hpp — sharded registers: one process per GPU, the top log2(world) qubits global.
#pragma once
#include <vector>

#include "planner.hpp"
#include "state.hpp"

int  dist_destroy(DistCtx* d);
// exchange the g global index bits with the top g local ones (in-place chunked all-to-all over NVLink),
// update st->perm, and relabel the not-yet-executed lowered gates (may be null)
int  dist_remap(qm_state* st, std::vector<qm::LGate>* gates, std::vector<char>* done);
int  dist_canonicalize(qm_state* st);
void dist_reset(qm_state* st);
int  dist_finish_z(qm_state* st, const double* pairs, const double* tot, double* out_host);
int  dist_finish_norm(qm_state* st, double local_tot, double* out);
int  dist_measure(qm_state* st, int t0, double thr, const double* R, double out[2]);
int  dist_reduced_density(qm_state* st, int keep_low, int k, double* out);
int  dist_sample(qm_state* st, uint64_t seed, int nshots, uint64_t* out);

#include "dist_bits.hpp"
