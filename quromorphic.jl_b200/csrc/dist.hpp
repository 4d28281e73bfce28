// dist.hpp — sharded registers: one process per GPU, the top log2(world) qubits global.
#pragma once
#include "state.hpp"

int  dist_destroy(DistCtx* d);
int  dist_run_circuit(qm_state* st, const qm_gate* gates, int ngates);
int  dist_canonicalize(qm_state* st);
int  dist_finish_z(qm_state* st, const double* pairs, const double* tot, double* out_host);
int  dist_finish_norm(qm_state* st, double local_tot, double* out);
int  dist_measure(qm_state* st, int t0, double thr, const double* R, double out[2]);
int  dist_reduced_density(qm_state* st, int keep_low, int k, double* out);
