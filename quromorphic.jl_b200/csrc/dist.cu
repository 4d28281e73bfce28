// dist.cu — sharded registers (placeholder until the single-GPU path is measured)
#include "dist.hpp"

struct DistCtx { int unused; };

int dist_destroy(DistCtx* d) { delete d; return QM_OK; }
int dist_run_circuit(qm_state*, const qm_gate*, int) { return set_err(QM_ERR_STATE, "sharded path not built"); }
int dist_canonicalize(qm_state*) { return QM_OK; }
int dist_finish_z(qm_state*, const double*, const double*, double*) { return set_err(QM_ERR_STATE, "sharded path not built"); }
int dist_finish_norm(qm_state*, double, double*) { return set_err(QM_ERR_STATE, "sharded path not built"); }
int dist_measure(qm_state*, int, double, const double*, double*) { return set_err(QM_ERR_STATE, "sharded path not built"); }
int dist_reduced_density(qm_state*, int, int, double*) { return set_err(QM_ERR_STATE, "sharded path not built"); }

extern "C" int32_t qm_dist_unique_id(uint8_t*) { return set_err(QM_ERR_STATE, "sharded path not built"); }
extern "C" int32_t qm_create_dist(int32_t, uint64_t, int32_t, int32_t, int32_t, const uint8_t*, qm_state**) {
    return set_err(QM_ERR_STATE, "sharded path not built");
}
