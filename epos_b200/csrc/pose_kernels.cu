// placeholder: replaced by corresp.cu / ransac.cu
#include "common.cuh"
using namespace epos;
extern "C" {
void epos_fit_params_default(epos_fit_params* p) {
  if (!p) return;
  p->threshold = 4.0; p->spatial_coherence_weight = 0.1; p->neighborhood_ball_radius = 20.0;
  p->scaling_from_millimeters = 0.1; p->min_triangle_area = 0.0; p->min_coverage = 0.5;
  p->max_iters = 400; p->min_iters = 10; p->min_iters_before_lo = 20; p->max_lo_trials = 20;
  p->max_graph_cuts = 10; p->max_lsq_iters = 10; p->max_unsuccessful = 100; p->max_neighbors = 5;
  p->apply_numerical_optimization = 1; p->reserved = 0;
}
int epos_corresp(const float*, const float*, const float*, int, int, int, int, int, const int32_t*, int, const double*,
                 const double*, double, float, float, int, int, double*, double*, float*, float*, float*, int32_t*,
                 int32_t*, int32_t*, void*, size_t, void*) { set_error("epos_corresp: not built"); return EPOS_ERR_UNSUPPORTED; }
size_t epos_corresp_workspace_bytes(int, int, int) { return 0; }
int epos_fit_poses(const double*, const double*, const int32_t*, const int32_t*, int, const double*, const uint64_t*,
                   const epos_fit_params*, double*, int32_t*, void*, size_t, void*) { set_error("epos_fit_poses: not built"); return EPOS_ERR_UNSUPPORTED; }
size_t epos_fit_workspace_bytes(int, int, const epos_fit_params*) { return 0; }
}
