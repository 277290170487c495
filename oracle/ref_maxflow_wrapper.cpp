// ORACLE support (test infrastructure): thin C wrapper that drives the REFERENCE's own vendored
// Boykov-Kolmogorov max-flow (energy.h / graph.h / graph.cpp / maxflow.cpp under
// /root/reference/external/progressive-x/graph-cut-ransac/src/pygcransac/include, compiled from where they
// lie via -I; no reference source is copied into this repository) exactly the way GCRANSAC.h:812-920 does:
// add_node x N, add_term1 per point, add_term2 per neighbourhood edge, minimize(), what_segment() == SINK.
// Built into oracle/_ref/libref_maxflow.so by oracle/Makefile; used to validate oracle/posefit.cpp's labeling.
#include "energy.h"
#include "graph.cpp"
#include "maxflow.cpp"

extern "C" int ref_bk_labeling(int N, const double* u0, const double* u1, int E, const int* ex, const int* ey,
                               const double* e00, const double* e01, const double* e10, const double* e11,
                               int* labels, double* energy_out) {
  typedef Energy<double, double, double> EnergyT;
  EnergyT* g = new EnergyT(N, E > 0 ? E : 1, NULL);
  for (int i = 0; i < N; ++i) g->add_node();
  for (int i = 0; i < N; ++i) g->add_term1(i, u0[i], u1[i]);
  for (int e = 0; e < E; ++e) g->add_term2(ex[e], ey[e], e00[e], e01[e], e10[e], e11[e]);
  double en = g->minimize();
  if (energy_out) *energy_out = en;
  int n = 0;
  for (int i = 0; i < N; ++i) {
    labels[i] = g->what_segment(i) == Graph<double, double, double>::SINK ? 1 : 0;
    n += labels[i];
  }
  delete g;
  return n;
}
