// ORACLE support (test infrastructure): thin C wrapper that drives the REFERENCE's own vendored alpha-expansion
// (GCoptimization.cpp / LinkedBlockList.cpp, which pull in graph.cpp / maxflow.cpp, under
// /root/reference/external/progressive-x/graph-cut-ransac/src/pygcransac/include, compiled from where they lie via -I;
// no reference source is copied into this repository) exactly the way PEARL::labeling does
// (/root/reference/external/progressive-x/src/pyprogressivex/include/PEARL.h:461-536): functor data and smooth costs,
// setLabelCost(cost), setNeighbors(i, j) per neighbour listing, optional initial labeling, expansion(iter, 1000).
// Built into oracle/_ref/libref_gco.so by oracle/Makefile; validates oracle/posefit.cpp's AlphaExpansion.
#include "GCoptimization.h"
#include "GCoptimization.cpp"
#include "LinkedBlockList.cpp"

namespace {
struct Info { const double* D; int L; double lambda; };
double data_fn(int p, int l, void* info) { const Info* I = (const Info*)info; return I->D[(size_t)p * I->L + l]; }
double smooth_fn(int, int, int l1, int l2, void* info) { return l1 != l2 ? ((const Info*)info)->lambda : 0.0; }
}  // namespace

extern "C" double ref_gco_expansion(int n, int L, const double* D, const int* nbr_offsets, const int* nbr_index,
                                    double lambda, double label_cost, int* labels, int use_initial, int max_iterations) {
  GCoptimizationGeneralGraph* gc = new GCoptimizationGeneralGraph(n, L);
  Info info = {D, L, lambda};
  gc->setDataCost(&data_fn, &info);
  if (lambda > 0.0) gc->setSmoothCost(&smooth_fn, &info);
  if (label_cost > 0.0) gc->setLabelCost(label_cost);
  for (int i = 0; i < n; ++i)
    for (int a = nbr_offsets[i]; a < nbr_offsets[i + 1]; ++a)
      if (nbr_index[a] != i) gc->setNeighbors(i, nbr_index[a]);
  if (use_initial)
    for (int i = 0; i < n; ++i) gc->setLabel(i, labels[i]);
  int iters = 0;
  double e = gc->expansion(iters, max_iterations);
  for (int i = 0; i < n; ++i) labels[i] = gc->whatLabel(i);
  delete gc;
  return e;
}
