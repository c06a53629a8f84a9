// TEST INFRASTRUCTURE ONLY.  C entry into oracle/_ref/libphylomap_ref.so = the UNMODIFIED reference sources
// (/root/reference/src/phylomap.cpp + RcppExports.cpp, compiled where they lie against oracle/standin/).
//
// ref_run() builds the R objects the reference's `.Call` entry points take (src/RcppExports.cpp:11-237: an
// ape/phytools-style tree list, Q, pid, B, Omega, nen, nodelist, root, N [, prior | lefts, rights, d]) from
// the same flat structs the oracle restatement takes (oracle/phylomap_oracle.cpp: orc_tree / orc_config),
// seeds the stand-in R generator like set.seed(seed), and calls phylomap_<fn>.  The reference has no site
// axis: S must be 1.  Q and B are copied back afterwards (the bf/ks/mt/DIC drivers mutate them in place).
#include <cstdint>
#include <cstdio>
#include <cstring>

#include <RcppArmadillo.h>

using namespace Rcpp;

extern "C" {
SEXP phylomap_SPARSEmaketreelistMCMC(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistMCMC(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistMCMC_bigtree(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistEXP(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistMCMCbf(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistMCMCks(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistMCMCmt(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistMCMCksmt(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistMCMC2sDICt(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
SEXP phylomap_maketreelistMCMCksDICt(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);

// same layout as orc_tree / orc_config in oracle/phylomap_oracle.cpp
struct ref_tree {
  int32_t T, E;
  const int32_t* edge; const int32_t* nen; const int32_t* nodelist; int32_t root;
  const int64_t* maps_off; const double* maps_len; const int32_t* maps_state;
  const int32_t* states; int64_t S;
  const double* edge_length;
};
struct ref_config {
  int32_t variant, n, N, ntrees;
  double Omega;
  const double* prior; int32_t nprior;
  int32_t rng_mode; uint64_t seed; int32_t want_log;
  const int64_t* tab_off; const double* tab_u; const double* host_tab; int64_t host_tab_n;
  const double* lefts; const double* rights; const double* d;
  int64_t site_offset;
};
}

namespace {

enum Variant { PLAIN = 0, SPARSE = 1, BIGTREE = 2, BF = 3, KS = 4, MT = 5, KSMT = 6, EXPV = 7, DIC2S = 8, DICKS = 9 };

SEXP real_vec(const double* p, long n) { NumericVector v(n); for (long i = 0; i < n; i++) v[i] = p[i]; return v; }
SEXP int_vec(const int32_t* p, long n) { IntegerVector v(n); for (long i = 0; i < n; i++) v[i] = p[i]; return v; }
SEXP real_mat(const double* p, int r, int c) { NumericMatrix m(r, c); for (long i = 0; i < (long)r * c; i++) m[i] = p[i]; return m; }

// the tree as the R wrappers hand it over (R/sumstatMCMC.R:21-27: maps, mapnames, edge, Nnode, node.states, states, edge.length)
SEXP tree_list(const ref_tree& t) {
  List x;
  List maps, mapnames;
  for (int e = 0; e < t.E; e++) {
    long a = t.maps_off[e], b = t.maps_off[e + 1];
    maps.push_named("", real_vec(t.maps_len + a, b - a));
    mapnames.push_named("", int_vec(t.maps_state + a, b - a));
  }
  x.push_named("maps", maps);
  x.push_named("mapnames", mapnames);
  IntegerMatrix edge(t.E, 2);
  for (long i = 0; i < 2L * t.E; i++) edge[i] = t.edge[i];  // column-major [E][2] like R
  x.push_named("edge", edge);
  x.push_named("Nnode", wrap((int)(t.T - 1)));
  IntegerMatrix ns(t.E, 2);  // read at entry, overwritten by updatenodestates before any use (phylomap.cpp:460-475)
  x.push_named("node.states", ns);
  if (t.S == 1) x.push_named("states", int_vec(t.states, t.T));
  else {  // site axis (shim only): ntips x nsites matrix, one column per character = the [site][tip] block as it is
    IntegerMatrix sm(t.T, (int)t.S);
    for (long i = 0; i < (long)t.T * t.S; i++) sm[i] = t.states[i];
    x.push_named("states", sm);
  }
  if (t.edge_length) x.push_named("edge.length", real_vec(t.edge_length, t.E));
  return x;
}

}  // namespace

extern "C" int ref_ncols(int variant, int n) {
  int k = n / 2 - 1;
  switch (variant) {
    case PLAIN: case SPARSE: case BIGTREE: case EXPV: return n + n * (n - 1);
    case BF: case MT: return n + n * n + 3;
    case DIC2S: return n + n * n + 4;
    case DICKS: return n + n * n + 2 + 3 * k + 2;
    default: return n + n * n + 2 + 3 * k + 1;
  }
}

extern "C" int ref_run(const ref_tree* trees, const ref_config* cfg, double* Q, const double* pid, double* B,
                       double* out /* N x ncols, column-major */, char* err, int errlen) {
  try {
    const int n = cfg->n, N = cfg->N, v = cfg->variant;
#ifndef PM_SHIM_DRIVER
    for (int t = 0; t < cfg->ntrees; t++)
      if (trees[t].S != 1) throw std::runtime_error("the reference handles one character per call (S must be 1)");
#endif
    if (cfg->rng_mode != 0) throw std::runtime_error("the reference consumes R's sequential stream only (rng_mode must be SEQUENTIAL)");
    standin::set_seed((unsigned)cfg->seed);
    standin::last_error().clear();

    NumericMatrix Qm(real_mat(Q, n, n)), Bm(real_mat(B, n, n));
    SEXP pidv = real_vec(pid, n);
    SEXP Omega = wrap(cfg->Omega), Nn = wrap((int)N);
    SEXP prior = cfg->prior ? real_vec(cfg->prior, cfg->nprior) : SEXP();
    SEXP res;
    if (v == MT || v == KSMT) {
      // R/sumstatMCMCmt.R:27-33: a list of trees, nen and nodelist as one ROW per tree, roots as a vector
      const int nt = cfg->ntrees, E = trees[0].E, nl = trees[0].T - 2;
      List x;
      IntegerMatrix nen(nt, E), nodelist(nt, nl);
      IntegerVector roots(nt);
      for (int t = 0; t < nt; t++) {
        x.push_named("", tree_list(trees[t]));
        for (int e = 0; e < E; e++) nen(t, e) = trees[t].nen[e];
        for (int i = 0; i < nl; i++) nodelist(t, i) = trees[t].nodelist[i];
        roots[t] = trees[t].root;
      }
      res = (v == MT ? phylomap_maketreelistMCMCmt : phylomap_maketreelistMCMCksmt)(x, Qm, pidv, Bm, Omega, nen, nodelist, roots, Nn, prior);
    } else {
      const ref_tree& t = trees[0];
      SEXP x = tree_list(t);
      SEXP nen = int_vec(t.nen, t.E), nodelist = int_vec(t.nodelist, t.T - 2), root = wrap((int)t.root);
      switch (v) {
        case PLAIN: res = phylomap_maketreelistMCMC(x, Qm, pidv, Bm, Omega, nen, nodelist, root, Nn); break;
        case SPARSE: res = phylomap_SPARSEmaketreelistMCMC(x, Qm, pidv, Bm, Omega, nen, nodelist, root, Nn); break;
        case BIGTREE: res = phylomap_maketreelistMCMC_bigtree(x, Qm, pidv, Bm, Omega, nen, nodelist, root, Nn); break;
        case BF: res = phylomap_maketreelistMCMCbf(x, Qm, pidv, Bm, Omega, nen, nodelist, root, Nn, prior); break;
        case KS: res = phylomap_maketreelistMCMCks(x, Qm, pidv, Bm, Omega, nen, nodelist, root, Nn, prior); break;
        case DIC2S: res = phylomap_maketreelistMCMC2sDICt(x, Qm, pidv, Bm, Omega, nen, nodelist, root, Nn, prior); break;
        case DICKS: res = phylomap_maketreelistMCMCksDICt(x, Qm, pidv, Bm, Omega, nen, nodelist, root, Nn, prior); break;
        case EXPV:
          if (!cfg->lefts || !cfg->rights || !cfg->d) throw std::runtime_error("maketreelistEXP needs lefts, rights, d");
          res = phylomap_maketreelistEXP(x, Qm, pidv, nen, nodelist, root, Nn, real_mat(cfg->lefts, n, n),
                                         real_mat(cfg->rights, n, n), real_mat(cfg->d, n, n));
          break;
        default: throw std::runtime_error("unknown variant");
      }
    }
    if (!res) throw std::runtime_error(standin::last_error().empty() ? "reference returned NULL" : standin::last_error());
    NumericMatrix R(res);
    if (R.nrow() != N || R.ncol() != ref_ncols(v, n)) throw std::runtime_error("unexpected result shape");
    for (long i = 0; i < (long)R.size(); i++) out[i] = R[i];
    for (int i = 0; i < n * n; i++) { Q[i] = Qm[i]; B[i] = Bm[i]; }
    return 0;
  } catch (std::exception& e) {
    snprintf(err, errlen, "%s", e.what());
    return 1;
  }
}

#ifdef PM_SHIM_DRIVER
// the two additions of the shim that are not `.Call` symbols of the reference
Rcpp::List phylomap_tree_order(Rcpp::IntegerMatrix& edge, int ntips);
extern "C" int shim_tree_order(const int32_t* edge, int E, int ntips, int32_t* nen, int32_t* nodelist, int32_t* root, char* err, int errlen) {
  try {
    IntegerMatrix em(E, 2);
    for (long i = 0; i < 2L * E; i++) em[i] = edge[i];
    List r = phylomap_tree_order(em, ntips);
    IntegerVector a = r["nen"], b = r["nodelist"];
    for (long i = 0; i < a.size(); i++) nen[i] = a[i];
    for (long i = 0; i < b.size(); i++) nodelist[i] = b[i];
    *root = as<int>(r["root"]);
    return 0;
  } catch (std::exception& e) { snprintf(err, errlen, "%s", e.what()); return 1; }
}
#endif

// R-level probes of the stand-in generator (the same stream the reference consumes), for the RNG tests.
extern "C" void ref_rng_probe(unsigned seed, int kind, int n, double a, double b, double* out) {
  standin::set_seed(seed);
  for (int i = 0; i < n; i++)
    out[i] = kind == 0 ? unif_rand() : kind == 1 ? exp_rand() : kind == 2 ? norm_rand() : kind == 3 ? Rf_rgamma(a, b) : as<double>(rexp(1, a));
}
