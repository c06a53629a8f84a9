// TEST INFRASTRUCTURE ONLY.  CPU oracle for the stochastic-mapping MCMC hot path of vnminin/phylomap.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
// library.  The product (phylomap_b200/) never links, imports or executes it.
//
// PINNED AGAINST THE REFERENCE ITSELF (round 2): the reference ships no golden vectors for this path, but its two source
// files compile unmodified against the stand-in Rcpp / RcppArmadillo / Armadillo headers of oracle/standin/
// (oracle/Makefile, target `ref` -> oracle/_ref/libphylomap_ref.so).  In R-sequential mode this restatement reproduces the
// rows of all ten `.Call` entry points bit for bit (tests/test_reference_pin.py, fixtures tests/golden/reference/*.json
// made by tests/golden/make_reference_golden.py); the DIC log-likelihood column agrees to 1e-12 (arma::expmat is a Pade
// scheme restated twice).  One known divergence: above 16 states RcppArmadillo::sample's std::sort is no longer an
// insertion sort and EXACT ties between weights may be ordered differently (a Jukes-Cantor-like 20-state model).
// What remains unpinned is the layer BELOW both: R's nmath RNG (restated in r_rng.hpp from the published algorithms,
// checked against widely published R outputs, tests/test_oracle_rng.py) and real Armadillo's summation order for n > 4
// (BLAS: implementation-defined).
//
// This file restates /root/reference/src/phylomap.cpp on flat arrays, keeping its order of floating-point operations
// and of random draws; third-party arithmetic (R nmath RNG, RcppArmadillo::sample, Armadillo tiny mat-vec / accu) is
// restated in r_rng.hpp and below from the published algorithms.
//
// Reference map (file = src/phylomap.cpp):
//   Branch / makeabranch ............ :18-34      -> struct Seg, Chain::init_from_maps
//   shortener (plain) ............... :44-73      -> Chain::merge_and_count (full=false)
//   shortenerbf / shortenermtNS ..... :997-1028 / :1881-1913 -> merge_and_count (full=true)
//   resamplebranchstates (+SPARSE,mt) :264-308 / :218-261 / :2050-2094 -> Chain::resample_states
//   sampleabranch* virtual jumps .... :379-410 (bf :1040-1071, SPARSE :329-360, mt :2106-2137) -> Chain::insert_virtual
//   mmmmvFORpl/Tvmmp/sp variants .... :431-457    -> matvec_left / matvec_right
//   makePLrcpp* ..................... :490-529, :1077-1088, :1938-1949 -> Chain::prune
//   sampleinternalnodes* ............ :535-738, :1091-1164, :1314-1403, :1952-2046 -> Chain::sample_nodes
//   updatenodestates* ............... :460-475, :1405-1420, :1918-1935 -> Chain::set_end_states
//   updatedwelltimes* ............... :745-757, :2142-2154 -> Chain::add_dwell
//   treesample* ..................... :761-797, :1169-1179, :1422-1432, :2157-2167 -> Chain::sweep
//   matTospmat ...................... :801-816    -> threshold in Run::init
//   maketreelistMCMC / SPARSE / _bigtree :822-986 -> Run::run_fixed
//   updatel01/10, recordQ, ...bf .... :1181-1305  -> Run::update_bf, Run::run_bf
//   updateks*, recordQks, ...ks ..... :1435-1872  -> Run::update_ks_*, Run::run_ks
//   mt / ksmt ....................... :2169-2362, :2371-2844 -> Run::run_mt
//   EXP comparator .................. :81-208, :2877-3051 -> Run::run_exp
//   DIC chains (2sDICt / ksDICt) .... :3064-3403  -> Run::loglik (PPmakePLD / PPmakePLksD), run_bf / run_ks with dic
//                                                   (arma::expmat is a Pade scheme in Armadillo; restated here as scaling and
//                                                   squaring of a Taylor series, accurate to ~1e-15)
//
// The "site" axis (tree->S > 1) does not exist in the reference: every site is an independent copy of the
// reference's chain state sharing Q; row i of the output holds the SUM over sites of the per-site statistics
// (what the conjugate rate updates need), and per-site columns (root state) report site 0.  S == 1 is the
// reference.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include "r_rng.hpp"

namespace orc {

enum Variant { PLAIN = 0, SPARSE = 1, BIGTREE = 2, BF = 3, KS = 4, MT = 5, KSMT = 6, EXPV = 7, DIC2S = 8, DICKS = 9 };
enum RngMode { SEQUENTIAL = 0, KEYED = 1, TABLE = 2 };
enum SlotKind { K_NODE = 0, K_BRSTATE = 1, K_BREXP = 2 };

struct Seg {
  double len;
  int st;  // 0-based
};

// ------------------------------------------------------------------------------------------------
// Uniform source used inside a sweep: same call sequence in all modes.
// ------------------------------------------------------------------------------------------------
struct SweepRng : UnifSource {
  int mode = SEQUENTIAL;
  RMersenne* mt = nullptr;  // sequential
  uint64_t seed = 0;        // keyed
  uint32_t site = 0, iter = 0, slot = 0, k = 0;
  // table mode / logging: linear slot id
  int64_t lin = 0;
  const int64_t* tab_off = nullptr;
  const double* tab_u = nullptr;
  std::vector<std::vector<double>>* log = nullptr;
  bool table_underrun = false;

  void begin(uint32_t kind, uint32_t idx, int64_t linear) {
    slot = (kind << 28) | idx;
    k = 0;
    lin = linear;
  }
  double unif() override {
    double u;
    if (mode == SEQUENTIAL) {
      u = mt->unif();
      if (log) (*log)[lin].push_back(u);
    } else if (mode == KEYED) {
      u = keyed_uniform(seed, site, iter, slot, k);
    } else {
      int64_t p = tab_off[lin] + k;
      if (p >= tab_off[lin + 1]) { table_underrun = true; u = 0.5; } else u = tab_u[p];
    }
    k++;
    return u;
  }
};

struct ListSource : UnifSource {  // host-side replay of a logged stream
  const double* u; int64_t n, pos = 0; bool underrun = false;
  ListSource(const double* u_, int64_t n_) : u(u_), n(n_) {}
  double unif() override { if (pos >= n) { underrun = true; return 0.5; } return u[pos++]; }
};
struct LoggingSource : UnifSource {  // wraps the MT so host-side draws are exported too
  UnifSource* base; std::vector<double>* sink;
  LoggingSource(UnifSource* b, std::vector<double>* s) : base(b), sink(s) {}
  double unif() override { double v = base->unif(); if (sink) sink->push_back(v); return v; }
};

// ------------------------------------------------------------------------------------------------
// RcppArmadillo::sample(sts, 1, TRUE, w)  (RcppArmadilloExtensions/sample.h: FixProb + ProbSampleReplace).
// Descending sort is an insertion sort: what std::sort does for n <= 16, i.e. ties keep index order.
// ------------------------------------------------------------------------------------------------
static int rcpp_sample(const double* w, int n, UnifSource& g) {
  double sum = 0.0;
  int npos = 0;
  for (int i = 0; i < n; i++) {
    double v = w[i];
    if (!std::isfinite(v)) throw std::range_error("NAs not allowed in probability");
    if (v < 0.0) throw std::range_error("Negative probabilities not allowed");
    if (v > 0.0) { npos++; sum += v; }
  }
  if (npos == 0) throw std::range_error("Not enough positive probabilities");
  double p[64];
  int perm[64];
  for (int i = 0; i < n; i++) { p[i] = w[i] / sum; perm[i] = i; }
  for (int i = 1; i < n; i++) {
    double pv = p[i]; int iv = perm[i]; int j = i - 1;
    while (j >= 0 && pv > p[j]) { p[j + 1] = p[j]; perm[j + 1] = perm[j]; j--; }
    p[j + 1] = pv; perm[j + 1] = iv;
  }
  for (int i = 1; i < n; i++) p[i] = p[i - 1] + p[i];
  double rU = g.unif();
  int jj;
  for (jj = 0; jj < n - 1; jj++) if (rU <= p[jj]) break;
  return perm[jj];
}

// sampleOnce :81-90
static int sample_once(const double* w, int n, double rU) {
  double total = 0;
  for (int i = 0; i < n; i++) total += w[i];
  double cum = 0;
  int i;
  for (i = 0; i < n; i++) { cum += w[i] / total; if (rU < cum) break; }
  return i;
}

// ------------------------------------------------------------------------------------------------
struct Model {
  int n = 0;
  double* Q = nullptr;  // caller's, column-major n x n (R layout), mutated in place by bf/ks/mt
  double* B = nullptr;  // caller's, column-major
  std::vector<double> Bs;  // SPARSE: thresholded copy (column-major)
  std::vector<double> pid;
  double Omega = 0;
  double q(int i, int j) const { return Q[i + j * n]; }
  double& q(int i, int j) { return Q[i + j * n]; }
  double b(int i, int j) const { return B[i + j * n]; }
  double& b(int i, int j) { return B[i + j * n]; }
  double bs(int i, int j) const { return Bs[i + j * n]; }
};

// v <- M v, left-to-right dot products (Armadillo tiny-square gemv / reference BLAS dgemv 'N' order)
static void matvec_left(const Model& m, bool sparse, double* v) {
  int n = m.n; double y[64];
  for (int i = 0; i < n; i++) {
    double acc = 0; bool first = true;
    for (int j = 0; j < n; j++) {
      double a = sparse ? m.bs(i, j) : m.b(i, j);
      if (sparse && a == 0.0) continue;
      double t = a * v[j];
      if (first) { acc = t; first = false; } else acc = acc + t;
    }
    y[i] = acc;
  }
  for (int i = 0; i < n; i++) v[i] = y[i];
}
// v <- M^T v  (Tvmmp with B4 = t(B); SPARSE: rowvec * sp_mat)
static void matvec_right(const Model& m, bool sparse, double* v) {
  int n = m.n; double y[64];
  for (int j = 0; j < n; j++) {
    double acc = 0; bool first = true;
    for (int i = 0; i < n; i++) {
      double a = sparse ? m.bs(i, j) : m.b(i, j);
      if (sparse && a == 0.0) continue;
      double t = a * v[i];
      if (first) { acc = t; first = false; } else acc = acc + t;
    }
    y[j] = acc;
  }
  for (int j = 0; j < n; j++) v[j] = y[j];
}

struct TreeIn {
  int T = 0, E = 0;
  const int* edge = nullptr;      // [2E] column-major, 1-based (parent col, child col)
  const int* nen = nullptr;       // [E] 1-based edge ids, sibling pairs adjacent
  const int* nodelist = nullptr;  // [T-2] 1-based node ids, top-down
  int root = 0;                   // 1-based
  const int64_t* maps_off = nullptr;
  const double* maps_len = nullptr;
  const int* maps_state = nullptr;  // 1-based
  const int* states = nullptr;      // [S*T] 1-based, site-major
  int64_t S = 1;
  const double* edge_length = nullptr;
  // "optimised CPU" baseline only (orc_set_fast_lookup): edge row of every non-root node, so that the node draws cost
  // O(1) per node instead of the reference's linear search (:643); empty = faithful.  Same draws, same results.
  std::vector<int> parent_edge;
  int parent(int e) const { return edge[e]; }
  int child(int e) const { return edge[E + e]; }
};

struct Flags {
  bool sparse = false, normalize = false, full_counts = false, redraw_tips = false, parity_tips = false;
};

// One site on one tree: exactly the state the reference keeps between sweeps.
struct Chain {
  const TreeIn* tr; int64_t site; int tree_idx;
  std::vector<std::list<Seg>> br;
  std::vector<std::vector<Seg>> merged;  // instrumentation: every branch right after shortener, before the new virtual jumps
  std::vector<double> PL;  // (2T-1) x n row-major
  std::vector<int> rm;     // node states 0-based, last sweep
  int n;

  void init(const TreeIn* t, int64_t s, int tidx, int nstates, const Flags& f) {
    tr = t; site = s; tree_idx = tidx; n = nstates;
    br.assign(t->E, {});
    merged.assign(t->E, {});
    for (int e = 0; e < t->E; e++)
      for (int64_t p = t->maps_off[e]; p < t->maps_off[e + 1]; p++) br[e].push_back({t->maps_len[p], t->maps_state[p] - 1});
    PL.assign((size_t)(2 * t->T - 1) * n, 0.0);
    const int* st = t->states + s * t->T;
    for (int i = 0; i < t->T; i++) {
      if (!f.parity_tips) PL[(size_t)i * n + (st[i] - 1)] = 1;
      else {  // :1838-1845 observed 1 -> even 0-based hidden states, 2 -> odd
        int par = st[i] - 2 * (st[i] / 2);
        if (par == 0) for (int j = 1; j < n; j += 2) PL[(size_t)i * n + j] = 1;
        if (par == 1) for (int j = 0; j < n; j += 2) PL[(size_t)i * n + j] = 1;
      }
    }
    rm.assign(2 * t->T - 1, 0);
  }

  int64_t lin_slot(int64_t base, int kind, int idx) const {
    int nn = 2 * tr->T - 1;
    return base + (kind == K_NODE ? idx : nn + 2 * idx + (kind - 1));
  }

  void prune(const Model& M, const Flags& f) {
    const TreeIn& t = *tr;
    int Nnode = t.T - 1;
    std::vector<double> first(n), second(n);
    for (int i = 0; i < Nnode; i++) {
      int ea = t.nen[2 * i] - 1, eb = t.nen[2 * i + 1] - 1;
      for (int j = 0; j < n; j++) first[j] = PL[(size_t)(t.child(eb) - 1) * n + j];
      for (int j = 0; j < n; j++) second[j] = PL[(size_t)(t.child(ea) - 1) * n + j];
      int kb = (int)br[eb].size() - 1, ka = (int)br[ea].size() - 1;
      for (int r = 0; r < kb; r++) matvec_left(M, f.sparse, first.data());
      for (int r = 0; r < ka; r++) matvec_left(M, f.sparse, second.data());
      double* row = &PL[(size_t)(t.parent(ea) - 1) * n];
      for (int j = 0; j < n; j++) row[j] = first[j] * second[j];
      if (f.normalize) {
        // arma::accu on a row subview: two interleaved accumulators
        double a1 = 0, a2 = 0; int j;
        for (j = 1; j < n; j += 2) { a1 += row[j - 1]; a2 += row[j]; }
        if (j - 1 < n) a1 += row[j - 1];
        double s = a1 + a2;
        for (int c = 0; c < n; c++) row[c] = row[c] / s;
      }
    }
  }

  // returns sampled root (0-based)
  int sample_nodes(const Model& M, const Flags& f, SweepRng& g, int64_t base) {
    const TreeIn& t = *tr;
    const int* st = t.states + site * t.T;
    for (int i = 0; i < t.T; i++) rm[i] = st[i] - 1;
    prune(M, f);
    std::vector<double> w(n);
    for (int j = 0; j < n; j++) w[j] = M.pid[j] * PL[(size_t)(t.root - 1) * n + j];
    g.begin(K_NODE, t.root - 1, lin_slot(base, K_NODE, t.root - 1));
    rm[t.root - 1] = rcpp_sample(w.data(), n, g);
    int rootstate = rm[t.root - 1];
    for (int i = 0; i < t.T - 2; i++) {
      int node = t.nodelist[i];
      int j = 0;
      if (!t.parent_edge.empty()) j = t.parent_edge[node - 1];
      else while (t.child(j) != node) j++;  // the reference's O(E) search, :643
      int ps = rm[t.parent(j) - 1];
      std::fill(w.begin(), w.end(), 0.0);
      w[ps] = 1;
      int kk = (int)br[j].size() - 1;
      for (int r = 0; r < kk; r++) matvec_right(M, f.sparse, w.data());
      for (int c = 0; c < n; c++) w[c] = w[c] * PL[(size_t)(node - 1) * n + c];
      g.begin(K_NODE, node - 1, lin_slot(base, K_NODE, node - 1));
      rm[node - 1] = rcpp_sample(w.data(), n, g);
    }
    if (f.redraw_tips) {  // :1385-1397, :2028-2040, edge-row order
      for (int e = 0; e < t.E; e++) {
        if (t.child(e) <= t.T) {
          int cn = t.child(e) - 1, ps = rm[t.parent(e) - 1];
          std::fill(w.begin(), w.end(), 0.0);
          w[ps] = 1;
          int kk = (int)br[e].size() - 1;
          for (int r = 0; r < kk; r++) matvec_right(M, false, w.data());
          for (int c = 0; c < n; c++) w[c] = w[c] * PL[(size_t)cn * n + c];
          g.begin(K_NODE, cn, lin_slot(base, K_NODE, cn));
          rm[cn] = rcpp_sample(w.data(), n, g);
        }
      }
    }
    return rootstate;
  }

  void set_end_states() {
    for (int e = 0; e < tr->E; e++) {
      br[e].front().st = rm[tr->parent(e) - 1];
      br[e].back().st = rm[tr->child(e) - 1];
    }
  }

  void resample_states(std::list<Seg>& b, const Model& M, const Flags& f, SweepRng& g) {
    int ss = (int)b.size();
    if (ss <= 2) return;
    std::vector<double> bp((size_t)n * ss, 0.0);  // column j = B^j e_end
    bp[b.back().st] = 1;
    std::vector<double> v(n);
    for (int j = 1; j < ss - 1; j++) {
      for (int c = 0; c < n; c++) v[c] = bp[(size_t)(j - 1) * n + c];
      matvec_left(M, f.sparse, v.data());
      for (int c = 0; c < n; c++) bp[(size_t)j * n + c] = v[c];
    }
    std::vector<double> w(n);
    auto it = b.begin();
    for (int i = 1; i < ss - 1; i++) {
      int prev = it->st;
      for (int c = 0; c < n; c++) w[c] = M.b(prev, c) * bp[(size_t)(ss - i - 1) * n + c];  // dense row even for SPARSE :254
      ++it;
      it->st = rcpp_sample(w.data(), n, g);
    }
  }

  // stats: [0,n) dwell; counts at n + ...
  void merge_and_count(std::list<Seg>& b, double* stats, bool full) {
    if (full) {
      auto a = b.begin(); auto c = a; ++c;
      for (; c != b.end(); ++a, ++c) stats[n + a->st * n + c->st] += 1;
    }
    auto it = b.begin();
    int cnt = (int)b.size();
    for (int i = 0; i < cnt - 1; i++) {
      auto nx = it; ++nx;
      if (it->st != nx->st) ++it; else { it->len = it->len + nx->len; b.erase(nx); }
    }
    if (!full) {
      auto a = b.begin(); auto c = a; ++c;
      for (; c != b.end(); ++a, ++c) {
        if (a->st < c->st) stats[n + a->st * (n - 1) + c->st - 1] += 1;
        if (a->st > c->st) stats[n + a->st * (n - 1) + c->st] += 1;
      }
    }
  }

  void insert_virtual(std::list<Seg>& b, const Model& M, SweepRng& g) {
    int cnt = (int)b.size();
    auto it = b.begin();
    for (int i = 0; i < cnt; i++) {
      double L = it->len, tot = 0;
      int s = it->st;
      double r = M.Omega + M.q(s, s);
      while (tot < L) {
        double rl = rexp_rate(g, r);
        if ((tot + rl) < L) { b.insert(it, {rl, s}); tot += rl; }
        else { it->len = L - tot; ++it; tot = L; }
      }
    }
  }

  void add_dwell(double* stats) {
    for (int e = 0; e < tr->E; e++) for (auto& s : br[e]) stats[s.st] += s.len;
  }

  int sweep(const Model& M, const Flags& f, SweepRng& g, int64_t base, double* stats) {
    int rootstate = sample_nodes(M, f, g, base);
    set_end_states();
    for (int e = 0; e < tr->E; e++) {
      g.begin(K_BRSTATE, e, lin_slot(base, K_BRSTATE, e));
      resample_states(br[e], M, f, g);
      merge_and_count(br[e], stats, f.full_counts);
      merged[e].assign(br[e].begin(), br[e].end());
      g.begin(K_BREXP, e, lin_slot(base, K_BREXP, e));
      insert_virtual(br[e], M, g);
    }
    add_dwell(stats);
    return rootstate;
  }
};

// ------------------------------------------------------------------------------------------------
struct Run {
  int variant = PLAIN;
  std::vector<TreeIn> trees;
  Model M;
  Flags F;
  std::vector<double> prior;
  int N = 0;
  int rng_mode = SEQUENTIAL;
  uint64_t seed = 0;
  int64_t site_offset = 0;  // keyed mode: global index of local site 0 (site-sharded runs)
  RMersenne mt{0};
  std::vector<std::vector<Chain>> chains;  // [tree][site]
  // logs (sequential mode)
  bool want_log = false;
  std::vector<std::vector<double>> slot_log;
  std::vector<double> host_log;
  // table (replay) mode
  const int64_t* tab_off = nullptr; const double* tab_u = nullptr;
  const double* host_tab = nullptr; int64_t host_tab_n = 0;
  // EXP inputs
  const double *lefts = nullptr, *rights = nullptr, *dmat = nullptr;
  int iters_done = 0;
  std::string err;

  int spi() const { return (2 * trees[0].T - 1) + 2 * trees[0].E; }
  int64_t base_of(int tree, int64_t site, int iter) const {
    return (((int64_t)tree * trees[0].S + site) * N + iter) * spi();
  }
  uint64_t tree_seed(int tree) const { return seed + (uint64_t)tree * 0x9E3779B97F4A7C15ull; }

  int ncols() const {
    int n = M.n, k = n / 2 - 1;
    switch (variant) {
      case PLAIN: case SPARSE: case BIGTREE: case EXPV: return n + n * (n - 1);
      case BF: case MT: return n + n * n + 3;
      case DIC2S: return n + n * n + 4;
      case DICKS: return n + n * n + 2 + 3 * k + 2;
      default: return n + n * n + 2 + 3 * k + 1;
    }
  }

  void init() {
    int n = M.n;
    F = Flags();
    F.sparse = (variant == SPARSE);
    const bool bf = (variant == BF || variant == DIC2S), ks = (variant == KS || variant == DICKS);
    F.normalize = (variant == BIGTREE || bf || ks);
    F.full_counts = (bf || ks || variant == MT || variant == KSMT);
    F.redraw_tips = (ks || variant == MT || variant == KSMT);
    F.parity_tips = (ks || variant == KSMT);
    if (F.sparse) {
      M.Bs.assign((size_t)n * n, 0.0);
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) if (M.b(i, j) > 1e-7) M.Bs[i + j * n] = M.b(i, j);
    }
    chains.resize(trees.size());
    for (size_t t = 0; t < trees.size(); t++) {
      chains[t].resize(trees[t].S);
      for (int64_t s = 0; s < trees[t].S; s++) chains[t][s].init(&trees[t], s, (int)t, n, F);
    }
    mt.set_seed((uint32_t)seed);
    if (want_log && rng_mode == SEQUENTIAL) slot_log.assign((size_t)trees.size() * trees[0].S * N * spi(), {});
  }

  // sweep every site of one tree for iteration `it`, summing statistics into stats[0 .. n+n*n)
  int sweep_tree(int tree, int it, double* stats) {
    int root0 = 0;
    for (int64_t s = 0; s < trees[tree].S; s++) {
      SweepRng g;
      g.mode = rng_mode; g.mt = &mt; g.seed = tree_seed(tree); g.site = (uint32_t)(site_offset + s); g.iter = (uint32_t)it;
      g.tab_off = tab_off; g.tab_u = tab_u;
      g.log = (want_log && rng_mode == SEQUENTIAL) ? &slot_log : nullptr;
      int r = chains[tree][s].sweep(M, F, g, base_of(tree, s, it), stats);
      if (g.table_underrun) throw std::runtime_error("replay table exhausted");
      if (site_offset + s == 0) root0 = r;
    }
    return root0;
  }

  // host-side generator for rate updates
  struct HostRng {
    UnifSource* src; LoggingSource logsrc; ListSource listsrc;
    HostRng(Run& r)
        : logsrc(&r.mt, (r.want_log && r.rng_mode == SEQUENTIAL) ? &r.host_log : nullptr),
          listsrc(r.host_tab, r.host_tab_n) {
      if (r.rng_mode == TABLE) src = &listsrc; else src = &logsrc;
    }
  };

  void set_b_from_q(int i, int j) { M.b(i, j) = (i == j) ? 1 + M.q(i, j) / M.Omega : M.q(i, j) / M.Omega; }

  // ---- fixed-Q drivers :822-986 ----
  void run_fixed(double* out) {
    int nc = ncols();
    std::vector<double> row(nc);
    for (int i = 0; i < N; i++) {
      std::fill(row.begin(), row.end(), 0.0);
      sweep_tree(0, i, row.data());
      for (int c = 0; c < nc; c++) out[i + (size_t)c * N] = row[c];
      iters_done = i + 1;
    }
  }

  // ---- 2-state rate updates :1190-1253 (bf) and :2192-2262 (mt) ----
  void update_2s(const double* st, UnifSource& h, bool which10, bool metropolis) {
    double Om = M.Omega;
    int n00 = (int)st[2], n01 = (int)st[3], n10 = (int)st[4], n11 = (int)st[5];
    double t0 = st[0], t1 = st[1], l01 = M.q(0, 1), l10 = M.q(1, 0);
    if (!which10) {
      double nw = rgamma(h, prior[0] + n01, 1 / (prior[1] + t0));
      if (nw > Om) return;
      double accept = std::pow((Om - nw) / (Om - l01), n00) * std::exp(t0 * (nw - l01));
      if (accept > 1) accept = 1;
      double compare = h.unif();
      if (metropolis && accept < compare) return;
      M.q(0, 0) = -nw; M.q(0, 1) = nw;
      M.b(0, 0) = 1 - nw / Om; M.b(0, 1) = nw / Om;
    } else {
      double nw = rgamma(h, prior[2] + n10, 1 / (prior[3] + t1));
      if (nw > Om) return;
      double accept = std::pow((Om - nw) / (Om - l10), n11) * std::exp(t1 * (nw - l10));
      if (accept > 1) accept = 1;
      double compare = h.unif();
      if (metropolis && accept < compare) return;
      M.q(1, 0) = nw; M.q(1, 1) = -nw;
      M.b(1, 0) = nw / Om; M.b(1, 1) = 1 - nw / Om;
    }
  }

  // ---- DIC: log p(y | Q) by matrix exponentiation, :3135-3178 (PPmakePLD), :3268-3297 (PPmakePLksD), :3242-3250 ----
  // exp(A) for a small dense matrix (row-major n x n): scaling and squaring of a degree-20 Taylor series
  static void expm(const std::vector<double>& A, int n, std::vector<double>& E) {
    double nrm = 0;
    for (int i = 0; i < n; i++) { double r = 0; for (int j = 0; j < n; j++) r += std::fabs(A[i * n + j]); nrm = std::max(nrm, r); }
    int sq = 0;
    while (nrm > 0.5) { nrm *= 0.5; sq++; }
    const double sc = std::ldexp(1.0, -sq);
    std::vector<double> X(n * n), T(n * n, 0.0), N2(n * n);
    for (int i = 0; i < n * n; i++) X[i] = A[i] * sc;
    E.assign(n * n, 0.0);
    for (int i = 0; i < n; i++) { E[i * n + i] = 1.0; T[i * n + i] = 1.0; }
    for (int k = 1; k <= 20; k++) {
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        double acc = 0; for (int l = 0; l < n; l++) acc += T[i * n + l] * X[l * n + j];
        N2[i * n + j] = acc / k;
      }
      T = N2;
      for (int i = 0; i < n * n; i++) E[i] += T[i];
    }
    for (int r = 0; r < sq; r++) {
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        double acc = 0; for (int l = 0; l < n; l++) acc += E[i * n + l] * E[l * n + j];
        N2[i * n + j] = acc;
      }
      E = N2;
    }
  }
  // sum over sites of log( sum_j pid_j PL[root][j] ) + S, with P(t_e) = expm(Q t_e)
  double loglik(int tree) {
    const TreeIn& t = trees[tree];
    int n = M.n;
    std::vector<double> Qr(n * n), TP((size_t)t.E * n * n), A(n * n), E;
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Qr[i * n + j] = M.q(i, j);
    for (int e = 0; e < t.E; e++) {
      for (int i = 0; i < n * n; i++) A[i] = Qr[i] * t.edge_length[e];
      expm(A, n, E);
      std::copy(E.begin(), E.end(), TP.begin() + (size_t)e * n * n);
    }
    double total = 0;
    std::vector<double> PL((size_t)(2 * t.T - 1) * n), x(n), y(n);
    for (int64_t s = 0; s < t.S; s++) {
      const int* st = t.states + s * t.T;
      std::fill(PL.begin(), PL.end(), 0.0);
      for (int i = 0; i < t.T; i++) {
        if (!F.parity_tips) PL[(size_t)i * n + (st[i] - 1)] = 1;
        else {
          int par = st[i] - 2 * (st[i] / 2);
          if (par == 0) for (int j = 1; j < n; j += 2) PL[(size_t)i * n + j] = 1;
          if (par == 1) for (int j = 0; j < n; j += 2) PL[(size_t)i * n + j] = 1;
        }
      }
      double S = 0;
      for (int i = 0; i < t.T - 1; i++) {
        int ea = t.nen[2 * i] - 1, eb = t.nen[2 * i + 1] - 1;
        const double* Pa = &TP[(size_t)ea * n * n]; const double* Pb = &TP[(size_t)eb * n * n];
        const double* ca = &PL[(size_t)(t.child(ea) - 1) * n]; const double* cb = &PL[(size_t)(t.child(eb) - 1) * n];
        double sum = 0;
        double* row = &PL[(size_t)(t.parent(ea) - 1) * n];
        for (int r = 0; r < n; r++) {
          double xa = 0, xb = 0;
          for (int c = 0; c < n; c++) { xa += Pa[r * n + c] * ca[c]; xb += Pb[r * n + c] * cb[c]; }
          x[r] = xa * xb;
          sum += x[r];
        }
        S += std::log(sum);
        for (int r = 0; r < n; r++) row[r] = x[r] / sum;
      }
      double X = 0;
      for (int j = 0; j < n; j++) X += PL[(size_t)(t.root - 1) * n + j] * M.pid[j];
      total += std::log(X) + S;
    }
    return total;
  }

  void run_bf(double* out) {
    int nc = ncols();  // 9 (10 with the log-likelihood column of 2sDICt)
    HostRng H(*this);
    std::vector<double> row(nc);
    for (int i = 0; i < N; i++) {
      std::fill(row.begin(), row.end(), 0.0);
      row[6] = M.q(0, 1); row[7] = M.q(1, 0);
      row[8] = sweep_tree(0, i, row.data());
      if (variant == DIC2S) row[9] = loglik(0);
      update_2s(row.data(), *H.src, false, false);
      update_2s(row.data(), *H.src, true, false);
      for (int c = 0; c < nc; c++) out[i + (size_t)c * N] = row[c];
      iters_done = i + 1;
    }
  }

  // ---- hidden-rate parameters read back from Q :1441-1450 ----
  struct KsPar { int n, k; double l0, l1; std::vector<double> rk, lk, ga; };
  KsPar ks_read() const {
    KsPar p; p.n = M.n; p.k = M.n / 2 - 1;
    p.l0 = M.q(0, 1); p.l1 = M.q(1, 0);
    p.rk.resize(p.k); p.lk.resize(p.k); p.ga.resize(p.k + 1);
    for (int i = 0; i < p.k; i++) p.rk[i] = M.q(2 * i, 2 * i + 2);
    for (int i = 0; i < p.k; i++) p.lk[i] = M.q(2 * i + 2, 2 * i);
    p.ga[0] = 1;
    for (int i = 1; i <= p.k; i++) p.ga[i] = M.q(2 * i, 2 * i + 1) / p.l0;
    return p;
  }
  void record_ks(double* row) const {  // :1789-1798
    int n = M.n, k = n / 2 - 1;
    row[n + n * n] = M.q(0, 1); row[n + n * n + 1] = M.q(1, 0);
    for (int i = 0; i < k; i++) row[n + n * n + 2 + i] = M.q(2 * i, 2 * i + 2);
    for (int i = 0; i < k; i++) row[n + n * n + 2 + k + i] = M.q(2 * i + 2, 2 * i);
    for (int i = 0; i < k; i++) row[n + n * n + 2 + 2 * k + i] = M.q(2 * (i + 1), 2 * (i + 1) + 1) / M.q(0, 1);
  }

  // updateksl01 / updateksl10 (:1435-1584) and their mt twins (:2371-2510).  side: 0 -> l01, 1 -> l10.
  void update_ks_lambda(const double* st, UnifSource& h, int side, bool mtv) {
    KsPar p = ks_read(); int n = p.n, k = p.k; double Om = M.Omega;
    auto soj = [&](int i) { return st[i]; };
    auto tc = [&](int idx) { return st[n + idx]; };
    double lam = side == 0 ? p.l0 : p.l1;
    int pa = (mtv && side == 1) ? 2 : 0, pb = (mtv && side == 1) ? 3 : 1;
    double alphaprime = prior[pa];
    for (int i = 0; i <= k; i++) alphaprime = alphaprime + (side == 0 ? tc(2 * i * n + 2 * i + 1) : tc((2 * i + 1) * n + 2 * i));
    double betaprime = prior[pb];
    for (int i = 0; i <= k; i++) betaprime = betaprime + p.ga[i] * soj(2 * i + side);
    double nw = rgamma(h, alphaprime, 1 / betaprime);
    double gammatimes = betaprime - prior[1];  // ksmt l10 keeps prior(1) here, :2472
    double logaccept = (nw - lam) * gammatimes;
    auto self = [&](int i) { return side == 0 ? tc(2 * i * n + 2 * i) : tc((2 * i + 1) * n + 2 * i + 1); };
    logaccept = logaccept + self(0) * std::log((Om - p.rk[0] - p.ga[0] * nw) / (Om - p.rk[0] - p.ga[0] * lam));
    for (int i = 1; i < k; i++)
      logaccept = logaccept + self(i) * std::log((Om - p.rk[i] - p.lk[i - 1] - p.ga[i] * nw) / (Om - p.rk[i] - p.lk[i - 1] - p.ga[i] * lam));
    logaccept = logaccept + self(k) * std::log((Om - p.lk[k - 1] - p.ga[k] * nw) / (Om - p.lk[k - 1] - p.ga[k] * lam));
    double compare = h.unif();
    if (nw + p.rk[0] > Om) return;
    for (int i = 1; i < k; i++) if (p.ga[i] * nw + p.rk[i] + p.lk[i - 1] > Om) return;
    if (p.ga[k] * nw + p.lk[k - 1] > Om) return;
    if (!mtv && nw < 1e-300) return;
    if (logaccept < std::log(compare)) return;
    int o = side;  // row offset: even rows for l01, odd rows for l10
    auto wr = [&](int i, double diag) {
      int r = 2 * i + o, c = 2 * i + (1 - o);
      M.q(r, r) = diag; M.q(r, c) = p.ga[i] * nw;
      set_b_from_q(r, r); set_b_from_q(r, c);
    };
    wr(0, -p.rk[0] - p.ga[0] * nw);
    for (int i = 1; i < k; i++) wr(i, -p.lk[i - 1] - p.rk[i] - p.ga[i] * nw);
    wr(k, -p.lk[k - 1] - p.ga[k] * nw);
  }

  // updaterkappas :1586-1650 / mt :2514-2575
  void update_ks_rkappa(const double* st, UnifSource& h, int j, bool mtv) {
    KsPar p = ks_read(); int n = p.n; double Om = M.Omega;
    auto tc = [&](int idx) { return st[n + idx]; };
    int pa = mtv ? 4 : 2, pb = mtv ? 5 : 3;
    double alphaprime = prior[pa] + tc((2 * j) * n + 2 * j + 2) + tc((2 * j + 1) * n + 2 * j + 3);
    double betaprime = prior[pb] + st[2 * j] + st[2 * j + 1];
    double nw = rgamma(h, alphaprime, 1 / betaprime);
    double logaccept = (nw - p.rk[j]) * (st[2 * j] + st[2 * j + 1]);
    if (j == 0) logaccept = logaccept + tc((2 * j) * n + 2 * j) * std::log((Om - nw - p.ga[j] * p.l0) / (Om - p.rk[j] - p.ga[j] * p.l0));
    if (j == 0) logaccept = logaccept + tc((2 * j + 1) * n + 2 * j + 1) * std::log((Om - nw - p.ga[j] * p.l1) / (Om - p.rk[j] - p.ga[j] * p.l1));
    if (j > 0) logaccept = logaccept + tc((2 * j) * n + 2 * j) * std::log((Om - p.lk[j - 1] - nw - p.ga[j] * p.l0) / (Om - p.lk[j - 1] - p.rk[j] - p.ga[j] * p.l0));
    if (j > 0) logaccept = logaccept + tc((2 * j + 1) * n + 2 * j + 1) * std::log((Om - p.lk[j - 1] - nw - p.ga[j] * p.l1) / (Om - p.lk[j - 1] - p.rk[j] - p.ga[j] * p.l1));
    double compare = h.unif();
    if (j == 0) if (nw + p.ga[j] * p.l0 > Om) return;
    if (j == 0) if (nw + p.ga[j] * p.l1 > Om) return;
    if (j > 0) if (nw + p.ga[j] * p.l0 + p.lk[j - 1] > Om) return;
    if (j > 0) if (nw + p.ga[j] * p.l1 + p.lk[j - 1] > Om) return;
    if (!mtv && nw < 1e-300) return;
    if (logaccept < std::log(compare)) return;
    M.q(2 * j, 2 * j + 2) = nw; M.q(2 * j + 1, 2 * j + 3) = nw;
    if (j == 0) { M.q(0, 0) = -nw - p.ga[j] * p.l0; M.q(1, 1) = -nw - p.ga[j] * p.l1; }
    if (j > 0) { M.q(2 * j, 2 * j) = -nw - p.lk[j - 1] - p.ga[j] * p.l0; M.q(2 * j + 1, 2 * j + 1) = -nw - p.lk[j - 1] - p.ga[j] * p.l1; }
    set_b_from_q(2 * j, 2 * j); set_b_from_q(2 * j + 1, 2 * j + 1);
    set_b_from_q(2 * j, 2 * j + 2); set_b_from_q(2 * j + 1, 2 * j + 3);
  }

  // updatelkappas :1655-1718 / mt :2580-2640
  void update_ks_lkappa(const double* st, UnifSource& h, int j, bool mtv) {
    KsPar p = ks_read(); int n = p.n, k = p.k; double Om = M.Omega;
    auto tc = [&](int idx) { return st[n + idx]; };
    int pa = mtv ? 4 : 2, pb = mtv ? 5 : 3;
    double alphaprime = prior[pa] + tc((2 * j) * n + 2 * j - 2) + tc((2 * j + 1) * n + 2 * j - 1);
    double betaprime = prior[pb] + st[2 * j] + st[2 * j + 1];
    double nw = rgamma(h, alphaprime, 1 / betaprime);
    double logaccept = (nw - p.lk[j - 1]) * (st[2 * j] + st[2 * j + 1]);
    if (j == k) logaccept = logaccept + tc((2 * j) * n + 2 * j) * std::log((Om - nw - p.ga[j] * p.l0) / (Om - p.lk[j - 1] - p.ga[j] * p.l0));
    if (j == k) logaccept = logaccept + tc((2 * j + 1) * n + 2 * j + 1) * std::log((Om - nw - p.ga[j] * p.l1) / (Om - p.lk[j - 1] - p.ga[j] * p.l1));
    if (j < k) logaccept = logaccept + tc((2 * j) * n + 2 * j) * std::log((Om - p.rk[j] - nw - p.ga[j] * p.l0) / (Om - p.rk[j] - p.lk[j - 1] - p.ga[j] * p.l0));
    if (j < k) logaccept = logaccept + tc((2 * j + 1) * n + 2 * j + 1) * std::log((Om - p.rk[j] - nw - p.ga[j] * p.l1) / (Om - p.rk[j] - p.lk[j - 1] - p.ga[j] * p.l1));
    double compare = h.unif();
    if (j == k) if (nw + p.ga[j] * p.l0 > Om) return;
    if (j == k) if (nw + p.ga[j] * p.l1 > Om) return;
    if (j < k) if (nw + p.ga[j] * p.l0 + p.rk[j] > Om) return;
    if (j < k) if (nw + p.ga[j] * p.l1 + p.rk[j] > Om) return;
    if (!mtv && nw < 1e-300) return;
    if (logaccept < std::log(compare)) return;
    M.q(2 * j, 2 * j - 2) = nw; M.q(2 * j + 1, 2 * j - 1) = nw;
    if (j == k) { M.q(2 * j, 2 * j) = -nw - p.ga[j] * p.l0; M.q(2 * j + 1, 2 * j + 1) = -nw - p.ga[j] * p.l1; }
    if (j < k) { M.q(2 * j, 2 * j) = -nw - p.rk[j] - p.ga[j] * p.l0; M.q(2 * j + 1, 2 * j + 1) = -nw - p.rk[j] - p.ga[j] * p.l1; }
    set_b_from_q(2 * j, 2 * j); set_b_from_q(2 * j + 1, 2 * j + 1);
    set_b_from_q(2 * j, 2 * j - 2); set_b_from_q(2 * j + 1, 2 * j - 1);
  }

  // updategammas :1722-1786 / mt :2643-2704
  void update_ks_gamma(const double* st, UnifSource& h, int j, bool mtv) {
    KsPar p = ks_read(); int n = p.n, k = p.k; double Om = M.Omega;
    auto tc = [&](int idx) { return st[n + idx]; };
    int pa = mtv ? 6 : 4, pb = mtv ? 7 : 5;
    double alphaprime = prior[pa] + tc((2 * j) * n + 2 * j + 1) + tc((2 * j + 1) * n + 2 * j);
    double betaprime = prior[pb] + st[2 * j] * p.l0 + st[2 * j + 1] * p.l1;
    double nw = rgamma(h, alphaprime, 1 / betaprime);
    double logaccept = (nw - p.ga[j]) * (st[2 * j] * p.l0 + st[2 * j + 1] * p.l1);
    if (j == k) logaccept = logaccept + tc((2 * j) * n + 2 * j) * std::log((Om - p.lk[j - 1] - nw * p.l0) / (Om - p.lk[j - 1] - p.ga[j] * p.l0));
    if (j == k) logaccept = logaccept + tc((2 * j + 1) * n + 2 * j + 1) * std::log((Om - p.lk[j - 1] - nw * p.l1) / (Om - p.lk[j - 1] - p.ga[j] * p.l1));
    if (j < k) logaccept = logaccept + tc((2 * j) * n + 2 * j) * std::log((Om - p.lk[j - 1] - p.rk[j] - nw * p.l0) / (Om - p.rk[j] - p.lk[j - 1] - p.ga[j] * p.l0));
    if (j < k) logaccept = logaccept + tc((2 * j + 1) * n + 2 * j + 1) * std::log((Om - p.lk[j - 1] - p.rk[j] - nw * p.l1) / (Om - p.rk[j] - p.lk[j - 1] - p.ga[j] * p.l1));
    double compare = h.unif();
    if (j == k) if (p.lk[j - 1] + nw * p.l0 > Om) return;
    if (j == k) if (p.lk[j - 1] + nw * p.l1 > Om) return;
    if (j < k) if (p.lk[j - 1] + nw * p.l0 + p.rk[j] > Om) return;
    if (j < k) if (p.lk[j - 1] + nw * p.l1 + p.rk[j] > Om) return;
    if (!mtv && nw < 1e-300) return;
    if (logaccept < std::log(compare)) return;
    M.q(2 * j, 2 * j + 1) = nw * p.l0; M.q(2 * j + 1, 2 * j) = nw * p.l1;
    if (j == k) { M.q(2 * j, 2 * j) = -p.lk[j - 1] - nw * p.l0; M.q(2 * j + 1, 2 * j + 1) = -p.lk[j - 1] - nw * p.l1; }
    if (j < k) { M.q(2 * j, 2 * j) = -p.lk[j - 1] - p.rk[j] - nw * p.l0; M.q(2 * j + 1, 2 * j + 1) = -p.lk[j - 1] - p.rk[j] - nw * p.l1; }
    set_b_from_q(2 * j, 2 * j); set_b_from_q(2 * j + 1, 2 * j + 1);
    set_b_from_q(2 * j, 2 * j + 1); set_b_from_q(2 * j + 1, 2 * j);
  }

  void ks_updates(const double* row, UnifSource& h, bool mtv) {
    int k = M.n / 2 - 1;
    update_ks_lambda(row, h, 0, mtv);
    update_ks_lambda(row, h, 1, mtv);
    for (int j = 0; j < k; j++) update_ks_rkappa(row, h, j, mtv);
    for (int j = 1; j <= k; j++) update_ks_lkappa(row, h, j, mtv);
    for (int j = 1; j <= k; j++) update_ks_gamma(row, h, j, mtv);
  }

  void run_ks(double* out) {
    int nc = ncols(), n = M.n, k = n / 2 - 1;
    HostRng H(*this);
    std::vector<double> row(nc);
    for (int i = 0; i < N; i++) {
      std::fill(row.begin(), row.end(), 0.0);
      record_ks(row.data());
      row[n + n * n + 2 + 3 * k] = sweep_tree(0, i, row.data());
      if (variant == DICKS) row[n + n * n + 2 + 3 * k + 1] = loglik(0);
      ks_updates(row.data(), *H.src, false);
      for (int c = 0; c < nc; c++) out[i + (size_t)c * N] = row[c];
      iters_done = i + 1;
    }
  }

  // ---- multiple trees :2267-2362 (2-state), :2722-2844 (k-state) ----
  void run_mt(double* out) {
    int nc = ncols(), n = M.n;
    bool ks = (variant == KSMT);
    int ntree = (int)trees.size();
    HostRng H(*this);
    std::vector<std::vector<double>> jodt(ntree, std::vector<double>(nc - 1, 0.0));
    std::vector<double> ones(ntree, 1.0);
    double wt = H.src->unif();  // :2332 / :2810
    (void)wt;
    for (int i = 0; i < N; i++) {
      for (int j = 0; j < ntree; j++) {
        std::fill(jodt[j].begin(), jodt[j].end(), 0.0);
        if (ks) record_ks(jodt[j].data()); else { jodt[j][6] = M.q(0, 1); jodt[j][7] = M.q(1, 0); }
        sweep_tree(j, i, jodt[j].data());
      }
      wt = H.src->unif();
      int j = sample_once(ones.data(), ntree, wt);
      if (j >= ntree) j = ntree - 1;
      for (int c = 0; c < nc - 1; c++) out[i + (size_t)c * N] = jodt[j][c];
      out[i + (size_t)(nc - 1) * N] = j;
      if (ks) ks_updates(jodt[j].data(), *H.src, true);
      else { update_2s(jodt[j].data(), *H.src, false, true); update_2s(jodt[j].data(), *H.src, true, true); }
      iters_done = i + 1;
    }
    (void)n;
  }

  // ---- EXP comparator :93-208, :2877-3051 (sequential stream only; statistical yardstick) ----
  static double dpois(int k, double lam) { return std::exp(-lam + k * std::log(lam) - std::lgamma(k + 1.0)); }

  void run_exp(double* out) {
    int n = M.n, nc = ncols();
    const TreeIn& t = trees[0];
    double rate = 0;
    for (int i = 0; i < n; i++) rate = (i == 0) ? M.q(0, 0) : std::min(rate, M.q(i, i));
    rate = -1.0 * rate;
    std::vector<double> B2((size_t)n * n);
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) B2[i + j * n] = (i == j ? 1.0 : 0.0) + M.q(i, j) / rate;
    auto b2 = [&](int i, int j) { return B2[i + j * n]; };
    // P(t_e) = | L exp(D t) R |
    std::vector<double> TP((size_t)t.E * n * n);
    std::vector<double> tmp((size_t)n * n);
    for (int e = 0; e < t.E; e++) {
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) tmp[i + j * n] = lefts[i + j * n] * std::exp(dmat[j + j * n] * t.edge_length[e]);
      for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
        double acc = 0;
        for (int l = 0; l < n; l++) acc += tmp[i + l * n] * rights[l + j * n];
        TP[(size_t)e * n * n + i + j * n] = std::fabs(acc);
      }
    }
    auto tp = [&](int e, int i, int j) { return TP[(size_t)e * n * n + i + j * n]; };
    RMersenne& g = mt;
    std::vector<double> row(nc), PL((size_t)(2 * t.T - 1) * n), w(n);
    std::vector<int> rm(2 * t.T - 1);
    for (int it = 0; it < N; it++) {
      std::fill(row.begin(), row.end(), 0.0);
      for (int64_t s = 0; s < t.S; s++) {
        const int* st = t.states + s * t.T;
        std::fill(PL.begin(), PL.end(), 0.0);
        for (int i = 0; i < t.T; i++) { PL[(size_t)i * n + st[i] - 1] = 1; rm[i] = st[i] - 1; }
        for (int i = 0; i < t.T - 1; i++) {
          int ea = t.nen[2 * i] - 1, eb = t.nen[2 * i + 1] - 1;
          for (int r = 0; r < n; r++) {
            double x = 0, y = 0;
            for (int c = 0; c < n; c++) x += tp(ea, r, c) * PL[(size_t)(t.child(ea) - 1) * n + c];
            for (int c = 0; c < n; c++) y += tp(eb, r, c) * PL[(size_t)(t.child(eb) - 1) * n + c];
            PL[(size_t)(t.parent(ea) - 1) * n + r] = x * y;
          }
        }
        for (int j = 0; j < n; j++) w[j] = M.pid[j] * PL[(size_t)(t.root - 1) * n + j];
        rm[t.root - 1] = rcpp_sample(w.data(), n, g);
        for (int i = 0; i < t.T - 2; i++) {
          int node = t.nodelist[i], j = 0;
          while (t.child(j) != node) j++;
          int ps = rm[t.parent(j) - 1];
          for (int c = 0; c < n; c++) w[c] = tp(j, ps, c) * PL[(size_t)(node - 1) * n + c];
          rm[node - 1] = rcpp_sample(w.data(), n, g);
        }
        for (int e = 0; e < t.E; e++) {
          int a = rm[t.parent(e) - 1], b = rm[t.child(e) - 1];
          double T = t.edge_length[e], P = tp(e, a, b);
          std::vector<std::vector<double>> bp(1, std::vector<double>(n, 0.0));
          bp[0][b] = 1;
          double rU = g.unif(), cum = 0;
          if (a == b) cum = dpois(0, rate * T) / P;
          bool notExceed = !(cum > rU);
          int nj = 0; bool broken = false;
          while (notExceed) {
            nj++;
            if (nj > 300) { broken = true; break; }
            std::vector<double> nx(n);
            for (int r = 0; r < n; r++) { double acc = 0; for (int c = 0; c < n; c++) acc += b2(r, c) * bp[nj - 1][c]; nx[r] = acc; }
            bp.push_back(nx);
            cum += dpois(nj, rate * T) * bp[nj][a] / P;
            if (cum > rU) notExceed = false;
          }
          if (broken) continue;  // reference returns without touching the branch or the counts
          std::vector<int> ps; std::vector<double> pt;
          if (nj == 0 || (nj == 1 && a == b)) { ps = {a, b}; pt = {0, T}; }
          else if (nj == 1) { ps = {a, b, b}; pt = {0, T * g.unif(), T}; }
          else {
            std::vector<double> jt(nj);
            for (int i = 0; i < nj; i++) jt[i] = T * g.unif();
            std::sort(jt.begin(), jt.end());
            std::vector<int> ds(nj + 1);
            ds[0] = a; ds[nj] = b;
            for (int i = 1; i < nj; i++) {
              for (int c = 0; c < n; c++) w[c] = b2(ds[i - 1], c) * bp[nj - i][c];
              ds[i] = sample_once(w.data(), n, g.unif());
              if (ds[i] >= n) ds[i] = n - 1;
            }
            ps.push_back(a); pt.push_back(0.0);
            for (int i = 1; i <= nj; i++) if (ds[i - 1] != ds[i]) { ps.push_back(ds[i]); pt.push_back(jt[i - 1]); }
            ps.push_back(b); pt.push_back(T);
          }
          int m = (int)ps.size() - 1;
          for (int i = 1; i < m; i++) {
            if (ps[i - 1] < ps[i]) row[n + ps[i - 1] * (n - 1) + ps[i] - 1] += 1;
            if (ps[i - 1] > ps[i]) row[n + ps[i - 1] * (n - 1) + ps[i]] += 1;
          }
          for (int i = 0; i < m; i++) row[ps[i]] += pt[i + 1] - pt[i];
        }
      }
      for (int c = 0; c < nc; c++) out[it + (size_t)c * N] = row[c];
      iters_done = it + 1;
    }
  }

  void run(double* out) {
    switch (variant) {
      case PLAIN: case SPARSE: case BIGTREE: run_fixed(out); break;
      case BF: case DIC2S: run_bf(out); break;
      case KS: case DICKS: run_ks(out); break;
      case MT: case KSMT: run_mt(out); break;
      case EXPV: run_exp(out); break;
      default: throw std::runtime_error("unknown variant");
    }
  }
};

}  // namespace orc

// ------------------------------------------------------------------------------------------------
// C interface for ctypes
// ------------------------------------------------------------------------------------------------
extern "C" {

struct orc_tree {
  int32_t T, E;
  const int32_t* edge; const int32_t* nen; const int32_t* nodelist; int32_t root;
  const int64_t* maps_off; const double* maps_len; const int32_t* maps_state;
  const int32_t* states; int64_t S;
  const double* edge_length;
};

struct orc_config {
  int32_t variant, n, N, ntrees;
  double Omega;
  const double* prior; int32_t nprior;
  int32_t rng_mode; uint64_t seed; int32_t want_log;
  const int64_t* tab_off; const double* tab_u; const double* host_tab; int64_t host_tab_n;
  const double* lefts; const double* rights; const double* d;
  int64_t site_offset;
};

void* orc_create(const orc_tree* trees, const orc_config* cfg, double* Q, const double* pid, double* B, char* err, int errlen) {
  try {
    auto* r = new orc::Run();
    r->variant = cfg->variant; r->N = cfg->N;
    for (int t = 0; t < cfg->ntrees; t++) {
      orc::TreeIn ti;
      ti.T = trees[t].T; ti.E = trees[t].E; ti.edge = trees[t].edge; ti.nen = trees[t].nen; ti.nodelist = trees[t].nodelist;
      ti.root = trees[t].root; ti.maps_off = trees[t].maps_off; ti.maps_len = trees[t].maps_len; ti.maps_state = trees[t].maps_state;
      ti.states = trees[t].states; ti.S = trees[t].S; ti.edge_length = trees[t].edge_length;
      r->trees.push_back(ti);
    }
    r->M.n = cfg->n; r->M.Q = Q; r->M.B = B; r->M.Omega = cfg->Omega;
    r->M.pid.assign(pid, pid + cfg->n);
    if (cfg->prior) r->prior.assign(cfg->prior, cfg->prior + cfg->nprior);
    r->site_offset = cfg->site_offset;
    r->rng_mode = cfg->rng_mode; r->seed = cfg->seed; r->want_log = cfg->want_log != 0;
    r->tab_off = cfg->tab_off; r->tab_u = cfg->tab_u; r->host_tab = cfg->host_tab; r->host_tab_n = cfg->host_tab_n;
    r->lefts = cfg->lefts; r->rights = cfg->rights; r->dmat = cfg->d;
    if (cfg->n > 64) throw std::runtime_error("oracle supports at most 64 states");
    r->init();
    return r;
  } catch (std::exception& e) { snprintf(err, errlen, "%s", e.what()); return nullptr; }
}

int orc_ncols(void* h) { return ((orc::Run*)h)->ncols(); }

// bench.py's "optimised CPU" leg: replace the reference's O(E) edge search per node by a lookup table
void orc_set_fast_lookup(void* h, int on) {
  auto* r = (orc::Run*)h;
  for (auto& t : r->trees) {
    t.parent_edge.clear();
    if (!on) continue;
    t.parent_edge.assign(2 * t.T - 1, 0);
    for (int e = t.E - 1; e >= 0; e--) t.parent_edge[t.child(e) - 1] = e;  // first matching row wins, like the search
  }
}

int orc_run(void* h, double* out, char* err, int errlen) {
  auto* r = (orc::Run*)h;
  try { r->run(out); return 0; }
  catch (std::exception& e) { snprintf(err, errlen, "%s", e.what()); return 1; }
}

// state after the last completed sweep (for history parity)
void orc_get_node_states(void* h, int tree, int32_t* out /* [S][2T-1] 0-based */) {
  auto* r = (orc::Run*)h; int nn = 2 * r->trees[tree].T - 1;
  for (int64_t s = 0; s < r->trees[tree].S; s++) for (int i = 0; i < nn; i++) out[s * nn + i] = r->chains[tree][s].rm[i];
}
void orc_get_piece_counts(void* h, int tree, int32_t* out /* [S][E] */) {
  auto* r = (orc::Run*)h; int E = r->trees[tree].E;
  for (int64_t s = 0; s < r->trees[tree].S; s++) for (int e = 0; e < E; e++) out[s * E + e] = (int)r->chains[tree][s].br[e].size();
}
// merged real path of one (site, branch) as the last sweep's shortener left it (before the virtual-jump insertion)
int orc_get_path(void* h, int tree, int64_t site, int e, double* len, int32_t* st, int cap) {
  auto* r = (orc::Run*)h; auto& b = r->chains[tree][site].merged[e];
  int k = 0;
  for (auto& s : b) { if (k < cap) { len[k] = s.len; st[k] = s.st; } k++; }
  return k;
}
// raw pieces (incl. virtual jumps)
int orc_get_pieces(void* h, int tree, int64_t site, int e, double* len, int32_t* st, int cap) {
  auto* r = (orc::Run*)h; auto& b = r->chains[tree][site].br[e]; int k = 0;
  for (auto& s : b) { if (k < cap) { len[k] = s.len; st[k] = s.st; } k++; }
  return k;
}
// partial likelihoods of the last sweep, row-major [2T-1][n]
void orc_get_pl(void* h, int tree, int64_t site, double* out) {
  auto* r = (orc::Run*)h; auto& c = r->chains[tree][site];
  std::memcpy(out, c.PL.data(), c.PL.size() * sizeof(double));
}

int64_t orc_log_nslots(void* h) { return (int64_t)((orc::Run*)h)->slot_log.size(); }
int64_t orc_log_total(void* h) { int64_t t = 0; for (auto& v : ((orc::Run*)h)->slot_log) t += (int64_t)v.size(); return t; }
void orc_log_export(void* h, int64_t* off, double* u) {
  auto* r = (orc::Run*)h; int64_t p = 0; size_t i = 0;
  for (; i < r->slot_log.size(); i++) { off[i] = p; for (double v : r->slot_log[i]) u[p++] = v; }
  off[i] = p;
}
int64_t orc_hostlog_n(void* h) { return (int64_t)((orc::Run*)h)->host_log.size(); }
void orc_hostlog_export(void* h, double* u) { auto* r = (orc::Run*)h; std::copy(r->host_log.begin(), r->host_log.end(), u); }

void orc_destroy(void* h) { delete (orc::Run*)h; }

// ---- RNG probes for the known-answer tests ----
void orc_rng_probe(uint32_t seed, int kind, int n, double a, double b, double* out) {
  orc::RMersenne g(seed);
  for (int i = 0; i < n; i++) {
    switch (kind) {
      case 0: out[i] = g.unif(); break;
      case 1: out[i] = orc::exp_rand(g); break;
      case 2: out[i] = orc::norm_rand(g); break;
      case 3: out[i] = orc::rgamma(g, a, b); break;
      case 4: out[i] = orc::rexp_rate(g, a); break;
    }
  }
}
void orc_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) { orc::philox4x32_10(ctr, key, out); }
double orc_keyed_uniform(uint64_t seed, uint32_t site, uint32_t iter, uint32_t slot, uint32_t k) {
  return orc::keyed_uniform(seed, site, iter, slot, k);
}
int orc_sample(const double* w, int n, double u, char* err, int errlen) {
  struct One : orc::UnifSource { double v; double unif() override { return v; } } g; g.v = u;
  try { return orc::rcpp_sample(w, n, g); } catch (std::exception& e) { snprintf(err, errlen, "%s", e.what()); return -1; }
}
}
