"""The package tutorial (vignettes/phylomap_tutorial.Rnw) step by step, on synthetic trees of the same sizes (the
tutorial draws its 50-tip tree with diversitree and reads 70-tip Cephalopod trees that are not shipped in extdata).

  Fixed rate matrices (:67-135): a 20-state tridiagonal Q, 50 tips, Omega = 0.2, N = 1000 histories from sumstatEXP,
  sumstatMCMC and SPARSEsumstatMCMC; the tutorial overlays the three histograms of the number of jumps -- here their
  means, standard deviations and two-sample KS p-values.
  Free rate matrices (:145-310): sumstatMCMCbf (2 states, prior2, Omega 10), sumstatMCMCks (4 states, prior4),
  sumstatMCMCmt and sumstatMCMCksmt over ten trees, N = 100.

    python examples/tutorial.py
"""
import os
import sys

import numpy as np
from scipy import stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import phylomap_b200 as pb  # noqa: E402
from phylomap_b200 import synth  # noqa: E402

# ---- fixed rate matrices ----
Q = np.zeros((20, 20))
for j in range(19):
    Q[j, j + 1] = 0.003
    Q[j + 1, j] = 0.003
np.fill_diagonal(Q, -Q.sum(1))
numtips, dimQ = 50, 20
phy = synth.yule_tree(numtips, seed=3, mean_branch=12.0)
pid = np.full(dimQ, 1.0 / dimQ)
tips = synth.simulate_tip_states(phy, Q, np.eye(dimQ)[0], 1, seed=11).numpy()[0].astype(np.int32)   # x0 = 1
z = phy.with_states(tips)
Omega, N = 0.2, 1000
clEXP = pb.sumstatEXP(z, Q, pid, N, seed=1)
clMCMC = pb.sumstatMCMC(z, Q, pid, Omega, N, seed=2)
clSPARSE = pb.SPARSEsumstatMCMC(z, Q, pid, Omega, N, seed=3)
jumps = {k: v[:, dimQ:].sum(1) for k, v in (("EXP", clEXP), ("MCMC", clMCMC[100:]), ("SPARSE", clSPARSE[100:]))}
print("number of jumps, 20-state model on %d tips (tip states %s...)" % (numtips, tips[:8].tolist()))
for k, v in jumps.items():
    print("  %-7s mean %.3f  sd %.3f" % (k, v.mean(), v.std()))
print("  KS p-values  EXP~MCMC %.3f   EXP~SPARSE %.3f   MCMC~SPARSE %.3f" % (
    stats.ks_2samp(jumps["EXP"], jumps["MCMC"][::5]).pvalue, stats.ks_2samp(jumps["EXP"], jumps["SPARSE"][::5]).pvalue,
    stats.ks_2samp(jumps["MCMC"][::5], jumps["SPARSE"][::5]).pvalue))

# ---- free rate matrices ----
Q2 = np.array([[-0.1, 0.1], [0.1, -0.1]])
pid2, prior2, Omega2 = np.array([0.5, 0.5]), np.array([0.55, 1, 0.56, 1.01]), 10.0
base = synth.yule_tree(70, seed=5, mean_branch=0.6)
rng = np.random.default_rng(7)
treelist = [pb.PhyloTree(base.edge, base.edge_length * rng.uniform(0.8, 1.25, size=base.E)) for _ in range(10)]
trait = synth.simulate_tip_states(base, Q2, pid2, 1, seed=13).numpy()[0].astype(np.int32)
treelist = [t.with_states(trait) for t in treelist]          # tip branches halved, named (1, tip state): :172-190
atree, N = treelist[0], 100
cephAnl2 = pb.sumstatMCMCbf(atree, np.asfortranarray(Q2.copy()), pid2, Omega2, N, prior2, seed=4)
print("sumstatMCMCbf   %s   last row %s" % (pb.colnames(pb.sumstatMCMCbf), np.round(cephAnl2[-1], 4).tolist()))
Q4 = synth.make2sQ(0.1, 0.1, 0.2, 0.2, 10)
pid4, prior4, Omega4 = np.full(4, 0.25), np.array([1.0, 10, 2, 10, 20, 2]), 10.0
cephAnl4 = pb.sumstatMCMCks(atree, np.asfortranarray(Q4.copy()), pid4, Omega4, N, prior4, seed=5)
print("sumstatMCMCks   rates (l01, l10, k01, k10, gamma) of the last row %s" % np.round(cephAnl4[-1, 20:25], 4).tolist())
cephAnl2mt = pb.sumstatMCMCmt(treelist[:10], np.asfortranarray(Q2.copy()), pid2, Omega2, N, prior2, seed=6)
print("sumstatMCMCmt   trees used %s" % np.bincount(cephAnl2mt[:, -1].astype(int), minlength=10).tolist())
prior4mt = np.array([1.0, 10, 1.1, 11, 2, 10, 20, 2])
cephAnl4mt = pb.sumstatMCMCksmt(treelist[:10], np.asfortranarray(Q4.copy()), pid4, Omega4, N, prior4mt, seed=7)
print("sumstatMCMCksmt trees used %s   mean l01 %.4f" % (np.bincount(cephAnl4mt[:, -1].astype(int), minlength=10).tolist(), cephAnl4mt[:, 20].mean()))
