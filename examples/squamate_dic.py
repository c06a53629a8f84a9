"""The reference's DIC vignette (vignettes/Squamate_DIC_model_selection.Rnw:70-125, "A single DIC comparison") on a B200:

    atree  <- readRDS(system.file("extdata/Squamate/phylomap_compatible_squamate_tree.RData", package = "phylomap"))
    atree2 <- simulate_2_state_tree(seed = 101, atree, Q2, pid2)
    SIMtstr <- sumstatMCMC2sDICt(atree2, Q2, pid2, Omega, N, prior2r)
    SIMfstr <- sumstatMCMCksDICt(atree2, Q4, pid4, Omega, N, prior4r)
    DIC_2_state <- make2stateDICbig(SIMtstr[-c(1:1000), ], atree2, pid2, ne)
    DIC_4_state <- make4stateDICbig(SIMfstr[-c(1:1000), ], atree2, pid4, ne)

"The computations take a long time to run so the code is not executed" there; here they are the two calls below.

    python examples/squamate_dic.py [tree.RData | tests/golden/squamate_tree.npz] [N] [out_prefix]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import phylomap_b200 as pb  # noqa: E402
from phylomap_b200 import rds, synth  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "tests", "golden", "squamate_tree.npz")
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
out = sys.argv[3] if len(sys.argv) > 3 else None
if src.endswith(".npz"):
    import cases
    atree = cases.squamate_tree()
else:
    atree = pb.PhyloTree.read_rds(src)

Q2 = np.array([[-0.001, 0.001], [0.006, -0.006]])          # matrix(c(-0.001, 0.006, 0.001, -0.006), nrow = 2)
Q4 = synth.make2sQ(0.001, 0.001, 0.001, 0.03, 16)
pid2, pid4 = np.array([0.5, 0.5]), np.full(4, 0.25)
prior2r, prior4r = np.array([0.55, 1, 0.55, 1.0]), np.array([1.0, 10, 2, 10, 20, 2])
Omega = 10.0
atree2 = synth.simulate_2_state_tree(101, atree, Q2, pid2, segments=100)
burn = min(int(np.ceil(N / 10)), 1000)

t0 = time.perf_counter()
SIMtstr = pb.sumstatMCMC2sDICt(atree2, np.asfortranarray(Q2.copy()), pid2, Omega, N, prior2r, precision="f64", seed=101)
t1 = time.perf_counter()
SIMfstr = pb.sumstatMCMCksDICt(atree2, np.asfortranarray(Q4.copy()), pid4, Omega, N, prior4r, precision="f64", seed=101)
t2 = time.perf_counter()
ne = pb.pruningwiseedgeorder(atree2)
DIC_2_state = pb.make2stateDICbig(SIMtstr[burn:], atree2, pid2, ne, precision="f64")
DIC_4_state = pb.make4stateDICbig(SIMfstr[burn:], atree2, pid4, ne, precision="f64")
t3 = time.perf_counter()
print("tips %d  branches %d  tree length %.1f  tips in state 2: %d" % (atree2.T, atree2.E, atree2.edge_length.sum(), int((atree2.states == 2).sum())))
print("sumstatMCMC2sDICt  N = %d: %.1f s (%.2f ms per sweep)   posterior mean l01 = %.5f  l10 = %.5f" %
      (N, t1 - t0, 1e3 * (t1 - t0) / N, SIMtstr[burn:, 6].mean(), SIMtstr[burn:, 7].mean()))
print("sumstatMCMCksDICt  N = %d: %.1f s (%.2f ms per sweep)   posterior mean l01 = %.5f  l10 = %.5f  k01 = %.5f  k10 = %.5f  gamma = %.3f" %
      ((N, t2 - t1, 1e3 * (t2 - t1) / N) + tuple(SIMfstr[burn:, 20 + i].mean() for i in range(5))))
print("DIC_2_state = %.3f   DIC_4_state = %.3f   (%.2f s)   [the vignette's R run: 2538.272 and 2536.056, with R's tip data]" %
      (DIC_2_state, DIC_4_state, t3 - t2))
if out:
    rds.write_rds(out + "_2s.rds", SIMtstr, colnames=pb.colnames(pb.sumstatMCMC2sDICt))
    rds.write_rds(out + "_4s.rds", SIMfstr, colnames=pb.colnames(pb.sumstatMCMCksDICt, 4))
