import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, cases
import phylomap_b200 as pb
from phylomap_b200 import capi
from oracle import bridge
z = cases.tree2(T=30, S=4, seed=9)
Om = 0.5
for k in range(1, 8):
    o = bridge.OracleRun(bridge.BF, [z.oracle_dict()], cases.Q2.copy(), cases.PID2, Om, k, prior=cases.PRIOR_BF, rng_mode=bridge.KEYED, seed=7)
    ref = o.run()
    ch = pb.Chain(capi.PM_V_BF, z, np.asfortranarray(cases.Q2.copy()), cases.PID2, Om, k, prior=cases.PRIOR_BF, seed=7, mode="deterministic", precision="f64")
    got = ch.run()
    bad = 0
    for s in range(4):
        for e in range(z.E):
            gl, gs = ch.path(s, e); ol, os_ = o.path(s, e)
            if not (np.array_equal(gs, os_) and np.array_equal(gl, ol)):
                if bad < 3:
                    print("iter", k, "site", s, "edge", e, "gpu", gl.tolist(), gs.tolist(), "orc", ol.tolist(), os_.tolist(), "pieces", o.pieces(s, e)[0].tolist())
                bad += 1
    print("k", k, "bad", bad, "Q", ch.Q.ravel().tolist(), "rows equal", np.array_equal(got[:, 2:6], ref[:, 2:6]), np.abs(got[:, :2] - ref[:, :2]).max())
