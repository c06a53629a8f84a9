"""One k_small_chain launch (BASELINE configs[0]: 100 tips, one character, 1 000 sweeps) for ncu:
    ncu --set full --import-source on --clock-control none -k regex:k_small_chain -c 1 -o gpurun_out/small python scripts/small_profile.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases
import phylomap_b200 as pb
from phylomap_b200 import capi
z = cases.tree2(T=100, S=int(os.environ.get("PM_SITES", 1)), seed=1, mean_branch=5.0)
ch = pb.Chain(capi.PM_V_PLAIN, z, cases.Q2, cases.PID2, 0.2, 1000, precision="f32", seed=5)
ch.run(1000)
ch.close()
