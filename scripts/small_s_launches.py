"""One character on the 100-tip tree of BASELINE configs[0], a few sweeps: run under
`ncu --metrics gpu__time_duration.sum` to see what each launch of a sweep costs when nothing is throughput-bound."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import cases
import phylomap_b200 as pb
from phylomap_b200 import synth
os.environ["PHYLOMAP_B200_GRAPH"] = "0"
z = synth.simulate_2_state_tree(101, synth.yule_tree(100, seed=1, mean_branch=5.0), cases.Q2, cases.PID2)
print(pb.sumstatMCMC(z, cases.Q2, cases.PID2, 0.2, 8, seed=5, precision="f32")[-1])
