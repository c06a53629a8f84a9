import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, cases, phylomap_b200 as pb
from phylomap_b200 import capi, synth
Q,pid=cases.q4(),np.full(4,.25)
tree=synth.yule_tree(10000,seed=4,mean_branch=0.1/1.2)
S=2048
st=synth.simulate_tip_states(tree,Q,pid,S,seed=7,device="cuda").cpu().numpy()
z=tree.with_states(st,segments=2)
ch=pb.Chain(capi.PM_V_BIGTREE,z,Q.copy(),pid,2.4,12,seed=3,precision="f32")
ch.run(12)
m=ch.piece_counts()
tot=m.size
for k in range(1,9): print(k, (m==k).sum()/tot)
print(">=9",(m>=9).sum()/tot)
