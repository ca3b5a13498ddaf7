import sys, ctypes as C
sys.path.insert(0,'/root/repo')
import torch, bench
from tscd_b200 import ops, selection, stage, weights, _lib as L
dev=torch.device('cuda',0); B=64
cfg = stage.StageConfig(num_classes=bench.C, selection=selection.SelectionConfig(mode="A", pre_k=750, top_k=30))
st = stage.AggregationStage(cfg, weights.random_state_dict(bench.C, 256, seed=2024), device=dev)
inp = bench.synth_s1(B, dev, seed=2024); head, feats = bench.views_of(inp, ops)
te = torch.cat([weights.timing_signal_1d(torch.arange(8), 256)] * B, 0).to(dev)
for _ in range(3): st.forward(head, feats, torch.float16, te, B, 32, 8)
torch.cuda.synchronize()
out=(C.c_longlong*8)(); L.lib().tscd_debug_chain_clocks(out,1)
st.forward(head, feats, torch.float16, te, B, 32, 8); torch.cuda.synchronize()
L.lib().tscd_debug_chain_clocks(out,0)
names=["reindex+copy","qin","qproj","normalise","attention","norms/state","carry"]
tot=sum(out[:7])
for n,v in zip(names,out): print(f"{n:14s} {v/1.9e3:8.1f} us  {100*v/tot:5.1f}%")
print("total us", tot/1.9e3)
