import os, sys, traceback
sys.path.insert(0, "/root/repo")
os.environ["SEMSEG_B200_GRAPH_DEBUG"] = "1"
import torch
from tests import util
from semseg_b200 import graphs
m = util.build_pspnet(50, 21).cuda().train()
opt = torch.optim.SGD(m.parameters(), lr=0.01, momentum=0.9)
x, y = util.synth(2, 65, 65, 21, seed=1, device="cuda")
for k in range(6):
    try:
        _, ml, al = m(x, y)
        (ml + 0.4 * al).backward()
        opt.step(); opt.zero_grad()
        print("step", k, ml.item(), graphs.launches_per_step(m), flush=True)
    except Exception:
        traceback.print_exc()
        break
