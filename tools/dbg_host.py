import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load
from epipolicy_rl_b200.engine import BatchedEpidemic
d, _ = load("COVID_B")
a = torch.rand((16, d.A), generator=torch.Generator().manual_seed(3))
e7, eh = BatchedEpidemic(d, 16, 0, "fp64"), BatchedEpidemic(d, 16, 0, "fp64")
o7, r7, _ = e7.step(a.cuda(), 7)
oh, rh, dh = eh.step_host(a.numpy(), 7)
print(r7.cpu().numpy()[:4], rh[:4], np.max(np.abs(r7.cpu().numpy()-rh)), dh)
print(e7.info.grid, e7.info.threads, eh.info.grid)
