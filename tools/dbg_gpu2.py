import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load, x_err
from epipolicy_rl_b200.engine import BatchedEpidemic
from oracle.oracle import Oracle
name, prec, target = sys.argv[1], sys.argv[2], int(sys.argv[3])
d, z = load(name)
eng = BatchedEpidemic(d, 1, 0, prec); eng.get_state("attempts")
o = Oracle(d)
acts = z["rnd_actions"]
day = 0
for w, a in enumerate(acts):
    at = torch.from_numpy(a[None]).cuda()
    for _ in range(7):
        if day > target: break
        eng.step(at, 1); o.step(a, 1)
        if day >= target - 1:
            x = eng.get_state("X")[0]
            print("day", day, "err vs oracle", x_err(x, o.X, d), "err vs ref", x_err(x, z["rnd_X"][day], d))
            n = len(o.attempts)
            print(" oracle attempts", o.attempts.tolist())
            print(" gpu attempts   ", eng.get_state("attempts")[0, :n].tolist())
            fl = eng.get_state("flat")[0]; of = o.flat
            rel = np.abs(fl - of) / (np.abs(of) + 1e-6)
            k = np.argsort(-rel)[:6]
            print(" worst flat idx", k, rel[k], fl[k], of[k])
        day += 1
