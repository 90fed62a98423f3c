import sys, numpy as np, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import load, x_err, cost_err
from epipolicy_rl_b200.engine import BatchedEpidemic
name, prec = sys.argv[1], sys.argv[2]
d, z = load(name)
eng = BatchedEpidemic(d, 2, 0, prec)
acts = np.stack([z["rnd_actions"], z["lax_actions"]], 1)
X = np.stack([z["rnd_X"], z["lax_X"]], 1); nf = np.stack([z["rnd_nfev"], z["lax_nfev"]], 1)
day = 0; shown = 0
for w, a in enumerate(acts):
    at = torch.from_numpy(a).cuda()
    for _ in range(7):
        if day >= d.horizon: break
        eng.step(at, 1)
        x = eng.get_state("X"); tr = eng.get_state("trace")
        for e in range(2):
            err = x_err(x[e], X[day, e], d)
            if (err > 1e-9 or tr[e,0] != nf[day,e]) and shown < 6:
                shown += 1
                pop = d.pop.reshape(1, d.L, d.G)
                rel = np.abs(x[e] - X[day, e]) / (np.abs(X[day, e]) + 1e-6 * pop)
                c, l, g = np.unravel_index(np.argmax(rel), rel.shape)
                print(f"day {day} env {e} err {err:.3e} at c={d.meta['compartments'][c]} l={l} g={g} gpu={x[e,c,l,g]!r} ref={X[day,e,c,l,g]!r} trace={tr[e]} ref nfev={nf[day,e]} lasterr={eng.get_state('last_err')[e]}")
        day += 1
print("done")
