"""Independent float64 torch re-derivation of the reference's NLP objective
(roswrapper/ros/src/avoid_mpc/tools/mpc_obstacle_casadi.py:162-214), written against
the script, not against oracle/nlp_oracle.c.  Used to cross-check the oracle's closed-form
gradient and Hessian with autograd, and to mint tests/golden/nlp_golden.npz."""
import numpy as np
import torch

torch.set_default_dtype(torch.float64)


def objective(N, K, w, p):
    ref = p[10:10 + 10 * N]
    obst = p[10 + 10 * N:10 + 10 * N + 3 * K * N]
    target = p[10 + 10 * N + 3 * K * N:20 + 10 * N + 3 * K * N]
    weights = p[-26:-1]
    radius = p[-1]
    Qg, Qp, Qu = torch.diag(weights[0:10]), torch.diag(weights[10:20]), torch.diag(weights[20:24])
    lam = weights[24]
    X = [w[14 * k:14 * k + 10] for k in range(N + 1)]
    U = [w[14 * k + 10:14 * k + 14] for k in range(N)]
    obj = 0
    for k in range(N):
        vi = X[k + 1][4:7]
        if k >= N - 1:
            d = X[k + 1] - target
            obj = obj + d @ Qg @ d
        else:
            xt = ref[10 * k:10 * k + 10]
            cy, sy = torch.cos(xt[3]), torch.sin(-xt[3])
            one, zero = torch.ones(()), torch.zeros(())
            rows = [[one if i == j else zero for j in range(10)] for i in range(10)]
            rows[0][0], rows[0][1], rows[1][0], rows[1][1] = cy, -sy, sy, cy
            rows[4][4], rows[4][5], rows[5][4], rows[5][5] = cy, -sy, sy, cy
            R = torch.stack([torch.stack(r) for r in rows])
            for j in range(K):
                po = obst[3 * K * k + 3 * j:3 * K * k + 3 * j + 3]
                v2o = po - X[k + 1][0:3]
                s = torch.abs(torch.dot(vi, v2o / torch.linalg.norm(v2o)))
                dist = torch.linalg.norm(v2o) - radius
                obj = obj + lam * torch.log(1 + torch.exp(dist * -32)) * s
            dp = X[k + 1] - xt
            obj = obj + (R @ dp) @ Qp @ (R @ dp)
        du = U[k] - torch.tensor([0, 0, 9.81, 0.0])
        obj = obj + du @ Qu @ du
    return obj


def f_grad_hess(N, K, w, p):
    pt = torch.tensor(np.asarray(p, dtype=np.float64))
    wt = torch.tensor(np.asarray(w, dtype=np.float64), requires_grad=True)
    f = objective(N, K, wt, pt)
    g, = torch.autograd.grad(f, wt)
    H = torch.autograd.functional.hessian(lambda ww: objective(N, K, ww, pt), wt.detach())
    return float(f.detach()), g.numpy(), H.numpy()
