"""Where the two CUDA solve kernels differ from the CPU oracle on the 512 golden instances (debugging aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import avoid_mpc_b200 as A
import helpers as H
for (N, K), g in H.golden2_groups().items():
    dt = 0.05 if N == 20 else 1.0 / N
    oW, ost, oit, ocost = H.oracle_solve_batch(N, K, dt, g['params'], g['W0'])
    for kern in ('warp', 'quad'):
        os.environ['AMPC_SOLVE_KERNEL'] = kern
        h = A.Handle(N=N, K=K, dt=dt, max_batch=len(g['ids']), max_points=16); h.set_solver_opts(max_iter=100)
        W, info = h.solve(g['prefix'], g['W0']); h.close()
        cl = H.golden2_classify(W, g)
        for j, sid in enumerate(g['sid']):
            e = np.abs(W[j] - oW[j]).max()
            relc = abs(info['cost'][j] - g['cost'][j]) / abs(g['cost'][j])
            if e > 1e-6 or info['status'][j] != ost[j] or (cl[j] == 'same' and relc > 1e-6):
                print(N, K, kern, sid, cl[j], 'gpu st', info['status'][j], 'it', info['iters'][j], 'cost %.9f kktd %.2e' % (info['cost'][j], info['kkt_dual'][j]),
                      '| oracle st', ost[j], 'it', oit[j], 'cost %.9f' % ocost[j], '| linf vs oracle %.2e' % e,
                      '| gold cost %.9f d_gold %.2e stage %d' % (g['cost'][j], np.abs(W[j] - g['w'][j]).max(), g['stage'][j]))
