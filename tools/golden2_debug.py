import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
import avoid_mpc_b200 as A
import helpers as H
g=H.golden2_groups()[(20,16)]
oW,ost,oit,ocost=H.oracle_solve_batch(20,16,0.05,g['params'],g['W0'])
for kern in ('warp','quad'):
    os.environ['AMPC_SOLVE_KERNEL']=kern
    h=A.Handle(N=20,K=16,dt=0.05,max_batch=len(g['ids']),max_points=16); h.set_solver_opts(max_iter=100)
    W,info=h.solve(g['prefix'],g['W0']); h.close()
    for sid in (1174,1327,1353,1097):
        j=g['sid'].index(sid)
        print(kern,sid,'gpu st',info['status'][j],'it',info['iters'][j],'cost %.6f kktd %.2e'%(info['cost'][j],info['kkt_dual'][j]),'| oracle st',ost[j],'it',oit[j],'cost %.6f'%ocost[j],'| linf %.2e'%np.abs(W[j]-oW[j]).max(), 'gold cost %.6f d_gold %.2e'%(g['cost'][j], np.abs(W[j]-g['w'][j]).max()))
    err=np.abs(W-oW).max(axis=1); print(kern,'n status differ',int((info['status']!=ost).sum()),'n err>1e-6',int((err>1e-6).sum()), 'iters differ', int((info['iters']!=oit).sum()))
