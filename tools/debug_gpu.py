import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import trajtrack_mpcndqn_rlboost_b200 as t
from tests import oracle_lib as O
cfg = t.Configurator().to_ttmpc()
g = np.load('tests/golden/problem_default.npz')
bs = t.BatchSolver(cfg)
ev = bs.evaluate(g['p'], g['u'], g['c'], g['y'])
rel = lambda a,b: np.abs(a-b).max()/max(1.0,np.abs(b).max())
print('vs golden: f', rel(ev['f'], g['f']), 'F1', rel(ev['F1'], g['F1']), 'F2', rel(ev['F2'], g['F2']),
      'psi', rel(ev['psi'], g['psi']), 'grad', rel(ev['grad'], g['grad_psi']))
nbit = 0
for i in range(48):
    f,F2,ps,gr = O.eval_warp(cfg,g['u'][i],g['p'][i],g['c'][i],g['y'][i])
    same = (f == ev['f'][i]) and (ps == ev['psi'][i]) and np.array_equal(gr, ev['grad'][i]) and np.array_equal(F2, ev['F2'][i])
    nbit += same
    if not same:
        print('  mismatch', i, f-ev['f'][i], ps-ev['psi'][i], np.abs(gr-ev['grad'][i]).max(), np.abs(F2-ev['F2'][i]).max())
print('eval bit-identical to warp oracle:', nbit, '/ 48')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
p = t.scenes.make_scenes(n, cfg, seed=3, n_static=4, n_dynamic=3)
sol = bs.run(p); ref = O.solve_batch(cfg, p, threads=8, warp=True)
bit = np.array([np.array_equal(sol.solution[i], ref['u'][i]) and sol.cost[i]==ref['cost'][i] for i in range(n)])
print('solve bit-identical:', int(bit.sum()), '/', n, ' status equal:', int((sol.exit_status==ref['exit_status']).sum()),
      ' inner equal:', int((sol.num_inner_iterations==ref['inner']).sum()), ' y equal:', int(np.all(sol.lagrange_multipliers==ref['y'],axis=1).sum()),
      ' pred equal:', int(np.all(sol.pred_states.reshape(n,-1)==ref['pred'].reshape(n,-1),axis=1).sum()))
for i in np.where(~bit)[0][:10]:
    print(i, sol.exit_status[i], ref['exit_status'][i], 'inner', sol.num_inner_iterations[i], ref['inner'][i], 'cost %.9f %.9f' % (sol.cost[i], ref['cost'][i]),
          'du %.2e' % np.abs(sol.solution[i]-ref['u'][i]).max())
print('stats', bs.read_stats())
