import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import parity
np.set_printoptions(linewidth=200, precision=6)
npx = (2,1,1)
nstep = 2
wg = parity.build_world('test.tpv8', npx, nstep)
wo = parity.build_world('test.tpv8', npx, nstep)
parity.run_gpu(wg); parity.run_oracle(wo)
for r in range(wg.size):
    g, o = wg.view(r), wo.view(r)
    print('rank', r)
    for name in ('velArr', 'dispArr', 'v1', 'stressArr', 'nodalForceArr', 'fnft'):
        a, b = getattr(g, name), getattr(o, name)
        d = np.abs(a - b)
        print('  ', name, 'max abs diff', d.max(), 'ref max', np.abs(b).max(), 'n bad', int((d > 1e-9 * np.abs(b).max()).sum()))
    k = int(g.nftnd[0])
    for sl in (70, 71, 73, 75, 76, 77, 78, 79):
        d = np.abs(g.fric[sl, :k, 0] - o.fric[sl, :k, 0])
        print('   fric', sl + 1, d.max(), np.abs(o.fric[sl, :k, 0]).max())
    # which elements have bad stress
    used = g.raw.stressUsed
    d = np.abs(g.stressArr[:used] - o.stressArr[:used])
    bad = np.argwhere(d > 1e-3).ravel()
    print('  bad stress entries', bad.size)
    if bad.size:
        sci = g.stressCompIndexArr
        el = np.unique(np.searchsorted(sci, bad, side='right') - 1)
        print('  bad elements', el.size, el[:10], 'types', g.elemTypeArr[el[:10]])
        for e in el[:4]:
            nodes = g.nodeElemIdRelation[:, e] - 1
            print('   elem', e, 'nodes', nodes, 'centroid', g.meshCoor[:, nodes].mean(axis=1))
            print('     g', g.stressArr[sci[e]:sci[e]+6]); print('     o', o.stressArr[sci[e]:sci[e]+6])
            print('     vel g', g.velArr[:, nodes].T.ravel()); print('     vel o', o.velArr[:, nodes].T.ravel())
