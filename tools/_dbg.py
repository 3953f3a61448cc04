import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import parity
np.set_printoptions(linewidth=220, precision=9)
case, npx, nstep = sys.argv[1], tuple(int(x) for x in sys.argv[2].split(',')), int(sys.argv[3])
wg = parity.build_world(case, npx, nstep); wo = parity.build_world(case, npx, nstep)
parity.run_gpu(wg); parity.run_oracle(wo)
for r in range(wg.size):
    g, o = wg.view(r), wo.view(r)
    k = int(g.nftnd[0])
    if not k: continue
    print('rank', r, 'pairs', k)
    for sl in (20, 23, 47, 48, 71, 74, 75, 76, 77, 78, 79, 80, 31, 34):
        a, b = g.fric[sl-1, :k, 0], o.fric[sl-1, :k, 0]
        d = np.abs(a - b)
        i = int(np.argmax(d))
        print('  fric(%d): max|d|=%.3e at pair %d g=%.12e o=%.12e  max|o|=%.3e  x=%s' % (sl, d[i], i, a[i], b[i], np.abs(b).max(), g.meshCoor[:, g.nsmp[0, i, 0]-1]))
    break
