import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import __graft_entry__ as ge
pkg = ge.load_package()
QI = pkg.qinfo
from test_gpu_qinfo_batch import random_density
for d in (8, 16):
    rng = np.random.default_rng(100 + d)
    count = 257
    full = np.array([random_density(rng, d) for _ in range(count)])
    defic = np.array([random_density(rng, d, rank=max(1, d // 2)) for _ in range(count)])
    got = QI.max_entropy_batch(defic)
    want = np.log2(max(1, d // 2))
    bad = np.nonzero(got != want)[0]
    print(d, 'bad', len(bad), np.unique(got)[:6])
    ev = QI.eigvalsh_batch(defic)
    for b in bad[:3]:
        print('  item', b, got[b], 'dev eig', ev[b], 'host eig', np.linalg.eigvalsh(defic[b]))
