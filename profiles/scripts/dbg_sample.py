import numpy as np, torch, sys, os
sys.path.insert(0, os.getcwd())
from vaura_b200.sampler import sample_logits
g = np.load('tests/golden/sampling_filters.npz')
logits = torch.from_numpy(g['logits'])
for key in g.files:
    if key.startswith('topk_'):
        temp, k = float(key.split('_t')[1].split('_k')[0]), int(key.split('_k')[1])
        _, probs = sample_logits(logits.cuda(), temp=temp, top_k=k, return_probs=True)
        p = probs.cpu().numpy().reshape(-1, 1024); r = g[key].reshape(-1, 1024)
        d = np.abs(p - r)
        print(key, 'max err', d.max(), 'rows bad', (d.max(1) > 2e-7).sum(), 'kept mine', (p > 0).sum(1)[:8], 'ref', (r > 0).sum(1)[:8])
        i = int(d.max(1).argmax()); j = int(d[i].argmax())
        print('   worst row', i, 'col', j, p[i, j], r[i, j], 'sum', p[i].sum())
