"""Generates tests/golden/checkpoint_schema.json from the six checkpoints the reference ships
(experiments/checkpoint/weights/Movielenslatest-small/<MODEL>/BPR/<args>/run_1/latest.pkl):
state_dict key names + shapes, Adam hyper-parameters, and the per-epoch metric trajectories.
Run in the authoring container only (needs /root/reference)."""
import glob
import json
import os

import numpy as np
import torch

ROOT = '/root/reference/experiments/checkpoint/weights/Movielenslatest-small'
out = {}
for path in sorted(glob.glob(os.path.join(ROOT, '*', 'BPR', '*', 'run_1', 'latest.pkl'))):
    model = path.split('/')[-5]
    ea = "'entity_aware': True" in path
    ck = torch.load(path, map_location='cpu', weights_only=False)
    sd = ck['model_states']['model']
    pg = ck['optim_states']['optim']['param_groups'][0]
    st = ck['optim_states']['optim']['state']
    first = st[sorted(st.keys())[0]]
    rec = ck['rec_metrics']
    out['%s/entity_aware=%s' % (model, ea)] = {
        'epoch': int(ck['epoch']),
        'state_dict': {k: list(v.shape) for k, v in sd.items()},
        'adam': {'lr': pg['lr'], 'weight_decay': pg['weight_decay'], 'betas': list(pg['betas']), 'eps': pg['eps'],
                 'step': int(first['step'])},
        'hr10_per_epoch': [float(v) for v in np.asarray(rec[0])[:, 5]],
        'ndcg10_per_epoch': [float(v) for v in np.asarray(rec[1])[:, 5]],
    }
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'checkpoint_schema.json'), 'w'), indent=1)
print({k: (len(v['state_dict']), v['adam']) for k, v in out.items()})
