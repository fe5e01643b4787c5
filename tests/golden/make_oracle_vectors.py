"""Freezes oracle outputs on the seeded 'tiny' synthetic HIN into tests/golden/oracle_vectors.pt:
for PEAGCN / PEAGAT / PEASage(entity-aware) - fused representation rows, BPR loss, a gradient
slice; plus the integer artefacts of the same seeds (CSR of user2item, sampled triples, evaluation
candidates, per-user ranks).  Run on CPU:  python tests/golden/make_oracle_vectors.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from helpers import oracle_model_for                                   # noqa: E402
from oracle import graph as ograph, solver as osolver                  # noqa: E402
from graph_recsys_benchmark_b200.datasets import SyntheticHIN          # noqa: E402


def build():
    out = {}
    for kind, ea in (('gcn', False), ('gat', False), ('sage', True)):
        ds = SyntheticHIN('tiny', seed=7, entity_aware=ea)
        torch.manual_seed(2020)
        model = oracle_model_for(ds, kind, entity_aware=ea)
        osolver.seed_everything(1)
        ds.cf_negative_sampling()
        batch = ds.get_batch(list(range(128)))
        model.train()
        loss = model.loss(batch)
        loss.backward()
        model.eval()
        np.random.seed(99)
        (hr, nd, auc, el), per = osolver.metrics(model, ds, 99, return_per_user=True)
        out[kind] = dict(
            entity_aware=ea, batch=batch.clone(), loss=float(loss.item()),
            repr_rows=model.cached_repr[:8].detach().clone(), repr_sum=float(model.cached_repr.double().sum()),
            x_grad_rows=model.x.grad[:4].clone(), x_grad_abs_sum=float(model.x.grad.double().abs().sum()),
            hr10=float(hr[5]), ndcg10=float(nd[5]), auc=float(auc[0]), eval_loss=float(el[0]),
            ranks=torch.from_numpy(per['ranks']).clone())
    ds = SyntheticHIN('tiny', seed=7)
    u2i = ds.edge_index_nps['user2item']
    rp, col, eid = ograph.csr_by_key(u2i[1].astype(np.int64), u2i[0].astype(np.int64), ds.num_nodes, True)
    out['csr_user2item_by_target'] = dict(rowptr=torch.from_numpy(rp), col=torch.from_numpy(col), eid=torch.from_numpy(eid))
    osolver.seed_everything(1)
    ds.cf_negative_sampling()
    out['train_triples_first_64'] = ds.train_data[:64].clone()
    np.random.seed(99)
    pos, neg = osolver.generate_candidates(ds, 0, 99)
    out['candidates_user0'] = torch.tensor(pos + [int(v) for v in neg])
    return out


if __name__ == '__main__':
    torch.save(build(), os.path.join(HERE, 'oracle_vectors.pt'))
    print('written', os.path.getsize(os.path.join(HERE, 'oracle_vectors.pt')), 'bytes')
