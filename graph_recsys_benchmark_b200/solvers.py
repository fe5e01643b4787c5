"""BPR solver - drop-in for reference ``graph_recsys_benchmark/solvers.py`` (``BaseSolver``):
``generate_candidates`` (:21-31), ``metrics`` (:33-104), ``run`` (:106-330: seeds, model + Adam,
resume, epoch loop, metapath ablation, checkpoints, loggers).  Same constructor, same
``train_args`` / ``model_args`` / ``dataset_args`` keys, same return values and log lines.

What changes underneath:
  * ``metrics`` draws every user's negatives with the same global-numpy-RNG stream as the
    per-user ``np.random.choice`` loop upstream (one broadcast ``randint`` - verified identical),
    then scores and ranks ALL users in one launch of the K7 kernel instead of >= 6
    host<->device crossings per user;
  * the train loop takes DataLoader's own index order (RandomSampler + BatchSampler) but gathers
    a batch with one indexing op instead of per-sample ``__getitem__`` + collate, and reads the
    loss back every ``loss_sync_every`` steps instead of every step.
"""
import os
import random as rd
import time

import numpy as np
import torch
import tqdm
from torch.utils.data import BatchSampler, RandomSampler

from .utils import (get_opt_class, load_dataset, load_global_logger, load_model, save_global_logger, save_model,
                    instantwrite, clearcache)
from . import functional as F_

_FMT = ('HR@5: {:.4f}, HR@10: {:.4f}, HR@15: {:.4f}, HR@20: {:.4f}, '
        'NDCG@5: {:.4f}, NDCG@10: {:.4f}, NDCG@15: {:.4f}, NDCG@20: {:.4f}, AUC: {:.4f}, ')


def _fmt_metrics(HRs, NDCGs, AUC):
    return _FMT.format(HRs[0], HRs[5], HRs[10], HRs[15], NDCGs[0], NDCGs[5], NDCGs[10], NDCGs[15], AUC[0])


class BaseSolver(object):
    def __init__(self, model_class, dataset_args, model_args, train_args):
        self.model_class = model_class

        self.dataset_args = dataset_args
        self.model_args = model_args
        self.train_args = train_args

    def generate_candidates(self, dataset, u_nid):
        """reference solvers.py:21-31 (one user; kept for API parity)."""
        pos_i_nids = dataset.test_pos_unid_inid_map[u_nid]
        neg_i_nids = list(np.random.choice(dataset.neg_unid_inid_map[u_nid], size=(self.train_args['num_neg_candidates'],)))
        return pos_i_nids, neg_i_nids

    def generate_all_candidates(self, dataset):
        """Candidates of every evaluation user, in the reference's user order (dict order,
        solvers.py:54), drawing from the global numpy RNG exactly as the per-user
        ``np.random.choice(neg_map[u], size=(num_neg,))`` calls do: with replacement, i.e.
        ``randint(0, len(neg_map[u]))`` per user - a broadcast randint consumes the MT19937
        stream element by element in the same order.  Returns (users [U], cand [U, n_pos + num_neg])."""
        num_neg = self.train_args['num_neg_candidates']
        u_nids = list(dataset.test_pos_unid_inid_map.keys())
        if hasattr(dataset, 'kth_unseen') and not isinstance(dataset.neg_unid_inid_map, dict):
            # large graphs keep no per-user lists: same draws, pool[j] computed as the j-th unseen item
            u0 = dataset.type_accs['uid']
            if u_nids != list(range(u0, u0 + dataset.num_uids)):
                raise NotImplementedError('lazy candidate pools need the users in node-id order')
            pos = np.asarray([dataset.test_pos_unid_inid_map[u] for u in u_nids], dtype=np.int64)
            n_pos = pos.shape[1]
            sizes = dataset.unseen_counts()
            if n_pos == 0 or (sizes <= 0).any():
                raise ValueError("No pos or neg samples found in evaluation!")
            idx = np.random.randint(0, sizes[:, None], size=(len(u_nids), num_neg))
            cand = np.concatenate([pos, dataset.kth_unseen(idx)], axis=1)
            return np.asarray(u_nids, dtype=np.int64), cand, n_pos
        pools = [dataset.neg_unid_inid_map[u] for u in u_nids]
        pos = [dataset.test_pos_unid_inid_map[u] for u in u_nids]
        n_pos = len(pos[0])
        if any(len(p) != n_pos for p in pos):
            raise NotImplementedError('the batched ranker needs the same number of positives per user '
                                      '(leave-one-out gives exactly one, datasets/movielens.py:307)')
        sizes = np.array([len(p) for p in pools], dtype=np.int64)
        if n_pos == 0 or (sizes == 0).any():
            raise ValueError("No pos or neg samples found in evaluation!")
        idx = np.random.randint(0, sizes[:, None], size=(len(u_nids), num_neg))
        cand = np.empty((len(u_nids), n_pos + num_neg), dtype=np.int64)
        cand[:, :n_pos] = np.asarray(pos, dtype=np.int64)
        for r, pool in enumerate(pools):
            cand[r, n_pos:] = np.asarray(pool, dtype=np.int64)[idx[r]]
        return np.asarray(u_nids, dtype=np.int64), cand, n_pos

    def metrics(self, run, epoch, model, dataset, return_per_user=False):
        """reference solvers.py:33-104.  Returns (HR[16], NDCG[16], AUC[1], eval_loss[1]) as fp64
        numpy arrays (column means over users; index 5 is @10)."""
        device = self.train_args['device']
        users, cand, n_pos = self.generate_all_candidates(dataset)     # identical draws on every rank
        world, rank = 1, 0
        if torch.distributed.is_available() and torch.distributed.is_initialized() and not return_per_user:
            world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
        n_users = users.shape[0]
        if world > 1:                                     # users are sharded; 36 partial sums are all-reduced
            users, cand = users[rank::world], cand[rank::world]
        users_t = torch.from_numpy(np.ascontiguousarray(users)).to(device)
        cand_t = torch.from_numpy(np.ascontiguousarray(cand)).to(device)
        per_user, means, _ = F_.eval_rank(model.cached_repr, users_t, cand_t, n_pos, model.fc1.weight,
                                          model.fc1.bias, model.fc2.weight, model.fc2.bias)
        if world > 1:
            sums = means * float(users.shape[0])
            torch.distributed.all_reduce(sums)
            means = sums / float(n_users)
        m = means.cpu().numpy()                          # the only device->host read of an evaluation
        out = (m[0:16].copy(), m[16:32].copy(), m[32:33].copy(), m[33:34].copy())
        if return_per_user:
            return out, per_user
        return out

    def _batches(self, dataset):
        """Index batches in torch.utils.data.DataLoader(shuffle=True) order (solvers.py:195-200)."""
        # DataLoader.__iter__ draws its worker base seed from the global torch generator before the
        # sampler draws the permutation seed; consume it too so the permutation is the same one.
        torch.empty((), dtype=torch.int64).random_()
        sampler = BatchSampler(RandomSampler(dataset), batch_size=self.train_args['batch_size'], drop_last=False)
        for indices in sampler:
            yield dataset.get_batch(indices) if hasattr(dataset, 'get_batch') else \
                torch.stack([dataset[i] for i in indices], dim=0)

    def _device_sampler(self, dataset, device, run):
        key = (id(dataset), str(device), run)
        cached = getattr(self, '_sampler_cache', None)
        if cached is None or cached[0] != key:
            from .sampling import DeviceBprSampler
            self._sampler_cache = cached = (key, DeviceBprSampler(dataset, device, seed=2019 + run))
        return cached[1]

    def train_epoch(self, run, epoch, model, optimizer, dataset, max_steps=None):
        device = self.train_args['device']
        sync_every = self.train_args.get('loss_sync_every', 50)
        model.train()
        losses, pending = [], []
        # the captured step is kept across epochs (same model, optimizer and batch shape): one capture per run
        cached = getattr(self, '_graphed', None)
        graphed = cached[1] if cached is not None and cached[0] == (id(model), id(optimizer)) else None
        bs = self.train_args['batch_size']
        if self.train_args.get('device_sampling', False):
            # train_args['device_sampling']: negatives and entity columns are drawn on the GPU (sampling.py);
            # same distributions as the host loops below, but not the reference's host RNG stream
            sampler = self._device_sampler(dataset, device, run)
            order = sampler.permutation(epoch)
            n_rows = len(sampler)
            batches = (sampler.rows(order[s:s + bs], epoch) for s in range(0, n_rows, bs))
        else:
            dataset.cf_negative_sampling()
            n_rows = len(dataset)
            batches = self._batches(dataset)
        n_batches = (n_rows + bs - 1) // bs
        train_bar = tqdm.tqdm(batches, total=n_batches, disable=self.train_args.get('quiet', False))
        for step, batch in enumerate(train_bar):
            if max_steps is not None and step >= max_steps:
                break
            batch = batch.to(device, non_blocking=True)
            if self.train_args.get('cuda_graph', False) and (step > 0 or graphed is not None):
                # train_args['cuda_graph']: after one eager step the whole step is replayed as one CUDA graph
                # (graphed.py; the optimizer must have been built with capturable=True)
                if graphed is None:
                    from .graphed import GraphedTrainStep
                    graphed = GraphedTrainStep(model, optimizer, batch)    # trains on this batch, then captures
                    self._graphed = ((id(model), id(optimizer)), graphed)
                    pending.append(graphed.first_loss)
                else:
                    pending.append(graphed(batch).clone())
            else:
                optimizer.zero_grad()
                loss = model.loss(batch)
                loss.backward()
                optimizer.step()
                pending.append(loss.detach())
                del loss        # a live loss keeps this step's autograd nodes (default stream) alive, which breaks a later capture
            if len(pending) >= sync_every:
                losses.extend(torch.stack(pending).cpu().tolist())
                pending = []
                train_bar.set_description('Run: {}, epoch: {}, train loss: {:.4f}'.format(run, epoch, np.mean(losses)))
        if pending:
            losses.extend(torch.stack(pending).cpu().tolist())
        return float(np.mean(losses)) if losses else float('nan'), losses

    def _post_run_ablation(self, dataset, logger_file):
        """reference solvers.py:333-392: after all runs, reload run 1's latest checkpoint and evaluate the model
        with each metapath's channel zeroed in turn ('exclude path' lines)."""
        run = 1
        epoch = 20 if self.dataset_args['dataset'] == 'Yelp' else 30
        seed = 2019 + run
        rd.seed(seed)
        np.random.seed(seed)
        torch.manual_seed(seed)
        torch.cuda.manual_seed(seed)
        self.model_args['num_nodes'] = dataset.num_nodes
        self.model_args['dataset'] = dataset
        model = self.model_class(**self.model_args).to(self.train_args['device'])
        optimizer = get_opt_class(self.train_args['opt'])(params=model.parameters(), lr=self.train_args['lr'],
                                                          weight_decay=self.train_args['weight_decay'])
        weights_path = os.path.join(self.train_args['weights_folder'], 'run_{}'.format(str(run)))
        os.makedirs(weights_path, exist_ok=True)
        model, optimizer, _, _ = load_model(os.path.join(weights_path, 'latest.pkl'), model, optimizer,
                                            self.train_args['device'])
        for metapath_idx in range(len(self.model_args['meta_path_steps'])):
            model.eval(metapath_idx)
            HRs, NDCGs, AUC, _ = self.metrics(run, epoch, model, dataset)
            msg = 'Run: {}, epoch: {}, exclude path:{}, '.format(run, epoch, metapath_idx) + _fmt_metrics(HRs, NDCGs, AUC) + '\n'
            print(msg)
            logger_file.write(msg)
            instantwrite(logger_file)
        del model, optimizer
        clearcache()

    def run(self):
        global_logger_path = self.train_args['logger_folder']
        if not os.path.exists(global_logger_path):
            os.makedirs(global_logger_path, exist_ok=True)
        global_logger_file_path = os.path.join(global_logger_path, 'global_logger.pkl')
        HRs_per_run_np, NDCGs_per_run_np, AUC_per_run_np, train_loss_per_run_np, eval_loss_per_run_np, last_run = \
            load_global_logger(global_logger_file_path)

        dataset = load_dataset(self.dataset_args)

        logger_file_path = os.path.join(global_logger_path, 'logger_file.txt')
        with open(logger_file_path, 'a') as logger_file:
            start_run = last_run + 1
            if start_run <= self.train_args['runs']:
                for run in range(start_run, self.train_args['runs'] + 1):
                    # Fix the random seed (solvers.py:123-127)
                    seed = 2019 + run
                    rd.seed(seed)
                    np.random.seed(seed)
                    torch.manual_seed(seed)
                    torch.cuda.manual_seed(seed)

                    if self.model_args['model_type'] == 'Graph':
                        if self.model_args['if_use_features']:
                            raise NotImplementedError('Feature not implemented!')
                        self.model_args['num_nodes'] = dataset.num_nodes
                        self.model_args['dataset'] = dataset
                    else:
                        raise NotImplementedError('only the PEAGNN graph models are in scope')

                    model = self.model_class(**self.model_args).to(self.train_args['device'])
                    # train_args['demand_driven_loss'] (default on): loss() computes only the representation rows
                    # its batch reads - same loss and gradients as the full propagation of models/base.py:45
                    model.demand_driven_loss = bool(self.train_args.get('demand_driven_loss', True))

                    opt_class = get_opt_class(self.train_args['opt'])
                    opt_kwargs = {'capturable': True} if self.train_args.get('cuda_graph', False) else {}
                    optimizer = opt_class(params=model.parameters(), lr=self.train_args['lr'],
                                          weight_decay=self.train_args['weight_decay'], **opt_kwargs)

                    weights_path = os.path.join(self.train_args['weights_folder'], 'run_{}'.format(str(run)))
                    if not os.path.exists(weights_path):
                        os.makedirs(weights_path, exist_ok=True)
                    weights_file = os.path.join(weights_path, 'latest.pkl')
                    model, optimizer, last_epoch, rec_metrics = load_model(weights_file, model, optimizer,
                                                                           self.train_args['device'])
                    HRs_per_epoch_np, NDCGs_per_epoch_np, AUC_per_epoch_np, train_loss_per_epoch_np, \
                        eval_loss_per_epoch_np = rec_metrics

                    torch.cuda.synchronize()
                    start_epoch = last_epoch + 1
                    if start_epoch == 1 and self.train_args['init_eval']:
                        model.eval()
                        with torch.no_grad():
                            HRs, NDCGs, AUC, eval_loss = self.metrics(run, 0, model, dataset)
                        msg = 'Initial performance ' + _fmt_metrics(HRs, NDCGs, AUC) + \
                              'eval loss: {:.4f} \n'.format(eval_loss[0])
                        print(msg)
                        logger_file.write(msg)
                        instantwrite(logger_file)
                        clearcache()

                    t_start = time.perf_counter()
                    train_loss, eval_loss = float('nan'), np.array([float('nan')])
                    if start_epoch <= self.train_args['epochs']:
                        for epoch in range(start_epoch, self.train_args['epochs'] + 1):
                            train_loss, _ = self.train_epoch(run, epoch, model, optimizer, dataset)

                            if model.__class__.__name__[:3] == 'PEA' and self.train_args['metapath_test']:
                                if (self.dataset_args['dataset'] == 'Movielens' and epoch == 30) or \
                                        (self.dataset_args['dataset'] == 'Yelp' and epoch == 20):
                                    for metapath_idx in range(len(self.model_args['meta_path_steps'])):
                                        model.eval(metapath_idx)
                                        HRs, NDCGs, AUC, eval_loss = self.metrics(run, epoch, model, dataset)
                                        msg = 'Run: {}, epoch: {}, exclude path:{}, '.format(run, epoch, metapath_idx) + \
                                              _fmt_metrics(HRs, NDCGs, AUC) + \
                                              'train loss: {:.4f}, eval loss: {:.4f} \n'.format(train_loss, eval_loss[0])
                                        print(msg)
                                        logger_file.write(msg)

                            model.eval()
                            with torch.no_grad():
                                HRs, NDCGs, AUC, eval_loss = self.metrics(run, epoch, model, dataset)

                            HRs_per_epoch_np = np.vstack([HRs_per_epoch_np, HRs])
                            NDCGs_per_epoch_np = np.vstack([NDCGs_per_epoch_np, NDCGs])
                            AUC_per_epoch_np = np.vstack([AUC_per_epoch_np, AUC])
                            train_loss_per_epoch_np = np.vstack([train_loss_per_epoch_np, np.array([train_loss])])
                            eval_loss_per_epoch_np = np.vstack([eval_loss_per_epoch_np, np.array([eval_loss])])

                            rec = (HRs_per_epoch_np, NDCGs_per_epoch_np, AUC_per_epoch_np, train_loss_per_epoch_np,
                                   eval_loss_per_epoch_np)
                            if epoch in self.train_args['save_epochs']:
                                save_model(os.path.join(weights_path, '{}.pkl'.format(epoch)), model, optimizer, epoch,
                                           rec_metrics=rec)
                            if epoch > self.train_args['save_every_epoch']:
                                save_model(os.path.join(weights_path, 'latest.pkl'), model, optimizer, epoch,
                                           rec_metrics=rec)
                            msg = 'Run: {}, epoch: {}, '.format(run, epoch) + _fmt_metrics(HRs, NDCGs, AUC) + \
                                  'train loss: {:.4f}, eval loss: {:.4f} \n'.format(train_loss, eval_loss[0])
                            print(msg)
                            logger_file.write(msg)
                            instantwrite(logger_file)
                            clearcache()

                        torch.cuda.synchronize()
                    t_end = time.perf_counter()

                    HRs_per_run_np = np.vstack([HRs_per_run_np, np.max(HRs_per_epoch_np, axis=0)])
                    NDCGs_per_run_np = np.vstack([NDCGs_per_run_np, np.max(NDCGs_per_epoch_np, axis=0)])
                    AUC_per_run_np = np.vstack([AUC_per_run_np, np.max(AUC_per_epoch_np, axis=0)])
                    train_loss_per_run_np = np.vstack([train_loss_per_run_np, np.mean(train_loss_per_epoch_np, axis=0)])
                    eval_loss_per_run_np = np.vstack([eval_loss_per_run_np, np.mean(eval_loss_per_epoch_np, axis=0)])

                    save_global_logger(global_logger_file_path, HRs_per_run_np, NDCGs_per_run_np, AUC_per_run_np,
                                       train_loss_per_run_np, eval_loss_per_run_np)
                    msg = 'Run: {}, Duration: {:.4f}, '.format(run, t_end - t_start) + \
                          _fmt_metrics(np.max(HRs_per_epoch_np, axis=0), np.max(NDCGs_per_epoch_np, axis=0),
                                       np.max(AUC_per_epoch_np, axis=0)) + \
                          'train_loss: {:.4f}, eval loss: {:.4f}\n'.format(train_loss_per_epoch_np[-1][0],
                                                                         eval_loss_per_epoch_np[-1][0])
                    print(msg)
                    logger_file.write(msg)
                    instantwrite(logger_file)

                    del model, optimizer, rec_metrics
                    from .graph import clear_cache
                    clear_cache()                      # the next run's model builds its own device copies of the relations
                    self._graphed = None
                    clearcache()

            if str(self.dataset_args.get('model', ''))[:3] == 'PEA' and self.train_args['metapath_test']:
                self._post_run_ablation(dataset, logger_file)

            msg = 'Overall HR@5: {:.4f}, HR@10: {:.4f}, HR@15: {:.4f}, HR@20: {:.4f}, ' \
                  'NDCG@5: {:.4f}, NDCG@10: {:.4f}, NDCG@15: {:.4f}, NDCG@20: {:.4f}, AUC: {:.4f}, ' \
                  'train loss: {:.4f}, eval loss: {:.4f}\n'.format(
                      HRs_per_run_np.mean(axis=0)[0], HRs_per_run_np.mean(axis=0)[5], HRs_per_run_np.mean(axis=0)[10],
                      HRs_per_run_np.mean(axis=0)[15], NDCGs_per_run_np.mean(axis=0)[0],
                      NDCGs_per_run_np.mean(axis=0)[5], NDCGs_per_run_np.mean(axis=0)[10],
                      NDCGs_per_run_np.mean(axis=0)[15], AUC_per_run_np.mean(axis=0)[0],
                      train_loss_per_run_np.mean(axis=0)[0], eval_loss_per_run_np.mean(axis=0)[0]) \
                if HRs_per_run_np.shape[0] else 'Overall: no runs\n'
            print(msg)
            logger_file.write(msg)
            instantwrite(logger_file)
