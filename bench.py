#!/usr/bin/env python
"""bench.py - BPR triples/s of one PEAGNN train step at the MovieLens-25M shape (BASELINE.json), with the
rooflines of the kernels that make up the step, or (``--phase eval``) users/s of the HR/NDCG evaluation.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl product|reference] [--phase train|eval]
                  [--workload ml-25m|yelp|ml-small|...] [--model gcn|gat|sage] [--batch 4096]

train: a "step" = one pass of the hot path over one batch of synthetic input: full-graph propagation over
every metapath (forward + backward), fused scoring + BPR loss, Adam step - what reference solvers.py:213-218
does per batch.  eval: a "step" = ``model.eval()`` (one propagation) + ``BaseSolver.metrics`` over ALL users
(reference solvers.py:246-248, :33-104).  One JSON line on stdout (rank 0).

  value     device-resident inputs, CUDA-event timed, max over ranks
  e2e       the same step through the public API with the inputs in pinned HOST memory (H2D copy inside
            the timed region) and the result read back to the host every step
  roofline / roofline_projection   per-launch CUDA events of the kernel families, taken INSIDE a timed region
            of their own: K replays of the step graph captured with external event nodes (N = 1) or K eager
            steps (N > 1); algorithmic bytes per SURVEY.md 8(d) / DESIGN.md section 4.  The aggregation is
            reported against the measured HBM copy peak (the contract figure) AND against the random-row
            gather rate measured live on this GPU by ``peagnn_probe_gather`` (its tables are L2-resident)
  cpu_baseline / --impl reference   the CPU oracle (oracle/: the reference's path restated on torch CPU ops)
            on the host cores, on the SAME graph, a bounded number of steps
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LITE = {'ml-25m': 'ml-25m-lite', 'yelp': 'yelp-lite'}
CPU_FULL_GRAPH_GB = {'ml-25m': 110.0, 'yelp': 24.0}     # host RAM the CPU oracle needs for one train step
CPU_BUDGET_S = 150.0                                     # wall-clock bound of the reference arm's stepping (18 s per full-graph step)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='product', choices=['product', 'reference'])
    ap.add_argument('--phase', default='train', choices=['train', 'eval'])
    ap.add_argument('--workload', default='ml-25m')
    ap.add_argument('--model', default='gcn', choices=['gcn', 'gat', 'sage'])
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-cuda-graph', action='store_true', help='launch every kernel eagerly instead of replaying the captured step')
    ap.add_argument('--no-strong', action='store_true', help='skip the fixed-global-batch leg at N > 1')
    ap.add_argument('--gather-dtype', default='fp32', choices=['fp32', 'bf16'],
                    help='bf16: the first-step aggregations gather a bf16 copy of their table (opt-in mode with a stated tolerance; '
                         'the headline stays fp32)')
    ap.add_argument('--full-propagation', action='store_true',
                    help='loss() propagates every row of the last step (the reference\'s literal schedule) instead of the batch rows only')
    ap.add_argument('--prewarm', type=float, default=2.0, help='seconds of untimed steps before the warm-up')
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index, enabled=True):
        self.gpu, self.lines, self.proc, self.enabled = gpu_index, [], None, enabled

    def __enter__(self):
        if not self.enabled:
            return self
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu=' + self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def make_batches(ds, B, count, seed):
    """count x [B, 3] BPR triples: positives drawn from user2item, negatives as the reference's
    'random' strategy (np.random.randint over the item id range, movielens.py:923-927)."""
    rng = np.random.RandomState(seed)
    u2i = ds.edge_index_nps['user2item']
    sel = rng.randint(0, u2i.shape[1], size=(count, B))
    neg = rng.randint(ds.type_accs['iid'], ds.type_accs['iid'] + ds.num_iids, size=(count, B))
    out = np.stack([u2i[0][sel].astype(np.int64), u2i[1][sel].astype(np.int64), neg.astype(np.int64)], axis=-1)
    return torch.from_numpy(out)


def n_metapaths(ds):
    from graph_recsys_benchmark_b200.utils import metapath_table
    return len(metapath_table({'dataset': ds.dataset, 'name': ds.name}))


def workload_config(args, ds=None, workload=None):
    wl = workload or args.workload
    what = 'BPR train step' if args.phase == 'train' else 'evaluation pass (propagation + ranking of every user, 1 + 99 candidates)'
    cfg = {'workload': '%s / PEA%s %s%s, emb 64, hidden 64, repr 16'
                       % (wl, args.model.upper(), what, (', %d metapaths x 2 steps' % n_metapaths(ds)) if ds is not None else ''),
           'l2_between_iterations': 'inputs larger than L2 (CSR + activations > 126 MB)',
           'arithmetic': 'fp32 storage and accumulation everywhere; the 64/16-wide projections run on the tensor '
                         'cores as a 3-pass TF32 split (hi*hi + hi*lo + lo*hi), fp32-accurate, same 1e-5 parity bound'}
    if getattr(args, 'gather_dtype', 'fp32') == 'bf16':
        cfg['arithmetic'] = ('OPT-IN bf16 mode: the two first-step tables (x and its gradient-side counterpart) are gathered from '
                             'bf16 copies, everything else as the fp32 mode; not the headline configuration')
    if args.phase == 'train':
        cfg.update(batch_per_gpu=args.batch, optimizer='Adam(lr=1e-3, weight_decay=1e-3, fused)', negatives='random',
                   last_step_rows=('every row (reference schedule)' if args.full_propagation else
                                   'the rows loss() reads (the batch\'s users and items, models/base.py:209-210); every '
                                   'computed value, the loss and all gradients equal the full propagation\'s'))
    return cfg


# ---------------------------------------------------------------------------------------------
# CPU legs: the oracle on the host cores.  These are the only places bench.py touches oracle/.
# ---------------------------------------------------------------------------------------------
def cpu_workload(workload):
    """The graph the CPU oracle can hold: the full workload when the host has the RAM for its [E, F]
    message tensors, otherwise the 1/10-edge 'lite' graph of the same schema - and then it SAYS so."""
    need = CPU_FULL_GRAPH_GB.get(workload)
    if need is None:
        return workload, None
    try:
        import psutil
        avail = psutil.virtual_memory().available / 2 ** 30
    except Exception:                                        # noqa: BLE001
        avail = 0.0
    if avail >= need:
        return workload, None
    return LITE[workload], 'host has %.0f GB available, the full graph needs ~%.0f GB' % (avail, need)


def oracle_model(ds, kind):
    from oracle import graph as ograph
    from oracle.models import OraclePEAModel
    from graph_recsys_benchmark_b200.utils.factory import default_model_args
    mk = default_model_args(ds, kind)
    eil = ograph.metapath_edge_index_list(ds, ds.dataset, ds.name)
    return OraclePEAModel(kind, ds.num_nodes, eil, mk['meta_path_steps'], emb_dim=mk['emb_dim'],
                          hidden_size=mk['hidden_size'], repr_dim=mk['repr_dim'], num_heads=mk.get('num_heads', 1),
                          dropout=0.0, channel_aggr=mk['channel_aggr'], entity_aware=False, entity_aware_coff=0.1)


def cpu_train_steps(ds, kind, B, steps, warmup, threads, budget_s):
    """seconds per oracle train step (zero_grad, loss, backward, Adam, loss.item()), steps actually timed."""
    torch.set_num_threads(threads)
    torch.manual_seed(2020)
    model = oracle_model(ds, kind)
    opt = torch.optim.Adam(model.parameters(), lr=1e-3, weight_decay=1e-3)
    batches = make_batches(ds, B, steps + warmup, seed=7)
    model.train()
    times, t_begin, done_warm, dt = [], time.perf_counter(), 0, 0.0
    for k in range(steps + warmup):
        elapsed = time.perf_counter() - t_begin
        if times and elapsed + dt > budget_s:                         # stay inside the wall-clock bound
            break
        t0 = time.perf_counter()
        opt.zero_grad()
        loss = model.loss(batches[k])
        loss.backward()
        opt.step()
        last = loss.item()
        dt = time.perf_counter() - t0
        if k >= warmup or elapsed + 3 * dt > budget_s:                # no time left to warm up any further
            times.append(dt)
        else:
            done_warm += 1
    return float(np.mean(times)), len(times), done_warm, last


def cpu_eval_pass(ds, kind, threads, sample_users):
    """(seconds of one no-grad propagation, seconds per user of the reference's per-user metrics loop)."""
    from oracle import solver as osolver
    torch.set_num_threads(threads)
    torch.manual_seed(2020)
    model = oracle_model(ds, kind)
    t0 = time.perf_counter()
    model.eval()
    t_prop = time.perf_counter() - t0

    class _Head(object):                                              # the first sample_users users of the dataset
        def __init__(self, ds, n):
            self.neg_unid_inid_map = ds.neg_unid_inid_map
            keys = list(ds.test_pos_unid_inid_map.keys())[:n]
            self.test_pos_unid_inid_map = {k: ds.test_pos_unid_inid_map[k] for k in keys}
    head = _Head(ds, sample_users)
    np.random.seed(5)
    t0 = time.perf_counter()
    osolver.metrics(model, head, 99)
    t_user = (time.perf_counter() - t0) / len(head.test_pos_unid_inid_map)
    return t_prop, t_user, len(head.test_pos_unid_inid_map)


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from graph_recsys_benchmark_b200.datasets import SyntheticHIN
    threads = os.cpu_count() or 1
    wl, why = cpu_workload(args.workload)
    ds = SyntheticHIN(wl, seed=1234)
    if args.phase == 'eval':
        t_prop, t_user, n_users = cpu_eval_pass(ds, args.model, threads, sample_users=min(ds.num_uids, 1500))
        t_pass = t_prop + ds.num_uids * t_user
        value = ds.num_uids / t_pass
        sample = ('CPU oracle on the %s graph: one no-grad propagation (%.2f s) + the per-user metrics loop on the first %d '
                  'of %d users (%.3f ms per user), pass time = propagation + users x per-user time'
                  % (wl, t_prop, n_users, ds.num_uids, t_user * 1e3))
        line = {'impl': 'reference', 'metric': 'eval_users_per_sec', 'value': value, 'unit': 'users/s', 'n_gpus': args.gpus,
                'steps': 1, 'warmup': 0, 'ms_per_step': t_pass * 1e3, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': workload_config(args, ds, wl)}
    else:
        t_step, steps, warm, last = cpu_train_steps(ds, args.model, args.batch, max(1, args.steps), max(0, args.warmup),
                                                    threads, CPU_BUDGET_S)
        value = args.batch / t_step
        sample = ('%d timed step(s) after %d warm-up step(s) of the CPU oracle (oracle/: the reference path on torch CPU ops, '
                  'pinned to the reference code by tests/test_reference_pinning.py) on the %s graph, %.2f s per step, '
                  'stepping bounded to %.0f s of wall clock' % (steps, warm, wl, t_step, CPU_BUDGET_S))
        line = {'impl': 'reference', 'metric': 'bpr_triples_per_sec', 'value': value, 'unit': 'triples/s',
                'n_gpus': args.gpus, 'steps': steps, 'warmup': warm, 'ms_per_step': t_step * 1e3,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': workload_config(args, ds, wl), 'last_loss': last}
    if why:
        line['config']['workload_note'] = 'NOT the full %s graph: %s' % (args.workload, why)
        sample += ' (NOT the full graph: %s)' % why
    line['cpu_baseline'] = {'value': value, 'unit': line['unit'], 'cores': threads, 'kind': 'port', 'sample': sample}
    line['e2e'] = {'value': value, 'unit': line['unit'], 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    emit(line)


def cpu_baseline_leg(args):
    """The bounded CPU sample printed next to the product's line (rank 0, N = 1)."""
    from graph_recsys_benchmark_b200.datasets import SyntheticHIN
    threads = os.cpu_count() or 1
    wl, why = cpu_workload(args.workload)
    ds = SyntheticHIN(wl, seed=1234)
    if args.phase == 'eval':
        t_prop, t_user, n_users = cpu_eval_pass(ds, args.model, threads, sample_users=min(ds.num_uids, 1000))
        value = ds.num_uids / (t_prop + ds.num_uids * t_user)
        sample = ('one propagation (%.2f s) + the per-user metrics loop on %d of %d users (%.3f ms per user) on the %s graph'
                  % (t_prop, n_users, ds.num_uids, t_user * 1e3, wl))
        unit = 'users/s'
    else:
        t_step, steps, warm, _ = cpu_train_steps(ds, args.model, args.batch, 2, 1, threads, 45.0)
        value = args.batch / t_step
        sample = '%d timed train step(s) after %d warm-up on the %s graph, %.2f s per step' % (steps, warm, wl, t_step)
        unit = 'triples/s'
    if why:
        sample += ' (NOT the full %s graph: %s)' % (args.workload, why)
    return {'value': value, 'unit': unit, 'cores': threads, 'kind': 'port', 'sample': sample, 'workload': wl}


# ---------------------------------------------------------------------------------------------
def family_of(name):
    if name.startswith('spmm_filtered'):
        return 'aggregation_filtered'          # data-dependent work: listed in the breakdown, kept out of the roofline figure
    if name.startswith('spmm') or name.startswith('gat_agg'):
        return 'aggregation'
    if name.startswith('linear') or name.startswith('wgrad'):
        return 'projection'
    name = name[len('peagnn_'):] if name.startswith('peagnn_') else name
    for prefix in ('gat_backward', 'gat_rowmax', 'gat_scores', 'fuse', 'bpr', 'entity', 'relu_backward', 'linear_wgrad'):
        if name.startswith(prefix):
            return 'bias_colsum' if prefix == 'linear_wgrad' else prefix
    return name


def summarise_profile(records, n_steps):
    """records: (name, algorithmic bytes, ms) per launch over n_steps steps -> per-family dict."""
    fam = {}
    for name, nbytes, ms in records:
        d = fam.setdefault(family_of(name), {'launches': 0, 'bytes': 0.0, 'ms': 0.0, 'by_name': {}})
        d['launches'] += 1
        d['bytes'] += nbytes
        d['ms'] += ms
        b = d['by_name'].setdefault(name, [0, 0.0, 0.0])
        b[0] += 1
        b[1] += nbytes
        b[2] += ms
    for d in fam.values():
        d['launches_per_step'] = d['launches'] / n_steps
        d['ms_per_step'] = d['ms'] / n_steps
        d['gbs'] = d['bytes'] / (d['ms'] * 1e-3) / 1e9 if d['ms'] > 0 else None
    return fam


def dominant_launch(fam, hbm_peak, gather):
    """The launch tag with the most algorithmic bytes in a family, with its own rate (the family figure averages
    it with the small, latency-bound relations)."""
    name, (cnt, nb, ms) = max(fam['by_name'].items(), key=lambda kv: kv[1][1] / kv[1][0])
    gbs = nb / (ms * 1e-3) / 1e9 if ms > 0 else None
    out = {'tag': name, 'launches': cnt, 'bytes_per_launch': nb / cnt, 'ms_per_launch': ms / cnt, 'achieved': gbs,
           'frac': gbs / hbm_peak if gbs else None, 'frac_of_l2_gather_ceiling': (gbs / gather['gbs']) if (gbs and gather) else None}
    import re
    m = re.search(r'_f(\d+)_e(\d+)_n(\d+)', name)
    if m:
        # what HAS to cross HBM once when the gathered table stays in L2 - the index stream, the gathered
        # table and the output table: E*4 + N*(F_g + F_o)*4 + (N+1)*4; `traffic` (ncu dram bytes) is to be read against this
        F, E, N = (int(v) for v in m.groups())
        out['compulsory_bytes_per_launch'] = E * 4 + N * (F + F) * 4 + (N + 1) * 4
    return out


def measure_gather_ceiling(model, ds, dev):
    """Random-row gather rate of this GPU for the dominant aggregation's own access pattern: the very index stream
    that launch walks (the column array of the forward CSR of the first metapath's first relation - user2item at the
    MovieLens shapes) addressing the embedding table it gathers from, 128-bit loads, and nothing else."""
    from graph_recsys_benchmark_b200 import _lib
    from graph_recsys_benchmark_b200.graph import _ptr, _stream, get_graph
    table = model.x.detach()
    g = get_graph(model.meta_path_edge_index_list[0][0], table.shape[0])
    idx = g.fwd.col
    out = torch.empty(int(_lib.query('peagnn_probe_out_floats')), dtype=torch.float32, device=dev)
    feat = table.shape[1]
    if feat not in (16, 32, 64, 128) or idx.numel() == 0:
        return None

    def launch():
        _lib.call('peagnn_probe_gather', _ptr(table), table.stride(0), feat, _ptr(idx), idx.numel(), _ptr(out), _stream())
    for _ in range(3):
        launch()
    torch.cuda.synchronize()
    evs = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); launch(); e1.record()
        evs.append((e0, e1))
    torch.cuda.synchronize()
    ms = float(np.median([a.elapsed_time(b) for a, b in evs]))
    nbytes = idx.numel() * (4 + 4 * feat)
    return {'gbs': nbytes / (ms * 1e-3) / 1e9, 'ms': ms, 'rows_gathered': int(idx.numel()), 'row_bytes': 4 * feat,
            'table_mb': table.numel() * 4 / 1e6,
            'how': 'peagnn_probe_gather over the column-index stream of the largest first-step relation (forward CSR) into the '
                   'embedding table: 128-bit row gathers and a running sum, no row bookkeeping; median of 10, warm L2'}


def traffic_from_profiles(kernel_key):
    """dram bytes per launch of the dominant kernel from a committed ncu capture (profiles/roofline_traffic.json,
    written by tools/ncu_traffic.py from the .csv next to it); None when no capture is committed."""
    path = os.path.join(ROOT, 'profiles', 'roofline_traffic.json')
    if not os.path.exists(path):
        return None, None
    try:
        table = json.load(open(path))
    except ValueError:
        return None, None
    hit = table.get(kernel_key)
    if not hit:
        return None, None
    return hit.get('dram_bytes_per_launch'), hit.get('source')


def run_product(args):
    import torch.distributed as dist
    from graph_recsys_benchmark_b200 import _lib, functional as F_
    from graph_recsys_benchmark_b200.datasets import SyntheticHIN
    from graph_recsys_benchmark_b200.utils.factory import build_model

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        # collectives run on NCCL's own stream; with record_stream bookkeeping the caching allocator
        # cannot reuse their buffers until a host sync, and an un-synced step loop keeps growing
        os.environ.setdefault('TORCH_NCCL_AVOID_RECORD_STREAMS', '1')
        dist.init_process_group('nccl', device_id=dev)

    ds = SyntheticHIN(args.workload, seed=1234)
    torch.manual_seed(2020)
    model = build_model(ds, args.model, device=dev)
    model.demand_driven_loss = not args.full_propagation
    model.gather_dtype = args.gather_dtype
    if world > 1:
        from graph_recsys_benchmark_b200.distributed import shard_model
        shard_model(model, world, rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    peaks = {}
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get('hbm_gbs', 6650.0))
    peak_src = 'MEASURED_PEAKS.json (torch copy_, burst)' if 'hbm_gbs' in peaks else 'fallback of B200_PROFILING.md'
    ctx = dict(args=args, ds=ds, model=model, dev=dev, world=world, rank=rank, local=local, barrier=barrier,
               hbm_peak=hbm_peak, peak_src=peak_src, dist=dist, lib=_lib, F_=F_)
    line = product_eval(ctx) if args.phase == 'eval' else product_train(ctx)
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            line['cpu_baseline'] = cpu_baseline_leg(args)
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    if line is not None and line.get('_fail'):
        sys.exit(3)


# ---------------------------------------------------------------------------------------------
def product_train(c):
    args, ds, model, dev, world, rank = c['args'], c['ds'], c['model'], c['dev'], c['world'], c['rank']
    dist, _lib, F_, barrier = c['dist'], c['lib'], c['F_'], c['barrier']
    from graph_recsys_benchmark_b200.graphed import GraphedTrainStep
    params = [p for p in model.parameters()]
    # the step - NCCL collectives included - is replayed as one CUDA graph (graphed.py captures in 'thread_local' mode so
    # that NCCL's watchdog thread cannot invalidate the capture); PEAGNN_BENCH_EAGER_MULTI=1 keeps N > 1 on eager launches
    use_graph = not args.no_cuda_graph and (world == 1 or not os.environ.get('PEAGNN_BENCH_EAGER_MULTI'))
    opt = torch.optim.Adam(params, lr=1e-3, weight_decay=1e-3, fused=True, capturable=use_graph)
    model.train()
    K, W, B = args.steps, args.warmup, args.batch
    host_batches = make_batches(ds, B, K + W, seed=100 + rank).pin_memory()
    dev_batches = host_batches.to(dev)

    def allreduce_grads():
        if world > 1:
            from graph_recsys_benchmark_b200.distributed import allreduce_gradients
            allreduce_gradients(params)

    def eager_step(batch):
        opt.zero_grad(set_to_none=True)
        loss = model.loss(batch)
        loss.backward()
        allreduce_grads()
        opt.step()
        return loss.detach()

    step = eager_step
    # bring the SM clocks out of idle before anything is timed (learning rate 0 while spinning: the same
    # kernels run, the weights stay at their initial values)
    for gr in opt.param_groups:
        gr['lr'] = 0.0
    t_spin = time.perf_counter()
    spin = torch.tensor([0.0], device=dev)
    while spin.item() < 0.5:
        step(dev_batches[0])
        torch.cuda.synchronize()
        spin[0] = 1.0 if time.perf_counter() - t_spin >= args.prewarm else 0.0
        if world > 1:
            dist.all_reduce(spin, op=dist.ReduceOp.MIN)       # every rank leaves the loop together
    for gr in opt.param_groups:
        gr['lr'] = 1e-3
    graphed = None
    if use_graph:
        ok = torch.ones(1, device=dev)
        try:
            graphed = GraphedTrainStep(model, opt, dev_batches[0], allreduce=allreduce_grads if world > 1 else None)
        except Exception as exc:                                  # noqa: BLE001
            sys.stderr.write('CUDA graph capture failed (%s: %s); running eager\n' % (type(exc).__name__, exc))
            ok.zero_()
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if ok.item() < 0.5:
            graphed = None
        else:
            step = graphed
    for k in range(W):
        step(dev_batches[k])
    barrier()

    # ---- leg 1: device-resident batches (value) ----------------------------------------------
    losses = torch.zeros(2 * K + 2, device=dev)                   # every step's loss, checked for finiteness at the end
    launches0 = _lib.load().peagnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(c['local'], enabled=(rank == 0 and not os.environ.get('PEAGNN_BENCH_NO_CLOCKS'))) as clocks:
        barrier()
        t_host = time.perf_counter()
        e0.record()
        for k in range(K):
            losses[k].copy_(step(dev_batches[W + k]))
        e1.record()
        host_ms = (time.perf_counter() - t_host) * 1e3 / K       # time the host needs to ENQUEUE a step
        barrier()
    launches = int(_lib.load().peagnn_launch_count() - launches0)
    if graphed is not None:
        launches = graphed.launches_per_replay * K                # replays do not pass through the C ABI's counter
    ms_total = e0.elapsed_time(e1)

    # ---- leg 2: end to end through the public API with host batches -------------------------
    barrier()
    t_e2e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_e2e[0].record()
    last, first = 0.0, None
    for k in range(K):
        if graphed is not None:
            lt = graphed(host_batches[W + k])                      # H2D into the graph's batch tensor
        else:
            lt = step(host_batches[W + k].to(dev, non_blocking=True))    # H2D inside the timed region
        losses[K + k].copy_(lt)
        last = lt.item()                                           # D2H read of the loss every step
        first = last if first is None else first
    t_e2e[1].record()
    barrier()
    ms_e2e = t_e2e[0].elapsed_time(t_e2e[1])

    # ---- leg 3: the kernel families, event-timed inside a timed region of their own -----------
    K_r = max(1, min(K, 10))
    records, ms_roof, roof_how = [], None, None
    if not os.environ.get('PEAGNN_BENCH_NO_PROFILE'):
        for gr in opt.param_groups:
            gr['lr'] = 0.0                                        # measurement only: do not move the weights further
        inst = None
        from graph_recsys_benchmark_b200 import engine as _engine
        _engine.PARALLEL_BRANCHES = False       # per-kernel events need the launches one after the other
        if graphed is not None:
            try:
                inst = GraphedTrainStep(model, opt, dev_batches[0], profile=True)
            except Exception as exc:                              # noqa: BLE001
                sys.stderr.write('instrumented capture failed (%s: %s); event-timing eager steps\n' % (type(exc).__name__, exc))
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        if inst is not None:
            inst(dev_batches[W])
            torch.cuda.synchronize()
            r0.record()
            for k in range(K_r):
                inst(dev_batches[W + k])
                torch.cuda.synchronize()                          # the graph's event nodes are re-recorded by every replay
                records.extend((n, b, a.elapsed_time(z)) for n, b, a, z in inst.profile_events)
            r1.record()
            roof_how = ('external CUDA event nodes inside %d replays of the step graph re-captured with them and with its '
                        'parallel branches serialised, so every launch is timed alone (one synchronize per replay to read '
                        'them); the event nodes themselves make this graph slower than the timed one' % K_r)
        else:
            _lib.profile = []
            r0.record()
            for k in range(K_r):
                eager_step(dev_batches[W + k])
            r1.record()
            barrier()
            records = [(n, b, a.elapsed_time(z)) for n, b, a, z in _lib.profile]
            _lib.profile = None
            roof_how = 'CUDA events around every launch of %d eager steps' % K_r
        barrier()
        _engine.PARALLEL_BRANCHES = True
        ms_roof = r0.elapsed_time(r1) / K_r
        for gr in opt.param_groups:
            gr['lr'] = 1e-3

    # ---- strong-scaling leg: the same GLOBAL batch as the 1-GPU run, split over the ranks -------
    strong = None
    if world > 1 and not args.no_strong and B % world == 0:
        small = make_batches(ds, B // world, K + W, seed=500 + rank).to(dev)
        sstep = eager_step
        if graphed is not None:                                   # the same captured form as the weak-scaling legs
            ok = torch.ones(1, device=dev)
            try:
                sgraph = GraphedTrainStep(model, opt, small[0], allreduce=allreduce_grads)
                sstep = sgraph
            except Exception as exc:                              # noqa: BLE001
                sys.stderr.write('strong-leg capture failed (%s: %s)\n' % (type(exc).__name__, exc))
                ok.zero_()
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 0.5:
                sstep = eager_step
        for k in range(W):
            sstep(small[k])
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        s0.record()
        for k in range(K):
            losses[2 * K].copy_(sstep(small[W + k]))
        s1.record()
        barrier()
        strong = s0.elapsed_time(s1)

    gather = measure_gather_ceiling(model, ds, dev) if rank == 0 else None
    finite = torch.isfinite(losses).all().float()
    t = torch.tensor([ms_total, ms_e2e, strong or 0.0, -float(finite.item())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e, strong_ms, all_finite = float(t[0]), float(t[1]), float(t[2]), float(t[3]) <= -0.5
    first_dev, last_dev = float(losses[0].item()), float(losses[K - 1].item())
    if rank != 0:
        return {'_fail': not all_finite}

    fam = summarise_profile(records, K_r) if records else {}
    agg, proj = fam.get('aggregation'), fam.get('projection')
    if os.environ.get('PEAGNN_BENCH_DUMP_SPMM'):
        for f, d in fam.items():
            for name, (cnt, nb, ms) in sorted(d['by_name'].items(), key=lambda kv: -kv[1][2]):
                sys.stderr.write('PROFILE %-12s %-44s x%6.1f/step %8.3f ms/launch %9.1f GB/s\n'
                                 % (f, name, cnt / K_r, ms / cnt, nb / ms / 1e6 if ms > 0 else 0.0))
    agg_kernel = 'csr_rows_kernel / csr_chunk_kernel (peagnn_spmm' + (', peagnn_gat_aggregate)' if args.model == 'gat' else ')')
    traffic, traffic_src = traffic_from_profiles('%s/%s/aggregation' % (args.workload, args.model)) if world == 1 else (None, None)
    line = {
        'metric': 'bpr_triples_per_sec', 'value': world * B * K / (ms_total * 1e-3), 'unit': 'triples/s',
        'n_gpus': world, 'steps': K, 'warmup': W, 'ms_per_step': ms_total / K, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32' if args.gather_dtype == 'fp32' else 'bf16 gathers, f32 accumulate',
        'data': 'synthetic',
        'config': dict(workload_config(args, ds), num_nodes=ds.num_nodes,
                       edges_user2item=int(ds.edge_index_nps['user2item'].shape[1]),
                       parallelism=('single GPU' if world == 1 else
                                    '%d-way destination-row sharded propagation (NCCL all-gather / reduce-scatter per step) '
                                    '+ data-parallel batches (NCCL all-reduce of gradients)' % world)),
        'e2e': {'value': world * B * K / (ms_e2e * 1e-3), 'unit': 'triples/s', 'h2d_bytes_per_step': B * 3 * 8,
                'd2h_bytes_per_step': 4, 'ms_per_step': ms_e2e / K, 'first_loss': first, 'last_loss': last},
        'loss': {'first': first_dev, 'last': last_dev, 'all_finite': all_finite},
        'gpu_launches': launches, 'host_enqueue_ms_per_step': host_ms, 'cuda_graph': graphed is not None,
        'clocks': clocks.summary(),
    }
    if agg:
        ach = agg['gbs']
        line['roofline'] = {
            'bound': 'hbm', 'kernel': agg_kernel, 'achieved': ach, 'peak': c['hbm_peak'], 'peak_source': c['peak_src'],
            'unit': 'GB/s', 'frac': ach / c['hbm_peak'], 'traffic': traffic, 'traffic_source': traffic_src,
            'note': ('achieved = algorithmic bytes (SURVEY 8d: every gathered row counted as if it came from HBM); the gathered '
                     'tables (<= 75 MB) live in the 126 MB L2, so a fraction above 1 of the HBM copy peak is expected - the '
                     'ceiling that binds is the L2 random-row gather rate, measured live below'),
            'l2_gather_ceiling': gather, 'frac_of_l2_gather_ceiling': (ach / gather['gbs']) if gather else None,
            'dominant_launch': dominant_launch(agg, c['hbm_peak'], gather),
            'launches_per_step': agg['launches_per_step'], 'ms_per_step': agg['ms_per_step'],
            'share_of_step': agg['ms_per_step'] / ms_roof if ms_roof else None,
            'events': roof_how, 'instrumented_ms_per_step': ms_roof,
        }
    if proj:
        line['roofline_projection'] = {
            'bound': 'hbm', 'kernel': 'linear_umma_ts / linear_umma / linear_tc / wgrad_umma / wgrad_tc, grouped across the metapaths where the step allows (peagnn_linear[_grouped], peagnn_linear_wgrad[_grouped])',
            'achieved': proj['gbs'], 'peak': c['hbm_peak'], 'unit': 'GB/s', 'frac': proj['gbs'] / c['hbm_peak'],
            'bytes': '4 N (K + M) + 4 K M per problem (+ 4 N M when accumulating / gating), summed over the problems of a grouped launch', 'traffic': None,
            'launches_per_step': proj['launches_per_step'], 'ms_per_step': proj['ms_per_step'],
            'share_of_step': proj['ms_per_step'] / ms_roof if ms_roof else None,
        }
    if fam:
        line['step_breakdown_ms'] = {f: round(d['ms_per_step'], 4) for f, d in sorted(fam.items())}
    if strong_ms > 0:
        line['strong'] = {'global_batch': B, 'batch_per_gpu': B // world, 'ms_per_step': strong_ms / K,
                          'value': B * K / (strong_ms * 1e-3), 'unit': 'triples/s',
                          'note': 'same global batch as the 1-GPU line; the propagation is strong-scaled in both legs'}
    if not all_finite:
        line['_fail'] = True
        sys.stderr.write('NON-FINITE LOSS in the timed region - the run is invalid\n')
    return line


# ---------------------------------------------------------------------------------------------
def product_eval(c):
    """users/s of one evaluation pass: model.eval() (propagation) + metrics() over every user."""
    args, ds, model, dev, world, rank = c['args'], c['ds'], c['model'], c['dev'], c['world'], c['rank']
    dist, _lib, F_, barrier = c['dist'], c['lib'], c['F_'], c['barrier']
    from graph_recsys_benchmark_b200.solvers import BaseSolver
    K, W = args.steps, args.warmup
    solver = BaseSolver(None, {}, {}, {'device': dev, 'num_neg_candidates': 99, 'batch_size': args.batch})
    np.random.seed(2020)
    users, cand, n_pos = solver.generate_all_candidates(ds)
    U = users.shape[0]
    mine_u, mine_c = (users[rank::world], cand[rank::world]) if world > 1 else (users, cand)
    users_h = torch.from_numpy(np.ascontiguousarray(mine_u)).pin_memory()
    cand_h = torch.from_numpy(np.ascontiguousarray(mine_c)).pin_memory()
    users_d, cand_d = users_h.to(dev), cand_h.to(dev)

    def device_pass(u, cnd):
        model.eval()                                              # models/base.py:88-96: one no-grad propagation
        per_user, means, _ = F_.eval_rank(model.cached_repr, u, cnd, n_pos, model.fc1.weight, model.fc1.bias,
                                          model.fc2.weight, model.fc2.bias)
        if world > 1:
            sums = means * float(u.shape[0])
            dist.all_reduce(sums)
            means = sums / float(U)
        return means

    for _ in range(max(W, 1)):
        device_pass(users_d, cand_d)
    barrier()
    launches0 = _lib.load().peagnn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(c['local'], enabled=rank == 0) as clocks:
        barrier()
        e0.record()
        for _ in range(K):
            means = device_pass(users_d, cand_d)
        e1.record()
        barrier()
    launches = int(_lib.load().peagnn_launch_count() - launches0)
    ms_total = e0.elapsed_time(e1)
    # e2e: candidates in pinned host memory -> device, metrics read back, every pass
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(K):
        m = device_pass(users_h.to(dev, non_blocking=True), cand_h.to(dev, non_blocking=True)).cpu().numpy()
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    # the public call as a user makes it (host-side candidate sampling included), once
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    np.random.seed(2020)
    model.eval()
    hr, ndcg, auc, eloss = solver.metrics(1, 0, model, ds)
    t_api = time.perf_counter() - t0
    t = torch.tensor([ms_total, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_e2e = float(t[0]), float(t[1])
    if rank != 0:
        return {}
    return {
        'metric': 'eval_users_per_sec', 'value': U * K / (ms_total * 1e-3), 'unit': 'users/s', 'n_gpus': world,
        'steps': K, 'warmup': max(W, 1), 'ms_per_step': ms_total / K, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': dict(workload_config(args, ds), users=int(U), candidates_per_user=int(cand.shape[1]),
                       parallelism='single GPU' if world == 1 else '%d-way row-sharded propagation + users sharded over ranks '
                                                                   '(all-reduce of 36 partial sums)' % world),
        'e2e': {'value': U * K / (ms_e2e * 1e-3), 'unit': 'users/s', 'h2d_bytes_per_step': int(cand_h.numel() * 8 + users_h.numel() * 8),
                'd2h_bytes_per_step': 36 * 8, 'ms_per_step': ms_e2e / K},
        'public_api_call': {'what': 'BaseSolver.metrics() incl. host-side candidate sampling (numpy, identical draws to the reference)',
                            'seconds': t_api, 'users_per_sec': U / t_api},
        'metrics': {'HR@10': float(hr[5]), 'NDCG@10': float(ndcg[5]), 'AUC': float(auc[0]), 'eval_loss': float(eloss[0]),
                    'device_HR@10': float(m[5])},
        'gpu_launches': launches, 'clocks': clocks.summary(),
    }


_REAL_STDOUT = None


def emit(line):
    """The ONE JSON line goes to the process's original stdout."""
    line = {k: v for k, v in line.items() if not k.startswith('_')}
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + '\n')
    out.flush()


def main():
    global _REAL_STDOUT
    args = parse_args()
    # Libraries (NCCL's version banner, tqdm, ...) may print to fd 1; keep stdout clean for the JSON line
    # by pointing fd 1 at stderr for the rest of the run.
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_product(args)


if __name__ == '__main__':
    main()
