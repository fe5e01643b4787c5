"""CPU oracle for the PEAGNN metapath message-passing hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``graph_recsys_benchmark_b200/`` may
import this package; only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do, and there only
as the checker / the timed CPU baseline - never as the thing shipped.

It is a pure-torch (CPU, fp32 or fp64) restatement of the reference
``ecml-peagnn/graph_recsys_benchmark`` path:

* ``oracle/pyg150.py``   - GCNConv / GATConv / SAGEConv of torch-geometric==1.5.0
  (un-vendored third-party dependency of the reference, requirements.txt:46;
  call sites models/peagcn.py:16-21, models/peagat.py:16-21, models/peasage.py:16-21)
* ``oracle/models.py``   - models/base.py:29-96,129-214 and the three channel files
* ``oracle/solver.py``   - solvers.py:21-104 (candidates + metrics) and :203-222 (train step)
* ``oracle/rec_utils.py``- utils/rec_utils.py:7-30
* ``oracle/sampling.py`` - datasets/movielens.py:920-940,994-997,1135-1182
* ``oracle/graph.py``    - CSR/CSC construction (stable by destination) and the
  metapath tables of utils/general_utils.py:280-395
* ``oracle/dense.py``    - dense-matrix closed forms (an independent derivation)
  that pin the conv restatements on small graphs

PARITY UNPINNED by reference tests: the reference ships no tests, no golden
vectors, and cannot be imported here (torch_geometric / torch_scatter absent,
numpy-2 removed aliases).  The oracle is pinned instead by (i) the dense
closed forms in ``oracle/dense.py``, (ii) the parameter names/shapes of the six
shipped checkpoints, (iii) the frozen vectors under ``tests/golden/``.
"""
