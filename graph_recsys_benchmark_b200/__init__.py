"""graph_recsys_benchmark_b200 - B200-native (sm_100a) implementation of PEAGNN's metapath
message-passing hot path behind the reference's own Python API.

Drop-in surface (reference file:line in each module's docstring):
  nn.PEAGCNConv / PEAGATConv / PEASageConv      the three conv layers inside the PEA channels
  models.PEAGCNRecsysModel / PEAGATRecsysModel / PEASageRecsysModel
  solvers.BaseSolver                            BPR training loop + HR/NDCG evaluation
All arithmetic runs in hand-written CUDA kernels (csrc/, C ABI in include/peagnn.h);
there is no CPU path and no PyG / torch-scatter / Triton dependency.
"""
__version__ = '0.1.0'
