"""PEAGCN + BPR - same command line as reference experiments/peagcn_solver_bpr.py, e.g.
  python3 peagcn_solver_bpr.py --dataset=Movielens --dataset_name=latest-small --sampling_strategy=unseen \
      --entity_aware=false --emb_dim=64 --repr_dim=16 --hidden_size=64 --runs=1 --epochs=2 --batch_size=1024"""
from pea_cli import run, models

if __name__ == '__main__':
    run('PEAGCN', models.PEAGCNRecsysModel, with_heads=False)
