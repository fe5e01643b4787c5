"""PEASage + BPR with the reference's command line, e.g.
  python3 peasage_solver_bpr.py --dataset=Movielens --dataset_name=latest-small --sampling_strategy=unseen \
      --entity_aware=false --runs=1 --epochs=2 --batch_size=1024 --synthetic=ml-small"""
from pea_cli import main

if __name__ == '__main__':
    main('PEASage')
