from .synthetic import SyntheticHIN, SHAPES, make_synthetic_dataset
from .hin_pickle import ProcessedHIN, dump_reference_pickle

__all__ = ['SyntheticHIN', 'SHAPES', 'make_synthetic_dataset', 'ProcessedHIN', 'dump_reference_pickle']
