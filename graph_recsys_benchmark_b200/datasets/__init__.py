from .synthetic import SyntheticHIN, SHAPES, make_synthetic_dataset

__all__ = ['SyntheticHIN', 'SHAPES', 'make_synthetic_dataset']
