"""PEAGCN (reference models/peagcn.py): see families.py - the three families share one channel class."""
from .families import PEAGCNChannel, PEAGCNRecsysModel  # noqa: F401
