"""PEAGAT (reference models/peagat.py): see families.py - the three families share one channel class."""
from .families import PEAGATChannel, PEAGATRecsysModel  # noqa: F401
