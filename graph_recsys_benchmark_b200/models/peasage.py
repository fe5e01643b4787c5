"""PEASage (reference models/peasage.py): see families.py - the three families share one channel class."""
from .families import PEASageChannel, PEASageRecsysModel  # noqa: F401
