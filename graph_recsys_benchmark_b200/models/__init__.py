from .base import GraphRecsysModel, PEABaseChannel, PEABaseRecsysModel
from .peagcn import PEAGCNChannel, PEAGCNRecsysModel
from .peagat import PEAGATChannel, PEAGATRecsysModel
from .peasage import PEASageChannel, PEASageRecsysModel

__all__ = [
    'GraphRecsysModel', 'PEABaseChannel', 'PEABaseRecsysModel',
    'PEAGCNChannel', 'PEAGCNRecsysModel', 'PEAGATChannel', 'PEAGATRecsysModel',
    'PEASageChannel', 'PEASageRecsysModel',
]
