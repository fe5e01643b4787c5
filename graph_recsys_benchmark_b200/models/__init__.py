from .base import GraphRecsysModel, PEABaseChannel, PEABaseRecsysModel
from .families import (PEAChannel, channel_layer_dims,
                       PEAGCNChannel, PEAGCNRecsysModel, PEAGATChannel, PEAGATRecsysModel,
                       PEASageChannel, PEASageRecsysModel)

__all__ = [
    'GraphRecsysModel', 'PEABaseChannel', 'PEABaseRecsysModel', 'PEAChannel', 'channel_layer_dims',
    'PEAGCNChannel', 'PEAGCNRecsysModel', 'PEAGATChannel', 'PEAGATRecsysModel',
    'PEASageChannel', 'PEASageRecsysModel',
]
