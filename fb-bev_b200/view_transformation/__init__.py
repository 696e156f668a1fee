"""Plugin classes of the forward-backward view transformation (mirror of
``mmdet3d/models/fbbev/view_transformation``).  Importing this package
registers them (see ``registry.py``)."""
from .forward_projection import *  # noqa: F401,F403
from .backward_projection import *  # noqa: F401,F403
from .temporal_fusion import *  # noqa: F401,F403
from .bevdet_lineage import *  # noqa: F401,F403
from .depth_net_tail import *  # noqa: F401,F403
