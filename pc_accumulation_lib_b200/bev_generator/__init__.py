from .bev_generator import BEVGenerator, DeviceWindow  # noqa: F401
from .rgb_bev import RGBBEVGenerator  # noqa: F401
from .sem_bev import SemBEVGenerator  # noqa: F401
