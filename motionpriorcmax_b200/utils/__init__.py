from .event_image_converter import EventImageConverter  # noqa: F401
from .flow import dense_flow_from_traj, list_to_grid  # noqa: F401
