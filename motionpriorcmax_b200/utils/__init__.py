from .event_image_converter import EventImageConverter  # noqa: F401
