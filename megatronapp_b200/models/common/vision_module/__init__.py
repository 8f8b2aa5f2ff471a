from .vision_module import VisionModule  # noqa: F401
