from .cli import retro  # noqa: F401
