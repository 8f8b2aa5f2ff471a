from .mamba_model import MambaModel  # noqa: F401
