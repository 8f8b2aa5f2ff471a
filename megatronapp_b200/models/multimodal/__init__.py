from .llava_model import DEFAULT_IMAGE_TOKEN_INDEX, IGNORE_INDEX, IMAGE_TOKEN, LLaVAModel  # noqa: F401
