from . import units  # noqa: F401
