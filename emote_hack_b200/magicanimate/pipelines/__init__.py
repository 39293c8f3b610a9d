from . import context  # noqa: F401
