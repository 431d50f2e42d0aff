from . import amp  # noqa: F401
