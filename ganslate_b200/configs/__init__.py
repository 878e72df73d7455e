from . import base  # noqa: F401
