"""`import_attr` with the semantics of ganslate/utils/io.py:73-76: dotted path -> attribute."""
import importlib


def import_attr(path: str):
    module, attr = path.rsplit(".", 1)
    return getattr(importlib.import_module(module), attr)
