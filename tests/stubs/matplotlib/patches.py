"""TEST INFRASTRUCTURE: see matplotlib/__init__.py"""
from . import _Anything


def __getattr__(name):
    return _Anything()
