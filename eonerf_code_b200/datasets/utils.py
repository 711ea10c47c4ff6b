"""/root/reference/datasets/utils.py:1-11"""
from collections import namedtuple

from .satellite import namedtuple_map  # noqa: F401

Rays = namedtuple("Rays", ("origins", "viewdirs"))
