"""Dev-only stand-in for `toolz` (absent offline): the handful of helpers gt4py.eve uses."""
import functools
import inspect
import itertools

from . import functoolz, itertoolz
from .functoolz import complement, compose, curry, identity
from .itertoolz import (
    diff,
    groupby,
    partition,
    partition_all,
    pluck,
    reduceby,
    take_nth,
    unique,
)
