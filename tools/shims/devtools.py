"""Dev-only stand-in for `devtools`."""
import pprint


def debug(*args, **kwargs):
    for a in args:
        pprint.pprint(a)
    return args[0] if len(args) == 1 else args


def pformat(obj, **kwargs):
    return pprint.pformat(obj)
