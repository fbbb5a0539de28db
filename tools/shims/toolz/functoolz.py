import functools
import inspect


def identity(x):
    return x


def complement(func):
    def _not(*args, **kwargs):
        return not func(*args, **kwargs)

    return _not


def compose(*funcs):
    if not funcs:
        return identity

    def composed(*args, **kwargs):
        res = funcs[-1](*args, **kwargs)
        for f in reversed(funcs[:-1]):
            res = f(res)
        return res

    return composed


class curry:
    def __init__(self, func, *args, **kwargs):
        if isinstance(func, curry):
            args = func.args + args
            kwargs = {**func.keywords, **kwargs}
            func = func.func
        self.func = func
        self.args = args
        self.keywords = kwargs
        functools.update_wrapper(self, func, updated=())

    def __call__(self, *args, **kwargs):
        all_args = self.args + args
        all_kwargs = {**self.keywords, **kwargs}
        try:
            sig = inspect.signature(self.func)
            sig.bind(*all_args, **all_kwargs)
        except TypeError:
            try:
                sig.bind_partial(*all_args, **all_kwargs)
            except TypeError:
                return self.func(*all_args, **all_kwargs)
            return curry(self.func, *all_args, **all_kwargs)
        except ValueError:
            pass
        return self.func(*all_args, **all_kwargs)

    def __get__(self, instance, owner):
        if instance is None:
            return self
        return curry(self, instance)
