class classproperty:
    def __init__(self, fn):
        self.fn = fn

    def __get__(self, instance, cls):
        return self.fn(cls)


def make_sentinel(name="_MISSING", var_name=None):
    class Sentinel:
        def __init__(self):
            self.name = name
            self.var_name = var_name

        def __repr__(self):
            return self.var_name or f"{self.__class__.__name__}({self.name!r})"

        def __bool__(self):
            return False

        def __copy__(self):
            return self

        def __deepcopy__(self, memo):
            return self

        if var_name:
            def __reduce__(self):
                return self.var_name

    return Sentinel()
