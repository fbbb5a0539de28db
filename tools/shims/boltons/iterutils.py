def is_collection(obj):
    return hasattr(obj, "__iter__") and not isinstance(obj, (str, bytes))


def flatten_iter(iterable):
    for item in iterable:
        if is_collection(item):
            yield from flatten_iter(item)
        else:
            yield item


def flatten(iterable):
    return list(flatten_iter(iterable))
