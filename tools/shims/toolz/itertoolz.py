import collections
import itertools
import operator

no_default = "__no__default__"
no_pad = "__no__pad__"


def take_nth(n, seq):
    return itertools.islice(seq, 0, None, n)


def unique(seq, key=None):
    seen = set()
    for item in seq:
        val = item if key is None else key(item)
        if val not in seen:
            seen.add(val)
            yield item


def _getter(index):
    if isinstance(index, list):
        if len(index) == 1:
            index = index[0]
            return lambda x: (x[index],)
        if index:
            return operator.itemgetter(*index)
        return lambda x: ()
    return operator.itemgetter(index)


def pluck(ind, seqs, default=no_default):
    if default == no_default:
        get = _getter(ind)
        return map(get, seqs)
    if isinstance(ind, list):
        return (tuple(_get_default(i, seq, default) for i in ind) for seq in seqs)
    return (_get_default(ind, seq, default) for seq in seqs)


def _get_default(i, seq, default):
    try:
        return seq[i]
    except (KeyError, IndexError, TypeError):
        return default


def partition(n, seq, pad=no_pad):
    args = [iter(seq)] * n
    if pad is no_pad:
        return zip(*args)
    return itertools.zip_longest(*args, fillvalue=pad)


def partition_all(n, seq):
    it = iter(seq)
    while True:
        chunk = tuple(itertools.islice(it, n))
        if not chunk:
            return
        yield chunk


def groupby(key, seq):
    if not callable(key):
        key = _getter(key)
    d = collections.defaultdict(list)
    for item in seq:
        d[key(item)].append(item)
    return dict(d)


def reduceby(key, binop, seq, init=no_default):
    is_no_default = isinstance(init, str) and init == no_default
    if not is_no_default and not callable(init):
        _init = init
        init = lambda: _init  # noqa: E731
    if not callable(key):
        key = _getter(key)
    d = {}
    for item in seq:
        k = key(item)
        if k not in d:
            if is_no_default:
                d[k] = item
                continue
            d[k] = init()
        d[k] = binop(d[k], item)
    return d


def diff(*seqs, **kwargs):
    n = len(seqs)
    if n == 1 and isinstance(seqs[0], list):
        seqs = seqs[0]
        n = len(seqs)
    if n < 2:
        raise TypeError("Too few sequences given (min 2 required)")
    default = kwargs.get("default", no_default)
    if default == no_default:
        iters = zip(*seqs)
    else:
        iters = itertools.zip_longest(*seqs, fillvalue=default)
    key = kwargs.get("key", None)
    if key is None:
        for items in iters:
            if items.count(items[0]) != n:
                yield items
    else:
        for items in iters:
            vals = tuple(map(key, items))
            if vals.count(vals[0]) != n:
                yield items
