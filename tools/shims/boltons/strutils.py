import re
import unicodedata


def a10n(string):
    if len(string) < 3:
        return string
    return f"{string[0]}{len(string) - 2}{string[-1]}"


def asciify(text, ignore=False):
    if isinstance(text, bytes):
        text = text.decode("utf-8", "ignore")
    norm = unicodedata.normalize("NFKD", text)
    return norm.encode("ascii", "ignore" if ignore else "replace")


def slugify(text, delim="_", lower=True, ascii=False):
    out = delim.join(re.split(r"[^\w]+", text.strip()))
    out = out.strip(delim)
    if lower:
        out = out.lower()
    return asciify(out) if ascii else out


def iter_splitlines(text):
    yield from text.splitlines()


def unwrap_text(text, ending="\n\n"):
    paras, cur = [], []
    for line in text.splitlines():
        line = line.strip()
        if line:
            cur.append(line)
        elif cur:
            paras.append(" ".join(cur))
            cur = []
    if cur:
        paras.append(" ".join(cur))
    return paras if ending is None else ending.join(paras)


def parse_int_list(range_string, delim=",", range_delim="-"):
    out = []
    for part in range_string.strip().split(delim):
        part = part.strip()
        if not part:
            continue
        if range_delim in part[1:]:
            lo, hi = part.rsplit(range_delim, 1) if part.count(range_delim) == 1 else part.split(range_delim, 1)
            out.extend(range(int(lo), int(hi) + 1))
        else:
            out.append(int(part))
    return sorted(out)


def format_int_list(int_list, delim=",", range_delim="-", delim_space=False):
    ints = sorted(set(int_list))
    parts, i = [], 0
    while i < len(ints):
        j = i
        while j + 1 < len(ints) and ints[j + 1] == ints[j] + 1:
            j += 1
        parts.append(str(ints[i]) if i == j else f"{ints[i]}{range_delim}{ints[j]}")
        i = j + 1
    return (delim + " " if delim_space else delim).join(parts)
