"""Stencil IR of the b200 backend: a plain-dict, JSON-serialisable restatement of gt4py's OIR.

The IR is the *input contract* of the CUDA emitter (`codegen.py`) and of the CPU oracle
(`oracle/numpy_oracle.py`).  It mirrors the node set of the reference's optimisable IR
(reference: src/gt4py/cartesian/gtc/oir.py:37-363, enums gtc/common.py:54-253) one to one, but is
made of dicts/lists/strings only so that a stencil lowered on a machine that has the gt4py frontend
can be shipped to (and compiled on) a machine that has not.

Node shapes (``t`` is the tag):

stencil    {"t":"stencil","name","params":[param],"temporaries":[temp],"loops":[loop],
            "field_info":{name:{access,boundary,axes,data_dims,dtype}|None},
            "parameter_info":{name:{access,dtype}|None},
            "domain_info":{"min_k":int}, "options":{...}}
param      {"t":"field","name","dtype","dims":[bool,bool,bool],"data_dims":[int]}
           {"t":"scalar","name","dtype"}
temp       {"name","dtype","dims","data_dims","extent":[[i0,i1],[j0,j1]]}
loop       {"order":"parallel"|"forward"|"backward","sections":[section],"caches":[cache]}
cache      {"t":"ij"|"k","name","fill":bool,"flush":bool}
section    {"interval":[[level,offset],[level,offset]],"hes":[he]}     level = "start"|"end"
he         {"locals":[{"name","dtype"}],"extent":[[i0,i1],[j0,j1]],"body":[stmt]}
stmt       {"t":"assign","left":expr,"right":expr}
           {"t":"mask","mask":expr,"body":[stmt]}
           {"t":"while","cond":expr,"body":[stmt]}
           {"t":"hregion","i":[bound|None,bound|None],"j":[...],"body":[stmt]}   bound=[level,offset]
expr       {"t":"field","name","dtype","off":[i,j,k],"data_index":[expr]}
           {"t":"field",...,"off":{"vk":expr}}      variable K offset
           {"t":"field",...,"off":{"abs_k":expr|int}} absolute K index
           {"t":"scalar","name","dtype"}  {"t":"lit","value":str,"dtype"}  {"t":"iter","axis","dtype"}
           {"t":"unary","op","expr","dtype"} {"t":"binary","op","left","right","dtype"}
           {"t":"ternary","cond","true","false","dtype"} {"t":"cast","expr","dtype"}
           {"t":"call","func","args":[expr],"dtype"}
dtype      "bool"|"int8"|"int16"|"int32"|"int64"|"float32"|"float64"
"""

from __future__ import annotations

import hashlib
import json
from typing import Any, Callable, Dict, Iterator, List

IR_VERSION = 1

DTYPES = ("bool", "int8", "int16", "int32", "int64", "float32", "float64")
ITEMSIZE = {"bool": 1, "int8": 1, "int16": 2, "int32": 4, "int64": 8, "float32": 4, "float64": 8}
CTYPE = {
    "bool": "bool",
    "int8": "signed char",
    "int16": "short",
    "int32": "int",
    "int64": "long long",
    "float32": "float",
    "float64": "double",
}


def dumps(stencil: Dict[str, Any]) -> str:
    return json.dumps(stencil, sort_keys=True, separators=(",", ":"))


def loads(text: str) -> Dict[str, Any]:
    ir = json.loads(text)
    if ir.get("t") != "stencil":
        raise ValueError("not a b200 stencil IR document")
    return ir


def load_file(path) -> Dict[str, Any]:
    with open(path, "r", encoding="utf-8") as fh:
        return loads(fh.read())


def save_file(stencil: Dict[str, Any], path) -> None:
    with open(path, "w", encoding="utf-8") as fh:
        json.dump(stencil, fh, sort_keys=True, indent=1)


def fingerprint(stencil: Dict[str, Any], extra: str = "") -> str:
    return hashlib.sha256((dumps(stencil) + "|" + extra).encode()).hexdigest()[:16]


# ---- traversal helpers -------------------------------------------------------------------------
def iter_hes(stencil) -> Iterator[tuple]:
    for li, loop in enumerate(stencil["loops"]):
        for si, sec in enumerate(loop["sections"]):
            for hi, he in enumerate(sec["hes"]):
                yield li, si, hi, loop, sec, he


def walk_exprs(node, fn: Callable[[dict], None]) -> None:
    """Call `fn` on every expression dict reachable from a stmt/expr (pre-order)."""
    if isinstance(node, list):
        for n in node:
            walk_exprs(n, fn)
        return
    if not isinstance(node, dict):
        return
    t = node.get("t")
    if t in ("assign",):
        walk_exprs(node["right"], fn)
        walk_exprs(node["left"], fn)
    elif t == "mask":
        walk_exprs(node["mask"], fn)
        walk_exprs(node["body"], fn)
    elif t == "while":
        walk_exprs(node["cond"], fn)
        walk_exprs(node["body"], fn)
    elif t == "hregion":
        walk_exprs(node["body"], fn)
    else:
        fn(node)
        if t == "field":
            off = node["off"]
            if isinstance(off, dict):
                for v in off.values():
                    if isinstance(v, dict):
                        walk_exprs(v, fn)
            walk_exprs(node.get("data_index", []), fn)
        elif t == "unary":
            walk_exprs(node["expr"], fn)
        elif t == "binary":
            walk_exprs(node["left"], fn)
            walk_exprs(node["right"], fn)
        elif t == "ternary":
            walk_exprs(node["cond"], fn)
            walk_exprs(node["true"], fn)
            walk_exprs(node["false"], fn)
        elif t == "cast":
            walk_exprs(node["expr"], fn)
        elif t == "call":
            walk_exprs(node["args"], fn)


def field_accesses(stmts) -> List[dict]:
    """All field-access records in a statement list: {"name","off","write":bool}."""
    out: List[dict] = []

    def visit_stmt(s):
        t = s["t"]
        if t == "assign":
            walk_exprs(s["right"], lambda e: e["t"] == "field" and out.append({"name": e["name"], "off": e["off"], "write": False}))
            left = s["left"]
            if left["t"] == "field":
                # index expressions on the lhs are reads
                for di in left.get("data_index", []):
                    walk_exprs(di, lambda e: e["t"] == "field" and out.append({"name": e["name"], "off": e["off"], "write": False}))
                if isinstance(left["off"], dict):
                    for v in left["off"].values():
                        if isinstance(v, dict):
                            walk_exprs(v, lambda e: e["t"] == "field" and out.append({"name": e["name"], "off": e["off"], "write": False}))
                out.append({"name": left["name"], "off": left["off"], "write": True})
        elif t == "mask":
            walk_exprs(s["mask"], lambda e: e["t"] == "field" and out.append({"name": e["name"], "off": e["off"], "write": False}))
            for b in s["body"]:
                visit_stmt(b)
        elif t == "while":
            walk_exprs(s["cond"], lambda e: e["t"] == "field" and out.append({"name": e["name"], "off": e["off"], "write": False}))
            for b in s["body"]:
                visit_stmt(b)
        elif t == "hregion":
            for b in s["body"]:
                visit_stmt(b)
        else:
            raise ValueError(f"unknown stmt {t}")

    for s in stmts:
        visit_stmt(s)
    return out


def ij_offset(off) -> tuple:
    """Horizontal part of an access offset (variable/absolute K accesses are IJ-centred)."""
    if isinstance(off, dict):
        return (0, 0)
    return (off[0], off[1])


def decl_table(stencil) -> Dict[str, dict]:
    """name -> declaration for API params and temporaries (adds "kind")."""
    table: Dict[str, dict] = {}
    for p in stencil["params"]:
        table[p["name"]] = {**p, "kind": "api" if p["t"] == "field" else "param"}
    for tmp in stencil["temporaries"]:
        table[tmp["name"]] = {**tmp, "t": "field", "kind": "temp"}
    return table
