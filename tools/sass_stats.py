"""Static SASS statistics of generated kernels (no GPU needed): registers, total instructions and,
per innermost loop that loads from global memory, (instructions, instructions per cell of a steady
trip, FP instructions, LDG, STG, SHFL, CCTL).  Uses nvcc (through gt4py_b200.jit) + cuobjdump.

    python tools/sass_stats.py hdiff_f32 '{}' '{"interior_loop": true, "static_pitch": 1056}'
    python tools/sass_stats.py --table          # the table committed in profiles/r01c_sass_static.json
"""
import subprocess, sys, re, collections
import pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
from gt4py_b200 import codegen, testing, jit
def analyze(name, opts, variant="staged", show=False):
    st = testing.load_ir(name,variant)
    src, plan = codegen.generate(st, opts)
    p = jit.cubin_path(src, opts, name="x"); jit.compile_cubin(src, opts, name="x")
    sass = subprocess.run(["cuobjdump","-sass",str(p)],capture_output=True,text=True).stdout
    res = subprocess.run(["cuobjdump","-res-usage",str(p)],capture_output=True,text=True).stdout
    out=[]
    for f in sass.split("Function : ")[1:]:
        fname=f.split()[0]
        lines=[l for l in f.splitlines() if re.search(r"^\s+/\*[0-9a-f]{4}\*/", l)]
        ins=[(int(re.search(r"/\*([0-9a-f]{4})\*/",l).group(1),16), re.sub(r"/\*[0-9a-f]{4}\*/","",l).strip().split(";")[0].strip()) for l in lines]
        loops=[]
        for addr,op in ins:
            if "BRA" in op and "BRA.DIV" not in op:
                m2=re.search(r"0x([0-9a-f]+)",op)
                if m2:
                    tgt=int(m2.group(1),16)
                    if tgt<addr: loops.append((tgt,addr))
        # innermost loops = those not containing another loop
        inner=[l for l in loops if not any((o!=l and l[0]<=o[0] and o[1]<=l[1]) for o in loops)]
        m = re.search(r"Function " + re.escape(fname) + r":\s*\n\s*REG:(\d+)", res)
        regs = [m.group(1)] if m else []
        U=plan["kernels"][0].get("period",1); V=plan["kernels"][0].get("vector",1)
        desc=[]
        for (a,b) in inner:
            body=[op for ad,op in ins if a<=ad<=b]
            if not any("LDG" in o for o in body): continue
            c=collections.Counter((o.split()[1] if o.startswith("@") else o.split()[0]).split(".")[0] for o in body)
            fp=sum(c[k] for k in ("FADD","FMUL","FFMA","FSEL","FSETP","FSET","DADD","DMUL","DFMA","MUFU","FCHK","FMNMX"))
            desc.append((len(body), round(len(body)/(U*V),1), fp, c["LDG"], c["STG"], c["SHFL"], c["CCTL"]))
            if show: print(dict(c.most_common(20)))
        out.append((fname, regs, len(ins), desc))
    return out
TABLE = [
    ("hdiff_f32", {}), ("hdiff_f32", {"static_pitch": 1056}), ("hdiff_f32", {"interior_loop": True}),
    ("hdiff_f32", {"interior_loop": "steady", "static_pitch": 1056}),
    ("hdiff_f32", {"interior_loop": True, "static_pitch": 1056}),
    ("hdiff_f32", {"interior_loop": True, "static_pitch": 1056, "row_pointers": True}),
    ("hdiff_f32", {"interior_loop": True, "static_pitch": 1056, "vector_width": 4}),
    ("upwind5_f32", {}), ("upwind5_f32", {"interior_loop": True, "static_pitch": 2080}),
    ("tridiagonal_f64", {"seq_cache": False}), ("tridiagonal_f64", {}),
    ("vadv_f64", {"seq_cache": False}), ("vadv_f64", {}),
]  # fmt: skip

if __name__ == "__main__":
    import json

    if sys.argv[1] == "--table":
        rows = []
        for name, opts in TABLE:
            variant = "default" if name in ("tridiagonal_f64", "vadv_f64") else "staged"
            for fname, regs, total, loops in analyze(name, opts, variant=variant):
                rows.append({"stencil": name, "options": opts, "kernel": fname, "registers": int(regs[0]) if regs else None,
                             "instructions": total,
                             "loops": [dict(zip(("instr", "instr_per_cell", "fp", "ldg", "stg", "shfl", "cctl"), d)) for d in loops]})
        print(json.dumps({"how": "nvcc 12.9 -O3 -fmad=false sm_100a, cuobjdump -sass; loops = innermost loops with global loads, "
                                 "first entry of a streaming kernel = its first steady loop", "rows": rows}, indent=1))
        sys.exit(0)
    name = sys.argv[1]
    for o in sys.argv[2:]:
        opts = json.loads(o)
        for r in analyze(name, opts):
            print(o, r)
