"""Breakdown of one distributed hdiff step (run under torchrun): exchange alone, interior alone, boundary alone."""
import os, sys, json, pathlib, ctypes
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np, torch, torch.distributed as dist
from gt4py_b200 import runtime, storage, testing
from gt4py_b200.distributed import HaloExchanger, SlabDecomposition
from gt4py_b200.stencil import B200Stencil

NI, NJ, NK, H = 1024, 1024, 80, 2
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
st = B200Stencil(testing.load_ir("hdiff_f32", "staged"), {"device_sync": False})
shape = (NI + 2 * H, NJ + 2 * H, NK); o = (H, H, 0)
org = {k: o for k in ("in_field", "out_field", "coeff")}
rng = np.random.default_rng(rank)
f = {"in_field": storage.from_array(rng.random(shape, dtype=np.float32), aligned_index=o),
     "coeff": storage.from_array(rng.random(shape, dtype=np.float32), aligned_index=o),
     "out_field": storage.zeros(shape, np.float32, aligned_index=o)}
fr = st.freeze(origin=org, domain=(NI, NJ, NK))
ex = HaloExchanger(SlabDecomposition(world, rank, NJ * world), NJ)
main = torch.cuda.current_stream().cuda_stream

def timeit(fn, n=30, stream_sync=True):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / n], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return round(float(t.item()) * 1e3, 1)

res = {"world": world}
res["exchange_on_main_us"] = timeit(lambda: ex.exchange([(f["in_field"], H, H)], stream=main))
res["full_kernel_us"] = timeit(lambda: fr(**f))
res["interior_us"] = timeit(lambda: fr(**f, subbox=(0, NI, H, NJ - H)))
res["boundary_lo_us"] = timeit(lambda: fr(**f, subbox=(0, NI, 0, H)))
res["boundary_both_us"] = timeit(lambda: (fr(**f, subbox=(0, NI, 0, H)), fr(**f, subbox=(0, NI, NJ - H, NJ))))
res["interior_tiles_us"] = timeit(lambda: fr(**f, subbox=(0, NI, 64, NJ - 64)))
res["boundary_tiles_both_us"] = timeit(lambda: (fr(**f, subbox=(0, NI, 0, 64)), fr(**f, subbox=(0, NI, NJ - 64, NJ))))
res["exchange_then_full_us"] = timeit(lambda: (ex.exchange([(f["in_field"], H, H)], stream=main), fr(**f)))
if rank == 0: print(json.dumps(res))
ex.close(); dist.barrier(); dist.destroy_process_group()
