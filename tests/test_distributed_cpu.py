"""world_size-2 gloo test of the J-slab decomposition + halo exchange logic (host transport):
two ranks each run the ORACLE on their slab after exchanging halos; gathered result must equal the
single-domain oracle run bit for bit."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gt4py_b200 import testing
from gt4py_b200.distributed import HaloExchanger, SlabDecomposition
from oracle import numpy_oracle

NI, NJ, NK, H = 24, 22, 3, 2


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _global_problem():
    rng = np.random.default_rng(7)
    shape = (NI + 2 * H, NJ + 2 * H, NK)
    return rng.random(shape, dtype=np.float32), (rng.random(shape, dtype=np.float32) * np.float32(0.1))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = testing.load_ir("hdiff_f32", "default")
    gin, gco = _global_problem()
    dec = SlabDecomposition(world, rank, NJ)
    lo, hi = dec.bounds()
    lin = dec.scatter(gin, H, H)
    lco = dec.scatter(gco, H, H)
    # wipe the halos that must come from the neighbours
    if dec.peer_lo >= 0:
        lin[:, :H] = np.nan
    if dec.peer_hi >= 0:
        lin[:, -H:] = np.nan
    ex = HaloExchanger(dec, transport="gloo")
    ex.exchange_host([(lin, H, H)])
    out = np.zeros_like(lin)
    org = {k: (H, H, 0) for k in ("in_field", "out_field", "coeff")}
    numpy_oracle.run(st, {"in_field": lin, "out_field": out, "coeff": lco}, {}, (NI, hi - lo, NK), org)
    np.save(os.path.join(out_dir, f"out{rank}.npy"), out[:, H : H + hi - lo])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slab_exchange_matches_single_domain(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    st = testing.load_ir("hdiff_f32", "default")
    gin, gco = _global_problem()
    out = np.zeros_like(gin)
    org = {k: (H, H, 0) for k in ("in_field", "out_field", "coeff")}
    numpy_oracle.run(st, {"in_field": gin, "out_field": out, "coeff": gco}, {}, (NI, NJ, NK), org)
    got = np.concatenate([np.load(tmp_path / f"out{r}.npy") for r in range(world)], axis=1)
    np.testing.assert_array_equal(got, out[:, H : H + NJ])


def test_decomposition_bounds_cover_domain():
    for n in (1, 2, 3, 8):
        spans = [SlabDecomposition(n, r, 1027).bounds() for r in range(n)]
        assert spans[0][0] == 0 and spans[-1][1] == 1027
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    with pytest.raises(ValueError):
        SlabDecomposition(4, 4, 100)


H2 = 4  # halo of two fused horizontal-diffusion steps


def _worker_fused(rank, world, port, out_dir):
    """communication-avoiding form of two time steps: ONE exchange of a 4-row halo, then the fused stencil
    (gt4py_b200/fuse.py), which recomputes step 1 on a 2-row rim of the slab"""
    from gt4py_b200 import fuse

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    st = testing.load_ir("hdiff_f32", "staged")
    f2 = fuse.repeat(st, 2, carry=("in_field", "out_field"))
    rng = np.random.default_rng(11)
    shape = (NI + 2 * H2, NJ + 2 * H2, NK)
    gin, gco = rng.random(shape, dtype=np.float32), rng.random(shape, dtype=np.float32) * np.float32(0.1)
    dec = SlabDecomposition(world, rank, NJ)
    lo, hi = dec.bounds()
    lin, lco = dec.scatter(gin, H2, H2), dec.scatter(gco, H2, H2)
    if dec.peer_lo >= 0:
        lin[:, :H2] = np.nan
    if dec.peer_hi >= 0:
        lin[:, -H2:] = np.nan
    HaloExchanger(dec, transport="gloo").exchange_host([(lin, H2, H2)])
    out = np.zeros_like(lin)
    org = {k: (H2, H2, 0) for k in ("in_field", "out_field", "coeff")}
    numpy_oracle.run(f2, {"in_field": lin, "out_field": out, "coeff": lco}, {}, (NI, hi - lo, NK), org)
    np.save(os.path.join(out_dir, f"fused{rank}.npy"), out[:, H2 : H2 + hi - lo])
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_fused_steps_with_one_wide_exchange_match_two_global_steps(tmp_path):
    world = 2
    mp.spawn(_worker_fused, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    st = testing.load_ir("hdiff_f32", "staged")
    rng = np.random.default_rng(11)
    shape = (NI + 2 * H2, NJ + 2 * H2, NK)
    gin, gco = rng.random(shape, dtype=np.float32), rng.random(shape, dtype=np.float32) * np.float32(0.1)
    # two separate global steps: step 1 wherever step 2 reads it (domain grown by 2), then step 2
    mid, out = np.zeros_like(gin), np.zeros_like(gin)
    grown = {k: (H2 - 2, H2 - 2, 0) for k in ("in_field", "out_field", "coeff")}
    numpy_oracle.run(st, {"in_field": gin, "out_field": mid, "coeff": gco}, {}, (NI + 4, NJ + 4, NK), grown)
    org = {k: (H2, H2, 0) for k in ("in_field", "out_field", "coeff")}
    numpy_oracle.run(st, {"in_field": mid, "out_field": out, "coeff": gco}, {}, (NI, NJ, NK), org)
    got = np.concatenate([np.load(tmp_path / f"fused{r}.npy") for r in range(world)], axis=1)
    np.testing.assert_array_equal(got, out[:, H2 : H2 + NJ])
