"""Sort-first multi-GPU plumbing on CPU: two gloo ranks each render their horizontal band (CpvkDrawState.bandY0/Y1)
with the oracle and all-gather the bands in place; the gathered frame must equal the single-rank frame byte for byte
(pixels are independent in the reference). The GPU path uses the same band fields and the same
torch.distributed call with the nccl backend (bench.py)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import ctypes as C

    from cpvulkan_b200 import capi, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = scenes.random_triangles(width=96, height=64, tris=150, seed=21)
    lib = capi.load_oracle()
    mem = scenes.HostMemory()
    m = scenes.materialize(scene, mem.alloc)
    rows = scene.color.height // world
    m.state.bandY0, m.state.bandY1 = rank * rows, (rank + 1) * rows
    for img, att in ((scene.color, m.color_attachment), (scene.depth, m.depth_attachment)):
        cv, is_ds = scenes.clear_value(img)
        lib.cpvk_oracle_clear(C.byref(att), C.byref(cv), is_ds)
    st = capi.DrawStats()
    assert lib.cpvk_oracle_draw(C.byref(m.desc), C.byref(m.state), C.byref(st)) == 0
    full = torch.from_numpy(mem.arrays["color"][:scene.color.nbytes])
    band = rows * scene.color.pitch
    chunks = list(full.split(band))
    dist.all_gather(chunks, full[rank * band:(rank + 1) * band].clone())
    gathered = torch.cat(chunks).numpy()
    cov = torch.tensor([int(st.fragmentsCovered)])
    dist.all_reduce(cov)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), gathered)
        np.save(os.path.join(out_dir, "cov.npy"), cov.numpy())
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("world", [2, 4])  # 64 rows: bands of 32 (one tile row each) and of 16 (two ranks share every tile row)
def test_band_render_matches_single_rank(built, tmp_path, world):
    from cpvulkan_b200 import scenes
    port = 29500 + (os.getpid() % 1000) + world
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    gathered = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    color, _, st = scenes.run_oracle(scenes.random_triangles(width=96, height=64, tris=150, seed=21))
    assert np.array_equal(gathered, color)
    assert int(np.load(os.path.join(str(tmp_path), "cov.npy"))[0]) == st.fragmentsCovered
