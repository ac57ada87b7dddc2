"""N>1 host path on CPU: tile partition -> one all-gather -> assemble, over torch.distributed gloo with world_size 2."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _tiles():
    import __graft_entry__ as ge
    ge.load_package()
    import importlib
    return importlib.import_module("monte-carlo-path-tracing_b200.tiles")


@pytest.mark.parametrize("width,height,world", [(64, 64, 2), (30, 21, 3), (8, 8, 4), (1920, 1080, 8), (5, 3, 2)])
def test_partition_is_a_bijection(width, height, world):
    tiles = _tiles()
    seen = np.zeros((height, width), dtype=np.int32)
    n = tiles.pixels_per_rank(width, height, world)
    for rank in range(world):
        i, j, valid = tiles.local_pixel_to_image(width, height, world, rank, np.arange(n))
        np.add.at(seen, (j[valid], i[valid]), 1)
    assert (seen == 1).all()  # every pixel owned by exactly one rank
    rng = np.random.RandomState(0)
    frame = rng.rand(height, width, 3).astype(np.float32)
    gathered = np.stack([tiles.extract(frame, world, r) for r in range(world)])
    assert np.array_equal(tiles.assemble(gathered, width, height), frame)


def _worker(rank, world, port, width, height, out_path):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tiles = _tiles()
    # every rank can compute the reference frame; it only contributes its own tiles
    rng = np.random.RandomState(1234)
    frame = rng.rand(height, width, 3).astype(np.float32)
    mine = torch.from_numpy(tiles.extract(frame, world, rank)).reshape(-1)  # flat tile buffer, as the CUDA path uses
    gathered = torch.zeros(world * mine.numel(), dtype=torch.float32)
    dist.all_gather_into_tensor(gathered, mine)  # the ONE collective of a step
    if rank == 0:
        result = tiles.assemble(gathered.numpy().reshape(world, -1, 3), width, height)
        np.save(out_path, np.array([np.array_equal(result, frame)]))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_and_assemble_world_size_2(tmp_path):
    out = str(tmp_path / "ok.npy")
    mp.spawn(_worker, args=(2, 29531, 52, 37, out), nprocs=2, join=True)
    assert bool(np.load(out)[0])
