"""world_size-2 (and 3) CPU run of the multi-GPU host logic over gloo: shard -> one gather -> de-tile
reproduces the single-process frame exactly.  The per-rank tiles come from the oracle frame cut with the
same tile mapping the CUDA shard kernel uses (the GPU version of this test is in test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, h, w, frame, out_path):
    import torch
    import torch.distributed as dist
    from raytracers_b200 import distributed as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tiles = torch.from_numpy(D.extract_rank_tiles(frame, rank, world))
        gathered = D.gather_tiles(tiles, dst=0)
        if rank == 0:
            img = D.detile_reference(gathered.numpy(), h, w, world)
            np.save(out_path, img)
        else:
            assert gathered is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,w", [(2, 64, 96), (3, 45, 50)])
def test_shard_gather_detile_roundtrip_gloo(tmp_path, oracle, world, h, w):
    import torch.multiprocessing as mp

    frame, _, _ = oracle.Scene.irreg().prepare(h, w).render(h, w)
    out = str(tmp_path / "frame.npy")
    mp.spawn(_worker, args=(world, _free_port(), h, w, frame, out), nprocs=world, join=True)
    np.testing.assert_array_equal(np.load(out), frame)
