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


class _SharedMemoryContext:
    """Stand-in for raytracers_b200.Context in the protocol test below: "device memory" is a multiprocessing shared-memory
    block (so two real processes share rank 0's frame ring like two GPUs share an IPC mapping), a "render" writes this rank's
    tiles of a known frame into the row-major frame and bumps the done flag, flag waits poll."""

    def __init__(self, frames, rank, world):
        self.frames, self.rank, self.world, self.shm, self.timeouts = frames, rank, world, None, 0

    def set_shard(self, rank, world):
        assert (rank, world) == (self.rank, self.world)

    def ipc_alloc(self, nbytes):
        from multiprocessing import shared_memory
        self.shm = shared_memory.SharedMemory(create=True, size=nbytes)
        self.shm.buf[:nbytes] = bytes(nbytes)
        return 0, self.shm.name.encode()          # "address" 0 = start of the block

    def ipc_open(self, handle):
        from multiprocessing import shared_memory
        self.shm = shared_memory.SharedMemory(name=bytes(handle).decode())
        return 0

    def _u32(self, addr):
        return np.frombuffer(self.shm.buf, dtype=np.uint32, count=1, offset=addr)

    def render_batch(self, jobs):
        from raytracers_b200 import distributed as D
        for j in jobs:
            if j.get("wait_flag") is not None:
                self.flag_wait(j["wait_flag"], j["wait_value"])
            h, w = j["h"], j["w"]
            src = self.frames[j["prepared"]]
            dst = np.frombuffer(self.shm.buf, dtype=np.int32, count=h * w, offset=j["out_dev"]).reshape(h, w)
            tiles_x = (w + 7) // 8
            for t in range(self.rank, ((h + 3) // 4) * tiles_x, self.world):   # this rank's tiles, interleaved
                ty, tx = divmod(t, tiles_x)
                dst[ty * 4:(ty + 1) * 4, tx * 8:(tx + 1) * 8] = src[ty * 4:(ty + 1) * 4, tx * 8:(tx + 1) * 8]
            import fcntl
            with open(f"/tmp/{self.shm.name}.lock", "w") as lk:   # the GPU version is an atomicAdd_system
                fcntl.lockf(lk, fcntl.LOCK_EX)
                self._u32(j["done_flag"])[0] += 1
                fcntl.lockf(lk, fcntl.LOCK_UN)

    def flag_wait(self, flag, value, stream=None, timeout_ms=20000):
        import time
        t0 = time.time()
        while int(self._u32(flag)[0]) < value:
            if time.time() - t0 > timeout_ms / 1e3:
                self.timeouts += 1
                return
            time.sleep(0.0005)

    def flag_set(self, flag, value, stream=None):
        self._u32(flag)[0] = value

    def copy_to_host_async(self, host_ptr, dev_ptr, nbytes, stream=None):
        import ctypes
        ctypes.memmove(host_ptr, ctypes.addressof(ctypes.c_char.from_buffer(self.shm.buf, dev_ptr)), nbytes)

    def flag_timeouts(self):
        return self.timeouts

    def sync(self):
        pass

    def ipc_free(self, ptr):
        self.shm.close(); self.shm.unlink()

    def ipc_close(self, ptr):
        self.shm.close()


def _peer_worker(rank, world, port, h, w, frames, out_path):
    import torch.distributed as dist
    from raytracers_b200 import distributed as D

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ctx = _SharedMemoryContext(frames, rank, world)
        pf = D.PeerFrameRenderer(ctx, rank, world, h, w, slots=2)
        got = []
        for step in range(3):                     # 6 frames through a 2-slot ring: every slot is re-used twice
            outs = pf.render([(step % len(frames), 1), ((step + 1) % len(frames), 1)])
            if rank == 0:
                got += [o.numpy().copy() for o in outs]
        pf.wait()
        dist.barrier()
        if rank == 0:
            np.save(out_path, np.stack(got))
        dist.barrier()
        pf.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_peer_frame_protocol_gloo(tmp_path, oracle, world):
    """PeerFrameRenderer's host logic (ring slots, use counts, done / ack flag values, handle broadcast) across `world`
    real processes: every rank writes only its own tiles into rank 0's frame, rank 0 hands complete frames out in order."""
    import torch.multiprocessing as mp

    h, w = 45, 50
    frames = [oracle.Scene.irreg().prepare(h, w).render(h, w)[0], oracle.Scene.rgbbox().prepare(h, w).render(h, w)[0]]
    out = str(tmp_path / "frames.npy")
    mp.spawn(_peer_worker, args=(world, _free_port(), h, w, frames, out), nprocs=world, join=True)
    got = np.load(out)
    want = [frames[i % 2] for step in range(3) for i in (step, step + 1)]
    assert got.shape[0] == 6
    for g, wnt in zip(got, want):
        np.testing.assert_array_equal(g, wnt)
