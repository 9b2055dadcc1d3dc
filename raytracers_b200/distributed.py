"""Multi-GPU host logic: one process per GPU (torchrun), image tiles sharded round-robin, one NCCL
gather of the final framebuffer.

The render path shards by pixels with no data-path exchange (pixels are independent, ray.fut:166-169):
the prepared scene is replicated (<= 66 MB even for 1 M spheres), 8x4-pixel tile t belongs to rank
t % world (contiguous bands would leave irreg's sky ranks idle), every rank renders its tiles into a
compact tile-major buffer with the single-GPU kernels, and ONE collective — a gather of int32 tiles to
rank 0 over NVLink — assembles the frame, which a de-tiling kernel turns back into [h][w].

torch is used here only for device buffers and torch.distributed (plumbing).
"""
import numpy as np

TILE_W, TILE_H, TILE_PIXELS = 8, 4, 32


def tile_layout(h, w, world):
    tiles_x = (w + TILE_W - 1) // TILE_W
    tiles_y = (h + TILE_H - 1) // TILE_H
    n_tiles = tiles_x * tiles_y
    padded = (n_tiles + world - 1) // world
    return tiles_x, tiles_y, n_tiles, padded


def rank_tile_count(h, w, rank, world):
    _, _, n_tiles, _ = tile_layout(h, w, world)
    return n_tiles // world + (1 if (n_tiles % world) > rank else 0)


def extract_rank_tiles(img, rank, world):
    """CPU statement of what ray_b200_render_shard_into produces for `rank`: int32[padded][32]
    (pixels outside the image and padding tiles are 0).  Used by tests and as documentation."""
    h, w = img.shape
    tiles_x, _, n_tiles, padded = tile_layout(h, w, world)
    out = np.zeros((padded, TILE_PIXELS), np.int32)
    for lt in range(padded):
        t = lt * world + rank
        if t >= n_tiles:
            break
        ty, tx = divmod(t, tiles_x)
        blk = img[ty * TILE_H:(ty + 1) * TILE_H, tx * TILE_W:(tx + 1) * TILE_W]
        full = np.zeros((TILE_H, TILE_W), np.int32)
        full[:blk.shape[0], :blk.shape[1]] = blk
        out[lt] = full.reshape(-1)
    return out


def detile_reference(gathered, h, w, world):
    """CPU statement of ray_b200_detile: gathered int32[world][padded][32] -> int32[h][w]."""
    tiles_x, _, _, padded = tile_layout(h, w, world)
    g = np.asarray(gathered).reshape(world, padded, TILE_PIXELS)
    jj, ii = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    t = (jj // TILE_H) * tiles_x + (ii // TILE_W)
    sub = (jj % TILE_H) * TILE_W + (ii % TILE_W)
    return g[t % world, t // world, sub].astype(np.int32)


def gather_tiles(local_tiles, dst=0, group=None):
    """The one collective of the path: gathers every rank's compact tile buffer on `dst`.
    Returns a [world][padded][32] tensor on dst, None elsewhere."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local_tiles.reshape(1, *local_tiles.shape)
    if rank == dst:
        out = torch.empty((world,) + tuple(local_tiles.shape), dtype=local_tiles.dtype, device=local_tiles.device)
        dist.gather(local_tiles, list(out.unbind(0)), dst=dst, group=group)
        return out
    dist.gather(local_tiles, None, dst=dst, group=group)
    return None


class ShardedRenderer:
    """Per-rank driver: render this rank's tiles on its GPU, gather on rank 0, de-tile there."""

    def __init__(self, ctx, rank, world):
        import torch

        self.ctx, self.rank, self.world = ctx, rank, world
        self.torch = torch
        ctx.set_shard(rank, world)
        self._tiles = None
        self._frame = None
        self._key = None
        self._batch = {}

    def _buffers(self, h, w):
        torch = self.torch
        if self._key != (h, w):
            _, _, _, padded = tile_layout(h, w, self.world)
            dev = torch.device("cuda", self.ctx.device)
            self._tiles = torch.empty((padded, TILE_PIXELS), dtype=torch.int32, device=dev)
            self._frame = torch.empty((h, w), dtype=torch.int32, device=dev) if self.rank == 0 else None
            self._key = (h, w)
        return self._tiles, self._frame

    def render(self, h, w, prepared, spp=1):
        """Returns the full frame (device int32[h][w]) on rank 0, None on the other ranks.  Asynchronous
        on torch's current stream (the context must have been pointed at it with ctx.set_stream)."""
        tiles, frame = self._buffers(h, w)
        self.ctx.render_shard_into(tiles.data_ptr(), h, w, prepared, spp)
        gathered = gather_tiles(tiles, dst=0)
        if self.rank != 0:
            return None
        self.ctx.detile(gathered.data_ptr(), frame.data_ptr(), h, w, self.world)
        return frame

    def render_batch(self, jobs):
        """jobs: sequence of (h, w, prepared, spp).  All shards are rendered by ONE ray_b200_render_batch call (two frames
        in flight: a frame's long-path tail is covered by the next frame's start), then gathered and de-tiled frame by
        frame.  Returns the list of full frames on rank 0 (each valid until the next call with the same job slot), None
        elsewhere.  Put the frame with the longest tail first."""
        torch = self.torch
        dev = torch.device("cuda", self.ctx.device)
        bufs = []
        for slot, (h, w, _, _) in enumerate(jobs):
            key = (slot, h, w)
            if key not in self._batch:
                _, _, _, padded = tile_layout(h, w, self.world)
                self._batch[key] = (torch.empty((padded, TILE_PIXELS), dtype=torch.int32, device=dev),
                                    torch.empty((h, w), dtype=torch.int32, device=dev) if self.rank == 0 else None)
            bufs.append(self._batch[key])
        self.ctx.render_batch([dict(prepared=pr, h=h, w=w, spp=spp, shard_layout=True, out_dev=b[0].data_ptr())
                               for (h, w, pr, spp), b in zip(jobs, bufs)])
        frames = []
        for (h, w, _, _), (tiles, frame) in zip(jobs, bufs):
            gathered = gather_tiles(tiles, dst=0)
            if self.rank == 0:
                self.ctx.detile(gathered.data_ptr(), frame.data_ptr(), h, w, self.world)
                frames.append(frame)
        return frames if self.rank == 0 else None


class PeerFrameRenderer:
    """The gather fused into the render kernel (include/ray_b200.h, "peer-memory frames").

    Rank 0 owns a ring of `slots` frames (int32[h][w] each, cudaMalloc + CUDA IPC) followed by two 32-bit flags per slot;
    every other rank maps that allocation over NVLink.  A frame is rendered by ALL ranks at once, each with its shard set
    and rank 0's frame slot as the row-major output: the kernels' pixel stores land in rank 0's memory, and the last warp
    of each rank's kernel bumps the slot's `done` flag.  Rank 0's copy stream waits for `done` to reach world x (use count),
    copies the frame to page-locked host memory and releases the slot by setting `ack`; the producers' next render into
    that slot waits for the ack.  No tile buffers, no ncclGather, no de-tiling kernel, no host synchronisation between
    frames: rendering, the NVLink transfer and the device-to-host copy of consecutive frames overlap.

    torch is used for the copy stream / pinned host buffers and for ONE object broadcast of the IPC handle at set-up."""

    FLAG_BYTES = 256  # per slot: done flag at +0, ack flag at +128 (separate lines)

    def __init__(self, ctx, rank, world, h, w, slots=4, group=None, same_process_base=None, pipeline=False):
        """pipeline: switch the context to pipelined submission (ray_b200_context_set_pipeline): consecutive `render` calls
        overlap on the GPU - the flags of this protocol are what orders the consumer.
        same_process_base: tests only - the rank-0 renderer's `base` when several "ranks" are contexts of ONE process on
        one GPU (a process cannot open its own IPC handle)."""
        import torch

        self.ctx, self.rank, self.world, self.h, self.w, self.slots = ctx, rank, world, h, w, slots
        self.torch = torch
        ctx.set_shard(rank, world)
        self.pipeline = bool(pipeline)
        if self.pipeline:
            ctx.set_pipeline(True)
        self.frame_bytes = (h * w * 4 + 255) // 256 * 256
        total = slots * (self.frame_bytes + self.FLAG_BYTES)
        self.mapped = False
        if rank == 0:
            self.base, handle = ctx.ipc_alloc(total)
        if same_process_base is not None:
            if rank != 0:
                self.base = same_process_base
        elif world > 1:
            import torch.distributed as dist
            box = [handle if rank == 0 else None]
            dist.broadcast_object_list(box, src=0, group=group)
            if rank != 0:
                self.base = ctx.ipc_open(box[0])
                self.mapped = True
        self.uses = [0] * slots          # how often each slot has been rendered into
        self.seq = 0
        if rank == 0:
            cuda = torch.cuda.is_available()   # (the CPU protocol test drives this class with a stand-in context)
            self.copy_stream = torch.cuda.Stream() if cuda else None
            self.watch_stream = torch.cuda.Stream() if cuda else None
            self.landed = torch.cuda.Event() if cuda else None   # recorded on the copy stream when the last consumed frame has fully arrived
            self.host = [torch.empty((h, w), dtype=torch.int32, pin_memory=cuda) for _ in range(slots)]

    def _slot(self, s):
        frame = self.base + s * self.frame_bytes
        flags = self.base + self.slots * self.frame_bytes + s * self.FLAG_BYTES
        return frame, flags, flags + 128

    def jobs(self, frames):
        """frames: sequence of (prepared, spp).  Returns render_batch job dicts for the next len(frames) ring slots and
        the slot numbers (call `consume` with them on rank 0 afterwards)."""
        out, used = [], []
        for prepared, spp in frames:
            s = self.seq % self.slots
            frame, done, ack = self._slot(s)
            out.append(dict(prepared=prepared, h=self.h, w=self.w, spp=spp, out_dev=frame, done_flag=done,
                            wait_flag=ack if self.uses[s] > 0 else None, wait_value=self.uses[s]))
            self.uses[s] += 1
            used.append(s)
            self.seq += 1
        return out, used

    def submit(self, frames):
        """Enqueues the frames on this rank (one ray_b200_render_batch: two in flight); returns their ring slots."""
        jobs, used = self.jobs(frames)
        self.ctx.render_batch(jobs)
        return used

    def render(self, frames):
        """submit + (rank 0) consume: the frames' hand-over to host memory on the copy stream.  Returns the pinned host
        tensors on rank 0 (valid after `wait`), else None.
        The flag waits are 1-thread spinning kernels: a wait only ever holds back work enqueued AFTER it on a stream that
        shares its hardware queue, and it only depends on work enqueued BEFORE it (on this rank) and on the other ranks, so
        one process per GPU cannot deadlock; several "ranks" inside ONE process (tests) must submit on all of them before
        any consume."""
        return self.consume(self.submit(frames))

    def consume(self, used):
        if self.rank != 0:
            return None
        cs = self.copy_stream.cuda_stream if self.copy_stream is not None else None
        if self.watch_stream is not None:   # "all frames have landed" without the device-to-host copies in between
            for s in used:
                self.ctx.flag_wait(self._slot(s)[1], self.world * self.uses[s], stream=self.watch_stream.cuda_stream)
            self.landed.record(self.watch_stream)   # every rank's pixels of these frames are in rank 0's HBM
        outs = []
        for s in used:
            frame, done, ack = self._slot(s)
            self.ctx.flag_wait(done, self.world * self.uses[s], stream=cs)
            self.ctx.copy_to_host_async(self.host[s].data_ptr(), frame, self.h * self.w * 4, stream=cs)
            self.ctx.flag_set(ack, self.uses[s], stream=cs)
            outs.append(self.host[s])
        return outs

    def flag_timeouts_safe(self):
        try:
            return self.ctx.flag_timeouts()
        except Exception:
            return 0

    def wait(self):
        """Blocks until every frame handed to `consume` is in host memory (rank 0); raises if a flag wait timed out."""
        if self.rank == 0 and self.copy_stream is not None:
            self.copy_stream.synchronize()
            self.watch_stream.synchronize()
        self.ctx.sync()
        if self.ctx.flag_timeouts():
            raise RuntimeError("PeerFrameRenderer: a peer-frame flag wait timed out (a rank did not deliver its pixels)")

    def close(self):
        self.ctx.sync()
        if self.pipeline:
            self.ctx.set_pipeline(False)
        if self.rank == 0:
            if self.copy_stream is not None:
                self.copy_stream.synchronize()
                self.watch_stream.synchronize()
            self.ctx.ipc_free(self.base)
        elif self.mapped:
            self.ctx.ipc_close(self.base)
        self.base = None
