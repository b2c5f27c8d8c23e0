"""Multi-GPU partitioning of the raster path (one process per GPU, torch.distributed plumbing).

The path shards in two natural ways (BASELINE.json north_star):
  * independent FRAMES of a camera sweep per GPU -- no data-path collective at all;
  * screen-space TILE-ROW RANGES of one frame per GPU -- geometry is replicated, every rank
    rasterises only its rows (rz_set_row_range) and the resolved strips are gathered (one
    all-gather of equal contiguous strips; NCCL over NVLink on GPUs, gloo in the CPU tests).
Nothing here touches pixels: it is index arithmetic plus one collective call.
"""
from __future__ import annotations


def tile_rows(height: int, tile_h: int) -> int:
    return (height + tile_h - 1) // tile_h


def strip_rows(height: int, world: int, tile_h: int) -> int:
    """Rows per rank: tile rows split evenly (rounded up) so every strip has the same size."""
    return ((tile_rows(height, tile_h) + world - 1) // world) * tile_h


def row_range(rank: int, world: int, height: int, tile_h: int) -> tuple[int, int]:
    """Pixel rows [begin, end) owned by `rank`; tile-aligned except at the bottom edge. May be empty
    (begin == end == height) for trailing ranks of a short image."""
    per = strip_rows(height, world, tile_h)
    return min(height, rank * per), min(height, (rank + 1) * per)


def frame_ids(rank: int, world: int, n_frames: int) -> list[int]:
    """Frames of a sweep rendered by `rank` (round-robin, so neighbouring camera angles spread evenly)."""
    return list(range(rank, n_frames, world))


def gather_strips(strip, height: int, group=None):
    """All-gather equal-sized strips [strip_rows, W] into the full image [height, W] (on every rank).
    `strip` is a torch tensor (CUDA for nccl, CPU for gloo)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    full = torch.empty((world * strip.shape[0],) + tuple(strip.shape[1:]), dtype=strip.dtype, device=strip.device)
    dist.all_gather_into_tensor(full, strip.contiguous(), group=group)
    return full[:height]


FLAG_STRIDE = 128  # bytes between two ranks' flags (one cache line each)


class PeerFrame:
    """Tile-row sharding of ONE frame without a gather step: every rank's tile kernel stores its resolved rows
    straight into the presenting rank's image over NVLink (CUDA IPC peer memory, include/rz.h), and completion
    flags travel the same way.  `n_buffers` images alternate so a rank may start frame f+1 while the presenter
    still consumes frame f.

        pf = PeerFrame(renderer, group=None, root=0)       # collective: exchanges the IPC handles
        for f in range(n): ... renderer.render(...); ptr = pf.finish_frame()   # presenter: device pointer of the
                                                           # complete image (stream-ordered), others: None
                           pf.release()                    # presenter: image consumed, peers may overwrite it
    """

    def __init__(self, renderer, group=None, root: int = 0, n_buffers: int = 2, timeout_ms: int = 2000,
                 interleave_band: int = 0):
        import torch.distributed as dist

        self.r, self.group, self.root, self.nbuf, self.timeout_ms = renderer, group, root, n_buffers, timeout_ms
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world > 16:
            raise ValueError("PeerFrame supports up to 16 ranks (one node)")
        W, H = renderer.width, renderer.height
        from .render import tile_size

        # interleave_band > 0: rank q owns the bands k of `interleave_band` tile rows with k % world == q (balanced
        # for a centred object; peer stores make the strided destination free); 0: contiguous tile-row ranges
        self.interleave_band = interleave_band if self.world > 1 else 0
        self.rows = (0, H) if self.interleave_band else row_range(self.rank, self.world, H, tile_size()[1])
        self.image_bytes = W * H * 4
        mine = {}
        self._owned = []
        if self.rank == root:
            self.images = []
            for _ in range(n_buffers):
                p, h = renderer.shared_alloc(self.image_bytes)
                self.images.append(p)
                self._owned.append(p)
                mine.setdefault("images", []).append(h)
            self.flags, mine["flags"] = renderer.shared_alloc(FLAG_STRIDE * self.world)
            self._owned.append(self.flags)
        self.ack, mine["ack"] = renderer.shared_alloc(FLAG_STRIDE)  # the presenter raises it: "frame consumed"
        self._owned.append(self.ack)
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=group)
        self._mapped = []
        if self.rank == root:
            self.peer_acks = []
            for q in range(self.world):
                if q != root:
                    p = renderer.shared_open(everyone[q]["ack"])
                    self.peer_acks.append(p)
                    self._mapped.append(p)
        else:
            self.images = [renderer.shared_open(h) for h in everyone[root]["images"]]
            self.flags = renderer.shared_open(everyone[root]["flags"])
            self._mapped += self.images + [self.flags]
        self.frame = 0
        if self.rows[0] < self.rows[1]:
            renderer.set_row_range(*self.rows)
        renderer.set_row_interleave(self.interleave_band, self.rank, self.world)

    def finish_frame(self):
        """Run the recorded draws for this rank's rows, storing them into the presenter's image."""
        f, r = self.frame, self.r
        img = self.images[f % self.nbuf]
        seq = f + 1
        if self.rank != self.root and f >= self.nbuf:
            r.wait_flags(self.ack, 1, FLAG_STRIDE, f - self.nbuf + 1, self.timeout_ms)  # image buffer is free again
        if self.rows[0] < self.rows[1]:
            r.framebuffer_async(img + self.rows[0] * r.width * 4)
        else:
            r.discard_frame()  # this rank owns no rows (more ranks than tile rows): drop the recorded draws
        self.frame += 1
        if self.rank != self.root:
            r.signal(self.flags + self.rank * FLAG_STRIDE, seq)
            return None
        others = [q for q in range(self.world) if q != self.root]
        if others and self.root == 0:  # flags[1..world) are contiguous: one wait kernel
            r.wait_flags(self.flags + FLAG_STRIDE, self.world - 1, FLAG_STRIDE, seq, self.timeout_ms)
        else:
            for q in others:
                r.wait_flags(self.flags + q * FLAG_STRIDE, 1, FLAG_STRIDE, seq, self.timeout_ms)
        return img

    def release(self):
        """Presenter: the image returned by the last finish_frame() has been consumed (stream-ordered)."""
        if self.rank == self.root and self.peer_acks:
            self.r.signal(self.peer_acks, self.frame)

    def close(self):
        self.r.sync()
        import torch.distributed as dist

        dist.barrier(group=self.group)  # nobody unmaps / frees memory a peer may still be writing
        self.r.set_row_interleave(0, 0, 1)
        for p in self._mapped:
            self.r.shared_close(p)
        dist.barrier(group=self.group)
        for p in self._owned:
            self.r.shared_free(p)
        self._mapped, self._owned = [], []
