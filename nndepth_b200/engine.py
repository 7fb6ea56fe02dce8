"""Inference engine: host pairs in, disparity maps out -- the call a user of the reference's
``evaluate.py`` / ``inference.py`` makes, minus the dataset plumbing.

Reference recipe (``nndepth/models/raft_stereo/scripts/evaluate.py:146-156``): ``Padder(divis_by=32)``
-> ``model(left, right)`` -> ``unpad(m_outputs[-1]["up_disp"])``.  ``Padder`` arithmetic follows
``nndepth/data/dataloaders/utils.py:5-21``: ``pad = (((d // k) + 1) * k - d) % k``, width split
``[pad//2, pad - pad//2]``, height all at the bottom, replicate mode.

One process drives one GPU.  Multi-GPU inference shards the batch of pairs (or, for the correlation
path alone, epipolar row bands) across ranks with no collective in the data path; ``gather_disparities``
is the single NCCL call, collecting the output maps.
"""
import torch
import torch.nn.functional as F


class Padder:
    """Pads images so both sides are divisible by ``divis_by`` (reference dataloaders/utils.py:5-21)."""

    def __init__(self, dims, divis_by=32):
        self.ht, self.wd = int(dims[-2]), int(dims[-1])
        # the reference's (((d // k) + 1) * k - d) % k is the distance to the next multiple of k: -d mod k
        extra_h, extra_w = -self.ht % divis_by, -self.wd % divis_by
        self.left, self.right = extra_w // 2, extra_w - extra_w // 2
        self.top, self.bottom = 0, extra_h

    @property
    def padded_size(self):
        return self.ht + self.top + self.bottom, self.wd + self.left + self.right

    def pad(self, *inputs):
        return [F.pad(x, (self.left, self.right, self.top, self.bottom), mode="replicate") for x in inputs]

    def unpad(self, x):
        return x[..., self.top:x.shape[-2] - self.bottom, self.left:x.shape[-1] - self.right]


def shard_range(total, rank, world_size):
    """Contiguous ``[begin, end)`` slice of ``total`` units for ``rank`` (remainder to the low ranks)."""
    base, extra = divmod(int(total), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def gather_disparities(local, world_size=None, group=None):
    """All-gather per-rank disparity maps ``(b_r, 1, H, W)`` along the batch axis (equal ``b_r``)."""
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size(group)
    if world_size == 1:
        return local
    out = torch.empty((world_size * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


class OverlappedGather:
    """``gather_disparities`` off the critical path: the all-gather of step ``i`` runs on its own stream while step
    ``i + 1`` computes.  ``submit(local)`` snapshots the rank's maps (the engine's graph output is overwritten by the next
    replay) and enqueues the collective behind that copy; ``result()`` makes the current stream wait for the latest one."""

    def __init__(self, world_size, group=None, device=None, depth=2):
        self.world_size, self.group = int(world_size), group
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self.depth, self._slots, self._next, self._last = depth, [], 0, None

    def submit(self, local):
        import torch.distributed as dist
        if self.world_size == 1:
            self._last = (local, None)
            return
        if not self._slots:
            for _ in range(self.depth):
                self._slots.append({"src": torch.empty_like(local),
                                    "dst": torch.empty((self.world_size * local.shape[0],) + tuple(local.shape[1:]),
                                                       dtype=local.dtype, device=local.device),
                                    "copied": torch.cuda.Event(), "done": torch.cuda.Event(), "used": False})
        slot = self._slots[self._next]
        self._next = (self._next + 1) % self.depth
        main = torch.cuda.current_stream(self.device)
        if slot["used"]:
            main.wait_event(slot["done"])          # the collective that last read this slot's snapshot has finished
        slot["src"].copy_(local)
        slot["copied"].record(main)
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(slot["copied"])
            dist.all_gather_into_tensor(slot["dst"], slot["src"], group=self.group)
            slot["done"].record(self.stream)
        slot["used"] = True
        self._last = (slot["dst"], slot["done"])

    def result(self):
        out, done = self._last
        if done is not None:
            torch.cuda.current_stream(self.device).wait_event(done)
        return out


def row_band(tensor, rank, world_size):
    """Rows ``[h0, h1)`` of a ``(B, C, H, W)`` tensor owned by ``rank`` (contiguous copy) and the range.

    The correlation path is independent per epipolar row: row ``h`` of the volume depends only on row
    ``h`` of the two feature maps and a lookup only on row ``h`` of the coordinates (reference
    ``raft_stereo/cost_volume.py:55-61,36-53``), so a single high-resolution pair is sharded across GPUs
    by row bands with no halo and no collective in the data path (BASELINE config 5).
    """
    h0, h1 = shard_range(tensor.shape[2], rank, world_size)
    return tensor[:, :, h0:h1].contiguous(), (h0, h1)


def gather_row_bands(local, height, world_size=None, group=None):
    """All-gather per-rank row bands ``(B, C, h_r, W)`` back into ``(B, C, height, W)``.

    Bands differ by at most one row (``shard_range``); short bands are padded to the tallest one for
    the fixed-size collective and trimmed afterwards.
    """
    import torch.distributed as dist
    if world_size is None:
        world_size = dist.get_world_size(group)
    if world_size == 1:
        return local
    tallest = -(-int(height) // world_size)
    B, C, h, W = local.shape
    padded = local if h == tallest else F.pad(local, (0, 0, 0, tallest - h))
    out = torch.empty((world_size * B, C, tallest, W), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded.contiguous(), group=group)
    out = out.view(world_size, B, C, tallest, W)
    bands = []
    for r in range(world_size):
        h0, h1 = shard_range(height, r, world_size)
        bands.append(out[r, :, :, :h1 - h0])
    return torch.cat(bands, dim=2)


class StereoEngine:
    """Host-to-host stereo inference on one GPU with CUDA-graph replay.

    ``infer(left, right)`` takes host (ideally pinned) or device ``(B,3,H,W)`` float32 image batches,
    normalised like the reference datasets (``(x - 127.5) / 127.5``), and returns the final disparity
    ``(B,1,H,W)`` on the host (pinned) -- H2D copy, pad, forward, unpad and D2H all on the current stream.
    """

    def __init__(self, model, device=None, use_cuda_graph=True, divis_by=32, final_only=False, channels_last_encoder=True,
                 cudnn_benchmark=True):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.model = model.to(self.device).eval()
        # final_only=False keeps the reference forward's behaviour (an upsampled map per iteration);
        # True upsamples only the last iteration, which is all evaluate.py:155 reads.
        self.model.final_only = final_only
        ub = getattr(self.model, "update_block", None)
        gru = getattr(ub, "gru", None)
        if hasattr(gru, "fuse_gates"):
            gru.fuse_gates()
        # the fused ConvGRU keeps the hidden state channels-last: give the heads that read it channels-last
        # weights once, instead of letting cuDNN convert them on every call
        for name in ("flow_head", "mask", "encoder"):
            head = getattr(ub, name, None)
            if head is not None:
                head.to(memory_format=torch.channels_last)
        if hasattr(getattr(ub, "encoder", None), "channels_last"):
            ub.encoder.channels_last = True
        # the feature encoder runs on cuDNN's tensor-core kernels, which are NHWC inside: channels-last weights and
        # a channels-last input image spare it a layout conversion around every convolution
        self.channels_last_encoder = channels_last_encoder
        if channels_last_encoder and hasattr(self.model, "fnet"):
            self.model.fnet.to(memory_format=torch.channels_last)
        # fixed shapes, replayed many times: let cuDNN time its candidates once per convolution shape (its heuristic
        # picks a 4x slower pre-Blackwell kernel for the ConvGRU's doubled-output convolutions otherwise)
        if cudnn_benchmark:
            torch.backends.cudnn.benchmark = True
        self.use_cuda_graph = use_cuda_graph
        self.divis_by = divis_by
        self._dev_in = {}
        self._host_out = {}
        self._pipe, self._pipe_next, self._copy_stream, self._h2d_stream = {}, {}, None, None

    @torch.no_grad()
    def infer_device(self, left, right):
        """Device tensors in, device disparity out (padded internally, un-padded on return)."""
        padder = Padder(left.shape, self.divis_by)
        left_p, right_p = padder.pad(left, right)
        if self.channels_last_encoder:
            left_p = left_p.contiguous(memory_format=torch.channels_last)
            right_p = right_p.contiguous(memory_format=torch.channels_last)
        if self.use_cuda_graph:
            outputs = self.model.forward_graphed(left_p, right_p)
        else:
            outputs = self.model(left_p, right_p)
        return padder.unpad(outputs[-1]["up_disp"])

    # ---- pipelined host-to-host inference: the next pair's H2D and the previous pair's D2H overlap the forward ----
    def _slots(self, key):
        slots = self._pipe.get(key)
        if slots is None:
            slots = []
            for _ in range(2):
                slots.append({
                    "dev_l": torch.empty(key, dtype=torch.float32, device=self.device),
                    "dev_r": torch.empty(key, dtype=torch.float32, device=self.device),
                    "dev_out": torch.empty((key[0], 1, key[2], key[3]), dtype=torch.float32, device=self.device),
                    "host_out": torch.empty((key[0], 1, key[2], key[3]), dtype=torch.float32).pin_memory(),
                    "h2d": torch.cuda.Event(), "consumed": torch.cuda.Event(), "out_ready": torch.cuda.Event(),
                    "d2h": torch.cuda.Event(), "used": False})
            self._pipe[key] = slots
            self._pipe_next[key] = 0
        return slots

    @torch.no_grad()
    def submit(self, left, right):
        """Enqueue one batch of (pinned) host images; returns a ticket for ``collect``.  Copies run on their own
        streams: with two batches in flight the H2D of batch i+1 and the D2H of batch i-1 hide behind forward i."""
        key = tuple(left.shape)
        slots = self._slots(key)
        i = self._pipe_next[key]
        self._pipe_next[key] = (i + 1) % len(slots)
        slot = slots[i]
        if self._copy_stream is None:
            # one stream per direction: a D2H queued behind forward i must not hold back the H2D of batch i+1
            self._copy_stream = torch.cuda.Stream(device=self.device)
            self._h2d_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        if slot["used"]:
            slot["d2h"].synchronize()                    # the host buffer of this slot has been read out
            self._h2d_stream.wait_event(slot["consumed"])    # and its device inputs are no longer being read
        with torch.cuda.stream(self._h2d_stream):
            slot["dev_l"].copy_(left, non_blocking=True)
            slot["dev_r"].copy_(right, non_blocking=True)
            slot["h2d"].record(self._h2d_stream)
        main.wait_event(slot["h2d"])
        disp = self.infer_device(slot["dev_l"], slot["dev_r"])
        slot["consumed"].record(main)
        slot["dev_out"].copy_(disp)                      # the graph's static output is overwritten by the next replay
        slot["out_ready"].record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(slot["out_ready"])
            slot["host_out"].copy_(slot["dev_out"], non_blocking=True)
            slot["d2h"].record(self._copy_stream)
        slot["used"] = True
        return slot

    def collect(self, ticket):
        """Wait for a submitted batch; the returned pinned tensor is reused two submissions later."""
        ticket["d2h"].synchronize()
        return ticket["host_out"]

    @torch.no_grad()
    def infer(self, left, right):
        key = tuple(left.shape)
        if key not in self._dev_in:
            self._dev_in[key] = (torch.empty(key, dtype=torch.float32, device=self.device),
                                 torch.empty(key, dtype=torch.float32, device=self.device))
            self._host_out[key] = torch.empty((key[0], 1, key[2], key[3]), dtype=torch.float32).pin_memory()
        dl, dr = self._dev_in[key]
        dl.copy_(left, non_blocking=True)
        dr.copy_(right, non_blocking=True)
        disp = self.infer_device(dl, dr)
        host = self._host_out[key]
        host.copy_(disp, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return host
