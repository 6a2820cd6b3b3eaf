"""Multi-GPU plumbing: one process per GPU, trajectories sharded, weights replicated.

The forward / eval recursion needs NO collective (every reduction is inside one trajectory,
SURVEY.md section 8e); the only exchange on the path is the gradient all-reduce of BPTT training
(NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
import ctypes as C
from typing import Dict, Iterable, Tuple

import torch
import torch.distributed as dist


def shard_bounds(num_trajectories: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of [0, N): the first N % W ranks get one extra trajectory."""
    assert 0 <= rank < world_size
    base, extra = divmod(num_trajectories, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tree, rank: int, world_size: int, batch_dim: int = 1):
    """Slice every tensor of a (possibly dict-valued) batch along its trajectory axis
    ((T, N, ...) -> batch_dim=1; (N, ...) -> batch_dim=0)."""
    if isinstance(tree, dict):
        return {k: shard_batch(v, rank, world_size, batch_dim) for k, v in tree.items()}
    lo, hi = shard_bounds(tree.shape[batch_dim], rank, world_size)
    return tree.narrow(batch_dim, lo, hi - lo)


def gather_estimates(local: torch.Tensor, num_trajectories: int, batch_dim: int = 1) -> torch.Tensor:
    """All-gather per-rank (T, N_local, sd) estimates back into (T, N, sd) (ragged shards allowed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(num_trajectories, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad_shape = list(local.shape)
    pad_shape[batch_dim] = width
    padded = local.new_zeros(pad_shape)
    padded.narrow(batch_dim, 0, local.shape[batch_dim]).copy_(local)
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded.contiguous())
    return torch.cat([p.narrow(batch_dim, 0, hi - lo) for p, (lo, hi) in zip(parts, sizes)], dim=batch_dim)


def allreduce_gradients(module: torch.nn.Module, average: bool = True) -> int:
    """One flat fp32 all-reduce of every parameter gradient (the BPTT step's only collective).
    Returns the number of elements reduced."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return 0
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset:offset + n].view_as(g))
        offset += n
    return flat.numel()


class FlatGradients:
    """All parameter gradients of a training step as views into ONE flat fp32 buffer (the layout DDP calls
    ``gradient_as_bucket_view``): autograd accumulates into the views in place, the collective and the optimiser read the
    same memory, and there is no concatenate / copy-back around the all-reduce.  Call ``zero()`` instead of
    ``optimizer.zero_grad(set_to_none=True)`` (which would detach the views)."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev = self.params[0].device
        self.flat = torch.zeros(sum(p.numel() for p in self.params), device=dev, dtype=torch.float32)
        offset = 0
        for p in self.params:
            assert p.dtype == torch.float32 and p.device == dev
            p.grad = self.flat[offset:offset + p.numel()].view_as(p)
            offset += p.numel()

    def zero(self):
        self.flat.zero_()

    def intact(self) -> bool:
        """True while every ``p.grad`` still is the view handed out at construction."""
        base = self.flat.data_ptr()
        end = base + self.flat.numel() * 4
        return all(p.grad is not None and base <= p.grad.data_ptr() < end for p in self.params)


class _NcclUniqueId(C.Structure):
    _fields_ = [("internal", C.c_ubyte * 128)]  # raw bytes (a c_char array would read back truncated at the first NUL)


def _loaded_nccl() -> str:
    """Path of the NCCL library torch has mapped into this process (the bundled one), else the SONAME."""
    try:
        with open("/proc/self/maps") as maps:
            for line in maps:
                if "libnccl.so" in line:
                    return line.split()[-1]
    except OSError:
        pass
    return "libnccl.so.2"


class StreamAllReduce:
    """The BPTT step's only collective, issued as ``ncclAllReduce`` DIRECTLY on the caller's CUDA stream (ctypes on the
    NCCL library torch has already loaded; its own communicator, bootstrapped once through ``torch.distributed``).  On the
    stream it is a plain stream-ordered kernel launch: it can be captured into the CUDA graph of the training step together
    with forward, backward and the optimiser, which torch's process-group wrapper (side stream, watchdog, work objects)
    could not (round 1: the capture hung).  NVLink / NVSwitch transport, in-switch reduction when NCCL selects NVLS.
    Single process / no process group: a no-op."""

    FLOAT32, SUM, AVG = 7, 0, 4

    def __init__(self, device: torch.device):
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.device = device
        self.comm = None
        if self.world == 1:
            return
        assert device.type == "cuda", "StreamAllReduce is the NCCL path; CPU tests use allreduce_gradients (gloo)"
        self.lib = C.CDLL(_loaded_nccl())
        self.lib.ncclGetErrorString.restype = C.c_char_p
        self.lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, _NcclUniqueId, C.c_int]
        self.lib.ncclAllReduce.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        uid = _NcclUniqueId()
        if dist.get_rank() == 0:
            self._check(self.lib.ncclGetUniqueId(C.byref(uid)))
        # the 128 id bytes travel as a tensor through the process group that already exists
        wire = torch.tensor(list(bytes(uid)), dtype=torch.uint8, device=device)
        dist.broadcast(wire, src=0)
        raw = bytes(wire.cpu().tolist())
        assert len(raw) == 128
        C.memmove(C.byref(uid), raw, 128)
        comm = C.c_void_p()
        with torch.cuda.device(device):
            self._check(self.lib.ncclCommInitRank(C.byref(comm), self.world, uid, dist.get_rank()))
        self.comm = comm

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"NCCL error {rc}: {self.lib.ncclGetErrorString(C.c_int(rc)).decode()}")

    def __call__(self, flat: torch.Tensor, average: bool = True) -> int:
        """In-place all-reduce of a contiguous fp32 tensor on the current stream of its device."""
        if self.comm is None:
            return 0
        assert flat.is_cuda and flat.dtype == torch.float32 and flat.is_contiguous()
        stream = torch.cuda.current_stream(flat.device).cuda_stream
        with torch.cuda.device(flat.device):
            self._check(self.lib.ncclAllReduce(flat.data_ptr(), flat.data_ptr(), flat.numel(), self.FLOAT32,
                                               self.AVG if average else self.SUM, self.comm, stream))
        return flat.numel()

    def close(self):
        """Drop the communicator handle.  ncclCommDestroy is NOT called: it blocks for good while a CUDA graph that captured
        collectives of this communicator is alive (observed on 2 x B200: both ranks hung here after a clean measurement),
        and the handle's resources go back to the driver when the process exits."""
        self.comm = None


def max_over_ranks(value: float, device) -> float:
    """Timing reduction used by bench.py: the slowest rank defines the step time."""
    t = torch.tensor([value], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
