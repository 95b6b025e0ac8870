"""Multi-GPU plumbing of the path: images are independent, so the batch is sharded across
ranks (one process per GPU) and the only exchange is the 8 fp64 loss partial sums
(``include/sihl_od.h``: sums layout) — one all-reduce per step over NCCL/NVLink
(SURVEY.md §8e).  Everything here also runs on the ``gloo`` backend for the CPU tests.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

NUM_SUMS = 8
TOTAL_WEIGHTS = (1.0, 10.0, 1.0, 1.0)


def shard_range(global_batch: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous image range ``[start, end)`` of ``rank`` (remainder spread over the first ranks)."""
    base, rem = divmod(int(global_batch), int(world_size))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_reduce_sums(sums: torch.Tensor, group=None, async_op: bool = False):
    """Global-batch normalisers: sum the 8 partial sums over the ranks (in place)."""
    assert sums.numel() == NUM_SUMS and sums.dtype == torch.float64, (sums.shape, sums.dtype)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return None
    return dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def losses_from_sums(sums: torch.Tensor) -> torch.Tensor:
    """ref object_detection.py:163-172, :180, :197, :208, :210 on the 8 sums ->
    ``[location, box, class, iou, total]`` (fp64).  Eight scalars of host-side glue used by the
    distributed tests and for logging; on the GPU path ``sihl_od_loss_finalize`` does this."""
    s = sums.to(torch.float64)
    loc = s[0] / s[1]
    if float(s[6]) == 0.0:
        z = torch.zeros((), dtype=torch.float64, device=s.device)
        return torch.stack([loc, z, z, z, loc])
    iou, box, cls = s[2] / s[3], s[4] / s[3], s[5] / s[3]
    return torch.stack([loc, box, cls, iou, loc + 10.0 * box + cls + iou])


def ddp_mean_of_local_losses(local_losses: torch.Tensor, group=None) -> torch.Tensor:
    """The "DDP-faithful" semantics: every rank normalises by its own shard (what the reference
    does under Lightning DDP) and the logged value is the mean over ranks."""
    out = local_losses.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        out /= dist.get_world_size(group)
    return out


def _check(status: int, what: str) -> None:
    from . import _native
    _native.check(status, what)


class PeerExchange:
    """Exchange regions for the loss-sum all-reduce that ``k_pos_loss_tiles`` performs itself over peer memory
    (``include/sihl_od.h``: sihl_od_pos_loss_tiles_exchange).  One block of ``n_regions`` regions per GPU — one
    region per step in flight — allocated by the library, its CUDA-IPC handle all-gathered through
    ``torch.distributed`` (plumbing only) and the peers' blocks mapped into this process.

    ``peer_array(i)`` is the device array of ``world`` device pointers (entry r = rank r's region i) that the C
    entry takes.  Collective: every rank of ``group`` must construct it at the same point of its program."""

    def __init__(self, device, n_regions: int = 1, group=None):
        import ctypes as C

        from . import _native
        assert dist.is_available() and dist.is_initialized(), "PeerExchange needs an initialised process group"
        self.lib = _native.load()
        self.device = torch.device(device)
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.n_regions = int(n_regions)
        self.region_bytes = int(self.lib.sihl_od_exchange_region_bytes(self.world))
        self._own, self._opened, self.blocks = C.c_void_p(), {}, []

        def agree(error: Optional[str], what: str) -> None:
            """Every rank reaches every checkpoint; if any rank failed, ALL raise (nobody is left in a collective)."""
            flag = torch.tensor([0 if error is None else 1], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
            if int(flag.item()) != 0:
                self._release()
                raise RuntimeError(f"PeerExchange: {what} failed on " + ("this rank: " + error if error else "another rank"))

        handle = (C.c_ubyte * 64)()
        error = None
        with torch.cuda.device(self.device):
            try:
                if self.region_bytes <= 0:
                    raise RuntimeError(f"world size {self.world} not supported by the fused exchange")
                _native.check(self.lib.sihl_od_exchange_create(self.world, self.n_regions, C.byref(self._own), handle),
                              "sihl_od_exchange_create")
            except Exception as exc:                      # noqa: BLE001 - reported to every rank below
                error = str(exc)
            agree(error, "allocating / exporting the exchange block")
            mine = torch.tensor(list(handle), dtype=torch.uint8, device=self.device)
            gathered = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(gathered, mine, group=group)
            try:
                for r in range(self.world):
                    if r == self.rank:
                        self.blocks.append(int(self._own.value))
                        continue
                    raw = (C.c_ubyte * 64)(*gathered[r].cpu().tolist())
                    ptr = C.c_void_p()
                    _native.check(self.lib.sihl_od_exchange_open(raw, C.byref(ptr)), "sihl_od_exchange_open")
                    self._opened[r] = ptr
                    self.blocks.append(int(ptr.value))
            except Exception as exc:                      # noqa: BLE001
                error = str(exc)
            agree(error, "mapping the peers' exchange blocks (CUDA IPC / peer access)")
        # per region: device array of `world` pointers (entry r = rank r's region), what the kernel indexes
        self._tables = torch.tensor([[b + i * self.region_bytes for b in self.blocks] for i in range(self.n_regions)],
                                    dtype=torch.int64, device=self.device)

    def _release(self) -> None:
        with torch.cuda.device(self.device):
            for ptr in self._opened.values():
                self.lib.sihl_od_exchange_close(ptr)
            self._opened = {}
            if self._own.value is not None:
                self.lib.sihl_od_exchange_destroy(self._own)
                self._own.value = None

    def device_barrier(self, region: int, stream=None) -> None:
        """Enqueue a device-side barrier over the ranks on ``region`` (a region no pipeline is attached to): the stream
        continues once every rank's GPU has reached the same point.  No host synchronisation."""
        st = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        with torch.cuda.device(self.device):
            _check(self.lib.sihl_od_exchange_barrier(self.peer_array(region), self.world, self.rank, st),
                   "sihl_od_exchange_barrier")

    def set_timeout(self, seconds: float) -> None:
        """Bound of the in-kernel wait for the peers' sums (default 120 s).  A rank that is merely slow — dataloader
        stall, checkpoint, first-step lazy initialisation — must not be mistaken for a dead one."""
        with torch.cuda.device(self.device):
            for i in range(self.n_regions):
                _check(self.lib.sihl_od_exchange_set_timeout(self.blocks[self.rank] + i * self.region_bytes, self.world,
                                                              int(seconds * 1e9)), "sihl_od_exchange_set_timeout")

    def check(self) -> None:
        """Raise if any exchange on this rank ever gave up waiting for a peer (the kernel then wrote NaN sums for that
        step and latched the step number).  Synchronises the device; call it where a host sync is acceptable
        (end of an epoch, before a checkpoint)."""
        import ctypes as C
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for i in range(self.n_regions):
                bad, done = C.c_uint64(0), C.c_uint64(0)
                _check(self.lib.sihl_od_exchange_status(self.blocks[self.rank] + i * self.region_bytes, self.world,
                                                         C.byref(bad), C.byref(done)), "sihl_od_exchange_status")
                if bad.value:
                    raise RuntimeError(f"fused loss-sum exchange: rank {self.rank} timed out waiting for a peer at step "
                                       f"{bad.value} of region {i} ({done.value} exchanges done); that step's losses are NaN")

    def peer_array(self, region: int) -> int:
        """Device address of region ``region``'s pointer table."""
        return self._tables[region].data_ptr()

    def close(self, group=None) -> None:
        """Unmap the peers' blocks, then (after a barrier: nobody still maps ours) free our own."""
        if self._own.value is None:
            return
        torch.cuda.synchronize(self.device)
        with torch.cuda.device(self.device):
            for ptr in self._opened.values():
                self.lib.sihl_od_exchange_close(ptr)
            self._opened = {}
            if dist.is_initialized():
                dist.barrier(group=group)
            self.lib.sihl_od_exchange_destroy(self._own)
        self._own.value = None
