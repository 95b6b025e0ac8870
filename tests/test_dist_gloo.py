"""N>1 host logic on CPU: world_size 2, gloo backend.  Every rank computes the 8 partial sums of ITS shard of the
batch (here with the CPU oracle standing in for the kernels), the sums are all-reduced (the path's only exchange,
SURVEY.md §8e) and finalised; the result must equal one process over the global batch."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import od_oracle as orc
from sihl_b200 import dist as sdist
from sihl_b200 import synth

H = W = 128
C, GB = 16, 6


def _inputs():
    levels = synth.level_sizes(H, W, mode="floor")
    off, sc, an = orc.anchors(levels, W, H)
    gt = synth.gt_batch_np(77, GB, H, W, C, 6, counts=[3, 0, 5, 2, 6, 1])
    maps = synth.dense_maps_np(78, GB, len(an), C)
    return off, sc, an, gt, maps


def _shard_sums(rank, world):
    off, sc, an, gt, maps = _inputs()
    s, e = sdist.shard_range(GB, rank, world)
    g0, g1 = int(gt.offsets[s]), int(gt.offsets[e])
    res = orc.train_losses(an, off, sc, W, H, gt.boxes[g0:g1], gt.classes[g0:g1], gt.offsets[s:e + 1] - gt.offsets[s],
                           maps.loc_logits[s:e], maps.iou_preds[s:e], maps.box_raw[s:e], maps.cls_logits[s:e])
    return res["sums"], res["losses"]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sums, local_losses = _shard_sums(rank, world)
        t = torch.from_numpy(sums.copy())
        sdist.all_reduce_sums(t)
        global_losses = sdist.losses_from_sums(t)
        ddp = sdist.ddp_mean_of_local_losses(torch.from_numpy(local_losses.astype(np.float64)))
        if rank == 0:
            out.put((t.numpy().tolist(), global_losses.numpy().tolist(), ddp.numpy().tolist()))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(120)
def test_two_ranks_equal_one_process_over_the_global_batch():
    ctx = mp.get_context("spawn")
    out = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    sums, global_losses, ddp = out.get()
    want_sums, want_losses = _shard_sums(0, 1)                      # one process, whole batch
    np.testing.assert_allclose(sums[:7], want_sums[:7], rtol=1e-12)
    np.testing.assert_allclose(global_losses, want_losses, rtol=1e-6)
    # the DDP-faithful number (mean of locally normalised losses) is a different, labelled quantity
    l0, l1 = _shard_sums(0, 2)[1], _shard_sums(1, 2)[1]
    np.testing.assert_allclose(ddp, (l0.astype(np.float64) + l1) / 2, rtol=1e-6)


def test_all_reduce_is_a_no_op_without_a_process_group():
    t = torch.arange(8, dtype=torch.float64)
    assert sdist.all_reduce_sums(t) is None and t.tolist() == list(range(8))
