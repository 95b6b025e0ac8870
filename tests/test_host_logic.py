"""Host-side logic that needs no GPU: synthetic generators, level sizes, CSR ground truth, the head's
constructor / state_dict contract, sharding helpers."""
import numpy as np
import pytest
import torch

from oracle import ref_loader
from sihl_b200 import dist as sdist
from sihl_b200 import synth
from sihl_b200.heads import ObjectDetection


def test_level_sizes_match_the_configs():
    assert synth.level_sizes(640, 640) == [(80, 80), (40, 40), (20, 20), (10, 10), (5, 5)]
    assert synth.num_anchors(synth.level_sizes(640, 640)) == 8525
    assert synth.num_anchors(synth.level_sizes(1024, 1024)) == 21824
    assert synth.num_anchors(synth.level_sizes(1280, 1280)) == 34100
    assert synth.level_sizes(320, 320) == [(40, 40), (20, 20), (10, 10), (5, 5), (3, 3)]      # FPN at 320²: P7 is 3x3
    assert synth.num_anchors(synth.level_sizes(320, 320)) == 2134
    assert synth.level_sizes(128, 128, mode="floor") == [(16, 16), (8, 8), (4, 4), (2, 2), (1, 1)]


def test_gt_generator_is_seeded_ragged_and_sane():
    a = synth.gt_batch_np(5, 6, 640, 640, 80, 100)
    b = synth.gt_batch_np(5, 6, 640, 640, 80, 100)
    np.testing.assert_array_equal(a.boxes, b.boxes)
    counts = np.diff(a.offsets)
    assert counts[0] == 0 and counts[1] == 100 and a.boxes.shape == (counts.sum(), 4)
    w, h = a.boxes[:, 2] - a.boxes[:, 0], a.boxes[:, 3] - a.boxes[:, 1]
    assert (w >= 1).all() and (h >= 1).all() and (a.boxes >= 0).all() and (a.boxes <= 640).all()
    assert (a.boxes != np.round(a.boxes)).mean() > 0.9            # non-lattice coordinates
    assert a.classes.dtype == np.int64 and a.classes.max() < 80


def test_gt_csr_from_lists_host_side():
    from sihl_b200 import ops
    boxes = [torch.zeros((0, 4)), torch.rand((3, 4)), torch.rand((1, 4))]
    classes = [torch.zeros((0,), dtype=torch.int64), torch.tensor([1, 2, 3]), torch.tensor([7])]
    gt = ops.GtBatch.from_lists(boxes, classes, "cpu")
    assert gt.counts == [0, 3, 1] and gt.total == 4 and gt.batch_size == 3
    assert gt.offsets.tolist() == [0, 0, 3, 4] and gt.offsets.dtype == torch.int32
    assert gt.boxes.shape == (4, 4) and gt.classes.tolist() == [1, 2, 3, 7]
    with pytest.raises(ValueError):
        ops.GtBatch.from_lists(boxes, classes[:2] + [torch.tensor([1, 2])], "cpu")


def test_head_constructor_and_state_dict_contract():
    head = ObjectDetection(in_channels=[3] + [256] * 7, num_classes=80, bottom_level=3, top_level=7)
    keys = list(head.state_dict().keys())
    assert len(keys) == 102 and sum(p.numel() for p in head.parameters()) > 1_400_000
    assert head.topk == 9 and head.max_instances == 100 and list(head.levels) == [3, 4, 5, 6, 7]
    assert set(head.output_shapes) == {"num_instances", "scores", "classes", "boxes"}
    assert float(head.loc_head[-2].bias.data[0]) == -5.0
    for bad in (dict(num_classes=0), dict(num_channels=30), dict(max_instances=0), dict(bottom_level=0)):
        kw = dict(in_channels=[3] + [8] * 7, num_classes=4, bottom_level=3, top_level=7, num_channels=8)
        kw.update(bad)
        with pytest.raises(AssertionError):
            ObjectDetection(**kw)
    with pytest.raises(AssertionError):
        ObjectDetection(in_channels=[3] * 4, num_classes=4, bottom_level=3, top_level=7)


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree only exists in the authoring container")
def test_state_dict_keys_equal_the_reference_head():
    Ref = ref_loader.ObjectDetection()
    kw = dict(in_channels=[3] + [32] * 7, num_classes=10, bottom_level=3, top_level=7, num_channels=32, num_layers=2)
    ours, ref = ObjectDetection(**kw), Ref(**kw)
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    assert [tuple(v.shape) for v in ours.state_dict().values()] == [tuple(v.shape) for v in ref.state_dict().values()]
    ours.load_state_dict(ref.state_dict())                                                    # checkpoints interchange


def test_shard_range_partitions_the_batch():
    for gb, ws in ((512, 8), (10, 4), (3, 8), (64, 1)):
        spans = [sdist.shard_range(gb, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
        sizes = [e - s for s, e in spans]
        assert max(sizes) - min(sizes) <= 1


def test_losses_from_sums_matches_oracle_finalize():
    from oracle import od_oracle as orc
    sums = torch.tensor([673.49, 97.0, 5614.77, 338.2, 292.5, 1650.1, 807.0, 0.0], dtype=torch.float64)
    np.testing.assert_allclose(sdist.losses_from_sums(sums).numpy(), orc.loss_finalize(sums.numpy()), rtol=1e-6)
    sums[6] = 0.0
    got = sdist.losses_from_sums(sums).numpy()
    assert got[1] == got[2] == got[3] == 0 and got[4] == got[0]                               # ref :165-172 early-out


def test_decode_mode_validation_needs_no_gpu():
    """Bad decode modes are rejected on the host before anything touches CUDA."""
    import pytest
    import torch

    from sihl_b200 import ops
    from sihl_b200.pipeline import DetectionHeadPipeline
    assert ops.DECODE_MODES == ("dense", "candidate_first")
    z = torch.zeros(1, 4)
    with pytest.raises(ValueError, match="mode="):
        ops.dense_decode(z, torch.zeros(1, 4, 8), torch.zeros(1, 4, 4), z, z, 8, 8, 0.05, None, mode="sparse")
    with pytest.raises(ValueError, match="decode_mode="):
        DetectionHeadPipeline([(2, 2)], 16, 16, 1, 8, 4, "cpu", decode_mode="sparse")
    # pinned host maps are only accepted by the gathering (candidate-first) decode: the dense scan is device-only
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        ops.dense_decode(z, torch.zeros(1, 4, 8), torch.zeros(1, 4, 4), z, z, 8, 8, 0.05, None, mode="dense")


def test_bench_config_is_the_workload_only_and_shared_by_both_arms():
    """The driver compares the `config` objects of `bench.py` and `bench.py --impl reference`: it must describe the workload
    and nothing of how this arm runs it (graphs, lanes, decode mode live in `execution`, input statistics in `workload_stats`)."""
    import argparse
    import inspect
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    args = argparse.Namespace(workload="cfg1", allreduce="fused")
    cfg = bench.config_dict(bench.WORKLOADS["cfg1"], args, 1)
    assert cfg["workload"].startswith("cfg1") and cfg["anchors"] == 8525 and cfg["global_batch"] == 64
    assert not {"decode_mode", "cuda_graph", "steps_in_flight", "streams_per_step", "positives_per_image",
                "candidates_per_image", "detections_per_image", "sample_images_per_step"} & set(cfg)
    assert "cache" in cfg                                   # the contract: say in `config` how the L2 is kept cold
    src = inspect.getsource(bench)
    assert src.count('"config": config_dict(w, args, world),') == 2      # both arms print the same object
    cfg8 = bench.config_dict(bench.WORKLOADS["cfg1"], args, 8)
    assert cfg8["global_batch"] == 512 and cfg8["batch_per_gpu"] == 64
