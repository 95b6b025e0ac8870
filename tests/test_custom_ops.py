"""The ``torch.library`` shim (sihl_b200/torch_ops.py): registration, fake implementations and traceability — checked
WITHOUT a GPU by exporting the head (``torch.export`` traces with FakeTensors only, no kernel runs), and on the GPU by
compiling it (tests marked ``gpu``)."""
import pytest
import torch

from sihl_b200 import torch_ops
from sihl_b200.heads import ObjectDetection

CH, NCLS, SIZE, TOP = 16, 6, 128, 5


def _head():
    torch.manual_seed(0)
    return ObjectDetection([3] + [CH] * TOP, NCLS, bottom_level=3, top_level=TOP, num_channels=CH, num_layers=1)


def _inputs(device="cpu", batch=2):
    g = torch.Generator().manual_seed(1)
    return [torch.randn((batch, 3, SIZE, SIZE), generator=g).to(device)] + [
        torch.randn((batch, CH, SIZE // 2 ** l, SIZE // 2 ** l), generator=g).to(device) for l in range(1, TOP + 1)]


def _custom_calls(graph_module):
    return [str(n.target) for n in graph_module.graph.nodes if n.op == "call_function" and "sihl_b200" in str(n.target)]


def test_every_op_is_registered_with_a_fake_implementation():
    for name in torch_ops.REGISTERED:
        op = getattr(torch.ops.sihl_b200, name)
        assert op.default._schema.name == "sihl_b200::" + name
    # fake implementations give shapes and dtypes without a device
    from torch._subclasses.fake_tensor import FakeTensorMode
    with FakeTensorMode():
        loc = torch.empty((3, 336), dtype=torch.bfloat16, device="cuda")
        top, idx = torch_ops.topk_locations(loc, 100)
        assert top.shape == (3, 100) and top.dtype == torch.float32 and idx.dtype == torch.int64 and idx.device.type == "cuda"
        lv = [16, 16, 8, 8, 4, 4]
        off, sc, an = torch_ops.anchor_tables(loc, lv, 128, 128)
        assert off.shape == (336, 4) and an.dtype == torch.float32
        num, scores, classes, boxes = torch_ops.dense_postprocess(loc, torch.empty((3, 336, 6), device="cuda"),
                                                                  torch.empty((3, 336, 4), device="cuda"), lv, 128, 128, 0.05, 0.5, 50)
        assert num.shape == (3,) and scores.shape == (3, 50) and classes.dtype == torch.int64 and boxes.shape == (3, 50, 4)
        a, r, p, m = torch_ops.train_assign(torch.empty((7, 4), device="cuda"), torch.empty((4,), dtype=torch.int32, device="cuda"),
                                            lv, 128, 128, 9, 63)
        assert a.shape == (3, 336) and a.dtype == torch.int64 and p.shape == (63,) and m.shape == (16 + 3 + 2,)


def test_forward_exports_with_the_custom_ops_in_the_graph():
    """ref tests/heads/test_object_detection.py:83-107 exports ``forward`` through dynamo; the replacement must trace too."""
    head = _head().eval()
    ep = torch.export.export(head, (_inputs(),), strict=True)
    calls = _custom_calls(ep.graph_module)
    assert any("topk_locations" in c for c in calls) and any("decode_rows" in c for c in calls), calls
    outs = [n for n in ep.graph_module.graph.nodes if n.op == "output"][0].args[0]
    shapes = [tuple(o.meta["val"].shape) for o in outs]
    assert shapes == [(2,), (2, head.max_instances), (2, head.max_instances), (2, head.max_instances, 4)]


def test_training_step_traces_with_autograd_registered():
    head = _head().train()
    boxes = [torch.tensor([[10.0, 12.0, 50.0, 70.0], [30.0, 5.0, 100.0, 60.0]]), torch.zeros((0, 4))]
    classes = [torch.tensor([1, 3]), torch.zeros((0,), dtype=torch.int64)]

    class Step(torch.nn.Module):
        def __init__(self, head):
            super().__init__()
            self.head = head

        def forward(self, inputs, classes, boxes):
            loss, metrics = self.head.training_step(inputs, classes, boxes)
            return loss, metrics["box_loss"]

    ep = torch.export.export(Step(head), (_inputs(), classes, boxes), strict=True)
    calls = _custom_calls(ep.graph_module)
    assert any("train_assign" in c for c in calls) and any("train_loss" in c for c in calls), calls


@pytest.mark.gpu
def test_compiled_forward_equals_eager_on_gpu():
    head = _head().to("cuda:0").eval()
    inputs = _inputs("cuda:0")
    with torch.no_grad():
        want = head(inputs)
        compiled = torch.compile(head, fullgraph=True, backend="aot_eager")
        got = compiled(inputs)
    for w, g in zip(want, got):
        assert w.dtype == g.dtype and torch.equal(w, g)


@pytest.mark.gpu
def test_custom_op_training_path_equals_the_eager_head_on_gpu():
    """The traced training step (custom ops + registered autograd) gives the loss and parameter gradients of the eager
    head (direct ctypes path), and ``torch.library.opcheck`` accepts the ops (schema, fake impl, autograd registration)."""
    head = _head().to("cuda:0").eval()
    inputs = _inputs("cuda:0")
    boxes = [torch.tensor([[10.0, 12.0, 50.0, 70.0], [30.0, 5.0, 100.0, 60.0]], device="cuda:0"), torch.zeros((0, 4), device="cuda:0")]
    classes = [torch.tensor([1, 3], device="cuda:0"), torch.zeros((0,), dtype=torch.int64, device="cuda:0")]
    loss, metrics = head.training_step(inputs, classes, boxes)
    loss.backward()
    want = {n: p.grad.clone() for n, p in head.named_parameters()}
    head.zero_grad()
    levels = head._level_sizes(inputs)
    from sihl_b200.heads.object_detection import _cat_gt
    gt_boxes, gt_classes, counts = _cat_gt(boxes, classes, torch.device("cuda:0"))
    loss2, metrics2 = head._training_step_traced(inputs, levels, SIZE, SIZE, gt_boxes, gt_classes, counts)
    loss2.backward()
    assert loss2.item() == loss.item()
    for n, p in head.named_parameters():
        assert torch.equal(p.grad, want[n]), n
    loc = torch.randn((2, 336), device="cuda:0")
    torch.library.opcheck(torch.ops.sihl_b200.topk_locations.default, (loc, 10))
    compiled = torch.compile(head.training_step, fullgraph=True, backend="aot_eager")
    head.zero_grad()
    loss3, _ = compiled(inputs, classes, boxes)
    loss3.backward()
    assert loss3.item() == loss.item()
    for n, p in head.named_parameters():
        assert torch.equal(p.grad, want[n]), n
