"""world_size-2 check (gloo, CPU) of the training step's distributed pieces (BASELINE.json config #5): the criterion's
num_masks all-reduce (criterion.py:221-227) and the DDP gradient average, against a single-process computation of
the same two-rank batch."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q, K, H, W, P = 6, 2, 12, 16, 24


class ToyModel(nn.Module):
    """image -> {'pred_logits','pred_masks'} + criterion, with the META_ARCH training contract (loss dict)."""

    def __init__(self, criterion):
        super().__init__()
        self.mask = nn.Conv2d(3, Q, 3, padding=1)
        self.cls = nn.Linear(Q, Q * (K + 1))
        self.criterion = criterion
        self.points = None

    def forward(self, batch):
        x = torch.stack([b["image"] for b in batch])
        masks = self.mask(x)
        logits = self.cls(masks.mean((2, 3))).view(len(batch), Q, K + 1)
        targets = [{"labels": b["labels"], "masks": b["masks"]} for b in batch]
        losses = self.criterion({"pred_logits": logits, "pred_masks": masks, "aux_outputs": []}, targets, self.points)
        return {k: v * self.criterion.weight_dict[k] for k, v in losses.items()}


class FixedPoints:
    def __init__(self, seed):
        self.g = torch.Generator().manual_seed(seed)

    def matcher_points(self, layers, batch, n, device):
        return torch.rand(layers, batch, n, 2, generator=self.g)

    def oversampled_points(self, layers, masks, n, device):
        return torch.rand(layers, masks, n, 2, generator=self.g)

    def random_points(self, layers, masks, n, device):
        return torch.rand(layers, masks, n, 2, generator=self.g)


def _build():
    sys.path.insert(0, ROOT)
    from unseenobjectswithmeanshift_b200.meanshiftformer.meanshiftformer_model import build_criterion
    torch.manual_seed(0)
    return ToyModel(build_criterion(K, deep_supervision=False, train_num_points=P))


def _rank_batch(rank):
    g = torch.Generator().manual_seed(100 + rank)
    T = 3 if rank == 0 else 1   # different numbers of ground-truth masks per rank: num_masks = (3 + 1) / 2
    masks = torch.zeros(T, H, W, dtype=torch.bool)
    for t in range(T):
        masks[t, 2 + 3 * t:6 + 3 * t, 1 + 4 * t:7 + 4 * t] = True
    return [{"image": torch.rand(3, H, W, generator=g), "labels": torch.randint(0, K, (T,), generator=g), "masks": masks}]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    model = _build()
    from unseenobjectswithmeanshift_b200 import sharding, training
    sharding.init_from_env("gloo")
    ddp = training.wrap_ddp(model)
    assert isinstance(ddp, nn.parallel.DistributedDataParallel)
    model.points = FixedPoints(7 + rank)
    losses = ddp(_rank_batch(rank))
    sum(losses.values()).backward()
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    q.put((rank, {k: float(v) for k, v in losses.items()}, {n: g.tolist() for n, g in grads.items()}))
    dist.destroy_process_group()


def test_two_rank_training_step_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=180) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    # single process: each rank's batch with num_masks forced to the two-rank average (3 + 1) / 2 = 2
    want_losses, want_grads = [], []
    for rank in range(world):
        model = _build()
        model.points = FixedPoints(7 + rank)
        batch = _rank_batch(rank)
        T = batch[0]["labels"].numel()
        losses = model(batch)
        # the criterion divides the mask losses by this rank's own count T; rescale them to the global average 2
        scaled = {k: (v * T / 2.0 if k != "loss_ce" else v) for k, v in losses.items()}
        model.zero_grad()
        sum(scaled.values()).backward()
        want_losses.append({k: float(v.detach()) for k, v in scaled.items()})
        want_grads.append({n: p.grad.clone() for n, p in model.named_parameters()})
    for rank in range(world):
        for k, v in want_losses[rank].items():
            assert abs(res[rank][1][k] - v) <= 1e-5 * max(1.0, abs(v)), (rank, k)
    for n in want_grads[0]:
        avg = (want_grads[0][n] + want_grads[1][n]) / 2   # DDP averages the gradients of the ranks
        for rank in range(world):
            torch.testing.assert_close(torch.tensor(res[rank][2][n]), avg, rtol=1e-4, atol=1e-6)


def test_train_step_updates_parameters_and_clips():
    from unseenobjectswithmeanshift_b200 import training
    model = _build()
    model.points = FixedPoints(3)
    opt = training.build_optimizer(model, lr=1e-2)
    assert sorted((g["lr"], g["weight_decay"]) for g in opt.param_groups) == [(1e-2, 0.05)]
    before = {n: p.detach().clone() for n, p in model.named_parameters()}
    losses = training.train_step(model, opt, _rank_batch(0), clip_value=0.01)
    assert set(losses) == {"loss_ce", "loss_mask", "loss_dice"} and not any(v.requires_grad for v in losses.values())
    assert any(not torch.equal(before[n], p) for n, p in model.named_parameters())
    gnorm = torch.sqrt(sum((p.grad ** 2).sum() for p in model.parameters()))
    assert float(gnorm) <= 0.01 * 1.001   # full-model clipping (Base-COCO-InstanceSegmentation.yaml:29-32)


def test_optimizer_groups_follow_the_reference_rules():
    """tabletop_train_net_pretrained.py:113-160: modules named *backbone* train at lr x BACKBONE_MULTIPLIER, norm layers
    and embeddings get their own weight decay; tensors with equal hyper-parameters share a group."""
    from unseenobjectswithmeanshift_b200 import training

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = nn.Sequential(nn.Conv2d(3, 4, 1), nn.GroupNorm(2, 4))
            self.head = nn.Linear(4, 4)
            self.norm = nn.LayerNorm(4)
            self.query_feat = nn.Embedding(5, 4)
            self.frozen = nn.Linear(2, 2).requires_grad_(False)

    net = Net()
    opt = training.build_optimizer(net, lr=1e-4, weight_decay=0.05, backbone_multiplier=0.1)
    groups = {(round(g["lr"], 9), g["weight_decay"]): {id(p) for p in g["params"]} for g in opt.param_groups}
    assert set(groups) == {(1e-5, 0.05), (1e-5, 0.0), (1e-4, 0.05), (1e-4, 0.0)}
    assert groups[(1e-5, 0.05)] == {id(p) for p in net.backbone[0].parameters()}
    assert groups[(1e-5, 0.0)] == {id(p) for p in net.backbone[1].parameters()}
    assert groups[(1e-4, 0.05)] == {id(p) for p in net.head.parameters()}
    assert groups[(1e-4, 0.0)] == {id(p) for p in list(net.norm.parameters()) + list(net.query_feat.parameters())}
    assert not any(id(p) in s for p in net.frozen.parameters() for s in groups.values())
