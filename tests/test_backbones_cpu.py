"""Host-side checks of the torch backbones that feed the path in the benchmark (backbones.py): the inference-time
BatchNorm folding (incl. the down-sample bias merged into the block's last convolution) must not change the function.
CPU only - the fused cuDNN / channels_last / side-stream paths are covered by tests/test_gpu_config2.py."""
import copy

import torch

from unseenobjectswithmeanshift_b200 import backbones


def _randomise_bn(model, seed):
    g = torch.Generator().manual_seed(seed)
    for m in model.modules():
        if isinstance(m, torch.nn.BatchNorm2d):
            m.running_mean.copy_(torch.randn(m.num_features, generator=g) * 0.1)
            m.running_var.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.weight.data.copy_(torch.rand(m.num_features, generator=g) + 0.5)
            m.bias.data.copy_(torch.randn(m.num_features, generator=g) * 0.1)
    return g


def test_resnet50_fold_is_function_preserving():
    plain = backbones.ResNet50Features(seed=3, fold_bn=False)
    g = _randomise_bn(plain, 1)
    fused = copy.deepcopy(plain).fold_()
    assert not any(isinstance(m, torch.nn.BatchNorm2d) for m in fused.modules())
    for blk in fused.res2:
        if blk.downsample is not None:
            assert blk.downsample[0].bias is None and blk.conv3.bias is not None   # one bias before the ReLU
    x = torch.randn(1, 3, 64, 96, generator=g)
    with torch.no_grad():
        want, got = plain(x), fused(x)
    assert set(want) == {"res2", "res3", "res4", "res5"}
    for k in want:
        assert got[k].shape == want[k].shape
        assert (got[k] - want[k]).abs().max().item() / want[k].abs().max().item() < 1e-5, k


def test_resnet50_never_leaves_eval_mode():
    m = backbones.ResNet50Features(seed=0)
    m.train()
    assert not m.training and all(not p.requires_grad for p in m.parameters())
