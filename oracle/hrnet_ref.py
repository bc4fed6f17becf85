"""ORACLE (test infrastructure, never shipped or timed as product): plain PyTorch
fp32 restatement of the reference's two HRNetV2-w48 heat-map networks.

Restates, with state_dict-compatible parameter names so reference checkpoints
(``nn_state_dict``) load unchanged:

* keypoint net: ``src/models/hrnet/hrnet.py:29-58`` (BasicBlock), ``:61-99``
  (Bottleneck), ``:102-246`` (HighResolutionModule, fuse at :222-246),
  ``:255-511`` (HighResolutionNet: stem :260-267, head :306-330, forward :437-511),
  wrapped by ``src/models/hrnet/model.py:130-150`` (HRNetHeatmap, 0 refinement
  stages as in every shipped config, ``train_config.yaml:34``).
* line net: ``src/models/line/hrnet.py:30-249`` (no stem skip, no upscale,
  Softmax head :86-102), wrapped by ``src/models/line/model.py:140-147``.

Pinned: ``tests/golden/make_golden_hrnet.py`` loads one state_dict into both the
real reference modules and these and requires bit-identical CPU outputs.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

W48 = dict(
    stem_width=64,
    stage1=dict(num_modules=1, num_branches=1, block="BOTTLENECK", num_blocks=[4], num_channels=[64]),
    stage2=dict(num_modules=1, num_branches=2, block="BASIC", num_blocks=[4, 4], num_channels=[48, 96]),
    stage3=dict(num_modules=4, num_branches=3, block="BASIC", num_blocks=[4, 4, 4], num_channels=[48, 96, 192]),
    stage4=dict(num_modules=3, num_branches=4, block="BASIC", num_blocks=[4, 4, 4, 4],
                num_channels=[48, 96, 192, 384]),
)
KEYPOINT_CFG = dict(W48, num_classes=58, upscale=2, head="logsoftmax")   # hrnet_w48.yaml (keypoints)
LINE_CFG = dict(W48, num_classes=23, upscale=1, head="softmax")          # line/model_config/hrnet_w48.yaml


def _variant(**stages):
    cfg = dict(W48)
    cfg.update(stages)
    return cfg


# the other shipped keypoint configs (hrnet_w18.yaml, hrnet_w64.yaml, hrnet_w48x4.yaml)
W18_CFG = dict(_variant(
    stage1=dict(num_modules=1, num_branches=1, block="BOTTLENECK", num_blocks=[1], num_channels=[32]),
    stage2=dict(num_modules=1, num_branches=2, block="BASIC", num_blocks=[2, 2], num_channels=[16, 32]),
    stage3=dict(num_modules=1, num_branches=3, block="BASIC", num_blocks=[2, 2, 2], num_channels=[16, 32, 64]),
    stage4=dict(num_modules=1, num_branches=4, block="BASIC", num_blocks=[2, 2, 2, 2], num_channels=[16, 32, 64, 128])),
    num_classes=58, upscale=2, head="logsoftmax")
W64_CFG = dict(_variant(
    stage2=dict(num_modules=1, num_branches=2, block="BASIC", num_blocks=[4, 4], num_channels=[64, 128]),
    stage3=dict(num_modules=4, num_branches=3, block="BASIC", num_blocks=[4, 4, 4], num_channels=[64, 128, 256]),
    stage4=dict(num_modules=3, num_branches=4, block="BASIC", num_blocks=[4, 4, 4, 4], num_channels=[64, 128, 256, 512])),
    num_classes=58, upscale=2, head="logsoftmax")
W48X4_CFG = dict(W48, num_classes=58, upscale=4, head="logsoftmax")


def _bn(c):
    # the reference aliases BatchNorm2d = SyncBatchNorm (hrnet.py:18); in eval mode
    # (the only mode on this path) both compute the same per-channel affine.
    return nn.BatchNorm2d(c, momentum=0.1)


class Basic(nn.Module):
    expansion = 1

    def __init__(self, cin, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 3, stride, 1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = _bn(planes)
        self.downsample = downsample

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        r = x if self.downsample is None else self.downsample(x)
        return F.relu(y + r)


class Bottle(nn.Module):
    expansion = 4

    def __init__(self, cin, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, planes, 1, bias=False)
        self.bn1 = _bn(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = _bn(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = _bn(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        r = x if self.downsample is None else self.downsample(x)
        return F.relu(y + r)


BLOCKS = {"BASIC": Basic, "BOTTLENECK": Bottle}


def _stack(block, cin, planes, n):
    ds = None
    if cin != planes * block.expansion:
        ds = nn.Sequential(nn.Conv2d(cin, planes * block.expansion, 1, bias=False),
                           _bn(planes * block.expansion))
    layers = [block(cin, planes, 1, ds)]
    layers += [block(planes * block.expansion, planes) for _ in range(1, n)]
    return nn.Sequential(*layers)


class HRModule(nn.Module):
    """One multi-resolution module: per-branch block stacks, then the all-to-all fuse."""

    def __init__(self, block, num_blocks: Sequence[int], channels: Sequence[int]):
        super().__init__()
        nb = len(channels)
        self.branches = nn.ModuleList(
            [_stack(block, channels[i], channels[i] // block.expansion, num_blocks[i]) for i in range(nb)])
        fuse = []
        for i in range(nb):
            row: List[nn.Module | None] = []
            for j in range(nb):
                if j == i:
                    row.append(None)
                elif j > i:      # lower resolution -> 1x1 + BN, upsampled in forward
                    row.append(nn.Sequential(nn.Conv2d(channels[j], channels[i], 1, 1, 0, bias=False),
                                             _bn(channels[i])))
                else:            # higher resolution -> chain of stride-2 3x3
                    chain = []
                    for k in range(i - j):
                        last = k == i - j - 1
                        co = channels[i] if last else channels[j]
                        mods = [nn.Conv2d(channels[j], co, 3, 2, 1, bias=False), _bn(co)]
                        if not last:
                            mods.append(nn.ReLU(inplace=True))
                        chain.append(nn.Sequential(*mods))
                    row.append(nn.Sequential(*chain))
            fuse.append(nn.ModuleList(row))
        self.fuse_layers = nn.ModuleList(fuse)

    def forward(self, xs: List[torch.Tensor]) -> List[torch.Tensor]:
        xs = [b(x) for b, x in zip(self.branches, xs)]
        out = []
        for i, row in enumerate(self.fuse_layers):
            y = xs[0] if i == 0 else row[0](xs[0])
            for j in range(1, len(xs)):
                if j == i:
                    y = y + xs[j]
                elif j > i:
                    y = y + F.interpolate(row[j](xs[j]), size=xs[i].shape[-2:], mode="bilinear",
                                          align_corners=True)
                else:
                    y = y + row[j](xs[j])
            out.append(F.relu(y))
        return out


def _transition(pre: Sequence[int], cur: Sequence[int]) -> nn.ModuleList:
    layers: List[nn.Module | None] = []
    for i, c in enumerate(cur):
        if i < len(pre):
            if c != pre[i]:
                layers.append(nn.Sequential(nn.Conv2d(pre[i], c, 3, 1, 1, bias=False), _bn(c),
                                            nn.ReLU(inplace=True)))
            else:
                layers.append(None)
        else:
            chain = []
            for j in range(i + 1 - len(pre)):
                co = c if j == i - len(pre) else pre[-1]
                chain.append(nn.Sequential(nn.Conv2d(pre[-1], co, 3, 2, 1, bias=False), _bn(co),
                                           nn.ReLU(inplace=True)))
            layers.append(nn.Sequential(*chain))
    return nn.ModuleList(layers)


class HRNetRef(nn.Module):
    """``HighResolutionNet`` of either flavour, selected by cfg['head'] / cfg['upscale']."""

    def __init__(self, cfg: Dict):
        super().__init__()
        sw = cfg["stem_width"]
        self.cfg = cfg
        self.conv1 = nn.Conv2d(3, sw, 3, 2, 1, bias=False)
        self.bn1 = _bn(sw)
        self.conv2 = nn.Conv2d(sw, sw, 3, 2, 1, bias=False)
        self.bn2 = _bn(sw)
        s1 = cfg["stage1"]
        blk1 = BLOCKS[s1["block"]]
        self.layer1 = _stack(blk1, 64, s1["num_channels"][0], s1["num_blocks"][0])
        pre = [blk1.expansion * s1["num_channels"][0]]
        for idx in (2, 3, 4):
            sc = cfg[f"stage{idx}"]
            blk = BLOCKS[sc["block"]]
            ch = [c * blk.expansion for c in sc["num_channels"]]
            setattr(self, f"transition{idx - 1}", _transition(pre, ch))
            setattr(self, f"stage{idx}", nn.Sequential(
                *[HRModule(blk, sc["num_blocks"], ch) for _ in range(sc["num_modules"])]))
            pre = ch
        self.upscale = cfg.get("upscale", 1)
        self.last_inp_channels = sum(pre) + (sw if self.upscale > 1 else 0)
        c = self.last_inp_channels
        act = nn.LogSoftmax(dim=1) if cfg["head"] == "logsoftmax" else nn.Softmax(dim=1)
        self.last_layer = nn.Sequential(nn.Conv2d(c, c, 1), _bn(c), nn.ReLU(inplace=True),
                                        nn.Conv2d(c, cfg["num_classes"], 1), act)

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        stem = x
        x = F.relu(self.bn2(self.conv2(x)))
        x = self.layer1(x)
        ys = [x]
        for idx in (2, 3, 4):
            tr = getattr(self, f"transition{idx - 1}")
            nprev = len(ys)
            xs = []
            for i, t in enumerate(tr):
                if t is None:
                    xs.append(ys[i])
                else:
                    xs.append(t(ys[i] if i < nprev else ys[-1]))
            for m in getattr(self, f"stage{idx}"):
                xs = m(xs)
            ys = xs
        h, w = int(ys[0].shape[2] * self.upscale), int(ys[0].shape[3] * self.upscale)
        feats = [stem] if self.upscale > 1 else []
        feats += list(ys)
        feats = [f if f.shape[-2:] == (h, w) else F.interpolate(f, size=(h, w), mode="bilinear",
                                                               align_corners=True) for f in feats]
        cat = torch.cat(feats, 1)
        return [self.last_layer(cat)], cat


class HRNetHeatmapRef(nn.Module):
    """``HRNetHeatmap`` with zero refinement stages (model.py:130-150 / line/model.py:127-147)."""

    def __init__(self, cfg: Dict):
        super().__init__()
        self.model = HRNetRef(cfg)

    def forward(self, x):
        return self.model(x)[0]


def randomize_bn_(module: nn.Module, gen: torch.Generator) -> None:
    """Fresh BN layers are identity (gamma=1, beta=0, mean=0, var=1, hrnet.py:513-517)
    which would hide BN-folding bugs: give every BN non-trivial eval statistics."""
    for m in module.modules():
        if isinstance(m, nn.modules.batchnorm._BatchNorm):
            c = m.num_features
            m.weight.data = 0.6 + 0.8 * torch.rand(c, generator=gen)
            m.bias.data = 0.2 * torch.randn(c, generator=gen)
            m.running_mean.data = 0.2 * torch.randn(c, generator=gen)
            m.running_var.data = 0.5 + torch.rand(c, generator=gen)


def make_model(kind: str = "keypoints", seed: int = 0) -> HRNetHeatmapRef:
    """Seeded random-init oracle network in eval mode (no trained weights ship with
    the reference)."""
    cfg = {"keypoints": KEYPOINT_CFG, "lines": LINE_CFG, "w18": W18_CFG, "w64": W64_CFG, "w48x4": W48X4_CFG}[kind]
    g = torch.Generator().manual_seed(seed)
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    try:
        m = HRNetHeatmapRef(cfg)
    finally:
        torch.random.set_rng_state(state)
    randomize_bn_(m, g)
    return m.eval()


@torch.no_grad()
def predict(model: nn.Module, x: torch.Tensor, size=(540, 960)) -> torch.Tensor:
    """``HRNetMetaModel.predict`` (metamodel.py:127-134): forward, take the last stage
    output, apply the prediction transform."""
    from oracle.decode_ref import keypoint_decode_torch
    model.eval()
    return keypoint_decode_torch(model(x)[-1], size)
