"""HRNetV2 heat-map networks on the B200 kernels, behind the reference's module API.

Mirrors (names, constructor arguments, call signature, state_dict keys):
  * ``HRNetHeatmap``  - src/models/hrnet/model.py:130-150 (keypoints, 58 log-prob maps at
    H/2 x W/2) and src/models/line/model.py:127-147 (lines, 23 prob maps at H/4 x W/4);
  * the layer schedule of ``HighResolutionNet.forward`` - src/models/hrnet/hrnet.py:437-511,
    src/models/line/hrnet.py:185-249, with ``HighResolutionModule.forward`` hrnet.py:222-246.

What runs where: this file only walks the architecture and launches kernels through the
C ABI (ops.py); every conv/BN/ReLU/add is one ``cal_conv2d`` launch (tcgen05 implicit GEMM,
BN folded, residual and ReLU in the epilogue), every multi-resolution sum one
``cal_fuse_combine`` launch.  Activations are fp16 NHWC, channels padded to 64.

Head restructuring (exact in real arithmetic): the 1x1 conv of the head
commutes with the bilinear upsampling, so it is applied per branch at the branch's own
resolution and only the stem part runs at full resolution; the concatenated 784-channel
tensor (hrnet.py:509) is never materialised.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict
from typing import Dict, List, Optional, Tuple

import torch

from . import ops, packing

EXPANSION = {"BASIC": 1, "BOTTLENECK": 4}


class Config(dict):
    """dict with attribute access; supports ``'upscale' in config`` and ``config.upscale``
    like the objects the reference builds from its YAML files (hrnet.py:306-315)."""

    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return Config(v) if isinstance(v, dict) and not isinstance(v, Config) else v


def _stage(nm, nb, bt, blocks, ch):
    return dict(num_modules=nm, num_branches=nb, block_type=bt, num_blocks=blocks, num_channels=ch)


def w48_config(kind: str = "keypoints") -> Config:
    """model_config/hrnet_w48.yaml of the keypoint (hrnet/) or line (line/) model."""
    cfg = dict(stem_width=64, final_conv_kernel=1, pretrain="",
               stage1=_stage(1, 1, "BOTTLENECK", [4], [64]),
               stage2=_stage(1, 2, "BASIC", [4, 4], [48, 96]),
               stage3=_stage(4, 3, "BASIC", [4, 4, 4], [48, 96, 192]),
               stage4=_stage(3, 4, "BASIC", [4, 4, 4, 4], [48, 96, 192, 384]))
    if kind == "keypoints":
        cfg.update(num_classes=58, upscale=2, internal_final_conv=0)
    elif kind == "lines":
        cfg.update(num_classes=23)
    else:
        raise ValueError(kind)
    return Config(cfg)


def w18_config() -> Config:
    """model_config/hrnet_w18.yaml (keypoints)."""
    return Config(dict(num_classes=58, stem_width=64, final_conv_kernel=1, internal_final_conv=0, upscale=2, pretrain="",
                       stage1=_stage(1, 1, "BOTTLENECK", [1], [32]),
                       stage2=_stage(1, 2, "BASIC", [2, 2], [16, 32]),
                       stage3=_stage(1, 3, "BASIC", [2, 2, 2], [16, 32, 64]),
                       stage4=_stage(1, 4, "BASIC", [2, 2, 2, 2], [16, 32, 64, 128])))


def w64_config() -> Config:
    """model_config/hrnet_w64.yaml (keypoints): the w48 schedule with 64/128/256/512 channels."""
    cfg = dict(w48_config("keypoints"))
    cfg.update(stage2=_stage(1, 2, "BASIC", [4, 4], [64, 128]),
               stage3=_stage(4, 3, "BASIC", [4, 4, 4], [64, 128, 256]),
               stage4=_stage(3, 4, "BASIC", [4, 4, 4, 4], [64, 128, 256, 512]))
    return Config(cfg)


def w48x4_config() -> Config:
    """model_config/hrnet_w48x4.yaml (keypoints): w48 with the head at four times the first branch's
    resolution (= the input resolution); the stem output is resampled to it (hrnet.py:494-498)."""
    cfg = dict(w48_config("keypoints"))
    cfg.update(upscale=4)
    return Config(cfg)


def config_from_yaml(path: str) -> Config:
    """One of the reference's ``model_config/*.yaml`` files (flat keys + stage1..stage4 mappings)."""
    import yaml
    with open(path) as f:
        return Config(yaml.safe_load(f))


# ------------------------------------------------------------------------- architecture walk
class _Conv:
    """One conv(+BN) of the reference, identified by its state_dict prefixes."""
    __slots__ = ("conv", "bn", "cin", "cout", "k", "s", "has_bias")

    def __init__(self, conv, bn, cin, cout, k, s, has_bias=False):
        self.conv, self.bn, self.cin, self.cout, self.k, self.s, self.has_bias = conv, bn, cin, cout, k, s, has_bias


def _walk(cfg: Config, kind: str):
    """Enumerates the network as nested python structures of _Conv, in state_dict order."""
    net: Dict[str, object] = {}
    sw = cfg.stem_width
    net["conv1"] = _Conv("conv1", "bn1", 3, sw, 3, 2)
    net["conv2"] = _Conv("conv2", "bn2", sw, sw, 3, 2)

    def block_list(prefix, block_type, cin, planes, n):
        blocks = []
        exp = EXPANSION[block_type]
        for i in range(n):
            p = f"{prefix}.{i}"
            ci = cin if i == 0 else planes * exp
            if block_type == "BOTTLENECK":
                convs = [_Conv(f"{p}.conv1", f"{p}.bn1", ci, planes, 1, 1),
                         _Conv(f"{p}.conv2", f"{p}.bn2", planes, planes, 3, 1),
                         _Conv(f"{p}.conv3", f"{p}.bn3", planes, planes * 4, 1, 1)]
            else:
                convs = [_Conv(f"{p}.conv1", f"{p}.bn1", ci, planes, 3, 1),
                         _Conv(f"{p}.conv2", f"{p}.bn2", planes, planes, 3, 1)]
            ds = None
            if i == 0 and ci != planes * exp:
                ds = _Conv(f"{p}.downsample.0", f"{p}.downsample.1", ci, planes * exp, 1, 1)
            blocks.append((convs, ds))
        return blocks

    s1 = cfg.stage1
    net["layer1"] = block_list("layer1", s1.block_type, 64, s1.num_channels[0], s1.num_blocks[0])
    pre = [EXPANSION[s1.block_type] * s1.num_channels[0]]
    for idx in (2, 3, 4):
        sc = cfg[f"stage{idx}"]
        exp = EXPANSION[sc["block_type"]]
        ch = [c * exp for c in sc["num_channels"]]
        # transition (hrnet.py:357-391)
        trans = []
        for i, c in enumerate(ch):
            tp = f"transition{idx - 1}.{i}"
            if i < len(pre):
                trans.append(None if c == pre[i] else [_Conv(f"{tp}.0", f"{tp}.1", pre[i], c, 3, 1)])
            else:
                chain = []
                for j in range(i + 1 - len(pre)):
                    co = c if j == i - len(pre) else pre[-1]
                    chain.append(_Conv(f"{tp}.{j}.0", f"{tp}.{j}.1", pre[-1], co, 3, 2))
                trans.append(chain)
        net[f"transition{idx - 1}"] = trans
        modules = []
        for m in range(sc["num_modules"]):
            mp = f"stage{idx}.{m}"
            branches = [block_list(f"{mp}.branches.{b}", sc["block_type"], ch[b], ch[b] // exp, sc["num_blocks"][b])
                        for b in range(len(ch))]
            fuse = []
            for i in range(len(ch)):
                row = []
                for j in range(len(ch)):
                    fp = f"{mp}.fuse_layers.{i}.{j}"
                    if j == i:
                        row.append(None)
                    elif j > i:
                        row.append([_Conv(f"{fp}.0", f"{fp}.1", ch[j], ch[i], 1, 1)])
                    else:
                        chain = []
                        for k in range(i - j):
                            co = ch[i] if k == i - j - 1 else ch[j]
                            chain.append(_Conv(f"{fp}.{k}.0", f"{fp}.{k}.1", ch[j], co, 3, 2))
                        row.append(chain)
                fuse.append(row)
            modules.append((branches, fuse))
        net[f"stage{idx}"] = modules
        pre = ch
    upscale = cfg["upscale"] if ("upscale" in cfg and cfg["upscale"] > 1) else 1
    if kind == "lines":
        upscale = 1
    last_in = sum(pre) + (sw if upscale > 1 else 0)
    net["last_in"] = last_in
    net["upscale"] = upscale
    net["branch_channels"] = pre
    net["head1"] = _Conv("last_layer.0", "last_layer.1", last_in, last_in, 1, 1, has_bias=True)
    net["head2"] = _Conv("last_layer.3", None, last_in, cfg.num_classes, cfg.final_conv_kernel, 1, has_bias=True)
    return net


def _all_convs(net) -> List[_Conv]:
    out: List[_Conv] = [net["conv1"], net["conv2"]]

    def blocks(bl):
        for convs, ds in bl:
            out.extend(convs)
            if ds is not None:
                out.append(ds)

    blocks(net["layer1"])
    for idx in (2, 3, 4):
        for t in net[f"transition{idx - 1}"]:
            if t:
                out.extend(t)
        for branches, fuse in net[f"stage{idx}"]:
            for b in branches:
                blocks(b)
            for row in fuse:
                for cell in row:
                    if cell:
                        out.extend(cell)
    out.extend([net["head1"], net["head2"]])
    return out


def conv_shapes(kind: str, H: int, W: int, cfg: Optional[Config] = None):
    """[(conv, Hout, Wout)] of every convolution of the reference's forward at an H x W input
    (hrnet.py:437-511 / line/hrnet.py:185-249), AS WRITTEN there (head 1x1 at full head
    resolution over the concatenated channels) - the algorithmic work list."""
    cfg = w48_config(kind) if cfg is None else cfg
    net = _walk(cfg, kind)
    dn = lambda v: (v - 1) // 2 + 1
    h1, w1 = dn(H), dn(W)
    h2, w2 = dn(h1), dn(w1)
    out = [(net["conv1"], h1, w1), (net["conv2"], h2, w2)]
    for convs, ds in net["layer1"]:
        out += [(c, h2, w2) for c in convs] + ([(ds, h2, w2)] if ds is not None else [])
    sizes = [(h2, w2)]
    for idx in (2, 3, 4):
        trans = net[f"transition{idx - 1}"]
        while len(sizes) < len(trans):
            sizes.append((dn(sizes[-1][0]), dn(sizes[-1][1])))
        for i, tr in enumerate(trans):
            if tr:
                out += [(c, *sizes[i]) for c in tr]
        for branches, fuse in net[f"stage{idx}"]:
            for b, bl in enumerate(branches):
                for convs, ds in bl:
                    out += [(c, *sizes[b]) for c in convs]
            for i, row in enumerate(fuse):
                for j, cell in enumerate(row):
                    if not cell:
                        continue
                    if j > i:
                        out.append((cell[0], *sizes[j]))
                    else:
                        for k, c in enumerate(cell):
                            out.append((c, *sizes[j + k + 1]))
    hh, wh = sizes[0][0] * net["upscale"], sizes[0][1] * net["upscale"]
    out += [(net["head1"], hh, wh), (net["head2"], hh, wh)]
    return out


def conv_gflop_per_frame(kind: str, H: int = 540, W: int = 960) -> float:
    """2*MAC of the reference's convolutions per frame, in GFLOP (507.82 for the
    keypoint net, 371.4 for the line net at 960x540).  roofline.achieved is computed from this."""
    return sum(2.0 * c.cin * c.cout * c.k * c.k * h * w for c, h, w in conv_shapes(kind, H, W)) / 1e9


def state_dict_schema(cfg: Config, kind: str, prefix: str = "model.") -> "OrderedDict[str, Tuple[int, ...]]":
    """Keys and shapes of the reference's ``nn_state_dict`` for this architecture
    (pinned by tests/golden/hrnet_state_keys.json)."""
    net = _walk(cfg, kind)
    sd: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def add(c: _Conv):
        sd[f"{prefix}{c.conv}.weight"] = (c.cout, c.cin, c.k, c.k)
        if c.has_bias:
            sd[f"{prefix}{c.conv}.bias"] = (c.cout,)
        if c.bn:
            for n in ("weight", "bias", "running_mean", "running_var"):
                sd[f"{prefix}{c.bn}.{n}"] = (c.cout,)
            sd[f"{prefix}{c.bn}.num_batches_tracked"] = ()
    # state_dict order: convN before bnN for blocks (conv1,bn1,conv2,bn2,...), which the
    # generic add() reproduces; the stem is conv1,bn1,conv2,bn2 as well.
    for c in _all_convs(net):
        add(c)
    return sd


def init_state_dict(cfg: Config, kind: str, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """Seeded random weights of the right shapes (no trained weights ship with the
    reference): He-style conv weights, non-trivial BN statistics."""
    g = torch.Generator().manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for k, shape in state_dict_schema(cfg, kind).items():
        if k.endswith("num_batches_tracked"):
            sd[k] = torch.zeros((), dtype=torch.long)
        elif len(shape) == 4:
            fan = shape[1] * shape[2] * shape[3]
            sd[k] = torch.randn(shape, generator=g) * math.sqrt(1.0 / fan)
        elif k.endswith("running_var"):
            sd[k] = 0.5 + torch.rand(shape, generator=g)
        elif k.endswith("running_mean") or ".bias" in k:
            sd[k] = 0.1 * torch.randn(shape, generator=g)
        else:  # BN weight
            sd[k] = 0.8 + 0.4 * torch.rand(shape, generator=g)
    return sd


# ------------------------------------------------------------------------------- the engine
class _Packed:
    __slots__ = ("w", "b", "rows", "cin_pad", "cout_pad", "k", "s", "cin", "slices", "w_k")


class HRNetHeatmap:
    """Drop-in for ``HRNetHeatmap`` (model.py:130-150 / line/model.py:127-147).

    ``hrnet_config`` is the reference's model config (mapping or attribute object);
    ``num_refinement_stages`` must be 0 (every shipped config, train_config.yaml:34).
    ``forward(x)`` takes (B,3,H,W) fp32 in [0,1] and returns a list with one fp32 tensor
    (B,num_classes,H/2,W/2) log-probs (keypoints) or (B,num_classes,H/4,W/4) probs (lines).
    """

    def __init__(self, hrnet_config, num_refinement_stages: int = 0, num_heatmaps: int = 38,
                 kind: Optional[str] = None):
        if num_refinement_stages != 0:
            raise NotImplementedError("refinement stages are disabled in every shipped config")
        cfg = hrnet_config if isinstance(hrnet_config, Config) else Config(
            {k: hrnet_config[k] for k in hrnet_config} if hasattr(hrnet_config, "keys") else vars(hrnet_config))
        if kind is None:
            kind = "keypoints" if ("upscale" in cfg and cfg["upscale"] > 1) else "lines"
        self.kind = kind
        self.cfg = cfg
        self.net = _walk(cfg, kind)
        self.num_classes = int(cfg["num_classes"])
        self._sd: Optional[Dict[str, torch.Tensor]] = None
        self._packed: Dict[str, _Packed] = {}
        self.device: Optional[torch.device] = None
        self.training = False
        self.fused_head = os.environ.get("CAL_FUSED_HEAD", "1") != "0"
        self.chained_head = os.environ.get("CAL_HEAD_CHAIN", "1") != "0"
        self.slice_major = os.environ.get("CAL_W_SLICES", "1") != "0"
        self.fuse_blocks = os.environ.get("CAL_BASICBLOCK", "1") != "0"     # cal_basicblock for the C <= 48 BasicBlocks
        self.branch_streams = os.environ.get("CAL_BRANCH_STREAMS", "0") != "0"   # measured neutral: off
        self._streams: List[torch.cuda.Stream] = []
        # the forward through the library's own engine (csrc/engine.cu: cal_hrnet_create / cal_hrnet_forward),
        # one C call per network instead of ~300 ctypes calls; CAL_ENGINE=0 walks the schedule here, op by op
        # (the readable twin: same kernels, same order, bit-identical output - and the per-launch profile)
        self.use_engine = os.environ.get("CAL_ENGINE", "1") != "0"
        self._engine = None

    # -- nn.Module-like surface used by the reference's callers
    def eval(self):
        self.training = False
        return self

    def state_dict(self):
        if self._sd is None:
            self.load_state_dict(init_state_dict(self.cfg, self.kind))
        return self._sd

    def load_state_dict(self, sd, strict: bool = True):
        self._engine_release()
        schema = state_dict_schema(self.cfg, self.kind)
        sd = {k.replace("_orig_mod.", ""): v for k, v in sd.items()}  # metamodel.py:113-117
        missing = [k for k in schema if k not in sd]
        if missing and strict:
            raise KeyError(f"missing {len(missing)} keys, e.g. {missing[:3]}")
        for k, shape in schema.items():
            if k in sd and tuple(sd[k].shape) != tuple(shape):
                raise ValueError(f"{k}: shape {tuple(sd[k].shape)} != {shape}")
        self._sd = OrderedDict((k, sd[k].detach().clone().cpu()) for k in schema if k in sd)
        self._packed = {}
        return self

    def to(self, device):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("HRNetHeatmap runs on CUDA (sm_100a) only; there is no CPU fallback")
        self.device = device
        self._pack()
        self._engine_release()
        return self

    # -- the C-ABI engine
    def engine_config(self) -> "ops._lib.HrnetConfig":
        """The CalHrnetConfig of this architecture (include/calib_b200.h)."""
        c = ops._lib.HrnetConfig()
        c.kind = 0 if self.kind == "keypoints" else 1
        c.num_classes, c.stem_width = self.num_classes, int(self.cfg["stem_width"])
        c.upscale = int(self.net["upscale"])
        for i in range(4):
            st = self.cfg[f"stage{i + 1}"]
            c.stage[i].num_modules, c.stage[i].num_branches = int(st["num_modules"]), int(st["num_branches"])
            c.stage[i].block_type = 1 if st["block_type"] == "BOTTLENECK" else 0
            for b in range(int(st["num_branches"])):
                c.stage[i].num_blocks[b], c.stage[i].num_channels[b] = int(st["num_blocks"][b]), int(st["num_channels"][b])
        return c

    def weight_blob(self) -> torch.Tensor:
        """Every float tensor of the state_dict in its own order, flattened and concatenated
        (num_batches_tracked skipped): cal_hrnet_create's input."""
        sd = self.state_dict()
        return torch.cat([sd[k].detach().reshape(-1).float() for k in state_dict_schema(self.cfg, self.kind)
                          if not k.endswith("num_batches_tracked")]).contiguous()

    def _engine_handle(self):
        if self._engine is None:
            import ctypes as C
            blob = self.weight_blob()
            h = C.c_void_p()
            cfg = self.engine_config()
            with torch.cuda.device(self.device):
                st = ops._lib.lib().cal_hrnet_create(C.byref(cfg), blob.data_ptr(), blob.numel(), C.byref(h))
            ops._lib.check(st, "cal_hrnet_create")
            self._engine = h
            self._engine_launches = 0
        return self._engine

    def _engine_release(self):
        if self._engine is not None:
            ops._lib.lib().cal_hrnet_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self._engine_release()
        except Exception:       # noqa: BLE001 - interpreter shutdown
            pass

    def _engine_forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        import ctypes as C
        h = self._engine_handle()
        u8 = x.dtype == torch.uint8
        if u8:
            B, H, W, _ = x.shape
        else:
            B, _, H, W = x.shape
        L = ops._lib.lib()
        nc, oh, ow = C.c_int32(), C.c_int32(), C.c_int32()
        ops._lib.check(L.cal_hrnet_output_shape(h, H, W, C.byref(nc), C.byref(oh), C.byref(ow)), "cal_hrnet_output_shape")
        heat = torch.empty((B, nc.value, oh.value, ow.value), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            st = L.cal_hrnet_forward(h, x.data_ptr(), int(u8), B, H, W, heat.data_ptr(), torch.cuda.current_stream().cuda_stream)
        ops._lib.check(st, "cal_hrnet_forward")
        n = int(L.cal_hrnet_launches(h))
        ops.LAUNCHES += n - self._engine_launches
        self._engine_launches = n
        return [heat]

    def cuda(self, idx: int = 0):
        return self.to(f"cuda:{idx}")

    # -- weight preparation
    def _bn(self, name):
        sd = self._sd
        return {k: sd[f"model.{name}.{k}"] for k in ("weight", "bias", "running_mean", "running_var")}

    def _fold(self, c: _Conv):
        sd = self._sd
        w = sd[f"model.{c.conv}.weight"]
        b = sd.get(f"model.{c.conv}.bias") if c.has_bias else None
        return packing.fold_bn(w, b, self._bn(c.bn) if c.bn else None)

    def _put(self, key, w, b, k, s, cout_pad=None):
        wp, bp, rows = packing.pack_conv(w, b, cout_pad=cout_pad)
        p = _Packed()
        p.w, p.b, p.rows = wp.to(self.device), bp.to(self.device), rows
        p.cin_pad, p.cout_pad, p.k, p.s = packing.pad_to(w.shape[1]), bp.numel(), k, s
        p.cin = int(w.shape[1])
        # 3x3 stride-1 layers: slice-major weights (every (tap, 64-channel chunk) slice contiguous)
        p.slices = bool(self.slice_major and k == 3)
        p.w_k = p.w                       # K-major copy for the generic kernel (fallback shapes)
        if p.slices:
            p.w = p.w.reshape(p.rows, k * k * p.cin_pad // 64, 64).permute(1, 0, 2).contiguous()
        self._packed[key] = p

    def _pack(self):
        if self._sd is None:
            self.load_state_dict(init_state_dict(self.cfg, self.kind))
        net = self.net
        for c in _all_convs(net):
            if c is net["conv1"] or c is net["head1"] or c is net["head2"]:
                continue
            w, b = self._fold(c)
            self._put(c.conv, w, b, c.k, c.s)
        # stem conv1: fp32 CUDA-core kernel, (64, 27) [co][ci*9+ky*3+kx]
        w, b = self._fold(net["conv1"])
        self.stem_w = w.reshape(w.shape[0], 27).float().contiguous().to(self.device)
        self.stem_b = b.float().contiguous().to(self.device)
        # head: split the first 1x1 conv by source (hrnet.py:509 concat order)
        w1, b1 = self._fold(net["head1"])
        cpad = packing.pad_to(net["last_in"])
        srcs = ([("stem", self.cfg.stem_width)] if net["upscale"] > 1 else []) + \
               [(f"b{i}", c) for i, c in enumerate(net["branch_channels"])]
        off = 0
        zero = torch.zeros(w1.shape[0], dtype=torch.float64)
        for name, c in srcs:
            self._put(f"head1.{name}", w1[:, off:off + c], zero, 1, 1, cout_pad=cpad)
            off += c
        b1p = torch.zeros(cpad, dtype=torch.float32)
        b1p[:w1.shape[0]] = b1.float()
        self.head_b1 = b1p.to(self.device)
        w2, b2 = self._fold(net["head2"])
        if w2.shape[-1] != 1:
            raise NotImplementedError("final_conv_kernel must be 1 (every shipped config)")
        self._put("head2", w2, b2, 1, 1, cout_pad=64)
        # the chained head kernel wants the final conv as a full 64-row K-major tile
        p2 = self._packed["head2"]
        w64 = torch.zeros((64, p2.w.shape[1]), dtype=torch.float16, device=self.device)
        w64[:p2.rows] = p2.w
        self.head2_w64 = w64

    # -- launch helpers
    def _conv(self, x, key, relu, res=None):
        p = self._packed[key]
        B, H, W, _ = x.shape
        pad = p.k // 2
        Ho, Wo = (H + 2 * pad - p.k) // p.s + 1, (W + 2 * pad - p.k) // p.s + 1
        y = torch.empty((B, Ho, Wo, p.cout_pad), dtype=torch.float16, device=x.device)
        if p.slices:
            try:
                return ops.conv2d(x, p.w, p.b, y, ksize=p.k, stride=p.s, cout_rows=p.rows, relu=relu, res=res,
                                  cin=p.cin, w_slices=True)
            except ops._lib.CalError as e:
                if "generic kernel" not in str(e):
                    raise
                p.slices = False            # this shape is served by the generic kernel: K-major from now on
                p.w = p.w_k
        return ops.conv2d(x, p.w, p.b, y, ksize=p.k, stride=p.s, cout_rows=p.rows, relu=relu, res=res, cin=p.cin)

    def _basicblock(self, x, convs):
        """Both convs of a BasicBlock of the full-resolution branch in one kernel (csrc/basicblock.cu); None when
        the shape is served by the two launches."""
        if not self.fuse_blocks or len(convs) != 2:
            return None
        p1, p2 = self._packed[convs[0].conv], self._packed[convs[1].conv]
        ok = all(p.slices and p.k == 3 and p.s == 1 and p.cin_pad == 64 and p.cout_pad == 64 and p.rows <= 48 for p in (p1, p2))
        if not ok or x.shape[-1] != 64 or p1.rows != p2.rows or p1.cin != p2.cin:
            return None
        try:
            return ops.basicblock(x, p1.w, p1.b, p2.w, p2.b, torch.empty_like(x), rows=p1.rows, c=p1.cin)
        except ops._lib.CalError as e:
            if "status -2" not in str(e):
                raise
            return None

    def _blocks(self, x, blocks):
        for convs, ds in blocks:
            if ds is None:
                y = self._basicblock(x, convs)
                if y is not None:
                    x = y
                    continue
            r = x if ds is None else self._conv(x, ds.conv, relu=False)
            t = x
            for c in convs[:-1]:
                t = self._conv(t, c.conv, relu=True)
            x = self._conv(t, convs[-1].conv, relu=True, res=r)
        return x

    def _module(self, xs, module):
        branches, fuse = module
        if self.branch_streams and len(xs) > 1:
            # the branches of a module are independent until the fuse: one stream each, so the
            # ramp-up / tail of one branch's (persistent) kernels is filled by another's
            main = torch.cuda.current_stream(self.device)
            while len(self._streams) < len(xs) - 1:
                self._streams.append(torch.cuda.Stream(device=self.device))
            outs_b = [None] * len(xs)
            for i in range(1, len(xs)):
                st = self._streams[i - 1]
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    outs_b[i] = self._blocks(xs[i], branches[i])
            outs_b[0] = self._blocks(xs[0], branches[0])
            for i in range(1, len(xs)):
                main.wait_stream(self._streams[i - 1])
                outs_b[i].record_stream(main)
            xs = outs_b
        else:
            xs = [self._blocks(x, b) for x, b in zip(xs, branches)]
        nb = len(xs)
        outs = []
        for i in range(nb):
            # same-resolution terms chained through conv epilogues: x_i + sum_{j<i} down_ij(x_j)
            has_up = i < nb - 1
            r = xs[i]
            for j in range(i):
                t = xs[j]
                chain = fuse[i][j]
                for c in chain[:-1]:
                    t = self._conv(t, c.conv, relu=True)
                last_term = (j == i - 1)
                r = self._conv(t, chain[-1].conv, relu=(last_term and not has_up), res=r)
            if has_up:
                ups = [self._conv(xs[j], fuse[i][j][0].conv, relu=False) for j in range(i + 1, nb)]
                y = torch.empty_like(xs[i])
                r = ops.fuse_combine(y, [r] + ups, None, relu=True, c=fuse[i][i + 1][0].cout)
            outs.append(r)
        return outs

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        if self.device is None:
            raise RuntimeError("call .to('cuda:N') first")
        if x.device != self.device:
            x = x.to(self.device, non_blocking=True)
        # (B,3,H,W) float in [0,1] (the reference's input), or (B,H,W,3) uint8 BGR frames as cv2.imread leaves
        # them: ToTensor's /255 then happens inside the stem kernel
        if x.dtype == torch.uint8:
            if x.dim() != 4 or x.shape[-1] != 3:
                raise ValueError("uint8 frames must be (B,H,W,3) BGR")
            x = x.contiguous()
            B, H, W, _ = x.shape
        else:
            x = x.contiguous().float()
            B, _, H, W = x.shape
        if self.use_engine and ops.PROFILE is None and not self.branch_streams and self.fused_head and self.chained_head \
                and self.slice_major:
            return self._engine_forward(x)
        net = self.net
        with torch.cuda.device(self.device):
            Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            stem = torch.empty((B, Ho, Wo, 64), dtype=torch.float16, device=self.device)
            ops.stem_conv(x, self.stem_w, self.stem_b, stem)
            t = self._conv(stem, "conv2", relu=True)
            t = self._blocks(t, net["layer1"])
            ys = [t]
            for idx in (2, 3, 4):
                trans = net[f"transition{idx - 1}"]
                nprev = len(ys)
                xs = []
                for i, tr in enumerate(trans):
                    if tr is None:
                        xs.append(ys[i])
                    else:
                        v = ys[i] if i < nprev else ys[-1]
                        for c in tr:
                            v = self._conv(v, c.conv, relu=True)
                        xs.append(v)
                for module in net[f"stage{idx}"]:
                    xs = self._module(xs, module)
                ys = xs
            return [self._head(stem, ys)]

    __call__ = forward

    def _head(self, stem, ys):
        net = self.net
        up = net["upscale"]
        h, w = ys[0].shape[1] * up, ys[0].shape[2] * up
        B = ys[0].shape[0]
        cpad = self.head_b1.numel()
        if up > 1:
            full, full_key, low = stem, "head1.stem", list(enumerate(ys))
            if stem.shape[1] != h or stem.shape[2] != w:
                # odd input sizes: the reference resamples x_stem to the head size (hrnet.py:495-498)
                rs = torch.empty((B, h, w, stem.shape[3]), dtype=torch.float16, device=self.device)
                full = ops.fuse_combine(rs, [stem], None, relu=False)
        else:
            full, full_key, low = ys[0], "head1.b0", list(enumerate(ys))[1:]
        proj = [self._conv(y, f"head1.b{i}", relu=False) for i, y in low]
        z = None
        pf = self._packed[full_key]
        p2 = self._packed["head2"]
        if self.fused_head and self.chained_head and full.shape[3] == 64:
            # the whole head in one kernel: first conv + interpolation GEMMs + ReLU, then the final
            # 1x1 conv and (Log)Softmax on the tile while it is still on chip
            heat = torch.empty((B, self.num_classes, h, w), dtype=torch.float32, device=self.device)
            out = ops.head_fused(full, pf.w, proj, self.head_b1, None, pf.rows, w2=self.head2_w64, bias2=p2.b, heat=heat,
                                 mode=1 if self.kind == "keypoints" else 2)
            if out is not None:
                return out
        if self.fused_head and full.shape[3] == 64:
            # one kernel: W1_full * full + interpolation GEMMs over the projections + bias + ReLU
            z = torch.empty((B, h, w, cpad), dtype=torch.float16, device=self.device)
            z = ops.head_fused(full, pf.w, proj, self.head_b1, z, pf.rows)
        if z is None:
            u = torch.empty((B, h, w, cpad), dtype=torch.float16, device=self.device)
            ops.fuse_combine(u, proj, self.head_b1, relu=False)
            z = self._conv(full, full_key, relu=True, res=u)
            del u
        del proj
        heat = torch.empty((B, self.num_classes, h, w), dtype=torch.float32, device=self.device)
        ops.conv2d(z, p2.w, p2.b, heat, ksize=1, stride=1, cout_rows=p2.rows, relu=False,
                   mode=1 if self.kind == "keypoints" else 2, n_classes=self.num_classes, cin=p2.cin)
        return heat
