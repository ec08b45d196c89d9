"""fp32 PyTorch restatement of the localization model.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

``CoordinateRegressor`` (``spine_vision/training/models/generic.py:286-391``)
is a timm ``convnext_base`` created with ``num_classes=0``
(``training/models/backbone.py:165-172``) followed by the inline head
``LayerNorm -> Dropout -> Linear(1024,256) -> GELU -> Dropout -> Linear(256,10)
-> Sigmoid`` (``generic.py:343-351``) and ``view(-1, 5, 2)``
(``generic.py:389-391``).

timm (1.0.22 in the reference's ``uv.lock``) is not installed in this image, so
the backbone is restated here from its published architecture with timm's
state-dict key names, so that a ``LocalizationTrainer`` checkpoint
(``training/trainers/base.py:695-706``) loads with ``strict=True``:

    backbone.stem.0 (Conv2d 3->C0 k4 s4)      backbone.stem.1 (LayerNorm2d)
    backbone.stages.S.downsample.0 (LayerNorm2d)  .downsample.1 (Conv2d k2 s2)   S=1..3
    backbone.stages.S.blocks.J.{gamma, conv_dw, norm, mlp.fc1, mlp.fc2}
    backbone.head.norm (LayerNorm2d after global average pool)
    head.{0,2,5}

ConvNeXt block: ``x + gamma * fc2(GELU_erf(fc1(LN_eps1e-6(dwconv7x7(x)))))``.
The arithmetic is checked for structural equivalence against
``torchvision.models.convnext_base`` in ``tests/test_oracle.py``.
"""

from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

VARIANTS = {
    "tiny": ((3, 3, 9, 3), (96, 192, 384, 768)),
    "small": ((3, 3, 27, 3), (96, 192, 384, 768)),
    "base": ((3, 3, 27, 3), (128, 256, 512, 1024)),
    "large": ((3, 3, 27, 3), (192, 384, 768, 1536)),
    "xlarge": ((3, 3, 27, 3), (256, 512, 1024, 2048)),
    # timm convnextv2_*: same macro design, GlobalResponseNorm inside the MLP, no layer scale
    "v2_tiny": ((3, 3, 9, 3), (96, 192, 384, 768)),
    "v2_small": ((3, 3, 27, 3), (96, 192, 384, 768)),
    "v2_base": ((3, 3, 27, 3), (128, 256, 512, 1024)),
    "v2_large": ((3, 3, 27, 3), (192, 384, 768, 1536)),
    "v2_huge": ((3, 3, 27, 3), (352, 704, 1408, 2816)),
}
LN_EPS = 1e-6


class LayerNorm2d(nn.LayerNorm):
    """LayerNorm over the channel dim of an NCHW tensor (timm ``LayerNorm2d``)."""

    def forward(self, x):  # type: ignore[override]
        x = x.permute(0, 2, 3, 1)
        x = F.layer_norm(x, self.normalized_shape, self.weight, self.bias, self.eps)
        return x.permute(0, 3, 1, 2)


class GlobalResponseNorm(nn.Module):
    """timm ``GlobalResponseNorm`` (ConvNeXt-V2), channels-last ``[B, H, W, C]``: L2 norm over the image's tokens per channel,
    divided by its mean over channels; ``x + bias + weight * (x * n)``.  Zero-initialised weight / bias (identity at init)."""

    def __init__(self, dim: int, eps: float = 1e-6):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.zeros(dim))
        self.bias = nn.Parameter(torch.zeros(dim))

    def forward(self, x):
        x_g = x.norm(p=2, dim=(1, 2), keepdim=True)
        x_n = x_g / (x_g.mean(dim=-1, keepdim=True) + self.eps)
        return x + torch.addcmul(self.bias.view(1, 1, 1, -1), self.weight.view(1, 1, 1, -1), x * x_n)


class Mlp(nn.Module):
    def __init__(self, dim: int, use_grn: bool = False):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        if use_grn:
            self.grn = GlobalResponseNorm(4 * dim)
        self.use_grn = use_grn
        self.fc2 = nn.Linear(4 * dim, dim)

    def forward(self, x):
        x = self.act(self.fc1(x))
        if self.use_grn:
            x = self.grn(x)
        return self.fc2(x)


class Block(nn.Module):
    def __init__(self, dim: int, ls_init: float | None = 1e-6, use_grn: bool = False):
        super().__init__()
        self.conv_dw = nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim)
        self.norm = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = Mlp(dim, use_grn)
        self.gamma = nn.Parameter(ls_init * torch.ones(dim)) if ls_init is not None else None

    def forward(self, x):
        shortcut = x
        x = self.conv_dw(x).permute(0, 2, 3, 1)
        x = self.mlp(self.norm(x)).permute(0, 3, 1, 2)
        if self.gamma is not None:
            x = x * self.gamma.reshape(1, -1, 1, 1)
        return shortcut + x


class Stage(nn.Module):
    def __init__(self, cin: int, cout: int, depth: int, first: bool, v2: bool = False):
        super().__init__()
        if first:
            self.downsample = nn.Identity()
        else:
            self.downsample = nn.Sequential(
                LayerNorm2d(cin, eps=LN_EPS), nn.Conv2d(cin, cout, kernel_size=2, stride=2)
            )
        self.blocks = nn.Sequential(*[Block(cout, None if v2 else 1e-6, use_grn=v2) for _ in range(depth)])

    def forward(self, x):
        return self.blocks(self.downsample(x))


class Head(nn.Module):
    """timm ``NormMlpClassifierHead`` with ``num_classes=0``: pool -> LayerNorm2d -> flatten."""

    def __init__(self, dim: int):
        super().__init__()
        self.norm = LayerNorm2d(dim, eps=LN_EPS)

    def forward(self, x):
        x = x.mean(dim=(2, 3), keepdim=True)
        return self.norm(x).flatten(1)


class ConvNeXt(nn.Module):
    def __init__(self, variant: str = "base"):
        super().__init__()
        depths, dims = VARIANTS[variant]
        self.num_features = dims[-1]
        self.stem = nn.Sequential(
            nn.Conv2d(3, dims[0], kernel_size=4, stride=4), LayerNorm2d(dims[0], eps=LN_EPS)
        )
        stages = []
        prev = dims[0]
        for i, (d, c) in enumerate(zip(depths, dims)):
            stages.append(Stage(prev, c, d, first=(i == 0), v2=variant.startswith("v2_")))
            prev = c
        self.stages = nn.Sequential(*stages)
        self.head = Head(prev)
        self.apply(self._init)

    @staticmethod
    def _init(m):
        # timm convnext._init_weights: trunc_normal_(std=.02) on conv/linear, zero bias
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    def forward(self, x):
        return self.head(self.stages(self.stem(x)))


class CoordinateRegressor(nn.Module):
    """generic.py:299-391, inference-relevant part only."""

    def __init__(self, variant: str = "base", num_levels: int = 5, num_outputs: int = 2, dropout: float = 0.2):
        super().__init__()
        self._num_levels = num_levels
        self._num_outputs = num_outputs
        self.backbone = ConvNeXt(variant)
        fd = self.backbone.num_features
        self.head = nn.Sequential(
            nn.LayerNorm(fd),
            nn.Dropout(dropout),
            nn.Linear(fd, 256),
            nn.GELU(),
            nn.Dropout(dropout / 2),
            nn.Linear(256, num_levels * num_outputs),
            nn.Sigmoid(),
        )

    def forward(self, x):
        return self.head(self.backbone(x)).view(-1, self._num_levels, self._num_outputs)


def make_model(variant: str = "base", seed: int = 0, trained_like: bool = False) -> CoordinateRegressor:
    """Random-init model (``torch.manual_seed(seed)``); ``trained_like`` replaces
    the near-identity init (layer-scale 1e-6) with gamma ~ U(0.1, 1), non-trivial
    LayerNorm affines and biases, so that reduced-precision error is visible."""
    torch.manual_seed(seed)
    m = CoordinateRegressor(variant).eval()
    if trained_like:
        g = torch.Generator().manual_seed(seed + 1000)
        with torch.no_grad():
            for name, p in m.named_parameters():
                if name.endswith("gamma"):
                    p.copy_(torch.rand(p.shape, generator=g) * 0.9 + 0.1)
                elif ".grn." in name:  # zero at init (identity): give the response norm something to do
                    p.copy_((0.5 if name.endswith("weight") else 0.05) * torch.randn(p.shape, generator=g))
                elif ".norm" in name or "stem.1" in name or "downsample.0" in name or name.startswith("head.0"):
                    if name.endswith("weight"):
                        p.copy_(1.0 + 0.2 * torch.randn(p.shape, generator=g))
                    else:
                        p.copy_(0.1 * torch.randn(p.shape, generator=g))
                elif name.endswith("bias"):
                    p.copy_(0.02 * torch.randn(p.shape, generator=g))
                elif name.endswith("weight") and p.dim() >= 2 and "conv_dw" not in name:
                    # widen the pointwise weights so block outputs are O(1)
                    p.mul_(2.0)
    return m
