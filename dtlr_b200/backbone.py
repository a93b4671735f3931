"""ResNet-50 trunk + FrozenBatchNorm + sine position encoding: the module (parameter-holding) side of the backbone.

State-dict layout identical to the reference: backbone.0.body.{conv1,bn1,layer{1-4}.{i}.{conv1-3,
bn1-3,downsample.{0,1}}}.*  -- reference models/dino/backbone.py:36-144 wraps torchvision's resnet50; the topology is
restated here (v1.5: stride on the 3x3) so the package does not depend on torchvision.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .misc import NestedTensor


class FrozenBatchNorm2d(nn.Module):
    """reference backbone.py:36-72: affine with fixed statistics, eps=1e-5 inside the rsqrt."""

    def __init__(self, n):
        super().__init__()
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)

    def scale_bias(self):
        scale = self.weight * (self.running_var + 1e-5).rsqrt()
        return scale, self.bias - self.running_mean * scale

    def forward(self, x):
        scale, bias = self.scale_bias()
        return x * scale.reshape(1, -1, 1, 1) + bias.reshape(1, -1, 1, 1)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = FrozenBatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride=stride, padding=1, bias=False)
        self.bn2 = FrozenBatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
        self.bn3 = FrozenBatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        idt = x
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        if self.downsample is not None:
            idt = self.downsample(x)
        return F.relu(y + idt)


class ResNet50Body(nn.Module):
    """conv1/bn1/maxpool/layer1-4; returns the feature maps named by return_interm_indices (1,2,3 -> layer2,3,4)."""

    def __init__(self, return_interm_indices=(1, 2, 3)):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 7, stride=2, padding=3, bias=False)
        self.bn1 = FrozenBatchNorm2d(64)
        inplanes = 64
        for li, (planes, blocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
            stride = 1 if li == 1 else 2
            layers = []
            for b in range(blocks):
                ds = None
                if b == 0:
                    ds = nn.Sequential(nn.Conv2d(inplanes, planes * 4, 1, stride=stride, bias=False),
                                       FrozenBatchNorm2d(planes * 4))
                layers.append(Bottleneck(inplanes, planes, stride if b == 0 else 1, ds))
                inplanes = planes * 4
            setattr(self, "layer%d" % li, nn.Sequential(*layers))
        self.return_layers = [i + 1 for i in return_interm_indices]
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")

    def forward(self, x):
        y = F.relu(self.bn1(self.conv1(x)))
        y = F.max_pool2d(y, kernel_size=3, stride=2, padding=1)
        outs = []
        for li in range(1, 5):
            y = getattr(self, "layer%d" % li)(y)
            if li in self.return_layers:
                outs.append(y)
        return outs


class Backbone(nn.Module):
    """reference backbone.py:75-128 (BackboneBase + Backbone), resnet50 only."""

    def __init__(self, name, train_backbone, dilation, return_interm_indices):
        super().__init__()
        if name != "resnet50":
            raise NotImplementedError("dtlr_b200 implements the resnet50 backbone only (every shipped config uses it)")
        if dilation:
            raise NotImplementedError("dilation=True is not used by any shipped config")
        assert list(return_interm_indices) in [[0, 1, 2, 3], [1, 2, 3], [3]]
        self.body = ResNet50Body(tuple(return_interm_indices))
        for pname, p in self.body.named_parameters():
            if not train_backbone or ("layer2" not in pname and "layer3" not in pname and "layer4" not in pname):
                p.requires_grad_(False)
        self.num_channels = [256, 512, 1024, 2048][4 - len(return_interm_indices):]

    def forward(self, tensor_list: NestedTensor):
        xs = self.body(tensor_list.tensors)
        out = []
        for x in xs:
            mask = F.interpolate(tensor_list.mask[None].float(), size=x.shape[-2:]).to(torch.bool)[0]
            out.append(NestedTensor(x, mask))
        return out


class PositionEmbeddingSineHW(nn.Module):
    """reference position_encoding.py:62-108 (normalize=True, scale 2*pi, temperatures from the config)."""

    def __init__(self, num_pos_feats=128, temperatureH=20, temperatureW=20):
        super().__init__()
        self.num_pos_feats = num_pos_feats
        self.temperatureH = temperatureH
        self.temperatureW = temperatureW
        self.scale = 2 * math.pi

    def forward(self, tensor_list: NestedTensor):
        mask = tensor_list.mask
        not_mask = ~mask
        y_embed = not_mask.cumsum(1, dtype=torch.float32)
        x_embed = not_mask.cumsum(2, dtype=torch.float32)
        eps = 1e-6
        y_embed = y_embed / (y_embed[:, -1:, :] + eps) * self.scale
        x_embed = x_embed / (x_embed[:, :, -1:] + eps) * self.scale
        i = torch.arange(self.num_pos_feats, dtype=torch.float32, device=mask.device)
        dim_tx = self.temperatureW ** (2 * (i // 2) / self.num_pos_feats)
        dim_ty = self.temperatureH ** (2 * (i // 2) / self.num_pos_feats)
        pos_x = x_embed[:, :, :, None] / dim_tx
        pos_y = y_embed[:, :, :, None] / dim_ty
        pos_x = torch.stack((pos_x[:, :, :, 0::2].sin(), pos_x[:, :, :, 1::2].cos()), dim=4).flatten(3)
        pos_y = torch.stack((pos_y[:, :, :, 0::2].sin(), pos_y[:, :, :, 1::2].cos()), dim=4).flatten(3)
        return torch.cat((pos_y, pos_x), dim=3).permute(0, 3, 1, 2)


class Joiner(nn.Sequential):
    """reference backbone.py:131-144"""

    def forward(self, tensor_list: NestedTensor):
        xs = self[0](tensor_list)
        pos = [self[1](x).to(x.tensors.dtype) for x in xs]
        return xs, pos


def build_position_encoding(args):
    if args.position_embedding not in ("v2", "sine"):
        raise NotImplementedError("dtlr_b200 implements position_embedding='sine' only")
    return PositionEmbeddingSineHW(args.hidden_dim // 2, args.pe_temperatureH, args.pe_temperatureW)


TORCHVISION_RESNET50_FILES = ("resnet50-0676ba61.pth", "resnet50-19c8e357.pth", "resnet50-11ad3fa6.pth")


def find_pretrained_resnet50(args=None):
    """Where ImageNet ResNet-50 weights can come from without a network: `args.backbone_pretrained` (a torchvision-format state-dict
    file), $DTLR_RESNET50_WEIGHTS, or a file torchvision already downloaded into the torch hub cache."""
    import os
    cand = [getattr(args, "backbone_pretrained", None), os.environ.get("DTLR_RESNET50_WEIGHTS")]
    hub = os.path.join(torch.hub.get_dir(), "checkpoints")
    cand += [os.path.join(hub, f) for f in TORCHVISION_RESNET50_FILES]
    return next((c for c in cand if c and os.path.isfile(c)), None)


def load_pretrained_resnet50(body, path):
    """reference backbone.py:118-121 builds torchvision resnet50(pretrained=is_main_process()): rank 0 holds ImageNet weights (incl. the
    FrozenBatchNorm statistics) and DDP's construction-time broadcast hands them to the other ranks.  Here: rank 0 reads `path`
    (torchvision key layout; `fc.*` and `num_batches_tracked` are dropped as IntermediateLayerGetter / FrozenBatchNorm2d do), then the
    tensors are broadcast when torch.distributed is initialised."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    if rank == 0:
        sd = torch.load(path, map_location="cpu")
        sd = sd.get("state_dict", sd.get("model", sd)) if isinstance(sd, dict) else sd
        sd = {k: v for k, v in sd.items() if not k.startswith("fc.") and not k.endswith("num_batches_tracked")}
        missing, unexpected = body.load_state_dict(sd, strict=False)
        if missing or unexpected:
            raise RuntimeError("pretrained ResNet-50 file %s does not match the torchvision layout: missing %s unexpected %s"
                               % (path, list(missing)[:5], list(unexpected)[:5]))
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        for t in list(body.parameters()) + list(body.buffers()):
            dist.broadcast(t.data, src=0)


def build_backbone(args):
    """reference backbone.py:147-219.  Divergence handled here: the reference DOWNLOADS ImageNet weights on rank 0
    (backbone.py:118-121); there is no network on the target boxes, so the weights are taken from a local file
    (find_pretrained_resnet50) and, when none exists, the body keeps its random init with a warning -- harmless for checkpoint-based
    inference / fine-tuning (the checkpoint overwrites every backbone tensor), wrong for from-scratch training."""
    import warnings
    train_backbone = args.lr_backbone > 0
    if not train_backbone:
        raise ValueError("Please set lr_backbone > 0")
    backbone = Backbone(args.backbone, train_backbone, args.dilation, args.return_interm_indices)
    path = find_pretrained_resnet50(args)
    if path is not None:
        load_pretrained_resnet50(backbone.body, path)
    elif getattr(args, "backbone_pretrained", None) or getattr(args, "require_pretrained_backbone", False):
        raise FileNotFoundError("pretrained ResNet-50 weights requested but not found: %r" % getattr(args, "backbone_pretrained", None))
    else:
        warnings.warn("dtlr_b200: no ImageNet ResNet-50 weights found (args.backbone_pretrained / $DTLR_RESNET50_WEIGHTS / torch hub "
                      "cache): the frozen stem + layer1 and the FrozenBatchNorm statistics stay at their random / identity init. "
                      "Load a checkpoint before use, or pass the torchvision resnet50 state dict for from-scratch training.")
    model = Joiner(backbone, build_position_encoding(args))
    model.num_channels = backbone.num_channels
    return model
