"""ResNet-18 / ResNet-50 backbones for the B200 VINCE hot path.

Mirrors /root/reference/models/building_blocks/backbone_models.py:21-75 (`Backbone`, `ResNet18`, `ResNet50`):
same constructor signature `(args, final_layer)`, same `.model` / `.output_channels` attributes and the same
parameter tree as `torchvision.models.resnet18/50` (so checkpoints keep the reference's
`feature_extractor.module.model.*` state-dict keys), but `forward` does not call a single torch op on the data:
it drives the sm_100a kernels through `EncoderRunner` (vince_b200/encoder.py).

The torch.nn modules below are PARAMETER CONTAINERS ONLY (initialisation, state_dict, .to(), train()/eval());
their own forward() is never used.
"""
from torch import nn

from .encoder import EncoderRunner

__all__ = ["ResNet18", "ResNet50"]


class _BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.relu = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride


class _Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, 1, 0, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)     # torchvision v1.5: stride on the 3x3
        self.bn2 = nn.BatchNorm2d(planes)
        self.conv3 = nn.Conv2d(planes, planes * 4, 1, 1, 0, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride


class ResNetParams(nn.Module):
    """Same module / parameter names, shapes and default initialisation as torchvision.models.ResNet
    (spec: /root/reference/models/building_blocks/resnet.py:140-229)."""

    def __init__(self, block, layers, num_classes=1000):
        super().__init__()
        self.inplanes = 64
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        self.layer1 = self._make_layer(block, 64, layers[0], 1)
        self.layer2 = self._make_layer(block, 128, layers[1], 2)
        self.layer3 = self._make_layer(block, 256, layers[2], 2)
        self.layer4 = self._make_layer(block, 512, layers[3], 2)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * block.expansion, num_classes)     # unused by VINCE but EMA'd / weight-decayed
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _make_layer(self, block, planes, blocks, stride):
        downsample = None
        if stride != 1 or self.inplanes != planes * block.expansion:
            downsample = nn.Sequential(nn.Conv2d(self.inplanes, planes * block.expansion, 1, stride, bias=False),
                                       nn.BatchNorm2d(planes * block.expansion))
        layers = [block(self.inplanes, planes, stride, downsample)]
        self.inplanes = planes * block.expansion
        for _ in range(1, blocks):
            layers.append(block(self.inplanes, planes, 1, None))
        return nn.Sequential(*layers)

    def forward(self, x):
        raise RuntimeError("ResNetParams is a parameter container; use Backbone.forward (CUDA kernels)")


class Backbone(nn.Module):
    """backbone_models.py:21-54.  Only final_layer == -2 (through layer4; what VinceModel asks for at
    vince_model.py:26) is implemented by the fused CUDA encoder."""

    def __init__(self, args, model, final_layer=None):
        super().__init__()
        self.args = args
        self.model = model
        self.final_layer = final_layer
        n_children = len(list(self.model.children()))
        if final_layer is not None and final_layer < 0:
            self.final_layer = n_children + final_layer
        if self.final_layer != n_children - 2:
            raise NotImplementedError(
                "vince_b200.Backbone implements final_layer=-2 (output of layer4) only; got %r" % (final_layer,))
        self.output_channels = self.model.output_channels
        self.runner = EncoderRunner(self.model, passes=getattr(args, "vince_b200_passes", 3))

    def forward(self, x, final_layer=None, gather_idx=None, scatter_idx=None, want_pooled=False, patch_grid=1,
                tape=False):
        """x: NCHW fp32 CUDA tensor.  Returns NCHW spatial features [B, C, h, w] (and the global-average-pooled
        [B, C] when want_pooled).  gather_idx / scatter_idx fold the MoCo batch shuffle / un-shuffle
        (vince_model.py:137-142,184-192) into the first load and the last store."""
        if final_layer is not None and final_layer != self.final_layer and final_layer != -2:
            raise NotImplementedError("vince_b200.Backbone: only final_layer=-2 is implemented")
        spatial, pooled = self.runner.forward(x, train=self.training, gather_idx=gather_idx, scatter_idx=scatter_idx,
                                              patch_grid=patch_grid, tape=tape)
        return (spatial, pooled) if want_pooled else spatial


class ResNet18(Backbone):
    def __init__(self, args, final_layer=None):
        model = ResNetParams(_BasicBlock, [2, 2, 2, 2])
        model.output_channels = 512
        super().__init__(args, model, final_layer)


class ResNet50(Backbone):
    def __init__(self, args, final_layer=None):
        model = ResNetParams(_Bottleneck, [3, 4, 6, 3])
        model.output_channels = 2048
        super().__init__(args, model, final_layer)


class DataParallelShim(nn.Module):
    """Stands in for dg_util's get_data_parallel (vince_model.py:35): keeps the `.module` level in the
    state-dict keys.  One process per GPU replaces nn.DataParallel's fan-out."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, *a, **k):
        return self.module(*a, **k)
