"""Drop-in for the reference's models/unet.py: same constructor, submodule tree, parameter names, state_dict keys
and default initialisation order -- but forward() runs the B200 kernel plan (engine.UNetPlan) instead of ATen ops.

Reference: models/unet.py:5-17 (BasicConv2d), :19-32 (UpSample2d), :35-92 (UNet.__init__), :94-156 (forward).
UNet.forward runs one plan for the whole network; BasicConv2d / UpSample2d also have a forward of their own (the same
kernels through engine.run_block) so that code using a sub-layer directly -- `net.down1(x)`, feature extraction, a
custom forward -- keeps working like with the reference modules.
"""
import torch.nn as nn

from .. import engine


class BasicConv2d(nn.Module):
    """conv3x3(pad 1) -> BatchNorm2d -> ReLU, held as `conv.0 / conv.1 / conv.2` like models/unet.py:10-14."""

    def __init__(self, in_channels, out_channels):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_channels, out_channels, 3, padding=1), nn.BatchNorm2d(out_channels),
                                  nn.ReLU(inplace=True))

    def forward(self, x):  # models/unet.py:16-17
        return engine.run_block(self, self.conv[0], self.conv[1], x)


class UpSample2d(nn.Module):
    """bilinear x2 (align_corners=True) followed by a BasicConv2d, models/unet.py:19-32."""

    def __init__(self, in_channels, out_channels, scale_factor=2.0):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = BasicConv2d(in_channels, out_channels)

    def forward(self, x):  # models/unet.py:28-32: upsample and conv block in one plan
        return engine.run_block(self, self.conv.conv[0], self.conv.conv[1], x, upsample=True)


class UNet(nn.Module):
    WIDTHS = (64, 128, 256, 512, 1024)

    def __init__(self, input_channels, class_num):
        super().__init__()
        self.input_channels, self.class_num = input_channels, class_num
        prev = input_channels
        for i, c in enumerate(self.WIDTHS):  # contracting path, models/unet.py:40-65
            setattr(self, f"down{i + 1}", nn.Sequential(BasicConv2d(prev, c), BasicConv2d(c, c)))
            prev = c
        for i, c in enumerate(reversed(self.WIDTHS[:-1])):  # expansive path, models/unet.py:67-89
            setattr(self, f"upsample{i + 1}", UpSample2d(2 * c, c))
            setattr(self, f"up{i + 1}", nn.Sequential(BasicConv2d(2 * c, c), BasicConv2d(c, c)))
        self.output = BasicConv2d(self.WIDTHS[0], class_num)  # models/unet.py:91
        self.maxpool = nn.MaxPool2d(2, 2)  # models/unet.py:92 (parameter-free; kept for attribute parity)

    def forward(self, x):
        return engine.run_module(self, engine.UNetPlan, x)
