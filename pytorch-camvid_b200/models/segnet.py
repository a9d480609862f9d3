"""Drop-in for the reference's models/segnet.py: same constructor, submodule tree (`conv / bn / relu` attribute
names, so state_dict keys match) and initialisation order; forward() runs engine.SegNetPlan.

Reference: models/segnet.py:5-17 (BasicConv), :19-80 (SegNet.__init__), :82-119 (forward).
"""
import torch.nn as nn

from .. import engine


class BasicConv(nn.Module):
    def __init__(self, input_channels, out_channels):
        super().__init__()
        self.conv = nn.Conv2d(input_channels, out_channels, 3, padding=1)
        self.bn = nn.BatchNorm2d(out_channels)
        self.relu = nn.ReLU()

    def forward(self, x):  # models/segnet.py:14-17, through engine.run_block (the whole network runs one plan instead)
        return engine.run_block(self, self.conv, self.bn, x)


class SegNet(nn.Module):
    # (stage name, channel chain) in construction order, models/segnet.py:23-77
    ENCODERS = ((64, 64), (128, 128), (256, 256, 256), (512, 512, 512), (512, 512, 512))

    def __init__(self, input_channels, class_num):
        super().__init__()
        self.input_channels, self.class_num = input_channels, class_num
        prev = input_channels
        for i, chain in enumerate(self.ENCODERS):
            layers = []
            for c in chain:
                layers.append(BasicConv(prev, c))
                prev = c
            setattr(self, f"encoder{i + 1}", nn.Sequential(*layers))
        # decoders mirror the encoders: decoderK keeps its width, its last conv narrows to encoder(K-1)'s width
        outs = {5: 512, 4: 256, 3: 128, 2: 64, 1: class_num}
        for k in (5, 4, 3, 2, 1):
            depth = len(self.ENCODERS[k - 1])
            layers = [BasicConv(prev, prev) for _ in range(depth - 1)] + [BasicConv(prev, outs[k])]
            setattr(self, f"decoder{k}", nn.Sequential(*layers))
            prev = outs[k]
        self.maxpool = nn.MaxPool2d(2, return_indices=True)  # models/segnet.py:79
        self.unpool = nn.MaxUnpool2d(2)  # models/segnet.py:80

    def forward(self, x):
        return engine.run_module(self, engine.SegNetPlan, x)
