"""B200-native multi-task joint ERFNet — drop-in for the reference's ``models/erfnet_multi_task.py`` (SURVEY 8a note and
8f-4): ONE shared encoder of plain ``non_bottleneck_1d`` blocks (no adapters, a single BatchNorm per position, Dropout2d
0.03 / 0.3) and one decoder head per domain.  Same class names, constructor order (hence the same default initialisation
under a seed), state_dict keys and ``Net.forward(input, task)`` as the reference; every block runs in the kernels of
libmdil_b200.so (the adapter-off, dropout-on mode of the fused block).  No fallback path.

Reference lines (models/erfnet_multi_task.py): DownsamplerBlock :13-25, non_bottleneck_1d :28-64, Encoder :73-100,
UpsamplerBlock :103-114, Decoder :117-145, Net :150-163.
"""
from __future__ import annotations

import torch.nn as nn

from . import erfnet_RA_parallel as _rap
from . import functional as F_

current_task = 0

non_bottleneck_1d = _rap.non_bottleneck_1d
UpsamplerBlock = _rap.UpsamplerBlock
Decoder = _rap.Decoder


class DownsamplerBlock(nn.Module):
    def __init__(self, ninput, noutput):
        super().__init__()
        self.conv = nn.Conv2d(ninput, noutput - ninput, (3, 3), stride=2, padding=1, bias=True)
        self.pool = nn.MaxPool2d(2, stride=2)
        self.bn = nn.BatchNorm2d(noutput, eps=1e-3)
        self._cache = F_.PackedCache()

    def forward(self, input):
        cfg = F_.SampConfig(self.training, _rap._bn_buffers(self.bn), self._cache, 0)
        out = F_.DownFn.apply(input, cfg, self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias)
        return out


class Encoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 16)
        self.layers = nn.ModuleList()
        self.layers.append(DownsamplerBlock(16, 64))
        for _ in range(0, 5):
            self.layers.append(non_bottleneck_1d(64, 0.03, 1))
        self.layers.append(DownsamplerBlock(64, 128))
        for _ in range(0, 2):
            self.layers.append(non_bottleneck_1d(128, 0.3, 2))
            self.layers.append(non_bottleneck_1d(128, 0.3, 4))
            self.layers.append(non_bottleneck_1d(128, 0.3, 8))
            self.layers.append(non_bottleneck_1d(128, 0.3, 16))

    def forward(self, input, predict=False, drop_noise=None):
        """``drop_noise``: optional per-layer list of Dropout2d noise tensors (parity tests replay the reference's
        RNG stream with it)."""
        output = self.initial_block(input)
        for i, layer in enumerate(self.layers):
            if drop_noise is not None and isinstance(layer, non_bottleneck_1d):
                output = layer(output, drop_noise[i])
            else:
                output = layer(output)
        return output


class Net(nn.Module):
    def __init__(self, num_classes=[20], nb_tasks=1, cur_task=0):
        super().__init__()
        print('hi, inside erfnet_multi_task.py', current_task, nb_tasks)
        self.encoder = Encoder()
        self.decoder = nn.ModuleList([Decoder(num_classes[i]) for i in range(nb_tasks)])

    def forward(self, input, task, drop_noise=None):
        global current_task
        current_task = task
        if drop_noise is not None:
            output = self.encoder(input, drop_noise=drop_noise)
        else:
            output = self.encoder(input)
        return self.decoder[task].forward(output)
