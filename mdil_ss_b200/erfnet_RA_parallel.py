"""B200-native ERFNet with parallel residual adapters — drop-in for the reference's
``models/erfnet_RA_parallel.py`` (same class names, constructor signatures, submodule tree,
state_dict keys, parameter order and default initialisation; ``Net.forward(input, task)``).

The torch.nn layers created in the constructors are parameter CONTAINERS only (they give the
reference's state_dict keys, init and ``repr`` for free); every forward/backward runs in the
hand-written sm_100a kernels of libmdil_b200.so through ``mdil_ss_b200.functional``.  There is no
PyTorch/cuDNN or CPU fallback: a CPU tensor or a missing library raises.

Reference lines (models/erfnet_RA_parallel.py): DownsamplerBlock :13-25, non_bottleneck_1d :28-64,
non_bottleneck_1d_RAP :67-113, Encoder :123-149, UpsamplerBlock :152-162, Decoder :165-190, Net :194-212.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import functional as F_

_PREPACK = os.environ.get("MDIL_PREPACK", "1") != "0"     # A/B switch: weight packing on a side stream at the step start

# Domain selector read by the blocks at call time, exactly like the reference's module-global (:11, :22, :91).
current_task = 0


def _bn_buffers(bn: nn.BatchNorm2d):
    # running statistics and the batch counter: all three are updated by the train-mode forward launch itself
    return (bn.running_mean, bn.running_var, bn.num_batches_tracked)


class DownsamplerBlock(nn.Module):
    def __init__(self, ninput, noutput, nb_tasks=1):
        super().__init__()
        self.conv = nn.Conv2d(ninput, noutput - ninput, (3, 3), stride=2, padding=1, bias=True)
        self.pool = nn.MaxPool2d(2, stride=2)
        self.bn_ini = nn.ModuleList([nn.BatchNorm2d(noutput, eps=1e-3) for _ in range(nb_tasks)])
        self._cache = F_.PackedCache()

    def forward(self, input):
        task = current_task
        bn = self.bn_ini[task]
        cfg = F_.SampConfig(self.training, _bn_buffers(bn), self._cache, 0)
        out = F_.DownFn.apply(input, cfg, self.conv.weight, self.conv.bias, bn.weight, bn.bias)
        return out

    def _prepack(self, task):
        F_.prepack_down(self._cache, 0, self.conv.weight, self.conv.in_channels)


class non_bottleneck_1d(nn.Module):
    def __init__(self, chann, dropprob, dilated):
        super().__init__()
        self.conv3x1_1 = nn.Conv2d(chann, chann, (3, 1), stride=1, padding=(1, 0), bias=True)
        self.conv1x3_1 = nn.Conv2d(chann, chann, (1, 3), stride=1, padding=(0, 1), bias=True)
        self.bn1 = nn.BatchNorm2d(chann, eps=1e-03)
        self.conv3x1_2 = nn.Conv2d(chann, chann, (3, 1), stride=1, padding=(1 * dilated, 0), bias=True,
                                   dilation=(dilated, 1))
        self.conv1x3_2 = nn.Conv2d(chann, chann, (1, 3), stride=1, padding=(0, 1 * dilated), bias=True,
                                   dilation=(1, dilated))
        self.bn2 = nn.BatchNorm2d(chann, eps=1e-03)
        self.dropout = nn.Dropout2d(dropprob)
        self._dil = dilated
        self._cache = F_.PackedCache()

    def _drop_noise(self, x):
        # F.dropout2d's noise: [N,C,1,1] bernoulli(1-p)/(1-p); skipped when p == 0 (:61) and in eval mode
        p = self.dropout.p
        if p == 0 or not self.training:
            return None
        return x.new_empty(x.shape[0], x.shape[1], 1, 1).bernoulli_(1 - p).div_(1 - p)

    def forward(self, input, drop_noise=None):
        noise = drop_noise if drop_noise is not None else self._drop_noise(input)
        cfg = F_.Nb1dConfig(self._dil, False, self.training, _bn_buffers(self.bn1), _bn_buffers(self.bn2),
                            self._cache, 0)
        out = F_.Nb1dFn.apply(input, noise, cfg,
                              self.conv3x1_1.weight, self.conv3x1_1.bias, self.conv1x3_1.weight, self.conv1x3_1.bias,
                              self.conv3x1_2.weight, self.conv3x1_2.bias, self.conv1x3_2.weight, self.conv1x3_2.bias,
                              self.bn1.weight, self.bn1.bias, self.bn2.weight, self.bn2.bias)
        return out

    def _prepack(self, task):
        F_.prepack_nb1d(self._cache, 0, False,
                        (self.conv3x1_1.weight, self.conv3x1_1.bias, self.conv1x3_1.weight, self.conv1x3_1.bias,
                         self.conv3x1_2.weight, self.conv3x1_2.bias, self.conv1x3_2.weight, self.conv1x3_2.bias,
                         self.bn1.weight, self.bn1.bias, self.bn2.weight, self.bn2.bias))


class non_bottleneck_1d_RAP(nn.Module):
    def __init__(self, chann, dropprob, dilated, nb_tasks=1):
        super().__init__()
        self.conv3x1_1 = nn.Conv2d(chann, chann, (3, 1), stride=1, padding=(1, 0), bias=True)
        self.conv1x3_1 = nn.Conv2d(chann, chann, (1, 3), stride=1, padding=(0, 1), bias=True)
        # domain-specific 1x1 adapters and BatchNorms
        self.parallel_conv_1 = nn.ModuleList(
            [nn.Conv2d(chann, chann, kernel_size=1, stride=1, padding=0, bias=True) for _ in range(nb_tasks)])
        self.bns_1 = nn.ModuleList([nn.BatchNorm2d(chann, eps=1e-03) for _ in range(nb_tasks)])
        self.conv3x1_2 = nn.Conv2d(chann, chann, (3, 1), stride=1, padding=(1 * dilated, 0), bias=True,
                                   dilation=(dilated, 1))
        self.conv1x3_2 = nn.Conv2d(chann, chann, (1, 3), stride=1, padding=(0, 1 * dilated), bias=True,
                                   dilation=(1, dilated))
        self.parallel_conv_2 = nn.ModuleList(
            [nn.Conv2d(chann, chann, kernel_size=1, stride=1, padding=0, bias=True) for _ in range(nb_tasks)])
        self.bns_2 = nn.ModuleList([nn.BatchNorm2d(chann, eps=1e-03) for _ in range(nb_tasks)])
        self.dropout = nn.Dropout2d(dropprob)
        self._dil = dilated
        self._cache = F_.PackedCache()

    _drop_noise = non_bottleneck_1d._drop_noise

    def forward(self, input, drop_noise=None):
        task = current_task
        noise = drop_noise if drop_noise is not None else self._drop_noise(input)
        bn1, bn2 = self.bns_1[task], self.bns_2[task]
        ad1, ad2 = self.parallel_conv_1[task], self.parallel_conv_2[task]
        cfg = F_.Nb1dConfig(self._dil, True, self.training, _bn_buffers(bn1), _bn_buffers(bn2), self._cache, task)
        out = F_.Nb1dFn.apply(input, noise, cfg,
                              self.conv3x1_1.weight, self.conv3x1_1.bias, self.conv1x3_1.weight, self.conv1x3_1.bias,
                              self.conv3x1_2.weight, self.conv3x1_2.bias, self.conv1x3_2.weight, self.conv1x3_2.bias,
                              bn1.weight, bn1.bias, bn2.weight, bn2.bias,
                              ad1.weight, ad1.bias, ad2.weight, ad2.bias)
        return out

    def _prepack(self, task):
        bn1, bn2 = self.bns_1[task], self.bns_2[task]
        ad1, ad2 = self.parallel_conv_1[task], self.parallel_conv_2[task]
        F_.prepack_nb1d(self._cache, task, True,
                        (self.conv3x1_1.weight, self.conv3x1_1.bias, self.conv1x3_1.weight, self.conv1x3_1.bias,
                         self.conv3x1_2.weight, self.conv3x1_2.bias, self.conv1x3_2.weight, self.conv1x3_2.bias,
                         bn1.weight, bn1.bias, bn2.weight, bn2.bias, ad1.weight, ad1.bias, ad2.weight, ad2.bias))


class Encoder(nn.Module):
    def __init__(self, nb_tasks=1):
        super().__init__()
        self.initial_block = DownsamplerBlock(3, 16, nb_tasks)
        self.layers = nn.ModuleList()
        self.layers.append(DownsamplerBlock(16, 64, nb_tasks))
        for _ in range(0, 5):
            self.layers.append(non_bottleneck_1d_RAP(64, 0.03, 1, nb_tasks))
        self.layers.append(DownsamplerBlock(64, 128, nb_tasks))
        for _ in range(0, 2):
            self.layers.append(non_bottleneck_1d_RAP(128, 0.3, 2, nb_tasks))
            self.layers.append(non_bottleneck_1d_RAP(128, 0.3, 4, nb_tasks))
            self.layers.append(non_bottleneck_1d_RAP(128, 0.3, 8, nb_tasks))
            self.layers.append(non_bottleneck_1d_RAP(128, 0.3, 16, nb_tasks))

    def _draw_noise(self, n, device):
        """The Dropout2d noise of every block of one training forward from ONE uniform draw ([N, C, 1, 1] per layer,
        bernoulli(1 - p) / (1 - p) like F.dropout2d): three small launches instead of two per block."""
        key = (n, str(device))
        plan = getattr(self, "_noise_plan", None)
        if plan is None or plan[0] != key:
            spans, keep = [], []
            off = 0
            for i, layer in enumerate(self.layers):
                if isinstance(layer, non_bottleneck_1d_RAP) and layer.dropout.p > 0:
                    c = layer.bns_1[0].num_features
                    spans.append((i, off, c))
                    keep += [1.0 - layer.dropout.p] * (n * c)
                    off += n * c
            keep_t = torch.tensor(keep, dtype=torch.float32, device=device)
            plan = (key, spans, keep_t, 1.0 / keep_t, off)
            self._noise_plan = plan
        _, spans, keep_t, inv_t, total = plan
        noise = [None] * len(self.layers)
        if total == 0:
            return noise
        flat = (torch.rand(total, device=device) < keep_t).to(torch.float32) * inv_t
        for i, off, c in spans:
            noise[i] = flat[off:off + n * c].view(n, c, 1, 1)
        return noise

    def forward(self, input, predict=False, drop_noise=None):
        """``drop_noise``: optional per-layer list of Dropout2d noise tensors (None entries = draw / skip), used by
        the parity tests to replay the reference's RNG stream."""
        if drop_noise is None and self.training and input.is_cuda:
            drop_noise = self._draw_noise(input.shape[0], input.device)
        output = self.initial_block(input)
        join = getattr(self, "_prepack_join", None)
        if join is not None:       # Net._prepack_async: the other blocks' weight packing ran under the initial block
            self._prepack_join = None
            torch.cuda.current_stream(input.device).wait_stream(join)
        for i, layer in enumerate(self.layers):
            if drop_noise is not None and isinstance(layer, non_bottleneck_1d_RAP):
                output = layer(output, drop_noise[i])
            else:
                output = layer(output)
        return output


class UpsamplerBlock(nn.Module):
    def __init__(self, ninput, noutput):
        super().__init__()
        self.conv = nn.ConvTranspose2d(ninput, noutput, 3, stride=2, padding=1, output_padding=1, bias=True)
        self.bn = nn.BatchNorm2d(noutput, eps=1e-3)
        self._cache = F_.PackedCache()

    def forward(self, input):
        cfg = F_.SampConfig(self.training, _bn_buffers(self.bn), self._cache, 0)
        out = F_.UpFn.apply(input, cfg, self.conv.weight, self.conv.bias, self.bn.weight, self.bn.bias)
        return out

    def _prepack(self, task):
        F_.prepack_up(self._cache, 0, self.conv.weight)


class Decoder(nn.Module):
    def __init__(self, num_classes):
        super().__init__()
        self.layers = nn.ModuleList()
        self.layers.append(UpsamplerBlock(128, 64))
        self.layers.append(non_bottleneck_1d(64, 0, 1))
        self.layers.append(non_bottleneck_1d(64, 0, 1))
        self.layers.append(UpsamplerBlock(64, 16))
        self.layers.append(non_bottleneck_1d(16, 0, 1))
        self.layers.append(non_bottleneck_1d(16, 0, 1))
        self.output_conv = nn.ConvTranspose2d(16, num_classes, 2, stride=2, padding=0, output_padding=0, bias=True)

    def forward(self, input):
        output = input
        for layer in self.layers:
            output = layer(output)
        return F_.OutConvFn.apply(output, self.output_conv.weight, self.output_conv.bias)


class Net(nn.Module):
    def __init__(self, num_classes=[20], nb_tasks=1, cur_task=0):
        super().__init__()
        global current_task
        current_task = cur_task
        print('hi, inside erfnet_RA_parallel', current_task, nb_tasks)
        self.encoder = Encoder(nb_tasks)
        self.decoder = nn.ModuleList([Decoder(num_classes[i]) for i in range(nb_tasks)])
        self.__dict__["_prepack_streams"] = {}     # device -> side stream (shared with nn.DataParallel's replicas)

    def _prepack_async(self, task, device):
        """Training step: the packed weight copies of every block but the first are refreshed on a side stream, under
        the initial downsampler, instead of one small launch in front of every block's forward (the optimiser changed
        all of them).  The blocks' own cache checks then find them current."""
        streams = self.__dict__.setdefault("_prepack_streams", {})
        side = streams.get(device)
        if side is None:
            side = streams[device] = torch.cuda.Stream(device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for layer in list(self.encoder.layers) + list(self.decoder[task].layers):
                layer._prepack(task)
        self.encoder._prepack_join = side

    def forward(self, input, task, drop_noise=None):
        global current_task
        current_task = task
        if self.training and input.is_cuda and torch.is_grad_enabled() and _PREPACK:
            self._prepack_async(task, input.device)
        if drop_noise is not None:
            output = self.encoder(input, drop_noise=drop_noise)
        else:
            output = self.encoder(input)
        output = self.decoder[task].forward(output)
        return output
