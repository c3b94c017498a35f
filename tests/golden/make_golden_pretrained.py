"""Known-answer fixture G1: the reference's SHIPPED trained weights (trained_models/erfnet_pretrained.pth,
a plain 20-class ERFNet) run through the UNMODIFIED reference on CPU.

    python tests/golden/make_golden_pretrained.py

1. the checkpoint is loaded (strict) into the reference's models/erfnet.py Net(20);
2. the same tensors are remapped (bn -> bn_ini.0, bn1 -> bns_1.0, bn2 -> bns_2.0, decoder. -> decoder.0.) into the
   reference's models/erfnet_RA_parallel.py Net([20], 1, 0) with every parallel_conv zeroed: the RAP network then IS the
   plain ERFNet, and the script asserts that the two reference networks agree bit-exactly;
3. the fixture stores the RAP state_dict, the seed of the input and the reference logits.

To keep the fixture small the float tensors are rounded to bfloat16 BEFORE either network sees them and stored as the
16-bit patterns (a trained network's value ranges and BatchNorm statistics are what matters here, not its accuracy).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import REF, load_by_path  # noqa: E402

OUT = os.path.join(HERE, "pretrained_eval.npz")
X_SEED, H, W = 700, 128, 256          # BASELINE configs[0]: N=1, 128 x 256, 1 task, eval forward


def to_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def main():
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    plain = load_by_path("ref_erfnet_plain", os.path.join(REF, "models", "erfnet.py"))
    rap = load_by_path("ref_erfnet_RA_parallel", os.path.join(REF, "models", "erfnet_RA_parallel.py"))
    ck = torch.load(os.path.join(REF, "trained_models", "erfnet_pretrained.pth"), map_location="cpu", weights_only=False)
    ck = {k[len("module."):] if k.startswith("module.") else k: to_bf16(v) if v.dtype.is_floating_point else v
          for k, v in ck.items()}

    net_plain = plain.Net(20)
    own = net_plain.state_dict()
    missing = [k for k in own if k not in ck and "num_batches_tracked" not in k]
    assert not missing and all(k in own for k in ck), (missing, [k for k in ck if k not in own][:5])
    net_plain.load_state_dict({**{k: v for k, v in own.items() if "num_batches_tracked" in k}, **ck}, strict=True)
    net_plain.eval()

    net_rap = rap.Net([20], 1, 0)
    sd = net_rap.state_dict()
    remapped = {}
    for k, v in ck.items():
        nk = k.replace("decoder.", "decoder.0.") if k.startswith("decoder.") else k
        if nk.startswith("encoder."):
            nk = nk.replace(".bn1.", ".bns_1.0.").replace(".bn2.", ".bns_2.0.").replace(".bn.", ".bn_ini.0.")
        assert nk in sd and sd[nk].shape == v.shape, (k, nk)
        remapped[nk] = v
    for k, v in sd.items():
        if k in remapped:
            continue
        if "parallel_conv" in k:
            remapped[k] = torch.zeros_like(v)
        else:
            assert "num_batches_tracked" in k, k
            remapped[k] = v.clone()
    net_rap.load_state_dict(remapped, strict=True)
    net_rap.eval()

    x = torch.rand(1, 3, H, W, generator=torch.Generator().manual_seed(X_SEED))
    with torch.no_grad():
        y_plain = net_plain(x)
        y_rap = net_rap(x, 0)
    assert torch.equal(y_plain, y_rap), float((y_plain - y_rap).abs().max())
    print("plain ERFNet == RAP with zero adapters: bit-exact; logits range", float(y_rap.min()), float(y_rap.max()))

    arrays = {"x_seed": np.int64(X_SEED), "hw": np.array([H, W]), "logits": y_rap.numpy(),
              "keys": np.array(list(sd.keys()))}
    for i, k in enumerate(sd.keys()):
        v = remapped[k]
        if v.dtype.is_floating_point:
            bits = v.contiguous().view(torch.int32).numpy()
            assert np.all((bits & 0xFFFF) == 0), k
            arrays[f"w{i}"] = (bits >> 16).astype(np.uint16)
        else:
            arrays[f"w{i}"] = v.numpy()
    np.savez_compressed(OUT, **arrays)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
