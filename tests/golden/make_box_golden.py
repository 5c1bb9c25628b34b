"""Generate tests/golden/box_head.npz from the live torchvision modules the reference reaches at
code/helpers/model.py:346 (RoIHeads.forward -> box_head -> box_predictor -> fastrcnn_loss; TV roi_heads.py:772-786).

Run in the build container:  python tests/golden/make_box_golden.py
The modules are torchvision's own TwoMLPHead(256*7*7, 1024) / FastRCNNPredictor(1024, 2) (code/helpers/model.py:13-20
builds exactly these) under torch.manual_seed(21); the 13 M weights are NOT stored -- the test rebuilds them from the
seed -- only the inputs, outputs, losses and gradient summaries are.
"""
import os

import numpy as np
import torch
import torchvision
from torchvision.models.detection.faster_rcnn import FastRCNNPredictor, TwoMLPHead
from torchvision.models.detection.roi_heads import fastrcnn_loss

HERE = os.path.dirname(os.path.abspath(__file__))
SEED, M = 21, 12


def build():
    torch.manual_seed(SEED)
    return TwoMLPHead(256 * 7 * 7, 1024), FastRCNNPredictor(1024, 2)


def inputs():
    g = torch.Generator().manual_seed(SEED + 1)
    x = torch.randn(M, 256, 7, 7, generator=g)
    labels = [torch.tensor([1, 0, 0, 1, 0, 1]), torch.tensor([0, 0, 1, 1, 0, 0])]
    # targets on both sides of beta = 1/9 (and one exactly large) so both smooth-L1 branches are pinned
    targets = [0.3 * torch.randn(6, 4, generator=g), 0.02 * torch.randn(6, 4, generator=g)]
    targets[0][0, 0] = 3.0
    return x, labels, targets


def main():
    head, pred = build()
    x, labels, targets = inputs()
    x.requires_grad_(True)
    feat = head(x)
    scores, deltas = pred(feat)
    l_cls, l_box = fastrcnn_loss(scores, deltas, labels, targets)
    (l_cls + l_box).backward()
    rec = {"x": x.detach().numpy(), "feat": feat.detach().numpy(), "scores": scores.detach().numpy(),
           "deltas": deltas.detach().numpy(), "loss_cls": np.float64(l_cls.item()), "loss_box": np.float64(l_box.item()),
           "gx": x.grad.numpy(), "tv_version": np.array(torchvision.__version__)}
    for i, (lab, tgt) in enumerate(zip(labels, targets)):
        rec[f"labels{i}"], rec[f"targets{i}"] = lab.numpy(), tgt.numpy()
    for mod, pre in ((head, "box_head."), (pred, "box_predictor.")):
        for n, p in mod.named_parameters():
            rec["gsum_" + pre + n] = np.float64(p.grad.double().sum().item())
            rec["gabs_" + pre + n] = np.float64(p.grad.double().abs().sum().item())
            rec["g64_" + pre + n] = p.grad.reshape(-1)[:: max(1, p.numel() // 64)][:64].numpy()
    np.savez_compressed(os.path.join(HERE, "box_head.npz"), **rec)
    print("box fixture: loss_cls", l_cls.item(), "loss_box", l_box.item())


if __name__ == "__main__":
    main()
