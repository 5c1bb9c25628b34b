import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch, conftest
conftest.force_ieee_fp32()
from torchvision.models.detection.mask_rcnn import MaskRCNNHeads as TVHeads, MaskRCNNPredictor as TVPredictor
from torchvision.models.detection.roi_heads import maskrcnn_loss as tv_loss
from sfvos_b200 import MaskRCNNHeads, MaskRCNNPredictor, maskrcnn_loss
from oracle import roi_oracle as ro

def nerr(a, b): return (a - b).abs().max().item() / (b.abs().max().item() + 1e-20)
for precision in ["fp32", "bf16"]:
    torch.manual_seed(3)
    hr, pr = TVHeads(256, (256,) * 4, 1).cuda(), TVPredictor(256, 256, 2).cuda()
    h, p = MaskRCNNHeads(256, (256,) * 4, 1).cuda(), MaskRCNNPredictor(256, 256, 2).cuda()
    h.load_state_dict(hr.state_dict()); p.load_state_dict(pr.state_dict())
    h.precision = p.precision = precision
    K = 9
    g = torch.Generator().manual_seed(1)
    x = torch.randn(K, 256, 14, 14, generator=g).cuda()
    xr, xo = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
    gt = (torch.rand(2, 96, 160, generator=g) > 0.5).to(torch.uint8).cuda()
    props = [torch.cat(ro.synthetic_rois(1, K, image_hw=(96, 160), seed=3, lo=8.0, hi=150.0)).cuda()]
    labels, matched = [torch.tensor([1, 1]).cuda()], [torch.randint(0, 2, (K,), generator=g).cuda()]
    fr = hr(xr); fr.retain_grad(); lr = pr(fr); lr.retain_grad()
    tv_loss(lr, props, [gt], labels, matched).backward()
    fo = h(xo); fo.retain_grad(); lo = p(fo); lo.retain_grad()
    maskrcnn_loss(lo, props, [gt], labels, matched).backward()
    print("==", precision, "logits", nerr(lo, lr), "glogits", nerr(lo.grad, lr.grad), "g_headout", nerr(fo.grad.float(), fr.grad), "gx", nerr(xo.grad, xr.grad))
    for (n1, p1), (n2, p2) in zip(list(p.named_parameters()) + list(h.named_parameters())[::-1], list(pr.named_parameters()) + list(hr.named_parameters())[::-1]):
        rel = (p1.grad - p2.grad).norm().item() / (p2.grad.norm().item() + 1e-20)
        print(f"  {n1:28s} maxnorm {nerr(p1.grad, p2.grad):.3e} relL2 {rel:.3e}")
