"""SegmentationModel -- the caller of the hot path, mirroring code/helpers/model.py:168-389 (same constructor,
attributes, helper-method names, forward contract) with SlowFastLayers and the ROIAlign / mask branch running on
libsfvos.so.  The frozen backbone, RPN, transform and box head are torchvision modules used as-is.

``from helpers.model import SegmentationModel`` keeps working for the reference's train.py / prediction.py through
the shim package in ``compat/helpers``."""
import os
import warnings
from collections import OrderedDict
from math import ceil, floor

import torch
import torchvision
from torch import nn
from torchvision.models.detection.faster_rcnn import FastRCNNPredictor
from torchvision.models.detection.image_list import ImageList

from . import backbone as sf_backbone
from . import roi_heads as sf_roi_heads
from .roi_heads import MaskRCNNPredictor
from .slowfast import SlowFastLayers


def _freeze_batchnorm(module):
    """nn.BatchNorm2d -> torchvision FrozenBatchNorm2d (same statistics / affine values, as buffers), recursively.  The
    pretrained Mask R-CNN the reference starts from has a frozen-BN ResNet (torchvision builds the backbone that way whenever
    it carries trained weights); ``weights=None`` builds trainable BatchNorm2d layers instead, which would change the
    parameter count (+53,120) and, in train mode, the arithmetic."""
    from torchvision.ops.misc import FrozenBatchNorm2d
    for name, child in list(module.named_children()):
        if isinstance(child, torch.nn.BatchNorm2d):
            frozen = FrozenBatchNorm2d(child.num_features, eps=child.eps)
            frozen.weight.copy_(child.weight.detach()); frozen.bias.copy_(child.bias.detach())
            frozen.running_mean.copy_(child.running_mean); frozen.running_var.copy_(child.running_var)
            setattr(module, name, frozen)
        else:
            _freeze_batchnorm(child)


def get_model_instance_segmentation(num_classes, pretrained=True):
    """code/helpers/model.py:12-27.  ``pretrained`` weights need the torchvision hub cache; offline we fall back to
    random init with FrozenBatchNorm2d so the parameter count matches the reference's report."""
    kwargs = {}
    if pretrained:
        try:
            model = torchvision.models.detection.maskrcnn_resnet50_fpn(weights="DEFAULT")
        except Exception as exc:  # no network / no cache
            warnings.warn(f"pretrained Mask R-CNN weights unavailable ({exc.__class__.__name__}); using random init")
            pretrained = False
    if not pretrained:
        kwargs = dict(weights=None, weights_backbone=None, norm_layer=torchvision.ops.misc.FrozenBatchNorm2d)
        try:
            model = torchvision.models.detection.maskrcnn_resnet50_fpn(**kwargs)
        except TypeError:
            kwargs.pop("norm_layer")
            model = torchvision.models.detection.maskrcnn_resnet50_fpn(**kwargs)
    if not pretrained:
        _freeze_batchnorm(model.backbone.body)      # what the pretrained model the reference loads has (model.py:13)
    in_features = model.roi_heads.box_predictor.cls_score.in_features
    model.roi_heads.box_predictor = FastRCNNPredictor(in_features, num_classes)
    in_features_mask = model.roi_heads.mask_predictor.conv5_mask.in_channels
    model.roi_heads.mask_predictor = MaskRCNNPredictor(in_features_mask, 256, num_classes)
    return model


class SegmentationModel(nn.Module):
    def __init__(self, device, slow_pathway_size, fast_pathway_size, maskrcnn_weights='maskrcnn/maskrcnn_model.pth',
                 pretrained=True):
        super().__init__()
        self.device = device
        self.maskrcnn_model = get_model_instance_segmentation(num_classes=2, pretrained=pretrained)
        if maskrcnn_weights and os.path.exists(maskrcnn_weights):
            self.maskrcnn_model.load_state_dict(torch.load(maskrcnn_weights, map_location="cpu"))
        elif maskrcnn_weights:
            warnings.warn(f"{maskrcnn_weights} not found; Mask R-CNN keeps its initial weights")
        for param in self.maskrcnn_model.backbone.parameters():
            param.requires_grad = False
        for param in self.maskrcnn_model.rpn.parameters():
            param.requires_grad = False
        self.slow_pathway_size = slow_pathway_size
        self.fast_pathway_size = fast_pathway_size
        self.slow_fast = SlowFastLayers(256, device=device, slow_pathway_size=slow_pathway_size,
                                        fast_pathway_size=fast_pathway_size)
        self.maskrcnn_model.roi_heads.detections_per_img = 10
        sf_roi_heads.install(self.maskrcnn_model.roi_heads)      # same parameters, libsfvos kernels
        if os.environ.get("SFVOS_NATIVE_BACKBONE", "1") != "0":
            # FPN + RPN head on libsfvos as well: channels-last bf16 features, no layout pass afterwards
            sf_backbone.install_backbone(self.maskrcnn_model)
        self.features_cache = {}
        self.use_caching = True
        # eval only: one temporal sweep per sequence chunk instead of one window per frame (same outputs, see
        # SlowFastLayers.temporally_enhance_sequence); False restores the reference's per-frame loop
        self.sequence_mode = True
        self.sequence_chunk = 32
        self.backbone_batch = 8              # eval sweep only: frames per backbone call (1 = one call per frame like the reference)
        self.batch_rpn = True                # eval sweep only: one RPN call per chunk instead of one per frame
        self.phase_times = {}                # filled when SFVOS_PIPE_TIMING is set (tools/bench_pipeline.py)
        self.transform_on_device = True      # eval sweep only: GeneralizedRCNNTransform on the GPU (the reference runs it on the host)

    # ---- feature cache / windowing (model.py:191-273) ----------------------------------------------------------------
    def compute_maskrcnn_features(self, images_tensors, indices):
        n = len(images_tensors)
        for key in [k for k in self.features_cache if k < indices[0]]:
            self.features_cache.pop(key)
        per_frame = []
        for idx in indices:
            if not 0 <= idx < n:
                continue
            if self.use_caching and idx in self.features_cache:
                feats = self._detach_features(self.features_cache[idx])
            else:
                feats = self.maskrcnn_model.backbone(images_tensors[idx:idx + 1].to(self.device))
                if self.use_caching:
                    self.features_cache[idx] = feats
            per_frame.append(feats)
        left = sum(1 for i in indices if i < 0)
        right = sum(1 for i in indices if i >= n)
        out = OrderedDict()
        for key in per_frame[0].keys():
            stacked = torch.cat([f[key] for f in per_frame])
            if left or right:        # out-of-sequence frames are all-zero feature maps (model.py:215-225)
                fmt = torch.channels_last if stacked.is_contiguous(memory_format=torch.channels_last) else torch.contiguous_format
                parts = [stacked]
                if left:
                    parts.insert(0, torch.zeros((left,) + stacked.shape[1:], dtype=stacked.dtype, device=stacked.device).contiguous(memory_format=fmt))
                if right:
                    parts.append(torch.zeros((right,) + stacked.shape[1:], dtype=stacked.dtype, device=stacked.device).contiguous(memory_format=fmt))
                stacked = torch.cat(parts)
            out[key] = stacked
        return out

    def batch_slice_features(self, features, begin, end):
        return OrderedDict((k, v[begin:end].to(self.device)) for k, v in features.items())

    def compute_rpn_proposals(self, image_tensors, image_sizes, features, target):
        batch_imgs = ImageList(image_tensors.to(self.device), image_sizes)
        return self.maskrcnn_model.rpn(batch_imgs, features, target)

    def _slice_features(self, features, image_feature_idx, pathway_size):
        lo = image_feature_idx - floor(pathway_size / 2)
        hi = image_feature_idx + ceil(pathway_size / 2)
        return OrderedDict((k, v[lo:hi]) for k, v in features.items())

    def _targets_to_device(self, targets, device):
        return [OrderedDict((k, v.to(device)) for k, v in t.items()) for t in targets]

    def _index_features(self, features, i_begin, i_end):
        return OrderedDict((k, v[i_begin:i_end]) for k, v in features.items())

    def _detach_features(self, features):
        return OrderedDict((k, v.detach()) for k, v in features.items())

    # ---- eval-mode sequence sweep ------------------------------------------------------------------
    @torch.no_grad()
    def _forward_eval_sequence(self, transformed_images, targets, valid, original_image_sizes):
        """Same detections as the per-frame loop of forward() in eval mode (model.py:318-348), computed per chunk of
        ``sequence_chunk`` frames: the frozen backbone once per frame (what features_cache achieves in the reference),
        SlowFast as one temporal sweep over the chunk plus its fp-1 halo frames, and ONE roi_heads call for all valid
        frames of the chunk (roi_heads has no cross-image interaction in eval mode)."""
        n = len(valid)
        fp = self.fast_pathway_size
        lo, hi = fp // 2, fp - fp // 2 - 1
        tensors, sizes = transformed_images.tensors, transformed_images.image_sizes
        cache = {}

        def backbone_range(lo_i, hi_i):
            """Per-frame features of frames [lo_i, hi_i): the frozen backbone runs once per frame (what features_cache achieves
            in the reference), ``backbone_batch`` frames per call."""
            missing = [i for i in range(lo_i, hi_i) if i not in cache]
            bb = max(1, int(self.backbone_batch))
            for j in range(0, len(missing), bb):
                grp = missing[j:j + bb]
                out = self.maskrcnn_model.backbone(tensors[grp[0]:grp[-1] + 1].to(self.device))     # missing frames are contiguous
                for n_, i in enumerate(grp):
                    cache[i] = OrderedDict((k, v[n_:n_ + 1]) for k, v in out.items())
            return [cache[i] for i in range(lo_i, hi_i)]

        detections = [{} for _ in range(n)]
        step = max(1, int(self.sequence_chunk))
        prof = self.phase_times if os.environ.get("SFVOS_PIPE_TIMING") else None   # opt-in per-phase wall clock (synchronising)

        def tick(name, t0):
            if prof is None:
                return 0.0
            import time
            torch.cuda.synchronize()
            now = time.perf_counter()
            if t0:
                prof[name] = prof.get(name, 0.0) + now - t0
            return now
        for c0 in range(0, n, step):
            c1 = min(n, c0 + step)
            idxs = [i for i in range(c0, c1) if valid[i]]
            if not idxs:
                continue
            f0, f1 = max(0, c0 - lo), min(n, c1 + hi)
            for key in [k for k in cache if k < f0]:
                cache.pop(key)
            t = tick("", 0.0)
            frames = backbone_range(f0, f1)
            feats = OrderedDict((k, torch.cat([f[k] for f in frames])) for k in frames[0].keys())
            t = tick("backbone", t)
            merged = self.slow_fast.temporally_enhance_sequence(feats, halo=(c0 - f0, f1 - c1))
            t = tick("slowfast_sweep", t)
            proposals = []
            if self.batch_rpn:          # one RPN call for the chunk's valid frames (per-image top-k / NMS inside, as per frame)
                if idxs == list(range(idxs[0], idxs[-1] + 1)):
                    centre = self._index_features(feats, idxs[0] - f0, idxs[-1] + 1 - f0)
                    imgs = tensors[idxs[0]:idxs[-1] + 1]
                else:
                    rel = torch.tensor([i - f0 for i in idxs], device=self.device)
                    centre = OrderedDict((k, v[rel]) for k, v in feats.items())
                    imgs = tensors[idxs]
                proposals, _ = self.compute_rpn_proposals(imgs, [sizes[i] for i in idxs], centre, None)   # eval: targets unused
            else:
                for i in idxs:
                    centre = self._index_features(feats, i - f0, i - f0 + 1)
                    props, _ = self.compute_rpn_proposals(tensors[i:i + 1], sizes[i:i + 1], centre, None)
                    proposals.append(props[0])
            if len(idxs) == c1 - c0:
                sel = merged
            else:
                pick = torch.tensor([i - c0 for i in idxs], device=self.device)
                sel = OrderedDict((k, v[pick]) for k, v in merged.items())
            t = tick("rpn", t)
            image_sizes = sizes[0:1] * len(idxs)        # all images of one sequence have the same size (model.py:342)
            dets, _ = self.maskrcnn_model.roi_heads(sel, proposals, image_sizes)
            t = tick("roi_heads", t)
            # one paste launch and one pinned D2H copy of the masks per chunk (the reference moves every detection to the host, model.py:348)
            dets = sf_roi_heads.postprocess(dets, image_sizes, [original_image_sizes[i] for i in idxs], to_cpu=True)
            for i, d in zip(idxs, dets):
                detections[i] = d
            t = tick("paste_d2h", t)
        return detections

    # ---- forward (model.py:275-389) --------------------------------------------------------------------------------------
    def forward(self, images, targets=None, optimizer=None):
        self.features_cache = {}
        original_image_sizes = [tuple(img.shape[-2:]) for img in images]
        if not self.training and self.sequence_mode:
            # eval sweep: targets only say which frames to skip (model.py:289-296) -- RPN and roi_heads ignore them in eval
            # mode -- so the reference's second transform pass (model.py:299) is not needed; with ``transform_on_device`` the
            # resize / normalise of GeneralizedRCNNTransform runs on the GPU instead of the host cores
            valid = [int('boxes' in t and len(t['boxes']) > 0) for t in targets]
            if self.transform_on_device:
                images = [img.to(self.device, non_blocking=True) for img in images]
            transformed_images, _ = self.maskrcnn_model.transform(images)
            return (0., self._forward_eval_sequence(transformed_images, targets, valid, original_image_sizes))
        transformed_images, _ = self.maskrcnn_model.transform(images)

        valid = [int('boxes' in targets[i] and len(targets[i]['boxes']) > 0) for i in range(len(transformed_images.tensors))]
        valid_imgs = [images[i] for i, v in enumerate(valid) if v]
        valid_targets = [targets[i] for i, v in enumerate(valid) if v]
        _, t_targets = self.maskrcnn_model.transform(valid_imgs, valid_targets)
        it = iter(t_targets)
        targets = [next(it) if v else {} for v in valid]
        images = transformed_images

        total_loss = 0.
        all_detections = []
        count = 0
        half_lo, half_hi = floor(self.fast_pathway_size / 2), ceil(self.fast_pathway_size / 2)
        centre = self.fast_pathway_size // 2
        for idx, ok in enumerate(valid):
            if not ok:
                continue
            indices = range(idx - half_lo, idx + half_hi)
            with torch.no_grad():
                window = self.compute_maskrcnn_features(transformed_images.tensors, indices)
            centre_feats = self._index_features(window, centre, centre + 1)
            target = self._targets_to_device(targets[idx:idx + 1], self.device)
            with torch.no_grad():
                rpn_proposals, proposal_losses = self.compute_rpn_proposals(
                    transformed_images.tensors[idx:idx + 1], transformed_images.image_sizes[idx:idx + 1], centre_feats, target)
            target[0]['proposals'] = rpn_proposals[0]
            slow_valid = [self._slice_features(window, centre, self.slow_pathway_size)]
            fast_valid = [window]
            slow_fast_features = self.slow_fast.temporally_enhance_features(slow_valid, fast_valid)
            batch_original_sizes = original_image_sizes[idx:idx + 1]
            batch_image_sizes = images.image_sizes[0:1] * len(batch_original_sizes)
            proposals = [t['proposals'] for t in target]
            detections, detector_losses = self.maskrcnn_model.roi_heads(slow_fast_features, proposals, batch_image_sizes, target)
            detections = self.maskrcnn_model.transform.postprocess(detections, batch_image_sizes, batch_original_sizes)
            all_detections.extend(self._targets_to_device(detections, torch.device('cpu')))
            if self.training:
                losses = sum(list(detector_losses.values()) + list(proposal_losses.values()))
                total_loss += losses.item()
                losses.backward()
                count += 1
                if count % 2 == 0:
                    optimizer.step()
                    optimizer.zero_grad()
        if not self.training:
            it = iter(all_detections)
            all_detections = [next(it) if v else {} for v in valid]
        return (total_loss, all_detections)
