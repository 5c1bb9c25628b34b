"""Synthetic DAVIS-shaped data ON DISK for the caller-level tests: the directory layout, file formats and the
``(imgs, targets, seq_name)`` item contract of the reference's DAVISDataset (code/helpers/dataset.py:16-137) -- JPEG frames,
palette-PNG annotations whose pixel value is the object id, ImageSets/2017/<subset>.txt -- with moving rectangles as objects.
``DavisLikeDataset`` re-implements that contract independently (the GPU box has no reference tree); the CPU test feeds the same
directory to the reference's own DAVISDataset."""
import os

import numpy as np
import torch
from PIL import Image


def write_davis_tree(root, sequences=(("seq_a", 4), ("seq_b", 3)), hw=(120, 160), subsets=("train", "val"), empty_frame=None):
    """root/JPEGImages/480p/<seq>/%05d.jpg, root/Annotations/480p/<seq>/%05d.png, root/ImageSets/2017/<subset>.txt.
    ``empty_frame=(seq, idx)``: that frame has no object (the reference skips such frames, model.py:289-296)."""
    h, w = hw
    rng = np.random.RandomState(0)
    for name, n in sequences:
        os.makedirs(os.path.join(root, "JPEGImages", "480p", name), exist_ok=True)
        os.makedirs(os.path.join(root, "Annotations", "480p", name), exist_ok=True)
        for i in range(n):
            img = (rng.rand(h, w, 3) * 255).astype(np.uint8)
            ann = np.zeros((h, w), dtype=np.uint8)
            if empty_frame != (name, i):
                x1, y1 = 20 + 10 * i, 30 + 5 * i
                ann[y1:y1 + 50, x1:x1 + 60] = 1
                img[y1:y1 + 50, x1:x1 + 60] //= 2
            Image.fromarray(img).save(os.path.join(root, "JPEGImages", "480p", name, f"{i:05d}.jpg"))
            Image.fromarray(ann, mode="P").save(os.path.join(root, "Annotations", "480p", name, f"{i:05d}.png"))
    os.makedirs(os.path.join(root, "ImageSets", "2017"), exist_ok=True)
    for s in subsets:
        with open(os.path.join(root, "ImageSets", "2017", f"{s}.txt"), "w") as f:
            f.write("\n".join(name for name, _ in sequences) + "\n")


class DavisLikeDataset(torch.utils.data.Dataset):
    """Item = (list of [3,H,W] float frames, tuple of target dicts (empty dict for a frame without objects), sequence name)."""

    def __init__(self, root, subset="val"):
        with open(os.path.join(root, "ImageSets", "2017", f"{subset}.txt")) as f:
            self.names = [x.strip() for x in f if x.strip()]
        self.root = root

    def __len__(self):
        return len(self.names)

    def __getitem__(self, idx):
        name = self.names[idx]
        img_dir = os.path.join(self.root, "JPEGImages", "480p", name)
        ann_dir = os.path.join(self.root, "Annotations", "480p", name)
        imgs, targets = [], []
        for i, fn in enumerate(sorted(os.listdir(img_dir))):
            img = np.array(Image.open(os.path.join(img_dir, fn)))
            ann = np.array(Image.open(os.path.join(ann_dir, fn.replace(".jpg", ".png"))))
            imgs.append(torch.from_numpy(img).permute(2, 0, 1).float() / 255.0)
            boxes, masks = [], []
            for obj in np.unique(ann)[1:]:
                m = ann == obj
                ys, xs = np.where(m)
                if xs.min() < xs.max() and ys.min() < ys.max():
                    boxes.append([xs.min(), ys.min(), xs.max(), ys.max()])
                    masks.append(m)
            if not boxes:
                targets.append({})
                continue
            b = torch.as_tensor(boxes, dtype=torch.float32)
            targets.append({"boxes": b, "labels": torch.ones(len(b), dtype=torch.int64),
                            "masks": torch.as_tensor(np.stack(masks), dtype=torch.uint8),
                            "image_id": torch.tensor([1000 * idx + i]), "area": (b[:, 3] - b[:, 1]) * (b[:, 2] - b[:, 0]),
                            "iscrowd": torch.zeros(len(b), dtype=torch.int64)})
        return imgs, tuple(targets), name
