"""Inference driver: images -> ``camera_<name>.json`` files in the challenge format.

Same inputs, outputs and constants as the reference's ``src/utils/make_submit.py:42-75``
(BGR ``cv2.imread`` + ``ToTensor`` scaling, ``model.predict``, ``CameraCreator`` with the
``iterative_voter`` settings, ``Camera.to_json_parameters`` dumped with indent 4, completeness
printed at the end).  The 16-process CPU pool of the reference (``:53, 69``) is replaced by
one batched camera-solve launch per image batch.

    python -m soccernet_calibration_sportlight_b200.make_submit --model ckpt.pth \\
        --img-dir frames/ --save-dir submit/ [--lines-file lines.pkl] [--batch-size 64]
"""
from __future__ import annotations

import argparse
import json
import os
from typing import List, Optional

import numpy as np
import torch

from .metamodel import HRNetMetaModel, load_model
from .pitch import PITCH_POINTS
from .prediction import MAKE_SUBMIT_KWARGS, CameraCreator


def frames_to_tensor(images: List[np.ndarray]) -> torch.Tensor:
    """``T.ToTensor`` on ``cv2.imread`` output (make_submit.py:66, transforms.py:59-68):
    HWC uint8 BGR -> CHW fp32 in [0,1]."""
    arr = np.stack(images, axis=0)
    return torch.from_numpy(np.ascontiguousarray(arr.transpose(0, 3, 1, 2))).float().div_(255.0)


def run(model: HRNetMetaModel, calibrator: CameraCreator, img_dir: str, save_dir: str, batch_size: int = 64,
        quiet: bool = False) -> float:
    """Processes every ``*.jpg`` of ``img_dir``; returns the completeness (make_submit.py:75)."""
    import cv2
    os.makedirs(save_dir, exist_ok=True)
    img_names = sorted(n for n in os.listdir(img_dir) if n.endswith(".jpg"))
    done = 0
    for i in range(0, len(img_names), batch_size):
        names = img_names[i:i + batch_size]
        images = [cv2.imread(os.path.join(img_dir, n)) for n in names]
        preds = model.predict(frames_to_tensor(images))
        cams = calibrator.batch(preds.cpu().numpy(), names)
        for name, cam in zip(names, cams):
            if cam is None:
                continue
            with open(os.path.join(save_dir, "camera_" + name.replace(".jpg", ".json")), "w") as f:
                json.dump(cam.to_json_parameters(), f, indent=4, default=float)
            done += 1
    completeness = done / max(len(img_names), 1)
    if not quiet:
        print(f"Completeness: {completeness:.2f}")
    return completeness


def main(argv: Optional[List[str]] = None) -> float:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--model", required=True, help="argus checkpoint of the keypoint model (make_submit.py:17)")
    ap.add_argument("--img-dir", required=True)
    ap.add_argument("--save-dir", required=True)
    ap.add_argument("--lines-file", default=None, help="pickle written by the line-model export (make_submit.py:21)")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--batch-size", type=int, default=64)
    a = ap.parse_args(argv)
    model = load_model(a.model, device=a.device)
    calibrator = CameraCreator(PITCH_POINTS, lines_file=a.lines_file, **MAKE_SUBMIT_KWARGS)
    calibrator.device = a.device
    return run(model, calibrator, a.img_dir, a.save_dir, a.batch_size)


if __name__ == "__main__":
    main()
