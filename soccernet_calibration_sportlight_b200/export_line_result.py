"""Line-model export: images -> the pickle ``CameraCreator(lines_file=...)`` reads.

Same inputs, outputs and defaults as the reference's ``src/utils/export_line_result.py:134-201``
(``cv2.imread`` + ``ToTensor`` scaling, ``model.nn_module(x)[-1]``, two-peak decode with ``--sigma``,
``get_line_data`` with ``--scale`` / ``--prob-thre``, ``{filename: {'lines': ..., 'points': ...}}``
pickled to ``--result-file``); the per-image loop of the reference becomes batches, the decode is the
CUDA kernel and the OpenCV preview window is left out.

    python -m soccernet_calibration_sportlight_b200.export_line_result --model line_ckpt.pth \\
        --image-folder frames/ --result-file results/lines.pkl [--sigma 3] [--prob-thre 0] [--scale 4]
"""
from __future__ import annotations

import argparse
import os
import pickle
from typing import Dict, List, Optional

import numpy as np
import torch

from .make_submit import frames_to_tensor
from .metamodel import HRNetMetaModel, load_model
from .transforms import EHMPredictionTransform, get_line_data


def export(model: HRNetMetaModel, image_folder: str, sigma: float = 3.0, prob_thre: float = 0.0, scale: int = 4,
           batch_size: int = 32, quiet: bool = True) -> Dict[str, dict]:
    """{file name: {'lines': {class: (slope, intercept)}, 'points': {class: [(x, y, p), ...]}}} for every
    ``*.jpg`` of the folder (export_line_result.py:171-188)."""
    import cv2
    names = sorted(n for n in os.listdir(image_folder) if n.endswith(".jpg"))
    out: Dict[str, dict] = {}
    decode = EHMPredictionTransform.mask_heat_points_gauss
    for i in range(0, len(names), batch_size):
        chunk = names[i:i + batch_size]
        images = [cv2.imread(os.path.join(image_folder, n), cv2.IMREAD_COLOR) for n in chunk]
        with torch.no_grad():
            heat = model.nn_module(frames_to_tensor(images).to(model.device))[-1]       # (B, 23, H/4, W/4)
            peaks = decode(heat, sigma=sigma).cpu().numpy()                              # (B, 23, 2, 3)
        for b, name in enumerate(chunk):
            lines, points = get_line_data(peaks, scale=scale, prob_thre=prob_thre, frame=b)
            out[name] = {"lines": lines, "points": points}
            if not quiet:
                print(f"{i + b}: {name}: {len(lines)} lines")
    return out


def write(result: Dict[str, dict], result_file: str) -> None:
    """export_line_result.py:165-167, 200-201."""
    d = os.path.dirname(result_file)
    if d and not os.path.exists(d):
        os.makedirs(d)
    with open(result_file, "wb") as f:
        pickle.dump(result, f)


def main(argv: Optional[List[str]] = None) -> Dict[str, dict]:
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--image-folder", default="/workdir/data/dataset/test")
    ap.add_argument("--result-file", default="results/result_on_test_set.pkl")
    ap.add_argument("--model", required=True, help="argus checkpoint of the line model")
    ap.add_argument("--device", default="cuda:0")
    ap.add_argument("--sigma", type=float, default=3.0)
    ap.add_argument("--prob-thre", type=float, default=0.0)
    ap.add_argument("--scale", type=int, default=4)
    ap.add_argument("--batch-size", type=int, default=32)
    ap.add_argument("--no-vis", action="store_true", help="accepted for compatibility; there is no preview window")
    a = ap.parse_args(argv)
    model = load_model(a.model, device=a.device)
    result = export(model, a.image_folder, a.sigma, a.prob_thre, a.scale, a.batch_size, quiet=False)
    write(result, a.result_file)
    return result


if __name__ == "__main__":
    main()
