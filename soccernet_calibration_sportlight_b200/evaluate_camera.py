"""The official camera-calibration metric on the GPU, behind the reference's names.

* ``get_polylines`` / ``distance_to_polyline`` / ``evaluate_camera_prediction`` - baseline/evaluate_camera.py:14-229
* ``mirror_labels`` - baseline/evaluate_extremities.py:24-34
* ``Evaluator`` / ``EvalAImetric`` - src/models/hrnet/metrics.py:87-230 (the argus ``Metric`` plumbing aside)

The reference evaluates one frame per call in a 16-process CPU pool (metrics.py:166, 186-188); here a batch of
camera records (``CameraCreator.batch_records``) and the batch's packed annotations go through ONE launch of
``evaluate_kernel`` (csrc/evaluate.cu, one thread block per frame): project the sampled pitch model, clip at
the image border, point-to-polyline distances for the annotated and the mirrored labelling, confusion
matrices, choice of the labelling.  No CPU fallback: the functions need the CUDA library and a device.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib, pitch

NC = 28                     # CAL_EVAL_CLASSES: the dataset's line classes (soccerpitch.py:15-44)
CLASS_INDEX = {n: i for i, n in enumerate(pitch.LINES_CLASSES)}


class EvalRecord(C.Structure):
    _fields_ = [("accuracy", C.c_double), ("confusion", C.c_double * 4), ("l2_sum", C.c_double), ("l2_count", C.c_int32),
                ("labelling", C.c_int32), ("valid", C.c_int32), ("pad", C.c_int32), ("per_class", (C.c_double * 4) * NC),
                ("touched", C.c_uint8 * NC), ("pad2", C.c_uint8 * 4)]


REC_BYTES = C.sizeof(EvalRecord)
_REC_DTYPE = np.dtype([("accuracy", "<f8"), ("confusion", "<f8", (4,)), ("l2_sum", "<f8"), ("l2_count", "<i4"), ("labelling", "<i4"),
                       ("valid", "<i4"), ("pad", "<i4"), ("per_class", "<f8", (NC, 4)), ("touched", "u1", (NC,)), ("pad2", "u1", (4,))])
assert _REC_DTYPE.itemsize == REC_BYTES


class _Tables:
    """Device copies of the sampled pitch model for one (device, sampling factor)."""
    _cache: Dict[Tuple[str, float], "_Tables"] = {}

    def __init__(self, device, sampling: float):
        polys = pitch.sample_field_points(sampling)
        self.names = list(polys)
        off = np.cumsum([0] + [len(polys[k]) for k in self.names]).astype(np.int32)
        self.max_poly = 2 * int(max(len(v) for v in polys.values())) + 2
        self.field = torch.from_numpy(np.concatenate([polys[k] for k in self.names], axis=0)).to(device)
        self.class_off = torch.from_numpy(off).to(device)
        self.class_id = torch.tensor([CLASS_INDEX[k] for k in self.names], dtype=torch.int32, device=device)
        self.mirror = torch.tensor([CLASS_INDEX[pitch.symmetric_class(k)] for k in pitch.LINES_CLASSES], dtype=torch.int32, device=device)
        self.is_circle = torch.tensor([1 if "Circle" in k else 0 for k in pitch.LINES_CLASSES], dtype=torch.uint8, device=device)

    @classmethod
    def get(cls, device, sampling: float) -> "_Tables":
        key = (str(device), float(sampling))
        if key not in cls._cache:
            cls._cache[key] = cls(device, sampling)
        return cls._cache[key]


def pack_annotations(annots: Sequence[Dict[str, List[Dict[str, float]]]], device) -> Tuple[torch.Tensor, torch.Tensor]:
    """List of annotation dicts {class: [{'x':, 'y':}, ...]} (pixels) -> (B, 28, max_gt, 2) fp64 points and
    (B, 28) int32 counts (-1 = class absent) on the device."""
    B = len(annots)
    max_gt = max([1] + [len(v) for a in annots for v in a.values()])
    pts = np.zeros((B, NC, max_gt, 2))
    cnt = np.full((B, NC), -1, dtype=np.int32)
    for b, a in enumerate(annots):
        for k, v in a.items():
            c = CLASS_INDEX[k]
            cnt[b, c] = len(v)
            for j, p in enumerate(v):
                pts[b, c, j] = (p["x"], p["y"])
    return torch.from_numpy(pts).to(device), torch.from_numpy(cnt).to(device)


def evaluate_records(records: torch.Tensor, gt_pts: torch.Tensor, gt_count: torch.Tensor, threshold: float = 5,
                     img_size: Tuple[int, int] = (960, 540), sampling_factor: float = 0.9,
                     polylines: Optional[Tuple[torch.Tensor, torch.Tensor]] = None, as_annotated: bool = False):
    """(B,16) fp64 camera records on the GPU + packed annotations -> structured numpy array of per-frame results
    (accuracy, confusion, per_class, l2_sum, l2_count, labelling, valid) and the polylines (device tensors)."""
    if not records.is_cuda:
        raise _lib.CalError("evaluate_records: tensors must live on a CUDA device (no CPU fallback)")
    dev = records.device
    B = records.shape[0]
    T = _Tables.get(dev, sampling_factor)
    n_proj = len(T.names)
    max_gt = gt_pts.shape[2]
    if polylines is None:
        poly = torch.empty((B, n_proj, T.max_poly, 2), dtype=torch.float64, device=dev)
        pcnt = torch.zeros((B, n_proj), dtype=torch.int32, device=dev)
        from_poly, max_poly = 0, T.max_poly
    else:
        poly, pcnt = polylines
        from_poly, max_poly = 1, poly.shape[2]
    if as_annotated:
        from_poly |= 2
    dist = torch.zeros((B, 2, NC, max_gt), dtype=torch.float64, device=dev)
    out = torch.zeros((B, REC_BYTES), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        st = _lib.lib().cal_evaluate_cameras(
            records.contiguous().data_ptr(), B, T.field.data_ptr(), T.class_off.data_ptr(), T.class_id.data_ptr(), n_proj,
            T.mirror.data_ptr(), T.is_circle.data_ptr(), gt_pts.contiguous().data_ptr(), gt_count.contiguous().data_ptr(), max_gt,
            int(img_size[0]), int(img_size[1]), float(threshold), poly.data_ptr(), pcnt.data_ptr(), max_poly, from_poly,
            dist.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    _lib.check(st, "cal_evaluate_cameras")
    return out.cpu().numpy().view(_REC_DTYPE).reshape(B), (poly, pcnt, T.names)


def _record_of(cam) -> np.ndarray:
    r = np.zeros(16)
    r[0:3], r[3:12], r[12], r[13] = np.asarray(cam.position), np.asarray(cam.rotation).reshape(9), cam.xfocal_length, cam.yfocal_length
    r.view(np.int32)[30] = 1
    return r


def get_polylines(camera_annotation, width, height, sampling_factor=0.2):
    """evaluate_camera.py:14-107: {class: [{'x':, 'y':}, ...]} of the pitch lines the camera sees."""
    from .camera import Camera
    cam = camera_annotation
    if not isinstance(cam, Camera):
        cam = Camera(width, height)
        cam.from_json_parameters(camera_annotation)
    dev = torch.device(getattr(cam, "device", "cuda:0"))
    rec = torch.from_numpy(_record_of(cam)[None]).to(dev)
    gp, gc = pack_annotations([{}], dev)
    _, (poly, pcnt, names) = evaluate_records(rec, gp, gc, 5, (width, height), sampling_factor)
    poly, pcnt = poly.cpu().numpy()[0], pcnt.cpu().numpy()[0]
    return {k: [{"x": float(x), "y": float(y)} for x, y in poly[s, :pcnt[s]]] for s, k in enumerate(names) if pcnt[s] > 0}


def mirror_labels(lines_dict):
    """evaluate_extremities.py:24-34."""
    return {pitch.symmetric_class(k): v for k, v in lines_dict.items()}


def _result_tuple(r):
    per_class = {pitch.LINES_CLASSES[c]: r["per_class"][c].reshape(2, 2).copy() for c in range(NC) if r["touched"][c]}
    return float(r["accuracy"]), r["confusion"].reshape(2, 2).astype(np.float32), per_class


def evaluate_camera_prediction(projected_lines, groundtruth_lines, threshold):
    """evaluate_camera.py:163-229 on given polylines: (global confusion, per-class confusions, per-class distances).
    The kernel evaluates both labellings at once; this entry point reports the one as annotated; the distances
    come back as their sum and count (what EvalAImetric accumulates)."""
    dev = torch.device("cuda:0")
    T = _Tables.get(dev, 0.9)
    names = T.names
    max_poly = max([2] + [len(v) for v in projected_lines.values()])
    poly = np.zeros((1, len(names), max_poly, 2))
    pcnt = np.zeros((1, len(names)), dtype=np.int32)
    for s, k in enumerate(names):
        v = projected_lines.get(k, [])
        pcnt[0, s] = len(v)
        for j, p in enumerate(v):
            poly[0, s, j] = (p["x"], p["y"])
    gp, gc = pack_annotations([groundtruth_lines], dev)
    rec = torch.zeros((1, 16), dtype=torch.float64, device=dev)
    res, _ = evaluate_records(rec, gp, gc, threshold, polylines=(torch.from_numpy(poly).to(dev), torch.from_numpy(pcnt).to(dev)),
                              as_annotated=True)
    r = res[0]
    _, conf, per_class = _result_tuple(r)
    return conf, per_class, {"sum": float(r["l2_sum"]), "count": int(r["l2_count"])}


class Evaluator:
    """metrics.py:87-137: per-frame (accuracy, confusion, per-class confusions, reprojection errors) or None."""

    def __init__(self, pred2cam: Callable, threshold: int = 5, img_size: Tuple[int, int] = (960, 540)):
        self.pred2cam, self.threshold, self.img_size = pred2cam, threshold, img_size

    def batch(self, preds, annots: Sequence[dict], names: Optional[Sequence[Optional[str]]] = None):
        """(B,57,3) predictions + B annotation dicts -> list of per-frame tuples / None (one solve launch and one
        metric launch)."""
        dev = torch.device(getattr(self.pred2cam, "device", "cuda:0"))
        p = preds if isinstance(preds, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(preds, dtype=np.float32))
        rec = self.pred2cam.batch_records(p.to(dev))
        gp, gc = pack_annotations(annots, dev)
        res, _ = evaluate_records(rec, gp, gc, self.threshold, self.img_size, 0.9)
        out = []
        for r in res:
            if not r["valid"]:
                out.append(None)
                continue
            acc, conf, per_class = _result_tuple(r)
            out.append((acc, conf, per_class, {"sum": float(r["l2_sum"]), "count": int(r["l2_count"])}))
        return out

    def __call__(self, x):
        pred, annot, name = x
        return self.batch(np.asarray(pred)[None], [annot], [name])[0]


class EvalAImetric:
    """metrics.py:142-230 without the argus base class: update(step_output) / compute() / epoch metrics."""
    name = "evalai"
    better = "max"

    def __init__(self, pred2cam: Callable, threshold: int = 5, img_size: Tuple[int, int] = (960, 540), max_workers: int = 16):
        self.pred2cam, self.threshold, self.img_size = pred2cam, threshold, img_size
        self.evaluator = Evaluator(pred2cam, threshold, img_size)
        self.reset()

    def reset(self):
        self.total_frames = self.missed_frames = 0
        self.tp = self.recall = self.accuracy = self.l2_proj_sum = 0.0
        self.n_precision = self.n_recall = self.n_accuracy = self.n_l2_proj = 0
        self.per_class_confusion = defaultdict(lambda: np.zeros((2, 2)))

    def update(self, step_output: dict):
        preds = step_output["prediction"]
        n = preds.shape[0]
        self.total_frames += n
        for res in self.evaluator.batch(preds, step_output["raw_annots"], step_output.get("img_name")):
            if res is None:
                self.missed_frames += 1
                continue
            accuracy, confusion, per_class_conf, reproj = res
            self.accuracy += accuracy
            self.n_accuracy += 1
            self.tp += confusion[0, 0]
            self.n_precision += confusion[0, :].sum()
            self.n_recall += confusion[0, 0] + confusion[1, 0]
            for k, m in per_class_conf.items():
                self.per_class_confusion[k] += m
            self.n_l2_proj += reproj["count"]
            self.l2_proj_sum += reproj["sum"]

    def compute(self) -> float:
        return (self.total_frames - self.missed_frames) / self.total_frames if self.total_frames > 0 else 0.0

    def epoch_metrics(self, prefix: str = "") -> Dict[str, float]:
        """The numbers epoch_complete stores in state.metrics (metrics.py:213-229)."""
        completeness = self.compute()
        precision = self.tp / self.n_precision if self.n_precision > 0 else 0.0
        recall = self.tp / self.n_recall if self.n_recall > 0 else 0.0
        accuracy = self.accuracy / self.n_accuracy if self.n_accuracy > 0 else 0.0
        l2 = self.l2_proj_sum / self.n_l2_proj if self.n_l2_proj > 0 else float("inf")
        return {f"{prefix}l2_reprojection": l2, f"{prefix}completeness": completeness, f"{prefix}eval_precision": float(precision),
                f"{prefix}eval_recall": float(recall), f"{prefix}eval_accuracy": accuracy, f"{prefix}{self.name}": completeness * accuracy}
