"""``HRNetMetaModel`` mirror: the reference's argus ``Model`` subclass reduced to what the
inference path uses (src/models/hrnet/metamodel.py:14-22, 88-134; line variant
src/models/line/metamodel.py): construction from the ``params`` dict, the checkpoint
format ``{'model_name','params','nn_state_dict'}`` and ``predict``.

    model = load_model(path, device='cuda:0')       # argus.load_model equivalent
    preds = model.predict(frames)                   # (B,3,H,W) fp32 [0,1] BGR -> (B,57,3)
"""
from __future__ import annotations

from typing import Optional

import torch

from .hrnet import Config, HRNetHeatmap, w48_config
from .transforms import EHMPredictionTransform, HRNetPredictionTransform


class HRNetMetaModel:
    nn_module_cls = HRNetHeatmap
    prediction_transform_cls = HRNetPredictionTransform
    kind = "keypoints"

    def __init__(self, params: dict):
        self.params = params
        nn_params = dict(params.get("nn_module", {}))
        cfg = nn_params.pop("hrnet_config", None) or w48_config(self.kind)
        self.nn_module = self.nn_module_cls(Config(dict(cfg)), kind=self.kind, **nn_params)
        pt = params.get("prediction_transform", None)
        self.prediction_transform = self.prediction_transform_cls(**pt) if pt is not None else None
        self.device: Optional[torch.device] = None
        dev = params.get("device", None)
        if dev is not None:
            self.set_device(dev)

    def set_device(self, device):
        self.device = torch.device(device)
        self.nn_module.to(self.device)
        return self

    def eval(self):
        self.nn_module.eval()
        return self

    def get_nn_module(self):
        return self.nn_module

    def _check_predict_ready(self):
        if self.nn_module is None or self.prediction_transform is None or self.device is None:
            raise AttributeError("predict needs nn_module, prediction_transform and a CUDA device")

    def predict(self, x: torch.Tensor) -> torch.Tensor:
        """metamodel.py:127-134: eval, move to device, forward, transform the last output."""
        self._check_predict_ready()
        with torch.no_grad():
            self.eval()
            x = x.to(self.device, non_blocking=True)
            prediction = self.nn_module(x)
            return self.prediction_transform(prediction[-1])

    def save(self, file_path: str, optimizer_state: bool = False):
        """metamodel.py:88-125 (no optimizer on this path)."""
        state = {"model_name": self.__class__.__name__, "params": self.params,
                 "nn_state_dict": {k.replace("_orig_mod.", ""): v.cpu()
                                   for k, v in self.nn_module.state_dict().items()}}
        torch.save(state, file_path)


class LineMetaModel(HRNetMetaModel):
    """src/models/line/metamodel.py: same wrapper around the 23-channel line network."""
    prediction_transform_cls = EHMPredictionTransform
    kind = "lines"


_MODELS = {"HRNetMetaModel": HRNetMetaModel, "LineMetaModel": LineMetaModel, "EHMMetaModel": LineMetaModel}


def load_model(file_path: str, device: Optional[str] = None, **_ignored):
    """``argus.load_model(path, loss=None, optimizer=None, device=...)`` for the checkpoints
    written by the reference (make_submit.py:51)."""
    state = torch.load(file_path, map_location="cpu", weights_only=False)
    cls = _MODELS.get(state.get("model_name", "HRNetMetaModel"), HRNetMetaModel)
    params = dict(state["params"])
    params.pop("device", None)
    model = cls(params)
    model.nn_module.load_state_dict(state["nn_state_dict"])
    if device is not None:
        model.set_device(device)
    return model
