"""Epoch-rate training callbacks with the reference's protocol (callbacks.py:9-153):
`on_epoch_begin(**kwargs) -> bool` / `on_epoch_end(**kwargs) -> bool` (True = stop), kwargs
`epoch, optimizer, device, model, logs`.  No compute happens here; kept so reference training scripts
run unchanged against the new models (state_dict wire format: torch.save(model.state_dict()) .pth).
"""
from __future__ import annotations

import abc
import os
from difflib import get_close_matches

import torch


class Callback(abc.ABC):
    @abc.abstractmethod
    def on_epoch_begin(self, **kwargs) -> bool:
        return False

    @abc.abstractmethod
    def on_epoch_end(self, **kwargs) -> bool:
        return False


class EarlyStopping(Callback):
    """callbacks.py:32-77.  Note: the reference looks up the key "val_loss", which the models
    never emit ("Loss/val_loss"), so it never triggers; that behaviour is preserved via `metric_name`."""

    def __init__(self, patience: int = 10, delta: float = 0) -> None:
        self.patience, self.delta = patience, delta
        self.counter = 0
        self.best_loss = float("inf")
        self.best_epoch = 0
        self.metric_name = "val_loss"

    def on_epoch_begin(self, **kwargs) -> bool:
        return False

    def on_epoch_end(self, **kwargs) -> bool:
        val_loss = kwargs.get("logs", {}).get(self.metric_name, float("inf"))
        if val_loss < self.best_loss - self.delta:
            self.best_loss, self.counter = val_loss, 0
        elif val_loss > self.best_loss + self.delta:
            self.counter += 1
            return self.counter >= self.patience
        return False


class ModelCheckpoint(Callback):
    """callbacks.py:80-153: saves `model.state_dict()` to <save_path>/<slurm_job_id>.pth on improvement of the
    monitored metric (fuzzy key match on the first epoch, :117-130)."""

    def __init__(self, slurm_job_id: str, save_path: str, monitor: str = "val_loss", mode: str = "min",
                 save_best_only: bool = True) -> None:
        self.slurm_job_id, self.save_path, self.monitor, self.mode = slurm_job_id, save_path, monitor, mode
        self.save_best_only = save_best_only
        self.best_metric = float("inf") if mode == "min" else float("-inf")
        self.best_epoch = 0

    def on_epoch_begin(self, **kwargs) -> bool:
        return False

    def _save(self, model, name):
        if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_rank() != 0:
            return                                  # data parallel: the replicas are identical, rank 0 writes the file
        os.makedirs(self.save_path, exist_ok=True)
        torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, os.path.join(self.save_path, name))

    def on_epoch_end(self, **kwargs) -> bool:
        logs = kwargs.get("logs", {})
        epoch = kwargs.get("epoch", 0)
        model = kwargs.get("model")
        if epoch == 1 and self.monitor not in logs:
            close = get_close_matches(self.monitor, logs.keys(), n=1, cutoff=0)
            if not close:
                raise ValueError(f"Monitor metric '{self.monitor}' not found in logs. Available metrics: {list(logs.keys())}")
            self.monitor = close[0]
        current = logs.get(self.monitor, float("inf"))
        if not self.save_best_only:
            self._save(model, f"{self.slurm_job_id}_epoch_{epoch}.pth")
        elif (self.mode == "min" and current < self.best_metric) or (self.mode == "max" and current > self.best_metric):
            self.best_metric, self.best_epoch = current, epoch
            self._save(model, f"{self.slurm_job_id}.pth")
        return False
