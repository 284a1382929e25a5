"""Lightning-free mirror of the reference's training wrapper (``/root/reference/lit.py:18-72``).

`lightning` / `torchmetrics` are not installable here, and the north star keeps the Lightning
``training_step`` contract unchanged rather than re-building Lightning: this class has the same
method names, return values, loss (``nn.MSELoss``, lit.py:24), metrics (MAE / MAPE / RMSE,
lit.py:36-38), optimizer (Adam 1e-3, lit.py:60) and scheduler (CosineAnnealingLR T_max=10,
lit.py:61), so it can be pasted under ``L.LightningModule`` unchanged.  The model wrapped is the
gwnet block itself (the UNet encoder/decoder stay upstream feature producers)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn


class _MSELoss(nn.Module):
    """``nn.MSELoss()`` (lit.py:24): one forward and one backward launch on fp32 CUDA tensors (`ops.mse_loss`), the stock
    functional otherwise."""

    def forward(self, input, target):
        from . import ops
        return ops.mse_loss(input, target)


class LitGWNet(nn.Module):
    def __init__(self, model: nn.Module):
        super().__init__()
        self.model = model
        self.loss_fn = _MSELoss()
        self.logged: Dict[str, float] = {}

    @property
    def device(self):
        return next(self.model.parameters()).device

    def log(self, name, value, prog_bar: bool = False):
        self.logged[name] = value

    def _metrics(self, yhat, y):
        with torch.no_grad():
            d = yhat - y
            mae = d.abs().mean()
            mape = (d.abs() / y.abs().clamp_min(1.17e-06)).mean()
            rmse = torch.sqrt((d * d).mean())
        return mae, mape, rmse

    def training_step(self, batch, batch_idx: int = 0):
        x, y = batch[0].to(self.device), batch[1].to(self.device)
        yhat = self.model(x)
        loss = self.loss_fn(yhat, y)
        mae, mape, rmse = self._metrics(yhat, y)
        self.log('train_loss', loss, prog_bar=True)
        self.log('train_mae', mae)
        self.log('train_mape', mape)
        self.log('train_rmse', rmse)
        return loss

    def validation_step(self, batch, batch_idx: int = 0):
        x, y = batch[0].to(self.device), batch[1].to(self.device)
        yhat = self.model(x)
        loss = self.loss_fn(yhat, y)
        mae, mape, rmse = self._metrics(yhat, y)
        self.log('val_loss', loss, prog_bar=True)
        self.log('val_mae', mae)
        self.log('val_mape', mape)
        self.log('val_rmse', rmse)
        return loss

    def configure_optimizers(self):
        optimizer = torch.optim.Adam(self.parameters(), lr=1e-3)
        scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(optimizer, T_max=10)
        return {'optimizer': optimizer,
                'lr_scheduler': {'scheduler': scheduler, 'monitor': 'val_loss', 'interval': 'epoch',
                                 'frequency': 1}}
