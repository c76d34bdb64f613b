"""Training step of the reference (C:780-786 / U:342-348) as a reusable, sync-free callable.

    optimizer.zero_grad(); out = model(x); loss = BCELoss(out, y); loss.backward(); optimizer.step()

Precision policy (the reference is fp32-only): ``precision="bf16"`` keeps fp32 master
weights and runs the forward under ``torch.autocast(bfloat16)`` — GEMMs, convolutions and the csb200
attention / SimAM kernels see bf16 activations, LayerNorm / softmax statistics stay fp32 — while the
sigmoid and the BCE loss are evaluated in fp32 outside the autocast region (CUDA autocast refuses
BCELoss on probabilities).  No ``.item()`` is called here: the reference's five host syncs per step
(C:797-806) are the caller's choice, not the step's.
"""
import contextlib
from typing import Optional, Tuple

import torch
import torch.nn.functional as F


def synthetic_batch(batch: int, size: int, device, seed: int = 0, first_index: int = 0,
                    pin: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """Images ~ U[0,1) (B,3,S,S) and binary masks (B,1,S,S), as BASELINE.md §4 specifies.

    Sample i of the GLOBAL batch is generated from (seed, first_index + i) alone, so an n-GPU run and
    a 1-GPU run see the same global batch however it is sharded.
    """
    imgs = torch.empty((batch, 3, size, size), dtype=torch.float32, pin_memory=pin)
    masks = torch.empty((batch, 1, size, size), dtype=torch.float32, pin_memory=pin)
    for i in range(batch):
        g = torch.Generator().manual_seed(seed * 1_000_003 + first_index + i)
        imgs[i] = torch.rand((3, size, size), generator=g)
        masks[i] = (torch.rand((1, size, size), generator=g) > 0.5).float()
    if torch.device(device).type == "cpu":
        return imgs, masks
    return imgs.to(device, non_blocking=pin), masks.to(device, non_blocking=pin)


def bce_from_logits_as_probabilities(probs: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.BCELoss() on sigmoid outputs, fp32 (C:936, C:782)."""
    return F.binary_cross_entropy(probs.float(), target.float())


class TrainStep:
    """One optimisation step; returns the loss as a device tensor (no host sync)."""

    def __init__(self, model: torch.nn.Module, optimizer: torch.optim.Optimizer, precision: str = "bf16",
                 reducer=None):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.model, self.optimizer, self.precision, self.reducer = model, optimizer, precision, reducer

    def _autocast(self, device_type: str):
        if self.precision == "bf16":
            return torch.autocast(device_type=device_type, dtype=torch.bfloat16)
        return contextlib.nullcontext()

    def forward_loss(self, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        with self._autocast(images.device.type):
            probs = self.model(images)
        return bce_from_logits_as_probabilities(probs, masks)

    def __call__(self, images: torch.Tensor, masks: torch.Tensor) -> torch.Tensor:
        if self.reducer is not None:
            self.reducer.begin_step()  # zeroes the flat gradient buckets (== zero_grad)
        else:
            self.optimizer.zero_grad(set_to_none=True)
        loss = self.forward_loss(images, masks)
        loss.backward()
        if self.reducer is not None:
            self.reducer.finish_step()  # wait for the bucketed all-reduces
        self.optimizer.step()
        return loss.detach()
