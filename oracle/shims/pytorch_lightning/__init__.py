"""ORACLE shim: the reference only uses pl.LightningModule as a base class (:1175)."""
import torch

LightningModule = torch.nn.Module
