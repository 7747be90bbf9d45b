"""BaseModel — plain (non nn.Module) base class; API mirror of the reference's model/base_model.py:31-127
for the members the shape branch touches (initialize, set_requires_grad, tocuda, eval/train switches)."""
from __future__ import annotations

import torch


class BaseModel:
    def name(self):
        return "BaseModel"

    def initialize(self, opt):
        self.opt = opt
        hyper = getattr(opt, "hyper", None)
        self.gpu_ids = getattr(hyper, "gpu_ids", 0) if hyper is not None else 0
        self.isTrain = getattr(hyper, "isTrain", False) if hyper is not None else False
        self.model_names = []
        self.epoch_labels = []
        self.optimizers = []

    def set_input(self, input):
        self.input = input

    def forward(self):
        pass

    def set_requires_grad(self, nets, requires_grad=False):
        if not isinstance(nets, list):
            nets = [nets]
        for net in nets:
            if net is not None:
                for param in net.parameters():
                    param.requires_grad = requires_grad

    def tocuda(self, var_names):
        for name in var_names:
            if isinstance(name, str):
                var = getattr(self, name)
                setattr(self, name, var.to(self.device, non_blocking=True))

    def get_current_errors(self):
        return {}

    def update_learning_rate(self):
        for scheduler in getattr(self, "schedulers", []):
            scheduler.step()
