"""LR schedules (reference: avssl/optim/scheduler.py:11-47) as LambdaLR multipliers, stepped once per optimizer step."""
from torch.optim import Optimizer
from torch.optim.lr_scheduler import LambdaLR


def _base_lr(optimizer: Optimizer) -> float:
    return optimizer.param_groups[0]["lr"]


def noam_scheduler(optimizer: Optimizer, warmup: int = 4000, last_epoch: int = -1) -> LambdaLR:
    return LambdaLR(optimizer, lambda step: (step + 1) / warmup if step < warmup else (warmup / (step + 1)) ** 0.5, last_epoch)


def linear_warmup_decay_scheduler(optimizer: Optimizer, warmup: int = 4000, max_step: int = 1000000, final_lr: float = 1e-8) -> LambdaLR:
    floor = final_lr / _base_lr(optimizer)

    def multiplier(step: int) -> float:
        if step < warmup:
            return (step + 1) / warmup
        return 1.0 - (1.0 - floor) * (step + 1 - warmup) / (max_step - warmup)

    return LambdaLR(optimizer, multiplier)


def get_scheduler(name: str, optimizer: Optimizer, **kwargs) -> LambdaLR:
    if name == "noam":
        return noam_scheduler(optimizer, **kwargs)
    if name == "linear_warmup_decay":
        return linear_warmup_decay_scheduler(optimizer, **kwargs)
    raise NotImplementedError(f"Unknown lr scheduler {name}")
