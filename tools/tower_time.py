"""Where a step's device time goes at a given per-GPU batch: each frozen tower alone (CUDA-graph replay), both concurrently, and
the trainable tail (weighted sum -> branch -> loss -> backward -> Adam) with the towers precomputed.

    python tools/tower_time.py --batch 32 [--config base]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from avssl.base import OrderedNamespace  # noqa: E402
from avssl.model import KWClip_GeneralTransformer  # noqa: E402
from speechclip_b200 import engine  # noqa: E402
from speechclip_b200.configs import parallel_config  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--config", default="base")
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda", 0)
engine.GRAPHS = True
model = KWClip_GeneralTransformer(OrderedNamespace(parallel_config(args.config))).to(dev).train()
opt = model.configure_optimizers()[0][0]
B = args.batch
g = torch.Generator().manual_seed(0)
S = 224
batch = {"wav": (0.1 * torch.randn(B, 102400, generator=g)).to(dev), "wav_len": torch.full((B,), 102400).to(dev),
         "image": torch.randn(B, 3, S, S, generator=g).to(dev), "id": torch.arange(B).to(dev)}
side = torch.cuda.Stream()


def audio():
    return model.audio_encoder.encode_frozen(batch["wav"], batch["wav_len"])


def image():
    return model.forward_image(batch["image"])


def both():
    cur = torch.cuda.current_stream()
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        image()
    audio()
    cur.wait_stream(side)


def tail(pre):
    b = dict(batch)
    b["_scb_towers"] = pre
    loss = model.training_step_end(model.training_step(b))["loss"]
    opt.zero_grad()
    loss.backward()
    model.on_after_backward()
    opt.step()


def timeit(fn, reps=args.reps):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


print(f"config {args.config}, {B} pairs: ms per call (CUDA events, graph replay of the towers)")
print(f"  speech tower alone      {timeit(audio):8.3f}")
print(f"  image tower alone       {timeit(image):8.3f}")
print(f"  both towers, 2 streams  {timeit(both):8.3f}")
pre = model.precompute_towers(batch, slot=1)
torch.cuda.synchronize()
print(f"  trainable tail alone    {timeit(lambda: tail(pre)):8.3f}")


def step():
    loss = model.training_step_end(model.training_step(batch))["loss"]
    opt.zero_grad()
    loss.backward()
    model.on_after_backward()
    opt.step()


print(f"  whole step, unpipelined {timeit(step):8.3f}")
