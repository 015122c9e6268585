#!/usr/bin/env python
"""Benchmark of the SpeechCLIP speech–image contrastive training step (BASELINE.json: Parallel SpeechCLIP-base, batch 256).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--scaling strong|weak]

Own arm: one process per GPU (torchrun for N > 1), the drop-in model ``avssl.model.KWClip_GeneralTransformer`` on the sm_100a
kernels.  A "step" = training_step (HuBERT + ViT towers, parallel branch) -> training_step_end (NCCL all-gather of the pooled
embeddings + masked InfoNCE over the global batch) -> backward -> gradient all-reduce -> fused clip+Adam -> LR schedule.
``value``  = pairs/s with the batch already resident in HBM; ``e2e`` = pairs/s through the same public API with HOST (pinned)
buffers: H2D of wav/image/ids and D2H of the loss inside the timed region.  ``roofline`` is measured live: every C-ABI call of the
timed steps is bracketed by CUDA events on the launching stream; the dominant kernel is the tcgen05 GEMM.

Reference arm (``--impl reference``): the reference is pure Python over fairseq / openai-CLIP, neither installable offline, so
its CPU implementation of the path is the torch fp32 oracle (oracle/, a restatement pinned against the reference's own torch-only
modules): a bounded sample (16 pairs per step) of the same training step on the host cores.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "speech-image pairs/sec (Parallel SpeechCLIP-base training step, Flickr8k-shape synthetic)"
N_SAMPLES = 102400      # max_audio_len crop = 6.4 s @ 16 kHz -> 319 frames (spchclp_p.yaml:104)
GLOBAL_BATCH = 256      # data.batch_size (spchclp_p.yaml:10)
GF_PER_PAIR_FWD = 106.3   # dense algorithmic GFLOP per pair, forward
GF_PER_PAIR_STEP = 116.0  # ... forward + branch backward


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return dict(tflops=float(d["bf16_tflops_sustained"]), tflops_burst=float(d["bf16_tflops"]), hbm=float(d["hbm_gbs"]), src="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback")


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i",
                                          str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:  # noqa: BLE001
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                continue
        sm.sort()
        busy = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def synth_batch(n, seed, pin):
    """Flickr8k-shape synthetic pairs: wav N(0, 0.1^2) x 102400 samples, image N(0,1) [3,224,224], unique ids."""
    g = torch.Generator().manual_seed(seed)
    wav = (0.1 * torch.randn(n, N_SAMPLES, generator=g))
    img = torch.randn(n, 3, 224, 224, generator=g)
    b = {"wav": wav, "wav_len": torch.full((n,), N_SAMPLES, dtype=torch.int64), "image": img, "id": torch.arange(n, dtype=torch.int64)}
    return {k: (v.pin_memory() if pin else v) for k, v in b.items()}


# =============================================================================================================== reference arm
def cpu_step_fn(pairs):
    """The oracle's training step on the host cores: forward of both towers + branch, masked InfoNCE, backward, clip, Adam."""
    from oracle import clip as oc
    from oracle import hubert as oh
    from oracle import speechclip as osc
    from speechclip_b200.init import seeded_init_
    torch.set_num_threads(os.cpu_count() or 1)
    model = osc.SpeechClipOracle(oh.HubertCfg.named("hubert"), oc.ClipCfg.named("ViT-B/32"),
                                 dict(n_layers=1, nhead=8, dim_feedforward=3072)).eval()
    seeded_init_(model, 7122)
    for n, p in model.named_parameters():
        p.requires_grad = n.startswith("parallel_branch") or "weightedsum" in n
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=1e-4, weight_decay=1e-6)
    b = synth_batch(pairs, 7122, False)
    wavs = list(b["wav"])

    def step():
        with torch.no_grad():
            enc = model.audio_encoder.encoder
            padded, mask = oh.preprocess_input(wavs, False)
            hs = enc.custom_forward(padded, mask)["layer_results"]
            image_feat = model.clip.model.encode_image(b["image"])
        feat = osc.weighted_sum(model.audio_encoder.weightedsum_layer.weights, hs)
        p = model.parallel_branch(feat, oh.feat_lengths([N_SAMPLES] * pairs, hs[0].shape[1]))
        feats = {"id": b["id"], "image_feat": image_feat / image_feat.norm(dim=-1, keepdim=True),
                 "parallel_audio_feat": p / p.norm(dim=-1, keepdim=True)}
        loss = model.compute_loss(feats)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_(params, 4.0)
        opt.step()
        return float(loss)

    return step


def time_cpu(pairs, steps, warmup):
    step = cpu_step_fn(pairs)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return pairs / dt, dt


CPU_PAIRS = 16   # the "steadier figure" is the B = 16 step (B = 2 is dominated by per-call overheads)


def time_cpu_config1(runs=3):
    """BASELINE.json configs[0]: Parallel SpeechCLIP-base FORWARD on 2 pairs, single process (the example.py flow:
    encode_speech + forward), eval mode.  Median of ``runs`` after one warm-up -> pairs/s."""
    from oracle import clip as oc
    from oracle import hubert as oh
    from oracle import speechclip as osc
    from speechclip_b200.init import seeded_init_
    torch.set_num_threads(os.cpu_count() or 1)
    model = osc.SpeechClipOracle(oh.HubertCfg.named("hubert"), oc.ClipCfg.named("ViT-B/32"), dict(n_layers=1, nhead=8, dim_feedforward=3072)).eval()
    seeded_init_(model, 7122)
    b = synth_batch(2, 7122, False)
    wavs = list(b["wav"])
    ts = []
    with torch.no_grad():
        for i in range(runs + 1):
            t0 = time.perf_counter()
            model.encode_speech(wavs)
            model(wavs, b["image"], b["id"])
            if i:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    return 2.0 / ts[len(ts) // 2]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = CPU_PAIRS
    value, dt = time_cpu(pairs, args.steps, args.warmup)
    cores = os.cpu_count() or 1
    config1 = time_cpu_config1()
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "Parallel SpeechCLIP-base (HuBERT-base + CLIP ViT-B/32) training step, batch 256, 102400-sample utterances",
                   "note": f"reference CPU arm times a bounded sample of {pairs} pairs per step (the full 256-pair step takes ~25 s on these cores)"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{pairs} pairs/step x {args.steps} steps: oracle fp32 training step (towers fwd, branch fwd+bwd, InfoNCE, clip+Adam), "
                                   f"torch.set_num_threads({cores})",
                         "config1_forward_pairs_per_s": config1,
                         "config1": "BASELINE.json configs[0]: Parallel-base forward on 2 pairs (example.py flow: encode_speech + forward), median of 3"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# =============================================================================================================== own arm
def run_own(args):
    import torch.distributed as dist
    from avssl.base import OrderedNamespace
    from avssl.model import KWClip_GeneralTransformer
    from speechclip_b200 import engine, lib, ops
    from speechclip_b200.configs import parallel_config

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the sm_100a extension is the only implementation of this path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib.load()

    gb = args.global_batch or GLOBAL_BATCH   # --global-batch 4096: BASELINE.json configs[4] (8 GPUs x 512 pairs, all-gathered negatives)
    per_gpu = gb // world if args.scaling == "strong" else gb
    if args.batch:
        per_gpu = args.batch
    global_batch = per_gpu * world
    # CUDA-graph replay of the frozen towers: needed once the per-GPU batch is small enough for python/ctypes launch time
    # (~10 ms per step) to bound the step, and used up to 256 pairs per GPU because it takes the host out of the timed region
    # (eager enqueue measured 8-28 ms per 35 ms step on the shared boxes, 3.5 ms with graphs; same device time).  With graphs
    # the per-kernel events cannot be taken inside the timed region, so the roofline numbers come from eager, event-
    # instrumented steps after it.  Above 256 pairs per GPU (config 5: 512) the graphs' private buffers would not fit beside
    # the pipeline's two tower slots.
    use_graphs = args.graphs == "on" or (args.graphs == "auto" and per_gpu <= 256)
    engine.GRAPHS = use_graphs

    if args.config == "cascaded":   # BASELINE.json configs[2]: keyword VQ + CLIP text tower, 8112-entry reduced vocabulary
        import tempfile
        from speechclip_b200.configs import cascaded_config, write_synthetic_vocab_usage
        cfg = cascaded_config("base", write_synthetic_vocab_usage(os.path.join(tempfile.mkdtemp(prefix="scb_bench_"), "vocab_usage.npy")))
    else:
        cfg = parallel_config(args.config)
    model = KWClip_GeneralTransformer(OrderedNamespace(cfg)).to(dev)
    model.train()
    opts, scheds = model.configure_optimizers()
    opt, sched = opts[0], scheds[0]["scheduler"]

    from speechclip_b200.runtime import DevicePrefetcher, TowerPipeline, bind_to_gpu_numa_node
    numa_node = bind_to_gpu_numa_node(local)  # pinned staging buffers local to the GPU's PCIe root (None: topology not visible)
    host = synth_batch(per_gpu, 7122 + rank, True)
    host["id"] += rank * per_gpu
    resident = {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    def step(batch):
        out = model.training_step(batch)
        loss = model.training_step_end(out)["loss"]
        opt.zero_grad()
        loss.backward()
        model.on_after_backward()
        opt.step()
        sched.step()
        return loss

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        gc.collect()
        gc.disable()   # a generation-2 collection inside a 100-400 ms timed region is a visible hiccup at 5-20 ms per step
        try:
            return _timed(fn, steps)
        finally:
            gc.enable()

    def _timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for _ in range(steps):
            fn()
        timed.host_ms = (time.perf_counter() - t_host) * 1e3 / steps  # python + launch time to ENQUEUE one step
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # Cross-step overlap (speechclip_b200.runtime.TowerPipeline): the frozen towers of batch i + 1 are enqueued on the tower
    # streams before the head / loss / backward / all-reduce / Adam tail of batch i, whose ~150 small dependent kernels then no
    # longer leave the GPU idle.  K timed steps are K complete steps: the first batch's towers run un-overlapped inside the
    # timed region, and nothing of step K + 1 is started.
    pipe = TowerPipeline(model) if args.pipeline == "on" else None

    def run_steps(batches):
        for batch in (pipe.iterate(batches) if pipe is not None else batches):
            step(batch)

    # warm-up: every CUDA-graph signature is run eagerly once, captured on its second use and replayed from the third; the
    # pipeline alternates two tower slots, so it needs six steps before every step is a replay
    n_warm = max(args.warmup, 3)
    setup_steps = max(0, (6 if pipe is not None else 3) - n_warm)   # graph warm / capture passes that the W warm-up steps do not cover
    if setup_steps:
        run_steps(resident for _ in range(setup_steps))
        barrier()
    run_steps(resident for _ in range(n_warm))
    barrier()

    # ---- device-resident timing.  The two towers run on two streams (and, at small per-GPU batch, as CUDA-graph replays), so a
    # per-kernel duration cannot be taken inside the timed region without changing it: the per-call CUDA events that feed the
    # roofline / breakdown are recorded in instrumented steps right after it — same process, same buffers, towers serialised on
    # one stream, every C-ABI call bracketed by events on its launching stream.
    import avssl.model.kwClip as kwclip_mod
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    n0 = lib.launch_count() + engine.graph_replayed_kernels()
    ops.PROFILE = None
    ms_total = timed(lambda: run_steps(resident for _ in range(args.steps)), 1)
    host_ms = timed.host_ms / args.steps
    launches = lib.launch_count() + engine.graph_replayed_kernels() - n0
    prof_steps = 3 if use_graphs else min(args.steps, 5)
    overlap, kwclip_mod.OVERLAP_TOWERS = kwclip_mod.OVERLAP_TOWERS, False
    ops.PROFILE = []
    for i in range(prof_steps + (1 if use_graphs else 0)):
        if use_graphs and i == 1:
            torch.cuda.synchronize()
            ops.PROFILE = []   # the first eager step after graph replay allocates the eager workspace: not counted
        if pipe is not None:   # same tower buffers as the timed steps (slot 1), both towers on this stream
            b = dict(resident)
            b["_scb_towers"] = model.precompute_towers(resident, slot=1, overlap=False)
            step(b)
        else:
            step(resident)
    torch.cuda.synchronize()
    prof, ops.PROFILE = ops.PROFILE, None
    kwclip_mod.OVERLAP_TOWERS = overlap
    ms_step = ms_total / args.steps
    value = global_batch / (ms_step * 1e-3)

    # ---- end to end through the public API with host buffers: every step's inputs are copied from pinned host memory
    # (double-buffered on a copy stream by speechclip_b200.runtime.DevicePrefetcher) and its loss is read back to the host
    host_times = []
    prefetcher = DevicePrefetcher(None, dev, lookahead=1 if pipe is not None else 0)  # one object for the warm-up and the timed pass: its two device slots are allocated once
    # host->device bandwidth of this process's pinned buffers (diagnostic: the copy must hide behind one step)
    probe = {k: torch.empty_like(v, device=dev) for k, v in host.items()}
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    c0.record()
    for k, v in host.items():
        probe[k].copy_(v, non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    h2d_gbs = h2d_bytes / (c0.elapsed_time(c1) * 1e-3) / 1e9
    del probe

    loss_slots = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]

    def e2e_run(n):
        # Every step's loss is copied to pinned host memory and read by the host — one step late (after step i+1 has been
        # enqueued), the way a training loop logs without draining the GPU between steps.
        last, pending = None, None
        staged = prefetcher.iterate(host for _ in range(n))
        for i, batch in enumerate(pipe.iterate(staged) if pipe is not None else staged):
            t_h = time.perf_counter()
            if dbg is not None:
                e_a = torch.cuda.Event(enable_timing=True)
                e_a.record()
            loss = step(batch)
            if dbg is not None:
                e_b = torch.cuda.Event(enable_timing=True)
                e_b.record()
                dbg.append((e_a, e_b, time.perf_counter()))
            host_times.append((time.perf_counter() - t_h) * 1e3)
            slot = loss_slots[i % 2]
            slot.copy_(loss.detach(), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            if pending is not None:
                pending[0].synchronize()
                last = float(pending[1])
            pending = (ev, slot)
        pending[0].synchronize()
        last = float(pending[1])
        return last

    dbg = []   # per-step device spans / gaps of the end-to-end loop (2 events per step): reported beside the end-to-end value
    e2e_run(6 if pipe is not None else 2)   # the prefetcher's device slots are new graph signatures: warm, capture, replay per tower slot
    # Three passes of K steps; the MEDIAN pass is reported and all are listed (`passes_ms_per_step`; the fastest is kept as a
    # diagnostic).  On the shared gpurun hosts the end-to-end loop was 39-41 ms / step in most runs and 48-59 ms in some, with the
    # device-resident number unchanged and no gap between steps: host-side interference (the H2D copy of 259 MB per step shares
    # the PCIe root / memory channels with other tenants), which several passes make visible instead of hiding.
    passes = []
    for _ in range(3):
        dbg.clear()
        ms = timed(lambda: e2e_run(args.steps), 1) / args.steps
        passes.append((ms, list(dbg)))
    ms_e2e, dbg = sorted(passes, key=lambda t: t[0])[len(passes) // 2]
    clocks = sampler.stop() if sampler else None
    torch.cuda.synchronize()
    spans = sorted(ea.elapsed_time(eb) for ea, eb, _ in dbg)
    gaps = [dbg[j - 1][1].elapsed_time(dbg[j][0]) for j in range(1, len(dbg))]
    e2e_diag = {"step_span_ms_median": spans[len(spans) // 2], "step_span_ms_max": spans[-1], "gap_ms_max": max(gaps) if gaps else 0.0,
                "outside_steps_ms": ms_e2e * args.steps - sum(spans) - sum(gaps)}
    if os.environ.get("SCB_E2E_DEBUG"):
        for j, (ea, eb, th) in enumerate(dbg):
            sys.stderr.write(f"e2e step {j}: device span {ea.elapsed_time(eb):.2f} ms, gap before {gaps[j - 1] if j else 0.0:.2f} ms, "
                             f"host t {1e3 * (th - dbg[0][2]):.1f} ms\n")

    # ---- per-entry-point breakdown and roofline of the dominant kernel
    agg = {}
    shapes = {}
    for name, flops, e0, e1, shape in prof:
        ms = e0.elapsed_time(e1)
        a = agg.setdefault(name, [0.0, 0.0, 0])
        a[0] += ms
        a[1] += flops
        a[2] += 1
        if shape:
            sh = shapes.setdefault(shape, [0.0, 0.0, 0])
            sh[0] += ms
            sh[1] += flops
            sh[2] += 1
    if args.dump_profile and rank == 0:
        with open(args.dump_profile, "w") as f:
            f.write("ms_per_step,calls_per_step,tflops,shape\n")
            for shp, (ms, fl, n) in sorted(shapes.items(), key=lambda kv: -kv[1][0]):
                f.write(f"{ms / prof_steps:.4f},{n / prof_steps:.1f},{fl / (ms * 1e-3) / 1e12 if ms > 0 else 0:.1f},{shp}\n")
    pk = peaks()
    # dense algorithmic GFLOP per pair per step (the cascaded branch's own work is < 1% of the towers')
    gf_step = 427.2 if args.config == "large" else GF_PER_PAIR_STEP
    g = agg.get("scb_gemm", [0.0, 0.0, 0])
    gemm_ms, gemm_flops, gemm_n = g
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # DRAM bytes per GEMM launch come from an ncu capture (profiles/gemm_dram_traffic.json), which cannot be repeated inside the
    # timed run: the number is only emitted when the capture was taken on THIS GEMM source (sha256 of csrc/gemm_tcgen05.cu stamped
    # into the file by tools/make_profiles.sh) and for exactly this workload; otherwise null.
    traffic, traffic_note = None, "no capture for this workload"
    tpath = os.path.join(ROOT, "profiles", "gemm_dram_traffic.json")
    if os.path.exists(tpath) and args.config == "base" and per_gpu == 256:
        import hashlib
        rec = json.load(open(tpath))
        cur = hashlib.sha256(open(os.path.join(ROOT, "speechclip_b200", "csrc", "gemm_tcgen05.cu"), "rb").read()).hexdigest()
        if rec.get("gemm_source_sha256") == cur:
            traffic, traffic_note = rec.get("dram_bytes_per_launch"), "ncu capture of this GEMM source: " + rec.get("captured", "profiles/gemm_dram_traffic.json")
        else:
            traffic_note = "profiles/gemm_dram_traffic.json was captured on a different csrc/gemm_tcgen05.cu: not reported"
    breakdown = {k: {"ms_per_step": v[0] / prof_steps, "calls_per_step": v[2] / prof_steps} for k, v in
                 sorted(agg.items(), key=lambda kv: -kv[1][0])}

    if rank != 0:
        dist.destroy_process_group()
        return
    line = {
        "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": n_warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f16",
        "data": "synthetic",
        "config": {"workload": {"base": "Parallel SpeechCLIP-base (HuBERT-base + CLIP ViT-B/32)",
                                "large": "Parallel SpeechCLIP-large (HuBERT-large + CLIP ViT-L/14)",
                                "cascaded": "Cascaded SpeechCLIP-base (HuBERT-base + keyword VQ over 8112 subwords + CLIP ViT-B/32 text and image towers)"}[args.config]
                               + f" training step, batch {global_batch}, 102400-sample utterances",
                   "global_batch": global_batch, "pairs_per_gpu": per_gpu, "frames": 319, "parallelism": f"dp{world}",
                   "mode": "training step (model.train()): branch dropout p=0.1 active (Philox masks, regenerated in backward), frozen towers in eval arithmetic, trainable branch "
                           + ("2.77 M params (+ frozen CLIP text tower in the differentiated path)" if args.config == "cascaded" else "7.48 M params"),
                   "l2": "inputs (259 MB) and activations (GBs) larger than the 126 MB L2; no flush needed",
                   "cuda_graphs": use_graphs, "tower_pipeline": pipe is not None, "setup_steps_before_warmup": setup_steps, "tower_streams": 2 if kwclip_mod.OVERLAP_TOWERS else 1},
        "e2e": {"value": global_batch / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": h2d_bytes,
                "d2h_bytes_per_step": 4, "h2d_gb_per_s_measured": h2d_gbs, "numa_node_bound": numa_node, "passes_ms_per_step": [t[0] for t in passes],
                "reported": "median pass", "fastest_pass_ms_per_step": min(t[0] for t in passes), "diagnostics": e2e_diag,
                "loss_read": "every step's loss is copied to pinned host memory and read by the host one step late (after the next step is enqueued)"},
        "hbm_peak_reserved_gb": torch.cuda.max_memory_reserved(dev) / 1e9,
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": sorted(host_times)[len(host_times) // 2],
        "host_enqueue_ms_per_step_timed": host_ms,
        "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (scb_gemm)", "achieved": achieved, "peak": pk["tflops"],
                     "unit": "TFLOP/s", "frac": achieved / pk["tflops"], "traffic": traffic, "traffic_source": traffic_note, "peak_source": pk["src"] + " (sustained cuBLAS bf16)",
                     "launches_per_step": gemm_n / prof_steps, "gemm_ms_per_step": gemm_ms / prof_steps,
                     "gemm_share_of_step": gemm_ms / max(sum(v[0] for v in agg.values()), 1e-9),
                     "events": f"{prof_steps} instrumented step(s) right after the timed region: every C-ABI call bracketed by CUDA events on its launching stream, towers serialised on one stream (inside the timed region they overlap on two streams" + (" and replay as CUDA graphs)" if use_graphs else ")"),
                     "instrumented_step_ms": sum(v[0] for v in agg.values()) / prof_steps,
                     "step_tflops_dense_algorithmic": value / world * gf_step / 1e3,
                     "step_frac_of_peak": value / world * gf_step / 1e3 / pk["tflops"]},
        "breakdown_ms_per_step": breakdown,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        t0 = time.perf_counter()
        v, dt = time_cpu(CPU_PAIRS, 3, 1)
        cores = os.cpu_count() or 1
        c1 = time_cpu_config1()
        line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port",
                                "sample": f"{CPU_PAIRS} pairs/step x 3 steps (+1 warm-up), oracle fp32 training step (towers fwd, branch fwd+bwd, InfoNCE, "
                                          f"clip+Adam), torch.set_num_threads({cores}), {time.perf_counter() - t0:.1f} s wall",
                                "config1_forward_pairs_per_s": c1,
                                "config1": "BASELINE.json configs[0]: Parallel-base forward on 2 pairs (example.py flow), median of 3"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line: dict):
    """Exactly ONE JSON line on the real stdout (libraries such as NCCL print banners to fd 1: it is parked on stderr)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = os.dup(1)


def main():
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"])
    ap.add_argument("--batch", type=int, default=0, help="pairs per GPU (default: 256 / N strong, 256 weak)")
    ap.add_argument("--global-batch", type=int, default=0, help="global batch under strong scaling (default 256; 4096 = BASELINE.json configs[4])")
    ap.add_argument("--config", default="base", choices=["base", "large", "cascaded"],
                    help="base = BASELINE.json configs[1] (headline); large = HuBERT-large + ViT-L/14 (configs[3])")
    ap.add_argument("--graphs", default="auto", choices=["auto", "on", "off"], help="replay the frozen towers as CUDA graphs")
    ap.add_argument("--pipeline", default="on", choices=["on", "off"],
                    help="run the frozen towers of batch i+1 under the head/backward/Adam tail of batch i (runtime.TowerPipeline)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-profile", default="", help="write the per-GEMM-shape timing table (CSV) here")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_own(args)


if __name__ == "__main__":
    main()
