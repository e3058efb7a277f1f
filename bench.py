"""bench.py -- samples/sec of the ViT-L/14 contrastive step (BASELINE.json) on N B200s.

One "step" = one pass of the hot path over one synthetic batch.  Default workload (the driver's headline) is BASELINE
configs[1]: ViT-L/14 image tower forward (patch embed -> 24 blocks -> ln_post/proj -> L2 norm), InfoNCE (ClipLoss) against
fixed CLIP text anchors, backward through every image-tower weight, gradient all-reduce (N > 1) and the fused AdamW step.
`--config 2|3|4` runs the three ViT-Lens recipes of BASELINE configs[2..4] (audio Lens with 128 latents / depth three-tower /
point-cloud Lens with 8192 points; TriCLIP with frozen image + text towers, TriClipLoss, frozen or partly unlocked ViT behind
the Lens) at 512 samples per GPU.  Random-init weights of the named architecture, synthetic inputs (no datasets / checkpoints
offline).  N > 1: one process per GPU (torchrun), fixed batch per GPU (weak scaling), one packed feature all-gather + LSE
exchange in the loss, bucketed gradient all-reduce overlapped with backward.

    python bench.py --gpus 1 --steps 10 --warmup 3            # this repo's CUDA path, configs[1]
    python bench.py --config 2 --steps 5                      # audio-Lens recipe
    torchrun ... bench.py --gpus 2                            # N > 1 adds the on-hardware multi-rank parity leg ("parity" in the JSON line)
    python bench.py --impl reference --steps 2 --warmup 1     # the reference's algorithm on the host CPUs
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "vit-lens_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

MODEL = "ViT-L-14"
IMG = 224
EMBED = 768
METRIC = "samples/sec ViT-L/14 contrastive step"
CPU_BATCH = 4

# BASELINE.json configs[1..4]; FLOP per sample fwd+bwd from SURVEY.md 8(d)
WORKLOADS = {
    1: dict(batch=256, flop=486.1e9, kind="image",
            name="ViT-L/14@224 image encoder fwd+bwd (all weights trainable) + ClipLoss vs fixed CLIP text anchors + AdamW (BASELINE configs[1])"),
    2: dict(batch=512, flop=434e9, kind="tri", modality="audio", overrides=dict(perceiver_num_latents=128), lock=dict(unlock_cls=True),
            name="ViT-L/14 + audio Lens (Perceiver, 128 latents, 600 AST tokens) + frozen image/text anchors, TriClipLoss, frozen ViT, AdamW on the Lens "
                 "(BASELINE configs[2])"),
    3: dict(batch=512, flop=525e9, kind="tri", modality="depth", overrides={}, lock=dict(unlock_cls=True, unlock_trans_first_n_layers=4),
            name="ViT-L/14 image + text + depth three-tower InfoNCE (depth adapter, Lens = identity, first 4 ViT blocks unlocked), TriClipLoss "
                 "(BASELINE configs[3]; 512 per GPU)"),
    4: dict(batch=512, flop=726e9, kind="tri", modality="pc", overrides={}, lock=dict(unlock_cls=True),
            name="ViT-L/14 + point-cloud Lens (8192 points -> 512 groups of 32, Perceiver 256 latents) + frozen image/text anchors, TriClipLoss, "
                 "frozen ViT, BatchNorm in training mode (BASELINE configs[4]; 512 per GPU)"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), tf_burst=float(p["bf16_tflops"]), tf=float(p["bf16_tflops_sustained"]), src="measured")
    except Exception:
        return dict(hbm=6650.0, tf_burst=1590.0, tf=1400.0, src="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- the reference arm (CPU)
def cpu_step_fn(batch: int):
    """The reference's algorithm for configs[1], restated by the oracle (kind = "port"; the reference itself is
    Python/PyTorch and does not exist on the GPU box).  Returns (step callable, cores)."""
    from oracle import vitlens_oracle as O
    from vitlens_b200 import synth
    import open_clip

    cores = pick_cpu_threads()
    model = open_clip.create_model(MODEL, device="cpu")
    sd = {k: v for k, v in synth.synth_state_dict(model.state_dict(), seed=0).items() if k.startswith("visual.") or k == "logit_scale"}
    del model
    for v in sd.values():
        v.requires_grad_(True)
    img = synth.synth_normal("image", (batch, 3, IMG, IMG), seed=1)
    anchors = O.l2_normalize(synth.synth_normal("anchors", (batch, EMBED), seed=2))

    def step():
        for v in sd.values():
            v.grad = None
        f = O.l2_normalize(O.image_tower(sd, "visual.", img, 16))
        loss = O.clip_loss(f, anchors, sd["logit_scale"].exp())
        loss.backward()
        return float(loss.detach())

    return step, cores


def pick_cpu_threads():
    """All the host threads the box can actually use: probe a GEMM at a few thread counts (shared / cgroup-limited hosts
    report more logical CPUs than they can run) and keep the fastest."""
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count()
    a = torch.randn(2048, 2048)
    best, best_t = avail, None
    for n in sorted({avail, max(1, avail // 2), max(1, avail // 4), min(avail, 32), min(avail, 16)}, reverse=True):
        torch.set_num_threads(n)
        a @ a
        t0 = time.perf_counter()
        for _ in range(3):
            a @ a
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t * 0.9:
            best, best_t = n, dt
    torch.set_num_threads(best)
    return best


def time_cpu(steps: int, warmup: int, batch: int = CPU_BATCH):
    step, cores = cpu_step_fn(batch)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    sec = statistics.median(ts)
    return batch / sec, cores, sec


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    sps, cores, sec = time_cpu(steps, warmup)
    sample = f"ViT-L/14 image tower + ClipLoss fwd+bwd, fp32, batch {CPU_BATCH} per step, {steps} timed steps (median)"
    out = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps, "warmup": warmup,
        "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ViT-L/14@224 image encoder fwd+bwd + ClipLoss vs fixed text anchors (BASELINE configs[1]), CPU sample batch {CPU_BATCH}"},
        "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------- workloads (this repo's arm)
class Workload:
    """Model + loss + synthetic inputs of one BASELINE config.  `inputs` are tuples of tensors; `step_loss(inputs)` runs the
    public-API forward + loss and returns (loss, features that the verify leg inspects)."""

    def __init__(self, cfg_id, batch, dev, rank, world):
        import open_clip
        from vitlens_b200 import synth

        w = WORKLOADS[cfg_id]
        self.id, self.B, self.dev, self.rank, self.world = cfg_id, batch, dev, rank, world
        self.kind = w["kind"]
        self.open_clip = open_clip
        seed = 100 + rank
        if self.kind == "image":
            model = open_clip.create_model(MODEL, device="cpu")
            model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0))
            self.tower = model.visual.to(dev)
            self.logit_scale = torch.nn.Parameter(model.logit_scale.detach().to(dev))
            del model
            self.named = [("visual." + n, p) for n, p in self.tower.named_parameters()] + [("logit_scale", self.logit_scale)]
            self.loss_fn = open_clip.ClipLoss(rank=rank, world_size=world)
            self.anchors = torch.nn.functional.normalize(synth.synth_normal("anchors", (batch, EMBED), seed=seed), dim=-1).to(dev)
            self.host_inputs = [(synth.synth_normal("image", (batch, 3, IMG, IMG), seed=seed + i).pin_memory(),) for i in range(2)]
            self.model = None
        else:
            from mm_vit_lens.model_cfg import training_args

            margs = training_args(w["modality"], **w["overrides"])
            model = open_clip.tri_create_model(MODEL, None, device="cpu", args=margs)
            model.load_state_dict(synth.synth_state_dict(model.state_dict(), seed=0), strict=True)
            model.lock_image_tower()
            model.lock_text_tower()
            model.lock_visual_tower(**w["lock"])
            model.train()  # BatchNorm of the point tokenizer: batch statistics, as in the reference's training loop
            self.model = model.to(dev)
            self.logit_scale = model.logit_scale
            self.named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
            self.loss_fn = open_clip.TriClipLoss(rank=rank, world_size=world)
            self.modality = w["modality"]
            self.host_inputs = []
            for i in range(2):
                img = synth.synth_normal("image", (batch, 3, IMG, IMG), seed=seed + i).pin_memory()
                txt = synth.synth_text(batch, 77, 49408, seed=seed + i).pin_memory()
                if self.modality == "audio":
                    vis = synth.synth_normal("audio", (batch, 512, 128), seed=seed + i)
                elif self.modality == "depth":
                    vis = synth.synth_normal("depth", (batch, 1, IMG, IMG), seed=seed + i)
                else:
                    vis, _ = synth.synth_shapes(batch, 8192, seed=seed + i)
                self.host_inputs.append((img, txt, vis.pin_memory()))
        self.params = [p for _, p in self.named]
        self.n_params = sum(p.numel() for p in self.params)
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in self.host_inputs[0])

    def features(self, inputs):
        oc = self.open_clip
        if self.kind == "image":
            return (oc.model._normalize(self.tower(inputs[0])), self.anchors)
        fi, ft, fv, _ = self.model(inputs[0], inputs[1], inputs[2])
        return (fi, ft, fv)

    def loss(self, feats):
        return self.loss_fn(*feats, self.logit_scale.exp())


def run_cuda(args):
    import torch.distributed as dist

    from vitlens_b200 import lib as L
    from vitlens_b200 import grad_sync, optim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU path in the product)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world} (launch with torchrun for N > 1)"

    L.load()
    wcfg = WORKLOADS[args.config]
    B = args.batch or wcfg["batch"]
    wl = Workload(args.config, B, dev, rank, world)
    opt = optim.AdamW(wl.named, lr=1e-6, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2)
    params = wl.params
    n_host = len(wl.host_inputs)
    dev_inputs = [tuple(t.to(dev) for t in hi) for hi in wl.host_inputs]
    host_loss = torch.zeros((), dtype=torch.float32).pin_memory()
    # N > 1: DDP-style bucketed gradient all-reduce, overlapped with backward (vitlens_b200/grad_sync.py)
    arena = None
    if world > 1:  # peer-memory arena: fused feature gather in the loss + copy-engine gradient exchange (VL_COMM=nccl switches it off)
        from vitlens_b200 import comm

        arena = comm.init_arena(nbytes=world * (wl.n_params + 64) * 4 + (1 << 20))
    reducer = grad_sync.GradReducer(params, arena=arena) if world > 1 else None

    def train_step(inputs):
        loss = wl.loss(wl.features(inputs))
        loss.backward()
        if reducer is not None:
            reducer.finish()
            opt.step(grad_scale=1.0 / world, n_src=reducer.n_src, src_stride=reducer.src_stride)
        else:
            opt.step()
        with torch.no_grad():
            wl.logit_scale.clamp_(0, 4.6052)
        opt.zero_grad()
        return loss.detach()

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- warm-up
    for i in range(args.warmup):
        train_step(dev_inputs[i % n_host])
    sync()
    peak_mem_gb = torch.cuda.max_memory_allocated() / 2 ** 30

    # ---- timed: device-resident inputs.  Working set per step (activations, tens of GB) >> 126 MB L2.
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    L.launch_count = 0
    L.GEMM_TIMING = [] if rank == 0 else None
    L.ATTN_TIMING = [] if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync()
    e0.record()
    for i in range(args.steps):
        train_step(dev_inputs[i % n_host])
    e1.record()
    sync()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = L.launch_count
    clocks = sampler.stop() if rank == 0 else None
    gemm_t, attn_t = L.GEMM_TIMING, L.ATTN_TIMING
    L.GEMM_TIMING = L.ATTN_TIMING = None
    value = world * B * args.steps / (ms / 1e3)

    # ---- timed: end to end through the public API with HOST buffers.  Every step copies its batch from pinned host memory
    # (on a copy stream, one step ahead: the input pipeline of a real training loop) and reads its loss back to the host
    # (asynchronously; the host waits for step i-1's value while step i is queued, and for the last one before the clock stops).
    copy_stream = torch.cuda.Stream()
    dev_in = [tuple(torch.empty_like(t) for t in dev_inputs[0]) for _ in range(2)]
    host_losses = torch.zeros(args.steps, dtype=torch.float32).pin_memory()

    def prefetch(i, after):
        buf = dev_in[i % 2]
        with torch.cuda.stream(copy_stream):
            if after is not None:
                copy_stream.wait_event(after)  # the step that last read this buffer has finished
            for d, h in zip(buf, wl.host_inputs[i % n_host]):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
        return buf, ev

    sync()
    e0.record()
    nxt = prefetch(0, None)
    read_ev, done = [], []
    for i in range(args.steps):
        imgs, ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        if i + 1 < args.steps:  # queued before step i's kernels so the copy runs underneath them
            nxt = prefetch(i + 1, done[i - 1] if i >= 1 else None)
        loss = train_step(imgs)
        d = torch.cuda.Event()
        d.record()
        done.append(d)
        host_losses[i:i + 1].copy_(loss.reshape(1), non_blocking=True)
        rev = torch.cuda.Event()
        rev.record()
        read_ev.append(rev)
        if i > 0:
            read_ev[i - 1].synchronize()  # the host consumes step i-1's loss here
    read_ev[-1].synchronize()
    host_loss.copy_(host_losses[-1])
    e1.record()
    sync()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e = world * B * args.steps / (ms_e2e / 1e3)
    final_loss = float(host_loss)

    # ---- diagnostic pass (outside both timed regions): CUDA events around every entry point of the C ABI for one step
    L.CALL_TIMING = []
    sync()
    e0.record()
    train_step(dev_inputs[0])
    e1.record()
    sync()
    calls, L.CALL_TIMING = L.CALL_TIMING, None
    by_kernel = {}
    for name, a, b, bound, amount in calls:
        ent = by_kernel.setdefault(name, [0, 0.0, bound, 0.0])
        ent[0] += 1
        ent[1] += a.elapsed_time(b)
        ent[3] += amount
    step_ms_diag = e0.elapsed_time(e1)
    pk_ = peaks()
    rows = {}
    for k, v in sorted(by_kernel.items(), key=lambda kv: -kv[1][1]):
        row = {"launches": v[0], "ms": round(v[1], 3)}
        if k == "vl_adamw_multi":
            v[2], v[3] = "hbm", wl.n_params * 30.0  # 16 B read + 14 B written per parameter
        if v[2] == "tensor" and v[1] > 0:
            row.update(bound="tensor", achieved=round(v[3] / v[1] / 1e9, 1), unit="TFLOP/s", frac=round(v[3] / v[1] / 1e9 / pk_["tf"], 3))
        elif v[2] == "hbm" and v[1] > 0:
            row.update(bound="hbm", achieved=round(v[3] / v[1] / 1e6, 1), unit="GB/s", frac=round(v[3] / v[1] / 1e6 / pk_["hbm"], 3))
        rows[k] = row
    breakdown = {"step_ms": round(step_ms_diag, 2), "sum_kernels_ms": round(sum(v[1] for v in by_kernel.values()), 2),
                 "note": "untimed diagnostic step, CUDA events around each C-ABI call; numerators = algorithmic FLOPs (2MNK) / bytes of each call "
                         "(bf16 operands read once, results written once)",
                 "by_entry_point": rows}

    # ---- multi-rank parity leg (N > 1, --verify): the N-rank CUDA loss and the overlapped gradient exchange against a
    # single-process fp32 recomputation on the gathered features / the explicit cross-rank sum of the local gradients
    parity = verify_multi_rank(wl, reducer, dev_inputs[0], world, rank, dev) if (world > 1 and args.verify) else None

    torch_base = None
    if args.torch_baseline and world == 1 and args.config == 1:
        del opt
        torch_base = torch_gpu_baseline(wl, B, dev)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    pk = peaks()
    roof = {"bound": "tensor", "achieved": None, "peak": pk["tf"], "unit": "TFLOP/s", "frac": None, "traffic": None,
            "kernel": "gemm_bf16_kernel (tcgen05, all linear layers fwd/dgrad/wgrad)", "peak_source": f"{pk['src']} bf16_tflops_sustained"}
    if gemm_t:
        torch.cuda.synchronize()
        tot_ms = sum(a.elapsed_time(b) for a, b, _, _ in gemm_t)
        tot_fl = sum(f for _, _, f, _ in gemm_t)
        by_shape = {}
        for a, b, f, key in gemm_t:
            d = by_shape.setdefault(key, [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += f
            d[2] += 1
        roof["by_shape"] = [{"MNK_epi_amn_bmn": list(k), "launches": v[2], "ms": round(v[0], 3), "tflops": round(v[1] / v[0] / 1e9, 1)}
                            for k, v in sorted(by_shape.items(), key=lambda kv: -kv[1][0])[:14]]
        roof["achieved"] = tot_fl / tot_ms / 1e9
        roof["frac"] = roof["achieved"] / pk["tf"]
        if args.config == 1:
            # DRAM bytes of one launch of the fc+GELU GEMM (M=65792 N=4096 K=1024, the largest single share of the step), from the
            # committed ncu --set full capture profiles/r01_ncu_gemm_fc_gelu_pair8.txt: 164.4 MB read + 1023.4 MB written, against
            # 143.1 MB + 1077.9 MB algorithmic (activations + weights in, activation + pre-activation out).  A capture constant
            # (ncu cannot run inside a timed bench); re-captured whenever the kernel changes.
            # profiles/r02_ncu_final_kernels.txt (ncu --set full, round-2 kernel): 155.5 MB read + 1022.0 MB written
            roof["traffic"] = int((155527424 + 1021996000) * B / 256)  # captured at batch 256; the GEMM is linear in tokens
            roof["traffic_kernel"] = "gemm2_bf16_kernel<256,8> fc+GELU, per launch; algorithmic 1221036032 B; from the committed ncu capture"
        roof["launches_timed"] = len(gemm_t)
        roof["share_of_step"] = tot_ms / ms
    mhsa = None
    if attn_t:
        by = {}
        for a, b, kind, byts in attn_t:
            d = by.setdefault(kind, [0.0, 0.0, 0])
            d[0] += a.elapsed_time(b)
            d[1] += byts
            d[2] += 1
        mhsa = {k: {"achieved": v[1] / v[0] / 1e6, "peak": pk["hbm"], "unit": "GB/s", "frac": v[1] / v[0] / 1e6 / pk["hbm"], "launches": v[2],
                    "share_of_step": v[0] / ms} for k, v in by.items()}
        if "fwd" in mhsa and args.config == 1:
            # DRAM bytes of one attn_fwd2_kernel launch at batch 256 from the committed ncu --set full capture
            # (profiles/r02_ncu_final_kernels.txt: 405.3 MB read + 113.0 MB written; algorithmic 539.0 MB, the last tiles'
            # output is still in L2 when the kernel ends): no re-reads of Q / K / V
            mhsa["fwd"]["traffic"] = int((405348352 + 112985344) * B / 256)
            mhsa["fwd"]["kernel"] = "attn_fwd2_kernel (persistent, P in TMEM), per launch; algorithmic %d B" % (4 * B * 257 * 1024 * 2)
        if "bwd" in mhsa and args.config == 1:
            # attn_bwd3_kernel at batch 256 (profiles/r02_ncu_attn_bwd3.txt: 679.6 MB read + 368.2 MB written; algorithmic 1078 MB):
            # Q / K / V / O / dO are read once (the L2 prefetch of the next wave's tiles adds no DRAM traffic)
            mhsa["bwd"]["traffic"] = int((679634944 + 368246784) * B / 256)
            mhsa["bwd"]["kernel"] = "attn_bwd3_kernel (keys on the TMEM lanes), per launch; algorithmic %d B" % (8 * B * 257 * 1024 * 2)
    cpu = None
    if not args.no_cpu_baseline and world == 1 and args.config == 1:
        sps, cores, sec = time_cpu(2, 1)
        cpu = {"value": sps, "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"same workload at batch {CPU_BATCH} per step on the host CPUs (oracle port, fp32), 2 timed steps, {sec:.1f} s/step"}
    out = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wcfg["name"], "baseline_config": args.config, "batch_per_gpu": B, "global_batch": world * B,
                   "parallelism": f"dp{world}", "trainable_params": wl.n_params,
                   "exchange": None if world == 1 else ("peer memory over NVLink: feature gather fused into the loss GEMMs, gradient buckets pushed by "
                                                        "the copy engines and summed inside AdamW" if arena is not None else "NCCL all-gather / all-reduce"),
                   "l2": "no flush needed: per-step working set (tens of GB of activations) >> 126 MB L2; inputs alternate between 2 buffers",
                   "tflops_per_gpu": wcfg["flop"] * B * args.steps / (ms / 1e3) / 1e12, "final_loss": final_loss,
                   "peak_mem_gb": round(peak_mem_gb, 1)},
        "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": wl.h2d_bytes * world, "d2h_bytes_per_step": 4 * world,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "breakdown": breakdown,
        "clocks": clocks,
        "roofline": roof,
        "roofline_mhsa": mhsa,
        "cpu_baseline": cpu,
    }
    if parity is not None:
        out["parity"] = parity
    if torch_base is not None:
        out["torch_gpu_baseline"] = torch_base
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def verify_multi_rank(wl, reducer, inputs, world, rank, dev):
    """On-hardware parity of the N-rank path.  (a) loss: every rank's loss value, d(loss)/d(local features)
    and d(loss)/d(logit_scale) from the CUDA kernels + NCCL exchange, against a plain fp32 torch recomputation of the reference's
    formula (loss.py:116-163, default flags local_loss=False, gather_with_grad=False) on the all-gathered features.  (b)
    gradients: the bucketed, overlapped all-reduce against the explicit sum of every rank's local gradients (gathered with one
    all_gather per tensor) for every 1-D parameter and the four largest matrices.  All errors are relative (max over ranks)."""
    import torch.distributed as dist

    def gather(t):
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t.contiguous())
        return out

    def relerr(a, b):
        return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))

    names = [n for n, _ in wl.named]
    mats = sorted((i for i, p in enumerate(wl.params) if p.dim() >= 2), key=lambda i: -wl.params[i].numel())[:4]
    sel = [i for i, p in enumerate(wl.params) if p.dim() < 2 and p.numel() > 1] + mats
    for p in wl.params:
        p.grad = None
    # local gradients (hooks disarmed), features kept for the loss check
    with reducer.no_sync():
        feats = wl.features(inputs)
        for f in feats:
            if f.requires_grad:
                f.retain_grad()
        loss = wl.loss(feats)
        loss.backward()
    local = {i: wl.params[i].grad.detach().clone() for i in sel}
    dfeat = [f.grad.detach().clone() if f.grad is not None else None for f in feats]
    dscale = wl.logit_scale.grad.detach().clone()
    loss_val = loss.detach().clone()
    for p in wl.params:
        p.grad = None
    # the same step with the reducer armed
    wl.loss(wl.features(inputs)).backward()
    reducer.finish(materialize=True)
    reduced = {i: wl.params[i].grad.detach().clone() for i in sel}
    for p in wl.params:
        p.grad = None
    grad_err, worst = 0.0, ""
    for i in sel:
        want = torch.stack(gather(local[i])).double().sum(0)
        e = relerr(reduced[i], want)
        if e > grad_err:
            grad_err, worst = e, names[i]
    # fp32 recomputation of the full-batch loss on the gathered features
    allf = [torch.cat(gather(f.detach().float())) for f in feats]
    Bl = feats[0].shape[0]
    sl = slice(rank * Bl, (rank + 1) * Bl)
    s = wl.logit_scale.detach().clone().float().requires_grad_(True)
    leaf = [a.clone().requires_grad_(True) for a in allf]

    def pair(x, y):
        lx = (s.exp() * x) @ y.t()
        tgt = torch.arange(x.shape[0], device=x.device)
        return (torch.nn.functional.cross_entropy(lx, tgt) + torch.nn.functional.cross_entropy(lx.t(), tgt)) / 2

    full = pair(leaf[0], leaf[1]) if len(leaf) == 2 else pair(leaf[0], leaf[2]) + pair(leaf[1], leaf[2])
    full.backward()
    errs = {"loss_err": abs(float(loss_val) - float(full)) / abs(float(full)), "dscale_err": abs(float(dscale) - float(s.grad)) / abs(float(s.grad))}
    dx = 0.0
    for f, g, lf in zip(feats, dfeat, leaf):
        if g is not None:  # gather_with_grad=False: only the local block carries gradient (loss.py:63-76)
            dx = max(dx, relerr(g, lf.grad[sl]))
    errs["dx_err"] = dx
    errs["grad_reduce_err"] = grad_err
    t = torch.tensor([errs["loss_err"], errs["dx_err"], errs["dscale_err"], errs["grad_reduce_err"]], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return {"loss_err": float(t[0]), "dx_err": float(t[1]), "dscale_err": float(t[2]), "grad_reduce_err": float(t[3]), "grad_reduce_worst": worst,
            "tensors_checked": len(sel), "loss": float(loss_val), "loss_full_batch_fp32": float(full), "world": world,
            "note": "max over ranks; loss / dx / dscale vs fp32 torch on the all-gathered features (bf16 operands in the logits GEMMs); "
                    "grad_reduce vs the explicit cross-rank sum of local gradients"}


def torch_gpu_baseline(wl, B, dev):
    """DIAGNOSTIC, not the headline and not this repo's path: the reference's own PyTorch stack for configs[1] on the same
    GPU -- nn.MultiheadAttention (SDPA fast path) / nn.LayerNorm / nn.GELU under torch.autocast(bfloat16), F.cross_entropy,
    torch.optim.AdamW(fused=True) -- i.e. what the reference's training loop executes on a CUDA device (the
    de-facto bar).  Same architecture, batch and step contents as the measured workload."""
    import gc

    import torch.nn as nn
    import torch.nn.functional as F

    del wl.tower
    gc.collect()
    torch.cuda.empty_cache()

    class Block(nn.Module):
        def __init__(self, d, h):
            super().__init__()
            self.ln_1, self.ln_2 = nn.LayerNorm(d), nn.LayerNorm(d)
            self.attn = nn.MultiheadAttention(d, h)
            self.c_fc, self.c_proj = nn.Linear(d, 4 * d), nn.Linear(4 * d, d)

        def forward(self, x):
            h = self.ln_1(x)
            x = x + self.attn(h, h, h, need_weights=False)[0]
            return x + self.c_proj(F.gelu(self.c_fc(self.ln_2(x))))

    class ViT(nn.Module):
        def __init__(self, d=1024, layers=24, heads=16, patch=14, out=EMBED):
            super().__init__()
            self.conv1 = nn.Conv2d(3, d, patch, patch, bias=False)
            self.cls = nn.Parameter(torch.randn(d) * d ** -0.5)
            self.pos = nn.Parameter(torch.randn(257, d) * d ** -0.5)
            self.ln_pre, self.ln_post = nn.LayerNorm(d), nn.LayerNorm(d)
            self.blocks = nn.ModuleList([Block(d, heads) for _ in range(layers)])
            self.proj = nn.Parameter(torch.randn(d, out) * d ** -0.5)

        def forward(self, img):
            x = self.conv1(img).flatten(2).transpose(1, 2)
            x = torch.cat([self.cls.to(x.dtype).expand(x.shape[0], 1, -1), x], 1) + self.pos.to(x.dtype)
            x = self.ln_pre(x).transpose(0, 1)  # LND like the reference
            for b in self.blocks:
                x = b(x)
            return self.ln_post(x[0]) @ self.proj.to(x.dtype)

    try:
        net = ViT().to(dev)
        scale = nn.Parameter(torch.tensor(2.6593, device=dev))
        opt = torch.optim.AdamW(list(net.parameters()) + [scale], lr=1e-6, betas=(0.9, 0.98), eps=1e-6, weight_decay=0.2, fused=True)
        img = torch.randn(B, 3, IMG, IMG, device=dev)
        anchors = F.normalize(torch.randn(B, EMBED, device=dev), dim=-1)
        tgt = torch.arange(B, device=dev)

        def step():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                f = F.normalize(net(img), dim=-1)
                lg = scale.exp() * f @ anchors.t()
                loss = (F.cross_entropy(lg, tgt) + F.cross_entropy(lg.t(), tgt)) / 2
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)

        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 5
        e0.record()
        for _ in range(n):
            step()
        e1.record()
        torch.cuda.synchronize()
        msb = e0.elapsed_time(e1) / n
        return {"value": B / (msb / 1e3), "unit": "samples/s", "ms_per_step": msb, "steps": n,
                "what": "plain PyTorch (nn.MultiheadAttention/SDPA, autocast bf16, fused AdamW) ViT-L/14 step on this GPU -- diagnostic only",
                "torch": torch.__version__}
    except Exception as e:  # a diagnostic must never take the bench line down
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=sorted(WORKLOADS), help="BASELINE.json configs[i] (1 = the driver's headline)")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the config's)")
    ap.add_argument("--no-verify", dest="verify", action="store_false", help="N > 1: skip the multi-rank parity leg (loss + gradient exchange)")
    ap.add_argument("--no-torch-baseline", dest="torch_baseline", action="store_false",
                    help="N = 1, config 1: skip the plain-PyTorch GPU step timed after the measurement (diagnostic field)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_cuda(args)


if __name__ == "__main__":
    main()
