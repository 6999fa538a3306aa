"""Benchmark of the north-star metric: image-pairs/sec (forward + backward) of DUSt3R ViT-L/16 + 12-layer two-view decoder +
linear pointmap head at 512x512 (BASELINE.json configs[2]/[3]), plus the other BASELINE configs as named workloads.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload linear512|linear224|dpt512|c5_518]
                    [--pairs-per-gpu 8] [--impl b200|reference] [--graph 1|0]

  linear512  configs[2]/[3]  DUSt3R ViT-L/16 + linear heads, 512^2 pairs (default; the metric BASELINE.json quotes)
  linear224  configs[1]      same model, 224^2 pairs
  dpt512                     DUSt3R ViT-L/16 + DPT heads, 512^2 pairs (SURVEY 8a a13/a14 inside the model)
  c5_518     configs[4]      ViT-L/14 intermediate-feature encoder + DPT dense depth, 518^2 images (unit: images/s)

One process per GPU (the driver launches torchrun for N > 1); weak scaling: 8 pairs (images) per GPU.  A step = zero grads
-> forward -> loss (`.sum()` of every output, the idiom of the reference's encoders/utils.py:29-31) -> backward (-> overlapped
NCCL all-reduce of the flat gradient buffer when N > 1).  Prints ONE JSON line on rank 0.

  value     units/s with the image batch already resident in HBM
  e2e       same metric through the public module call with HOST inputs: pinned-host -> device copy of the image batches
            and a device -> host read of the loss inside the timed region, every step
  roofline  dominant kernel = the tcgen05 GEMM: sum of its algorithmic FLOPs / sum of its CUDA-event durations over
            instrumented steps, vs the measured cuBLAS bf16 peak in MEASURED_PEAKS.json (sustained figure)
  gpu_eager_baseline
            the UNMODIFIED reference (baseline/_ref or /root/reference; else the oracle port routed through the same torch
            functionals) run eagerly on the same GPU under torch.autocast(bf16): F.scaled_dot_product_attention (the
            reference default, utils/config.py:13-17), cuBLAS/cuDNN, the PyTorch RoPE fallback with its host sync per
            call (libs/croco/pos_embed.py:149) -- the denominator of north_star's ">= 6x the reference GPU-eager" target
  cpu_baseline / --impl reference
            the reference's own CPU implementation on the host cores (kind "reference" when the reference package is
            present, else the oracle port: kind "port"), bounded sample.  That arm never imports `uniception_b200`.
"""
import argparse
import glob
import hashlib
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# fwd+bwd FLOPs per unit (dpt512 = linear512 - linear heads + 2 DPT heads x 0.2489 T fwd x 3)
WORKLOADS = {
    "linear512": dict(S=512, kind="dust3r", head="linear", unit="pairs/s", flop=6.2137e12,
                      metric="image-pairs/sec (fwd+bwd) DUSt3R ViT-L/16 512^2",
                      desc="DUSt3R ViT-L/16 + 12-layer 2-view decoder + linear head, 512x512 pairs, fwd+bwd"),
    "linear224": dict(S=224, kind="dust3r", head="linear", unit="pairs/s", flop=1.0218e12,
                      metric="image-pairs/sec (fwd+bwd) DUSt3R ViT-L/16 224^2",
                      desc="DUSt3R ViT-L/16 + 12-layer 2-view decoder + linear head, 224x224 pairs, fwd+bwd"),
    "dpt512": dict(S=512, kind="dust3r", head="dpt", unit="pairs/s", flop=6.2137e12 - 3 * 2 * 0.0032e12 + 3 * 2 * 0.2489e12,
                   metric="image-pairs/sec (fwd+bwd) DUSt3R ViT-L/16 + DPT heads 512^2",
                   desc="DUSt3R ViT-L/16 + 12-layer 2-view decoder + DPT heads, 512x512 pairs, fwd+bwd"),
    "c5_518": dict(S=518, kind="depth", head="dpt", unit="images/s", flop=3.967e12,
                   metric="images/sec (fwd+bwd) ViT-L/14 + DPT dense depth 518^2",
                   desc="ViT-L/14 intermediate-feature encoder + DPTFeature + DPTRegressionProcessor + DepthAdaptor, 518x518 images, fwd+bwd"),
}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


def _latest_profile(pattern):
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", pattern)))
    return files[-1] if files else None


def _gemm_traffic():
    """(DRAM bytes per uc_gemm launch, source file) from the newest committed ncu pass (dram__bytes_read.sum +
    dram__bytes_write.sum averaged over the GEMM launches of one step); (None, None) if absent."""
    f = _latest_profile("r0*_dram_traffic_per_kernel.json")
    try:
        with open(f) as fh:
            return float(json.load(fh)["gemm_dram_bytes_per_launch"]), os.path.relpath(f, ROOT)
    except Exception:
        return None, None


def attn_source_hash():
    """sha256 over the attention kernel sources: ties a committed ncu capture to the build it was taken from."""
    h = hashlib.sha256()
    for name in sorted(glob.glob(os.path.join(ROOT, "uniception_b200", "csrc", "attention*.cu")) +
                       [os.path.join(ROOT, "uniception_b200", "csrc", "common.cuh")]):
        with open(name, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def _attn_tensor_pipe():
    """BASELINE.json's secondary metric (fused-attention tensor-pipe %): the newest committed ncu capture of the attention
    kernels at the encoder shape (never measured under the bench's own timing).  `current_build` says whether the capture's
    source hash equals the attention sources of THIS tree; a stale capture is reported as such, not silently reused."""
    f = _latest_profile("r0*_attn_tensor_pipe.json")
    try:
        with open(f) as fh:
            d = json.load(fh)
        out = {"metric": d["metric"], "source": os.path.relpath(f, ROOT) + " (ncu, not the timed run)",
               "kernels": {k: v.get("tensor_pipe_pct") for k, v in d["encoder_16x16x1024"].items()},
               "capture_source_hash": d.get("attn_source_hash"), "tree_source_hash": attn_source_hash()}
        out["current_build"] = out["capture_source_hash"] == out["tree_source_hash"]
        return out
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            busy = [s for s in sm if s > 0]
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def _synthetic_batch(B, S, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1)
    b = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1)
    return a, b


# ------------------------------------------------------------------------------------------------------------------
# The reference side (CPU arm and GPU-eager baseline).  Nothing here imports `uniception_b200`.
# ------------------------------------------------------------------------------------------------------------------
def _reference_root():
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("UC_REFERENCE_ROOT", "/root/reference")):
        if cand and os.path.isdir(os.path.join(cand, "uniception")):
            return cand
    return None


def build_reference_step(wl, device, B):
    """(step(img1, img2) -> loss, kind, note).  kind "reference": the UNMODIFIED castacks/UniCeption modules through their own
    public API (installed into baseline/_ref by `pip install --no-deps --target`); kind "port": the oracle
    restatement routed through the same torch functionals, random weights of the same architecture."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    S = wl["S"]
    root = _reference_root()
    torch.manual_seed(42)
    if root is not None:
        os.environ["UC_REFERENCE_ROOT"] = root
        import ref_import

        ref_import.REFERENCE_ROOT = root
        ref_import.import_reference()
        if wl["kind"] == "dust3r":
            from uniception.models.factory import DUSt3R

            m = DUSt3R(name="dust3r", img_size=(S, S), patch_embed_cls="PatchEmbedDust3R", pred_head_type=wl["head"]).to(device)
            inst1, inst2 = [str(2 * i) for i in range(B)], [str(2 * i + 1) for i in range(B)]

            def step(img1, img2):
                m.zero_grad(set_to_none=True)
                r1, r2 = m({"img": img1, "instance": inst1, "data_norm_type": "dust3r"},
                           {"img": img2, "instance": inst2, "data_norm_type": "dust3r"})
                loss = r1["pts3d"].sum() + r1["conf"].sum() + r2["pts3d_in_other_view"].sum() + r2["conf"].sum()
                loss.backward()
                return loss
        else:
            from uniception.models.encoders import ViTEncoderInput
            from uniception.models.encoders.croco import CroCoIntermediateFeatureReturner
            from uniception.models.prediction_heads.adaptors import DepthAdaptor
            from uniception.models.prediction_heads.base import AdaptorInput, PredictionHeadLayeredInput
            from uniception.models.prediction_heads.dpt import DPTFeature, DPTRegressionProcessor

            enc = CroCoIntermediateFeatureReturner(name="enc", data_norm_type="dust3r", img_size=(S, S), patch_size=14,
                                                   indices=[5, 11, 17, 23], intermediates_only=True).to(device)
            feat = DPTFeature(patch_size=14, hooks=[0, 1, 2, 3], input_feature_dims=[1024] * 4).to(device)
            reg = DPTRegressionProcessor(input_feature_dim=256, output_dim=1).to(device)
            ad = DepthAdaptor(name="depth", mode="exp")
            mods = torch.nn.ModuleList([enc, feat, reg])

            def step(img1, img2):
                mods.zero_grad(set_to_none=True)
                feats = [o.features.float() for o in enc(ViTEncoderInput(image=img1, data_norm_type="dust3r"))]
                with torch.autocast(device.type, enabled=False):  # heads in fp32, as factory/dust3r.py:309 runs them
                    raw = reg(feat(PredictionHeadLayeredInput(list_features=feats, target_output_shape=(S, S)))).decoded_channels
                    loss = ad(AdaptorInput(adaptor_feature=raw, output_shape_hw=(S, S))).value.sum()
                loss.backward()
                return loss
        return step, "reference", f"unmodified castacks/UniCeption 0.1.7 from {os.path.relpath(root, ROOT) if root.startswith(ROOT) else root}"
    # ---- port fallback ----
    import dust3r_oracle as O
    import model_shapes as MS

    shapes = MS.dust3r_shapes(wl["head"]) if wl["kind"] == "dust3r" else MS.c5_shapes()
    sd = {k: v.to(device).requires_grad_(True) for k, v in MS.random_init(shapes).items()}

    def step(img1, img2):
        for v in sd.values():
            v.grad = None
        with O.reference_functionals():
            if wl["kind"] == "dust3r":
                r1, r2 = O.dust3r_forward(sd, img1, img2, head=wl["head"])
                loss = O.bench_loss(r1, r2)
            else:
                _, inter = O.croco_encoder(sd, "encoder.", img1, 24, 16, 14, indices=[5, 11, 17, 23])
                with torch.autocast(device.type, enabled=False):
                    raw = O.dpt_regressor(sd, "dpt_regressor_head.", O.dpt_feature(sd, "dpt_feature_head.", [t.float() for t in inter]), (S, S))
                    loss = O.depth_adaptor(raw, "exp").sum()
        loss.backward()
        return loss

    return step, "port", "oracle/dust3r_oracle.py under reference_functionals() (reference package not present)"


def cpu_reference(wl, steps: int, warmup: int, budget_s: float):
    """The reference path on the host CPUs, fp32, all host threads, 1 unit per step.  Returns a dict."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device("cpu")
    step, kind, note = build_reference_step(wl, dev, 1)
    a, b = _synthetic_batch(1, wl["S"], 1234)
    t0 = time.time()
    step(a, b)  # first step doubles as warm-up / calibration
    t_first = time.time() - t0
    spent = t_first
    for _ in range(max(0, warmup - 1)):
        if spent + t_first > 0.3 * budget_s:
            break
        t0 = time.time()
        step(a, b)
        spent += time.time() - t0
    n = max(1, min(steps, int(max(0.0, budget_s - spent) / max(t_first, 1e-3))))
    ts = []
    for _ in range(n):
        t0 = time.time()
        step(a, b)
        ts.append(time.time() - t0)
    t = statistics.mean(ts)
    unit = wl["unit"].split("/")[0].rstrip("s")
    return {"value": 1.0 / t, "cores": cores, "kind": kind, "sec_per_step": t, "steps_run": n,
            "sample": f"1 {unit} {wl['S']}x{wl['S']} fwd+bwd per step, {n} timed step(s) after warm-up, fp32, "
                      f"{torch.get_num_threads()} threads; {note}"}


def gpu_eager_reference(wl, device, B, steps=4, warmup=2):
    """The reference's own eager GPU path under bf16 autocast on THIS GPU (same batch as the product arm)."""
    import torch

    step, kind, note = build_reference_step(wl, device, B)
    a, b = _synthetic_batch(B, wl["S"], 1234)
    a, b = a.to(device), b.to(device)

    def run():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return step(a, b)

    for _ in range(warmup):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    sdp = {"flash": torch.backends.cuda.flash_sdp_enabled(), "mem_efficient": torch.backends.cuda.mem_efficient_sdp_enabled(),
           "cudnn": torch.backends.cuda.cudnn_sdp_enabled(), "math": torch.backends.cuda.math_sdp_enabled()}
    return {"value": B / (ms / 1e3), "unit": wl["unit"], "ms_per_step": ms, "kind": kind, "units_per_step": B, "steps": steps,
            "warmup": warmup, "dtype": "bf16 autocast (heads fp32)",
            "attention": "F.scaled_dot_product_attention, torch's own backend choice; enabled backends: " +
                         ",".join(k for k, v in sdp.items() if v),
            "rope": "PyTorch RoPE2D fallback (one host sync per call)" if kind == "reference" else "reference fallback formula (one host sync per call)",
            "note": note}


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import contextlib

    with contextlib.redirect_stdout(sys.stderr):  # the reference prints while it builds its modules; stdout carries ONE JSON line
        r = cpu_reference(wl, args.steps, args.warmup, budget_s=170.0)
    assert "uniception_b200" not in sys.modules, "the reference arm must not load the product package"
    v = r["value"]
    line = {
        "impl": "reference", "metric": wl["metric"], "value": v, "unit": wl["unit"],
        "n_gpus": args.gpus, "steps": r["steps_run"], "steps_requested": args.steps, "warmup": args.warmup,
        "ms_per_step": r["sec_per_step"] * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "workload_key": args.workload, "reference_sample": r["sample"]},
        "cpu_baseline": {"value": v, "unit": wl["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": v, "unit": wl["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------------------
# The product arm
# ------------------------------------------------------------------------------------------------------------------
def run_b200(args, wl):
    import torch
    import torch.distributed as dist

    import uniception_b200 as U
    from uniception_b200 import _lib, dp, ops

    if args.graph and args.graph_multi and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # capturing NCCL collectives into a CUDA graph: the process group's async error watchdog must not poll events of a
        # capturing stream (torch CUDA-graphs notes, "Usage with DistributedDataParallel")
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")
        os.environ.setdefault("NCCL_ASYNC_ERROR_HANDLING", "0")
    rank, local, world = dp.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S, B = wl["S"], args.pairs_per_gpu
    torch.manual_seed(42)
    if wl["kind"] == "dust3r":
        model = U.DUSt3R(name="dust3r", img_size=(S, S), pred_head_type=wl["head"]).to(dev)
    else:
        model = U.ViTDPTDepth(img_size=(S, S)).to(dev)
    pk = model.pack()
    if world > 1:  # identical weights on every rank, then overlapped gradient all-reduce
        # communicator creation makes NCCL print its version banner on stdout: keep stdout to the ONE JSON line
        sys.stdout.flush()
        keep = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.broadcast(pk.flat, src=0)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(keep, 1)
            os.close(keep)
        if args.dp_mode != "none":  # "none": diagnostic, ranks run independently (no gradient exchange)
            pk.grad_sync = dp.GradSync(pk.flat_grad, pk.index, max_bucket_elems=args.bucket_mb * 1024 * 1024 // 4,
                                       compress_bf16=bool(args.dp_bf16))
    a_host, b_host = _synthetic_batch(B, S, 1234 + rank)
    a_host, b_host = a_host.pin_memory(), b_host.pin_memory()
    a_dev, b_dev = a_host.to(dev), b_host.to(dev)
    two_inputs = wl["kind"] == "dust3r"
    inst1 = [str(2 * i) for i in range(B)]
    inst2 = [str(2 * i + 1) for i in range(B)]

    def step(img1, img2):
        pk.zero_grad()
        if two_inputs:
            r1, r2 = model({"img": img1, "instance": inst1, "data_norm_type": "dust3r"},
                           {"img": img2, "instance": inst2, "data_norm_type": "dust3r"})
            loss = r1["pts3d"].sum() + r1["conf"].sum() + r2["pts3d_in_other_view"].sum() + r2["conf"].sum()
        else:
            loss = model(img1).sum()
        if args.dp_mode == "end" and pk.grad_sync is not None:
            with pk.grad_sync.no_sync():  # diagnostic: nothing overlaps, one all-reduce of the whole buffer after the backward
                loss.backward()
            pk.grad_sync.finish()
        else:
            loss.backward()
            if pk.grad_sync is not None:
                pk.grad_sync.finish()
        return loss

    graph = None
    static_loss = None

    def step_resident():
        if graph is not None:
            graph.replay()
            return static_loss
        return step(a_dev, b_dev)

    def step_e2e():
        if graph is not None:  # the graph reads the resident buffers: this step's host batch is copied into them first
            a_dev.copy_(a_host, non_blocking=True)
            if two_inputs:
                b_dev.copy_(b_host, non_blocking=True)
            graph.replay()
            return float(static_loss.item())
        img1 = a_host.to(dev, non_blocking=True)
        img2 = b_host.to(dev, non_blocking=True) if two_inputs else None
        return float(step(img1, img2).item())  # device -> host read of the step's result

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(ms.item())

    n_warm = max(3, args.warmup)
    n0 = _lib.launch_count()
    for _ in range(n_warm):
        step_resident()
    launches_per_step = (_lib.launch_count() - n0) // n_warm  # host-side count of libuc_b200 kernel launches
    graph_note = "eager launches (2 decoder view streams, PDL)"
    if args.graph and (world == 1 or args.graph_multi):
        # One CUDA graph of the whole step (zero grads, forward, loss, backward, and -- for N > 1 -- the bucketed NCCL
        # all-reduces on their side stream): every launch and every programmatic-dependent-launch edge is replayed
        # without host work.  Falls back to eager launches if capture is refused.
        try:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss_g = step(a_dev, b_dev)
            graph, static_loss = g, loss_g
            for _ in range(2):
                step_resident()
            torch.cuda.synchronize()
            graph_note = "one CUDA graph per step (same kernels, side streams, PDL edges" + (", NCCL all-reduce buckets" if world > 1 else "") + " captured)"
        except Exception as exc:  # pragma: no cover
            graph, static_loss = None, None
            torch.cuda.synchronize()
            graph_note = f"eager launches (CUDA graph capture refused: {type(exc).__name__}: {str(exc)[:120]})"
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(step_resident, args.steps)
    launches = launches_per_step * args.steps  # the graph replays exactly the launches counted above
    clocks = sampler.stop() if sampler else {}
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel (tcgen05 GEMM), CUDA events on the launching stream ----
    # Each GEMM is timed ALONE on its stream: the per-view decoder streams and the column-sum side stream are switched off for the
    # instrumented steps, otherwise an event pair around one view's GEMM also spans the other view's kernels that share the GPU
    # with it (the decoder shapes read 2x too slow, the earlier lines' 0.60-0.63).  The timed region above keeps all streams on.
    from uniception_b200 import engine as _E

    _vs, _cs = _E.ViewStreams.enabled, _E._ColsumSide.enabled
    _E.ViewStreams.enabled = _E._ColsumSide.enabled = False
    ops.PROFILE = []
    try:
        for _ in range(0 if args.skip_instrumented else 2):
            step(a_dev, b_dev)  # eager: per-launch CUDA events
        torch.cuda.synchronize()
    finally:
        _E.ViewStreams.enabled, _E._ColsumSide.enabled = _vs, _cs
    gemm_ms = sum(rec[0].elapsed_time(rec[1]) for rec in ops.PROFILE)
    gemm_flop = sum(rec[2] for rec in ops.PROFILE)
    n_gemm = len(ops.PROFILE)
    if args.gemm_breakdown and rank == 0:
        by = {}
        for rec in ops.PROFILE:
            t = by.setdefault(rec[3], [0, 0.0, 0.0])
            t[0] += 1
            t[1] += rec[0].elapsed_time(rec[1])
            t[2] += rec[2]
        for key, (cnt, t_ms, fl) in sorted(by.items(), key=lambda kv: -kv[1][1]):
            print(f"[gemm] m,n,k,aL,bL,epi={key}: n={cnt} total {t_ms:.2f} ms  avg {t_ms/cnt*1e3:.1f} us  {fl/t_ms/1e9:.0f} TFLOP/s", file=sys.stderr)
    ops.PROFILE = None
    sustained, burst, peak_src = _peaks()

    if rank == 0:
        units_per_s = world * B * args.steps / (ms / 1e3)
        e2e_units = world * B * args.steps / (ms_e2e / 1e3)
        achieved = gemm_flop / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        flop_unit = wl["flop"]
        traffic, traffic_src = _gemm_traffic()
        n_img = 2 if two_inputs else 1
        line = {
            "metric": wl["metric"], "value": units_per_s, "unit": wl["unit"],
            "n_gpus": world, "steps": args.steps, "warmup": n_warm, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": wl["desc"], "workload_key": args.workload,
                       "units_per_gpu": B, "global_units": B * world, "tokens_per_view": (S // (14 if wl["kind"] == "depth" else 16)) ** 2,
                       "parallelism": f"dp{world}",
                       "l2": "activations per step (several GB) far exceed the 126 MB L2; no flush needed",
                       "grad_allreduce": (f"flat {'bf16-compressed' if args.dp_bf16 else 'fp32'} buffer, {args.bucket_mb} MB buckets "
                                          "overlapped with backward") if world > 1 else "none",
                       "launch": graph_note},
            "e2e": {"value": e2e_units, "unit": wl["unit"], "h2d_bytes_per_step": int(a_host.numel() * 4 * n_img),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches * world),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                         "frac": achieved / sustained if sustained else None, "traffic": traffic,
                         "traffic_note": f"avg DRAM bytes per uc_gemm launch (ncu, {traffic_src}); "
                                         "algorithmic FLOPs per launch = 2*m*n*k, avg %.3e" % (gemm_flop / max(n_gemm, 1)),
                         "kernel": "uc::gemm2_kernel<EPI,F32,BN> (CTA-pair tcgen05.mma cta_group::2 kind::f16, TMA, TMEM double-buffered "
                                   "epilogue) + uc::gemm_kernel<BN> for narrow n",
                         "how": f"sum of algorithmic FLOPs over {n_gemm} uc_gemm / uc_conv3x3 / uc_patch_embed launches of 2 instrumented steps / sum of their "
                                "CUDA-event durations on the launching stream (instrumented steps run single-stream, eager, so that every "
                                "launch is timed alone)",
                         "peak_source": peak_src, "frac_of_burst_peak": achieved / burst if burst else None,
                         "step_tflops": (units_per_s / world) * flop_unit / 1e12,
                         "step_frac_of_peak": (units_per_s / world) * flop_unit / 1e12 / sustained},
        }
        line["attn_tensor_pipe"] = _attn_tensor_pipe()
        import contextlib

        if world == 1 and not args.no_gpu_eager_baseline:
            del graph
            try:
                with contextlib.redirect_stdout(sys.stderr):
                    line["gpu_eager_baseline"] = gpu_eager_reference(wl, dev, B)
                line["gpu_eager_baseline"]["speedup_value_over_eager"] = units_per_s / line["gpu_eager_baseline"]["value"]
            except Exception as exc:  # pragma: no cover
                line["gpu_eager_baseline"] = {"unavailable": f"{type(exc).__name__}: {str(exc)[:200]}"}
        if not args.no_cpu_baseline and world == 1:
            with contextlib.redirect_stdout(sys.stderr):
                r = cpu_reference(wl, 1, 0, budget_s=25.0)
            line["cpu_baseline"] = {"value": r["value"], "unit": wl["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        print(json.dumps(line), flush=True)
    if world > 1:
        _shutdown_distributed(graph)


def _shutdown_distributed(graph) -> None:
    """Leave a multi-rank run without hanging.  A CUDA graph that captured NCCL all-reduces keeps references into the
    communicator: ncclCommDestroy underneath `destroy_process_group()` was observed to block forever after such a capture
    (both ranks had printed / finished; the launcher had to be killed).  So: drain the device, drop the graph, rendezvous,
    then try the graceful teardown on a helper thread and leave the process regardless after a few seconds -- every result has
    been printed and flushed by then."""
    import threading

    import torch
    import torch.distributed as dist

    torch.cuda.synchronize()
    if graph is not None:
        try:
            graph.reset()
        except Exception:  # pragma: no cover
            pass
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    sys.stderr.flush()
    t = threading.Thread(target=dist.destroy_process_group, daemon=True)
    t.start()
    t.join(timeout=10.0)
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--size", type=int, default=None, help="legacy alias: 512 -> linear512, 224 -> linear224")
    ap.add_argument("--pairs-per-gpu", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="1 (default): replay the step from one captured CUDA graph; 0: eager")
    ap.add_argument("--graph-multi", type=int, default=1, help="capture the step (incl. NCCL buckets) when N > 1 too")
    ap.add_argument("--bucket-mb", type=int, default=128, help="gradient all-reduce bucket size (N > 1)")
    ap.add_argument("--dp-bf16", type=int, default=0, help="1: all-reduce bf16-compressed gradient buckets (halves NVLink bytes)")
    ap.add_argument("--dp-mode", default="overlap", choices=["overlap", "end", "none"],
                    help="overlap (default): buckets all-reduced as the backward retires them; end / none: diagnostics")
    ap.add_argument("--skip-instrumented", action="store_true", help="skip the per-GEMM CUDA-event pass (roofline block reports 0)")
    ap.add_argument("--gemm-breakdown", action="store_true", help="print per-shape GEMM timings of the instrumented steps to stderr")
    args = ap.parse_args()
    if args.workload is None:
        args.workload = {None: "linear512", 512: "linear512", 224: "linear224", 518: "c5_518"}[args.size]
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_b200(args, wl)


if __name__ == "__main__":
    main()
