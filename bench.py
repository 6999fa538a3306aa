"""Benchmark of the north-star metric: image-pairs/sec (forward + backward) of DUSt3R ViT-L/16 +
12-layer two-view decoder + linear pointmap head at 512x512 (BASELINE.json configs[2]/[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--pairs-per-gpu 8] [--size 512] [--impl b200|reference]

One process per GPU (the driver launches torchrun for N > 1); weak scaling: 8 pairs per GPU.  A step =
zero grads -> forward -> loss (sum of the four outputs, the `.sum().backward()` idiom of the
reference's encoders/utils.py:29-31) -> backward (-> overlapped NCCL all-reduce of the flat gradient
buffer when N > 1).  Prints ONE JSON line on rank 0.

  value     pairs/s with the image batch already resident in HBM
  e2e       same metric through the public `DUSt3R.forward(view1, view2)` call with HOST inputs:
            pinned-host -> device copy of both image batches and a device -> host read of the loss
            inside the timed region, every step
  roofline  dominant kernel = the tcgen05 GEMM (82 % of the path's FLOPs): sum of its algorithmic
            FLOPs / sum of its CUDA-event durations over instrumented steps, vs the measured cuBLAS
            bf16 peak in MEASURED_PEAKS.json (sustained figure: the kernel is timed inside a long step)
  cpu_baseline / --impl reference
            the oracle port of the reference path (oracle/dust3r_oracle.py, fp32, all host threads) on a
            bounded sample (1 pair at 512x512, forward + backward)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_PAIR = {512: 6.2137e12, 224: 1.0218e12}  # fwd+bwd, SURVEY.md 8d / BASELINE.md section 3


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["bf16_tflops_sustained"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1400.0, 1590.0, "fallback (B200_PROFILING.md)"


def _gemm_traffic():
    """DRAM bytes per uc_gemm launch (read + write, averaged over the launches of one step) from the committed ncu pass
    profiles/r01i_dram_traffic_per_kernel.json (dram__bytes_read.sum + dram__bytes_write.sum); None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01i_dram_traffic_per_kernel.json")) as f:
            return float(json.load(f)["gemm_dram_bytes_per_launch"])
    except Exception:
        return None


def _attn_tensor_pipe():
    """BASELINE.json's secondary metric (fused-attention tensor-pipe %): the committed ncu capture of the attention kernels
    (profiles/r01n_attn_tensor_pipe.json, sm__pipe_tensor_cycles_active of attn_fwd_kernel / attn_bwd_pipe_kernel at the
    encoder shape); never measured under the bench's own timing.  None if absent."""
    try:
        with open(os.path.join(ROOT, "profiles", "r01n_attn_tensor_pipe.json")) as f:
            d = json.load(f)
        enc = d["encoder_16x16x1024"]
        return {"fwd_pct": enc["attn_fwd_kernel"]["tensor_pipe_pct"], "bwd_pct": enc["attn_bwd_pipe_kernel"]["tensor_pipe_pct"],
                "metric": d["metric"], "source": "profiles/r01n_attn_tensor_pipe.json (ncu, not the timed run)"}
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


def _synthetic_pair_batch(B, S, seed):
    import torch

    g = torch.Generator().manual_seed(seed)
    a = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1)
    b = torch.randn(B, 3, S, S, generator=g).clamp_(-1, 1)
    return a, b


def cpu_oracle_pairs_per_sec(S: int, steps: int, warmup: int, budget_s: float):
    """Reference path on the host CPUs: fp32 oracle port, 1 pair per step, forward + backward."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import dust3r_oracle as O
    import uniception_b200 as U

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(42)
    m = U.DUSt3R(name="dust3r", img_size=(S, S))  # parameter container only (CPU); arithmetic below is the oracle's
    sd = {k: v.detach().requires_grad_(True) for k, v in m.state_dict().items()}
    a, b = _synthetic_pair_batch(1, S, 1234)

    def step():
        for v in sd.values():
            v.grad = None
        r1, r2 = O.dust3r_forward(sd, a, b)
        O.bench_loss(r1, r2).backward()

    t0 = time.time()
    step()  # first step doubles as warm-up / calibration
    t_first = time.time() - t0
    n = max(1, min(steps, int(max(0.0, budget_s - t_first) / max(t_first, 1e-3))))
    ts = []
    for _ in range(n):
        t0 = time.time()
        step()
        ts.append(time.time() - t0)
    t = statistics.mean(ts)
    return 1.0 / t, cores, f"1 pair {S}x{S} fwd+bwd per step, {n} timed step(s) after 1 warm-up, fp32, {torch.get_num_threads()} threads", t, n


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, cores, sample, t, n = cpu_oracle_pairs_per_sec(args.size, args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference", "metric": f"image-pairs/sec (fwd+bwd) DUSt3R ViT-L/16 {args.size}^2", "value": v, "unit": "pairs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"DUSt3R ViT-L/16 + 12-layer 2-view decoder + linear head, {args.size}x{args.size} pairs, fwd+bwd",
                   "pairs_per_step": 1, "timed_steps_run": n},
        "cpu_baseline": {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import torch
    import torch.distributed as dist

    import uniception_b200 as U
    from uniception_b200 import _lib, dp, ops

    rank, local, world = dp.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    S, B = args.size, args.pairs_per_gpu
    torch.manual_seed(42)
    model = U.DUSt3R(name="dust3r", img_size=(S, S)).to(dev)
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
        pk.grad_sync = dp.GradSync(pk.flat_grad, pk.index)
    a_host, b_host = _synthetic_pair_batch(B, S, 1234 + rank)
    a_host, b_host = a_host.pin_memory(), b_host.pin_memory()
    a_dev, b_dev = a_host.to(dev), b_host.to(dev)
    inst1 = [str(2 * i) for i in range(B)]
    inst2 = [str(2 * i + 1) for i in range(B)]

    def step(img1, img2):
        pk.zero_grad()
        r1, r2 = model({"img": img1, "instance": inst1, "data_norm_type": "dust3r"},
                       {"img": img2, "instance": inst2, "data_norm_type": "dust3r"})
        loss = r1["pts3d"].sum() + r1["conf"].sum() + r2["pts3d_in_other_view"].sum() + r2["conf"].sum()
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
            b_dev.copy_(b_host, non_blocking=True)
            graph.replay()
            return float(static_loss.item())
        img1 = a_host.to(dev, non_blocking=True)
        img2 = b_host.to(dev, non_blocking=True)
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

    n0 = _lib.launch_count()
    for _ in range(max(3, args.warmup)):
        step_resident()
    launches_per_step = (_lib.launch_count() - n0) // max(3, args.warmup)  # host-side count of libuc_b200 kernel launches
    graph_note = "eager launches (2 decoder view streams, PDL)"
    if args.graph and world == 1:
        # One CUDA graph of the whole step (zero grads, forward, loss, backward; both decoder view streams and every
        # programmatic-dependent-launch edge are captured): ~1550 launches replayed without host work.  Falls back to
        # eager launches if capture is refused.
        try:
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                loss_g = step(a_dev, b_dev)
            graph, static_loss = g, loss_g
            for _ in range(2):
                step_resident()
            torch.cuda.synchronize()
            graph_note = "one CUDA graph per step (same kernels, 2 decoder view streams and PDL edges captured)"
        except Exception as exc:  # pragma: no cover
            graph, static_loss = None, None
            torch.cuda.synchronize()
            graph_note = f"eager launches (CUDA graph capture refused: {type(exc).__name__})"
    sampler = ClockSampler(local) if rank == 0 else None
    ms = timed(step_resident, args.steps)
    launches = launches_per_step * args.steps  # the graph replays exactly the launches counted above
    clocks = sampler.stop() if sampler else {}
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # ---- roofline of the dominant kernel (tcgen05 GEMM), CUDA events on the launching stream ----
    ops.PROFILE = []
    for _ in range(2):
        step(a_dev, b_dev)  # eager: per-launch CUDA events
    torch.cuda.synchronize()
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
        pairs_per_s = world * B * args.steps / (ms / 1e3)
        e2e_pairs = world * B * args.steps / (ms_e2e / 1e3)
        achieved = gemm_flop / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        flop_pair = FLOP_PER_PAIR.get(S)
        line = {
            "metric": f"image-pairs/sec (fwd+bwd) DUSt3R ViT-L/16 {S}^2", "value": pairs_per_s, "unit": "pairs/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"DUSt3R ViT-L/16 + 12-layer 2-view decoder + linear head, {S}x{S} pairs, fwd+bwd",
                       "pairs_per_gpu": B, "global_pairs": B * world, "tokens_per_view": (S // 16) ** 2,
                       "parallelism": f"dp{world}",
                       "l2": ("activations per step (>10 GB) far exceed the 126 MB L2; no flush needed" if S >= 512 else
                              "activations per step (~3 GB at 224^2) exceed the 126 MB L2; no flush needed"),
                       "grad_allreduce": "flat fp32 buffer, per-block buckets overlapped with backward" if world > 1 else "none",
                       "launch": graph_note},
            "e2e": {"value": e2e_pairs, "unit": "pairs/s", "h2d_bytes_per_step": int(a_host.numel() * 4 * 2),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches * world),
            "clocks": clocks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": sustained, "unit": "TFLOP/s",
                         "frac": achieved / sustained if sustained else None, "traffic": _gemm_traffic(),
                         "traffic_note": "avg DRAM bytes per uc_gemm launch (ncu, profiles/r01i_dram_traffic_per_kernel.json); "
                                         "algorithmic FLOPs per launch = 2*m*n*k, avg %.3e" % (gemm_flop / max(n_gemm, 1)),
                         "kernel": "uc::gemm2_kernel<EPI,F32,BN> (CTA-pair tcgen05.mma cta_group::2 kind::f16, TMA, TMEM double-buffered "
                                   "epilogue) + uc::gemm_kernel<BN> for narrow n",
                         "how": f"sum of 2*m*n*k over {n_gemm} uc_gemm launches of 2 instrumented steps / sum of CUDA-event durations on the launching stream",
                         "peak_source": peak_src, "frac_of_burst_peak": achieved / burst if burst else None,
                         "step_tflops": (pairs_per_s / world) * flop_pair / 1e12 if flop_pair else None,
                         "step_frac_of_peak": (pairs_per_s / world) * flop_pair / 1e12 / sustained if flop_pair else None},
        }
        line["attn_tensor_pipe"] = _attn_tensor_pipe()
        if not args.no_cpu_baseline and world == 1:
            v, cores, sample, _t, _n = cpu_oracle_pairs_per_sec(S, 1, 0, budget_s=25.0)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--pairs-per-gpu", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=1, help="1 (default): replay the step from one captured CUDA graph (single GPU); 0: eager")
    ap.add_argument("--gemm-breakdown", action="store_true", help="print per-shape GEMM timings of the instrumented steps to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
