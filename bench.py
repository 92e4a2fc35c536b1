#!/usr/bin/env python
"""Headline benchmark: SAM ViT-B image encoder + mask decoder at 1024x1024, images/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32] [--model vit_b]

A "step" is one pass of the hot path over one batch of synthetic images on every rank: ImageEncoderViT forward
followed by MaskDecoder forward with one box-prompt-shaped prompt set per image (configs[1] of BASELINE.json: batch
32 per GPU).  N > 1 is launched by torchrun, one rank per GPU; the batch is sharded (weak scaling: 32 images per
rank), weights are replicated, and the only collective is one NCCL all-gather of the mask logits and IoU scores per
step.  Rank 0 prints ONE JSON line.

  value     whole-job images/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the public API with HOST (pinned) buffers: every step's H2D image copy and
            D2H mask/IoU read are inside the timed region (copies of step s+1 overlap the compute of step s)
  roofline  the dominant kernel family (the tcgen05 GEMM behind qkv / proj / MLP), timed live with CUDA events the
            library records between launches on the launching stream, against the measured cuBLAS bf16 peak
  cpu_baseline  the CPU oracle (a port of the reference's PyTorch path) timed on this box's host cores, rank 0, N=1

--impl reference times the reference's CPU implementation of the path (the oracle port: the reference is pure
Python and /root/reference does not exist on the GPU box) on the host cores; each step is ONE image (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly ONE JSON line: keep NCCL's version / debug banner off it
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

METRIC = "images/sec/GPU SAM ViT-B enc+dec @1024² bf16; tensor-pipe util %; 1/2/4/8 GPU"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_burst": p["bf16_tflops"], "bf16_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_burst": 1590.0, "bf16_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def flops_per_image(cfg):
    """Algorithmic (useful) FLOPs per image and per kernel category; SURVEY.md §8(d) / BASELINE.md §2."""
    D, T, mlp = cfg.embed_dim, 4096, int(cfg.embed_dim * cfg.mlp_ratio)
    n_glob = len(cfg.global_attn_indexes)
    n_win = cfg.depth - n_glob
    heads = cfg.num_heads
    f = {
        "patch_embed": 2.0 * T * 768 * D,
        "gemm_qkv": cfg.depth * 2.0 * T * D * 3 * D,
        "gemm_proj": cfg.depth * 2.0 * T * D * D,
        "gemm_mlp_lin1": cfg.depth * 2.0 * T * D * mlp,
        "gemm_mlp_lin2": cfg.depth * 2.0 * T * D * mlp,
        # QK^T + PV (4*T*T*64 per head) + the two rel-pos contractions (2 * 2*T*64*64 per head)
        "attn_global": n_glob * heads * (4.0 * T * T * 64 + 4.0 * T * 64 * 64),
        # per window: 196 real keys; useful query rows = 4096 per image
        "attn_windowed": n_win * heads * (4.0 * T * 196 * 64 + 4.0 * T * 14 * 64),
        "neck": 2.0 * T * D * 256 + 2.0 * T * 9 * 256 * 256,
        "decoder": 3.62e9,
    }
    f["total"] = sum(f.values())
    return f


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.thread = [], None, None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                          str(index), "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.3] or [r for _, r in self.rows]
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except (ValueError, IndexError):
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def workload_config(model: str, B: int, world: int) -> dict:
    """The `config` object of the JSON line (shared by the CUDA arm and the reference arm)."""
    return {"workload": f"SAM {model} image encoder + mask decoder, batch {B} of 1024x1024 per GPU, 1 box-shaped "
                        "prompt per image, random-init weights (configs[1] of BASELINE.json)",
            "images_per_gpu_per_step": B, "global_batch": world * B,
            "parallelism": f"dp{world} (batch-sharded, weights replicated)",
            "collective": "all_gather(mask logits, iou) per step" if world > 1 else "none",
            "l2": "inputs (403 MB of images per step) and activations exceed the 126 MB L2; no flush needed",
            "residual_stream": "fp32", "gemm_operands": "bf16", "accumulate": "fp32"}


def run_reference(args):
    """The reference's CPU implementation of the path (oracle port), all host threads, one image per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle import sam_oracle as O
    from ullsam_b200 import synth

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = O.config(args.model)
    sd = synth.make_state_dict(cfg, 1234)
    images, _ = synth.make_images(1, 1024, seed=7)
    sparse, pe, dvec = synth.make_prompt_embeddings(1)
    dense = dvec.reshape(1, 256, 1, 1).expand(1, 256, 64, 64)

    def step():
        with torch.no_grad():
            emb = O.encoder_forward(sd, images, cfg)
            O.decoder_forward(sd, emb, pe, sparse, dense, False, cfg)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps / dt
    sample = "1 image 1024x1024 + 1 box-shaped prompt per step, fp32, torch CPU threads = all cores"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the same workload as the CUDA arm (its `config`); a step here is a bounded sample of it: ONE image + prompt
        "config": dict(workload_config(args.model, args.batch, args.gpus),
                       reference_arm="the reference's CPU path (fp32 oracle port of its PyTorch modules, all host threads); "
                                     "each step is a bounded sample of the workload: 1 image + 1 prompt"),
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--model", default="vit_b", choices=["vit_b", "vit_l"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from ullsam_b200 import synth
    from ullsam_b200.config import config as model_config
    from ullsam_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = model_config(args.model)
    B = args.batch
    eng = Engine(model=args.model, device=local)
    eng.load_state_dict(synth.make_state_dict(cfg, 1234))
    assert eng.missing(0) == 0 and eng.missing(1) == 0

    # synthetic microscopy-shaped inputs: a few distinct images tiled to the batch (generation is CPU-bound)
    base, _ = synth.make_images(4, 1024, seed=7 + rank)
    images_host = base.repeat((B + 3) // 4, 1, 1, 1)[:B].contiguous().pin_memory()
    sparse_h, pe_h, dvec_h = synth.make_prompt_embeddings(B, Ns=2, seed=13 + rank)
    images = images_host.to(dev)
    sparse, pe, dvec = sparse_h.to(dev), pe_h.to(dev), dvec_h.to(dev)[None]
    gathered_m = torch.empty((world * B, 1, 256, 256), device=dev) if world > 1 else None
    gathered_s = torch.empty((world * B, 1), device=dev) if world > 1 else None

    def step(img):
        emb = eng.encoder_forward(img)
        masks, iou = eng.decoder_forward(emb, pe, sparse, dvec, False, prompts_per_image=1)
        if world > 1:
            dist.all_gather_into_tensor(gathered_m, masks)
            dist.all_gather_into_tensor(gathered_s, iou)
        return masks, iou

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident timing
    for _ in range(args.warmup):
        step(images)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    eng.profile_begin()
    ev0.record()
    for _ in range(args.steps):
        step(images)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    prof = eng.profile_end()
    launches = eng.launch_count() - launches0
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = ms.item()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    value = world * B * args.steps / (ms_total / 1e3)

    # ------------------------------------------------------------------ end-to-end through the public API, host buffers
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream(dev)
        dev_in = [torch.empty_like(images), torch.empty_like(images)]
        host_m = [torch.empty((B, 1, 256, 256)).pin_memory() for _ in range(2)]
        host_s = [torch.empty((B, 1)).pin_memory() for _ in range(2)]
        in_ready = [torch.cuda.Event(), torch.cuda.Event()]
        in_free = [torch.cuda.Event(), torch.cuda.Event()]
        out_ready = [torch.cuda.Event(), torch.cuda.Event()]
        main_stream = torch.cuda.current_stream(dev)

        def h2d(s):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(in_free[s & 1])
                dev_in[s & 1].copy_(images_host, non_blocking=True)
                in_ready[s & 1].record(copy_stream)

        def e2e_loop(n):
            for b in range(2):
                in_free[b].record(main_stream)
            h2d(0)
            pending = []
            for s in range(n):
                if s + 1 < n:
                    h2d(s + 1)
                main_stream.wait_event(in_ready[s & 1])
                masks, iou = step(dev_in[s & 1])
                in_free[s & 1].record(main_stream)
                out_ready[s & 1].record(main_stream)
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(out_ready[s & 1])
                    host_m[s & 1].copy_(masks, non_blocking=True)
                    host_s[s & 1].copy_(iou, non_blocking=True)
                    masks.record_stream(copy_stream)
                    iou.record_stream(copy_stream)
                pending.append(s)
            torch.cuda.synchronize()

        e2e_loop(2)
        barrier()
        t0 = time.perf_counter()
        e2e_loop(args.steps)
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * B * args.steps / dt.item(), "unit": "images/s",
               "h2d_bytes_per_step": images_host.numel() * 4, "d2h_bytes_per_step": (host_m[0].numel() + host_s[0].numel()) * 4,
               "api": "Engine.encoder_forward + Engine.decoder_forward on pinned host tensors, copies double-buffered"}

    # ------------------------------------------------------------------ roofline of the dominant kernel family
    peaks = load_peaks()
    fl = flops_per_image(cfg)
    n_img = B * args.steps
    gemm_cats = ["gemm_qkv", "gemm_proj", "gemm_mlp_lin1", "gemm_mlp_lin2"]
    gemm_ms = sum(prof[c][0] for c in gemm_cats)
    gemm_launches = sum(prof[c][1] for c in gemm_cats)
    gemm_flops = sum(fl[c] for c in gemm_cats) * n_img
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_detail = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.model == "vit_b" and B == 32:  # the capture is of this exact workload
        with open(tpath) as f:
            traffic_detail = json.load(f)
        traffic = traffic_detail.get("gemm_mlp_lin1_bytes_per_launch")
    kernels = {}
    dec_ms = sum(prof[c][0] for c in prof if c.startswith("dec_"))
    for c, (t, n) in prof.items():
        if c.startswith("dec_") or n == 0:
            continue
        key = c
        kernels[key] = {"ms_per_step": t / args.steps, "launches_per_step": n / args.steps,
                        "tflops": (fl[key] * n_img / (t / 1e3) / 1e12) if key in fl and t > 0 else None}
    kernels["decoder_parts"] = {c: {"ms_per_step": prof[c][0] / args.steps, "launches_per_step": prof[c][1] / args.steps}
                                for c in prof if c.startswith("dec_")}
    kernels["decoder"] = {"ms_per_step": dec_ms / args.steps, "launches_per_step": sum(prof[c][1] for c in prof if c.startswith("dec_")) / args.steps,
                          "tflops": fl["decoder"] * n_img / (dec_ms / 1e3) / 1e12 if dec_ms > 0 else None}
    roofline = {
        "bound": "tensor", "kernel": "gemm_kernel<BN,EPI> (tcgen05): qkv + proj + mlp.lin1 + mlp.lin2 launches",
        "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
        "frac": achieved / peaks["bf16_sustained"] if achieved else None, "traffic": traffic,
        "traffic_note": "DRAM read+write bytes of ONE mlp.lin1 launch (largest GEMM of the family; algorithmic 1011 MB) from "
                        "ncu --set full; per-kernel figures in traffic_detail / profiles/traffic.json",
        "traffic_detail": traffic_detail,
        "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
        "launches": gemm_launches, "ms_per_launch": gemm_ms / max(gemm_launches, 1),
        "flops_per_launch": gemm_flops / max(gemm_launches, 1),
        "step": {"achieved": value / world * fl["total"] / 1e12, "frac_of_sustained": value / world * fl["total"] / 1e12 / peaks["bf16_sustained"],
                 "frac_of_burst": value / world * fl["total"] / 1e12 / peaks["bf16_burst"], "flops_per_image": fl["total"]},
        "kernels": kernels,
    }
    # the single kernel symbol with the largest share of the step (profiles/r1g_launches.csv) is the fused global
    # attention, MUFU-co-bound (softmax of d = 64 heads); reported next to the GEMM family so neither hides the other
    ag = kernels.get("attn_global")
    if ag and ag.get("tflops"):
        roofline["largest_single_kernel"] = {
            "kernel": "attention_kernel<64> (fused global attention with decomposed rel-pos bias)",
            "share_of_step": ag["ms_per_step"] / (ms_total / args.steps), "achieved": ag["tflops"], "unit": "TFLOP/s",
            "frac": ag["tflops"] / peaks["bf16_sustained"], "ms_per_launch": ag["ms_per_step"] / max(ag["launches_per_step"], 1e-9),
            "note": "exp2 on the MUFU pipe bounds it near 50 % tensor utilisation (DESIGN.md section 4)"}

    # ------------------------------------------------------------------ CPU baseline (oracle port), rank 0 at N == 1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import sam_oracle as O

        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd = synth.make_state_dict(cfg, 1234)
        img1 = images_host[:1]
        dense1 = dvec_h.reshape(1, 256, 1, 1).expand(1, 256, 64, 64)
        n_cpu = 2
        t0 = time.perf_counter()
        with torch.no_grad():
            for _ in range(n_cpu):
                emb = O.encoder_forward(sd, img1, cfg)
                O.decoder_forward(sd, emb, pe_h, sparse_h[:1], dense1, False, cfg)
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": "images/s", "cores": cores, "kind": "port",
               "sample": f"{n_cpu} x (1 image 1024x1024 + 1 prompt), fp32 oracle port of the reference, torch threads = {cores}"}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "images/s", "per_gpu": value / world, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.model, B, world),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
