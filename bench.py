#!/usr/bin/env python
"""Headline benchmark: SAM image encoder + mask decoder at 1024x1024, images/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 32]
                    [--model vit_b|vit_l] [--workload box|llm_dense]

A "step" is one pass of the hot path over one batch of synthetic images on every rank: ImageEncoderViT forward followed
by MaskDecoder forward with one prompt set per image.  N > 1 is launched by torchrun, one rank per GPU; the batch is
sharded (weak scaling: --batch images per rank), weights are replicated, and the only collective is one NCCL all-gather
of the mask logits and IoU scores per step, issued on a side stream so it overlaps the next step's encoder.
Rank 0 prints ONE JSON line.

Workloads (BASELINE.json configs): `box` = configs[1]/[3]: a box-shaped prompt (2 sparse tokens) and the no-mask dense
embedding; `llm_dense` = configs[2]: 4 point tokens + padding (5 sparse tokens) and a per-image LLM-conditioned dense
prompt [256,64,64] (PromptEncoder._normalize_llm_hidden_states of a [B,256,64,64] map, inside the step).
`--model vit_l` is configs[3].

  value     whole-job images/s with the inputs already resident in HBM (CUDA events, max over ranks)
  e2e       the same metric through the public API with HOST (pinned) buffers: every step's H2D image copy and
            D2H mask/IoU read are inside the timed region (copies of step s+1 overlap the compute of step s);
            e2e.c_abi_host: the same through ONE synchronous C-ABI call per step (ullsam_segment_host)
  roofline  frac = the WHOLE STEP's algorithmic FLOP/s over the measured sustained bf16 peak (the north-star quantity);
            roofline.kernels = per-kernel-category times (CUDA events the library records between launches on the
            launching stream) with achieved TFLOP/s; roofline.gemm_family / roofline.dominant_kernel single them out
  latency_b1      one image + one prompt, CUDA-graph replay vs plain launches (every real caller runs batch 1)
  gpu_eager_bf16  the reference's own GPU recipe (its PyTorch modules, bf16 weights + autocast) on the SAME GPU
  cpu_baseline    the reference's modules (oracle/_ref, kind "reference"; else the oracle port) on this box's host cores

--impl reference times the reference's CPU implementation of the path on the host cores; each step is ONE image
(a bounded sample of the workload).
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


def workload_config(model: str, B: int, world: int, workload: str = "box") -> dict:
    """The `config` object of the JSON line (shared by the CUDA arm and the reference arm)."""
    which = {("vit_b", "box"): "configs[1]", ("vit_b", "llm_dense"): "configs[2]", ("vit_l", "box"): "configs[3]"}.get(
        (model, workload), "a variant of configs[1..3]")
    prompt = ("1 box-shaped prompt per image (2 sparse tokens, no-mask dense embedding)" if workload == "box" else
              "1 LLM-conditioned prompt per image (5 sparse point tokens + dense prompt [256,64,64] normalised from a "
              "synthetic LLM map inside the step)")
    return {"workload": f"SAM {model} image encoder + mask decoder, batch {B} of 1024x1024 per GPU, {prompt}, "
                        f"random-init weights ({which} of BASELINE.json)",
            "prompt_kind": workload, "images_per_gpu_per_step": B, "global_batch": world * B,
            "parallelism": f"dp{world} (batch-sharded, weights replicated)",
            "collective": "all_gather(mask logits, iou) per step on a side stream" if world > 1 else "none",
            "l2": "inputs (403 MB of images per step) and activations exceed the 126 MB L2; no flush needed",
            "residual_stream": "fp32", "gemm_operands": "bf16", "accumulate": "fp32"}


# ---------------------------------------------------------------------------------------------------------------------
# the reference on the host cores: its own modules from oracle/_ref when staged, else the oracle port
# ---------------------------------------------------------------------------------------------------------------------
def make_cpu_reference(model: str, workload: str):
    """-> (step_fn(), kind).  One step = one 1024x1024 image + one prompt through the reference's encoder, prompt
    encoder and mask decoder in fp32 (its CPU path)."""
    import torch

    from oracle import sam_oracle as O
    from ullsam_b200 import synth

    cfg = O.config(model)
    sd = synth.make_state_dict(cfg, 1234)
    images, boxes = synth.make_images(1, 1024, seed=7)
    pts, lab = synth.make_points(1, 1024, seed=8)
    llm = synth.make_llm_map(1, seed=9)
    try:
        from oracle import make_ref
        _, build_sam = make_ref.import_reference() if make_ref.available() else (None, None)
    except Exception:  # pragma: no cover - a broken staged copy must not take the bench down
        build_sam = None
    if build_sam is not None:
        torch.manual_seed(0)
        sam = build_sam.sam_model_registry[model](None)
        sam.load_state_dict(sd, strict=False)

        def step():
            with torch.no_grad():
                emb = sam.image_encoder(images)
                if workload == "box":
                    sp, de = sam.prompt_encoder(points=None, boxes=boxes, masks=None)
                else:
                    sp, de = sam.prompt_encoder(points=(pts, lab), boxes=None, masks=None, llm_hidden_states=llm)
                sam.mask_decoder(image_embeddings=emb, image_pe=sam.prompt_encoder.get_dense_pe(),
                                 sparse_prompt_embeddings=sp, dense_prompt_embeddings=de, multimask_output=False)
        return step, "reference"

    pe = O.dense_pe(sd, cfg)

    def step():
        with torch.no_grad():
            emb = O.encoder_forward(sd, images, cfg)
            if workload == "box":
                sp, de = O.embed_boxes(sd, boxes, cfg), O.dense_no_mask(sd, 1, cfg)
            else:
                sp, de = O.embed_points(sd, pts, lab, cfg), O.dense_from_llm(sd, llm)
            O.decoder_forward(sd, emb, pe, sp, de, False, cfg)
    return step, "port"


def run_reference(args):
    """The reference's CPU implementation of the path, all host threads, one image per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = make_cpu_reference(args.model, args.workload)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = args.steps / dt
    what = ("the reference's own PyTorch modules (unmodified files staged in oracle/_ref)" if kind == "reference" else
            "the fp32 oracle port of the reference's PyTorch modules (oracle/_ref not staged)")
    sample = f"1 image 1024x1024 + 1 prompt per step, fp32, torch CPU threads = {cores}: {what}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # exactly the CUDA arm's `config` (same workload); a step here is a bounded sample of it: ONE image + prompt
        "config": workload_config(args.model, args.batch, args.gpus, args.workload),
        "reference_arm": f"the reference's CPU path ({what}, all host threads); each step is a bounded sample of the "
                         "workload: 1 image + 1 prompt",
        "cpu_baseline": {"value": val, "unit": "images/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def gpu_eager_bf16(model: str, dev, n_img: int = 8):
    """SURVEY.md §8(d) "second comparison": the reference's GPU recipe — its PyTorch modules with bf16 weights under
    autocast (train_joint_v2.py:1665,1675-1676) — on the same B200, box workload, images/s."""
    import torch

    from oracle import sam_oracle as O
    from ullsam_b200 import synth

    cfg = O.config(model)
    sd = synth.make_state_dict(cfg, 1234)
    images, boxes = synth.make_images(n_img, 1024, seed=7)
    images, boxes = images.to(dev), boxes.to(dev)
    try:
        from oracle import make_ref
        _, build_sam = make_ref.import_reference() if make_ref.available() else (None, None)
    except Exception:  # pragma: no cover
        build_sam = None
    if build_sam is not None:
        torch.manual_seed(0)
        sam = build_sam.sam_model_registry[model](None)
        sam.load_state_dict(sd, strict=False)
        sam = sam.to(dev).to(torch.bfloat16).eval()
        kind = "reference modules (oracle/_ref), weights bf16 + autocast"

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                emb = sam.image_encoder(images.to(torch.bfloat16))
                pe = sam.prompt_encoder.get_dense_pe()
                for i in range(n_img):     # per-image prompts => the caller loops (sam.py:101)
                    sp, de = sam.prompt_encoder(points=None, boxes=boxes[i:i + 1].to(torch.bfloat16), masks=None)
                    sam.mask_decoder(image_embeddings=emb[i:i + 1], image_pe=pe, sparse_prompt_embeddings=sp,
                                     dense_prompt_embeddings=de, multimask_output=False)
    else:
        sdg = {k: v.to(dev) for k, v in sd.items()}
        pe = O.dense_pe(sdg, cfg)
        kind = "oracle port of the reference modules, fp32 weights + autocast (oracle/_ref not staged)"

        def run():
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
                emb = O.encoder_forward(sdg, images, cfg)
                for i in range(n_img):
                    O.decoder_forward(sdg, emb[i:i + 1].float(), pe, O.embed_boxes(sdg, boxes[i:i + 1], cfg),
                                      O.dense_no_mask(sdg, 1, cfg), False, cfg)
    run()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / reps
    return {"value": n_img / ms * 1e3, "unit": "images/s", "ms_per_image": ms / n_img, "images": n_img, "kind": kind}


def latency_b1(eng, images, pe, sparse, dense_arg, multimask, dev):
    """One image + one prompt: plain launches vs one CUDA-graph replay of the same launch sequence."""
    import torch

    img1, sp1 = images[:1].contiguous(), sparse[:1].contiguous()
    d1 = dense_arg[:1].contiguous() if dense_arg.shape[0] > 1 else dense_arg

    def one():
        emb = eng.encoder_forward(img1)
        return eng.decoder_forward(emb, pe, sp1, d1, multimask, prompts_per_image=1)

    out = {}
    n0 = eng.launch_count()
    one()
    out["kernels_per_image"] = eng.launch_count() - n0
    for _ in range(2):
        one()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    t0 = time.perf_counter()
    e0.record()
    for _ in range(reps):
        one()
    e1.record()
    torch.cuda.synchronize(dev)
    out["ms_plain_launches"] = e0.elapsed_time(e1) / reps
    out["ms_plain_launches_wall"] = 1e3 * (time.perf_counter() - t0) / reps
    try:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            one()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                res = one()
        torch.cuda.current_stream(dev).wait_stream(side)
        for _ in range(3):
            graph.replay()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(reps):
            graph.replay()
        e1.record()
        torch.cuda.synchronize(dev)
        out["ms_cuda_graph"] = e0.elapsed_time(e1) / reps
        del res
    except Exception as exc:  # pragma: no cover - reported, never fatal
        out["cuda_graph_error"] = f"{type(exc).__name__}: {exc}"[:300]
        torch.cuda.synchronize(dev)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per GPU per step")
    ap.add_argument("--model", default="vit_b", choices=["vit_b", "vit_l"])
    ap.add_argument("--workload", default="box", choices=["box", "llm_dense"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip latency_b1 and gpu_eager_bf16")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from ullsam_b200 import ops, synth
    from ullsam_b200.config import config as model_config
    from ullsam_b200.engine import Engine
    from ullsam_b200.parallel import AsyncGather

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
    llm_mode = args.workload == "llm_dense"
    eng = Engine(model=args.model, device=local)
    eng.load_state_dict(synth.make_state_dict(cfg, 1234))
    assert eng.missing(0) == 0 and eng.missing(1) == 0

    # synthetic microscopy-shaped inputs: a few distinct images tiled to the batch (generation is CPU-bound)
    base, _ = synth.make_images(4, 1024, seed=7 + rank)
    images_host = base.repeat((B + 3) // 4, 1, 1, 1)[:B].contiguous().pin_memory()
    sparse_h, pe_h, dvec_h = synth.make_prompt_embeddings(B, Ns=5 if llm_mode else 2, seed=13 + rank)
    images = images_host.to(dev)
    sparse, pe, dvec = sparse_h.to(dev), pe_h.to(dev), dvec_h.to(dev)[None]
    if llm_mode:
        # the [B,256,64,64] map InternVLSAMModel.text_aware_dense_feature hands to the prompt encoder: it is produced on
        # the device by the LLM side, so it is device-resident here as well (distinct maps per image)
        llm_map = synth.make_llm_map(4, seed=21 + rank).repeat((B + 3) // 4, 1, 1, 1)[:B].contiguous().to(dev)
        llm_scale, llm_bias = torch.full((1,), 0.1, device=dev), torch.zeros(1, device=dev)   # prompt_encoder.py:51-52
    gather = AsyncGather(dev) if world > 1 else None
    gathered_m = [torch.empty((world * B, 1, 256, 256), device=dev) for _ in range(2)] if world > 1 else None
    gathered_s = [torch.empty((world * B, 1), device=dev) for _ in range(2)] if world > 1 else None
    step_no = [0]

    def dense_for_step():
        # llm_dense: PromptEncoder._normalize_llm_hidden_states runs inside the step (its own kernel, N1)
        return ops.prompt_llm_dense(llm_map, llm_scale, llm_bias) if llm_mode else dvec

    def step(img):
        emb = eng.encoder_forward(img)
        masks, iou = eng.decoder_forward(emb, pe, sparse, dense_for_step(), False, prompts_per_image=1)
        if world > 1:
            par = step_no[0] & 1
            step_no[0] += 1
            gather.submit([masks, iou], [gathered_m[par], gathered_s[par]])
        return masks, iou

    def barrier():
        if world > 1:
            gather.wait()
            dist.barrier()
        torch.cuda.synchronize()

    # ------------------------------------------------------------------ device-resident timing
    for _ in range(args.warmup):
        step(images)
    barrier()
    if gather:
        gather.comm_ms()
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    for _ in range(args.steps):
        step(images)
    if gather:
        gather.wait()     # the timed region ends when the last gather has landed
    ev1.record()
    barrier()
    t_wall1 = time.time()
    launches = eng.launch_count() - launches0 + (args.steps if llm_mode else 0)
    # Per-kernel breakdown: a SECOND pass of the same K steps with a CUDA event after every launch.  It is kept out of the
    # timed region above because an event between two kernels undoes their programmatic dependent launch (the next grid's
    # prologue no longer overlaps the previous grid's tail), i.e. the instrumented pass is slightly slower than the real one.
    pv0, pv1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.profile_begin()
    pv0.record()
    for _ in range(args.steps):
        step(images)
    if gather:
        gather.wait()
    pv1.record()
    barrier()
    prof = eng.profile_end()
    ms_prof_total = pv0.elapsed_time(pv1)
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    comm = torch.tensor([gather.comm_ms() if gather else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(comm, op=dist.ReduceOp.MAX)
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
            if gather:
                gather.wait()
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
        if not llm_mode:
            # the same through ONE synchronous C-ABI call per step (include/ullsam.h: ullsam_segment_host), host in, host out
            pe_host = pe_h.reshape(-1).contiguous().pin_memory()
            sp_host = sparse_h.contiguous().pin_memory()
            dv_host = dvec_h.contiguous().pin_memory()
            for _ in range(2):
                eng.segment_host(images_host, pe_host, sp_host, dv_host, False, host_m[0], host_s[0])
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                eng.segment_host(images_host, pe_host, sp_host, dv_host, False, host_m[0], host_s[0])
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            e2e["c_abi_host"] = {"value": world * B * args.steps / dt.item(), "unit": "images/s",
                                 "api": "ullsam_segment_host: one synchronous call per step, sub-batches of 4 / 12 / 16 images pipelined on "
                                        "two internal streams (no gather)"}
            barrier()

    # ------------------------------------------------------------------ roofline: the whole step, then its parts
    peaks = load_peaks()
    fl = flops_per_image(cfg)
    n_img = B * args.steps
    gemm_cats = ["gemm_qkv", "gemm_proj", "gemm_mlp_lin1", "gemm_mlp_lin2"]
    gemm_ms = sum(prof[c][0] for c in gemm_cats)
    gemm_launches = sum(prof[c][1] for c in gemm_cats)
    gemm_flops = sum(fl[c] for c in gemm_cats) * n_img
    gemm_tf = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
    traffic, traffic_detail = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and args.model == "vit_b" and B == 32:  # the capture is of this exact workload
        with open(tpath) as f:
            traffic_detail = json.load(f)
        traffic = traffic_detail.get("attn_global_bytes_per_launch")
    kernels = {}
    dec_ms = sum(prof[c][0] for c in prof if c.startswith("dec_"))
    for c, (t, n) in prof.items():
        if c.startswith("dec_") or n == 0:
            continue
        kernels[c] = {"ms_per_step": t / args.steps, "launches_per_step": n / args.steps,
                      "share_of_step": t / ms_prof_total,
                      "tflops": (fl[c] * n_img / (t / 1e3) / 1e12) if c in fl and t > 0 else None}
    kernels["decoder_parts"] = {c: {"ms_per_step": prof[c][0] / args.steps, "launches_per_step": prof[c][1] / args.steps}
                                for c in prof if c.startswith("dec_")}
    kernels["decoder"] = {"ms_per_step": dec_ms / args.steps, "launches_per_step": sum(prof[c][1] for c in prof if c.startswith("dec_")) / args.steps,
                          "share_of_step": dec_ms / ms_prof_total,
                          "tflops": fl["decoder"] * n_img / (dec_ms / 1e3) / 1e12 if dec_ms > 0 else None}
    step_tf = value / world * fl["total"] / 1e12
    roofline = {
        "bound": "tensor", "scope": "whole step: encoder + decoder, every kernel and the gaps between them",
        "achieved": step_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": step_tf / peaks["bf16_sustained"],
        "frac_of_burst": step_tf / peaks["bf16_burst"], "flops_per_image": fl["total"],
        "peak_source": peaks["source"] + ", sustained figure (kernels timed inside a long step)",
        "traffic": traffic,
        "traffic_note": "DRAM read+write bytes of ONE launch of the dominant kernel (global attention; algorithmic q,k,v,o = "
                        "0.81 GB) from ncu --set full; per-kernel figures in traffic_detail / profiles/traffic.json",
        "traffic_detail": traffic_detail,
        "gemm_family": {"kernel": "gemm_kernel<BN,EPI> (tcgen05): qkv + proj + mlp.lin1 + mlp.lin2 launches",
                        "achieved": gemm_tf, "frac": gemm_tf / peaks["bf16_sustained"] if gemm_tf else None,
                        "launches": gemm_launches, "ms_per_launch": gemm_ms / max(gemm_launches, 1),
                        "flops_per_launch": gemm_flops / max(gemm_launches, 1), "share_of_step": gemm_ms / ms_prof_total},
        "kernels": kernels,
    }
    ag = kernels.get("attn_global")
    if ag and ag.get("tflops"):
        roofline["dominant_kernel"] = {
            "kernel": "global attention (fused, decomposed rel-pos bias in-kernel): the single kernel symbol with the largest share",
            "share_of_step": ag["share_of_step"], "achieved": ag["tflops"], "unit": "TFLOP/s",
            "frac": ag["tflops"] / peaks["bf16_sustained"], "ms_per_launch": ag["ms_per_step"] / max(ag["launches_per_step"], 1e-9),
            "flops_per_launch": fl["attn_global"] * B / max(len(cfg.global_attn_indexes), 1),
            "note": "softmax exponentials on the MUFU pipe bound a d = 64 head near 50 % tensor utilisation (DESIGN.md section 4)"}

    # ------------------------------------------------------------------ rank 0, N == 1: B = 1 latency, same-GPU eager bf16, CPU
    lat, eager, cpu = None, None, None
    if rank == 0 and world == 1 and not args.no_extras:
        lat = latency_b1(eng, images, pe, sparse, dense_for_step(), False, dev)
        try:
            eager = gpu_eager_bf16(args.model, dev)
            eager["speedup_of_this_library"] = value / eager["value"]
        except Exception as exc:  # pragma: no cover - reported, never fatal
            eager = {"error": f"{type(exc).__name__}: {exc}"[:300]}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        cpu_step, kind = make_cpu_reference(args.model, args.workload)
        cpu_step()
        n_cpu = 2
        t0 = time.perf_counter()
        for _ in range(n_cpu):
            cpu_step()
        dt = time.perf_counter() - t0
        cpu = {"value": n_cpu / dt, "unit": "images/s", "cores": cores, "kind": kind,
               "sample": f"{n_cpu} x (1 image 1024x1024 + 1 prompt) after 1 warm-up, fp32, torch threads = {cores}; "
                         + ("the reference's own modules (oracle/_ref)" if kind == "reference" else "oracle port of the reference")}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": "images/s", "per_gpu": value / world, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": workload_config(args.model, B, world, args.workload),
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "comm_ms_per_step": comm.item() / args.steps if world > 1 else 0.0,
            "profile_pass": {"ms_per_step": ms_prof_total / args.steps,
                             "note": "the per-kernel figures in roofline.kernels come from a second pass of the same steps with a "
                                     "CUDA event after every launch (not part of the timed region)"},
            "latency_b1": lat, "gpu_eager_bf16": eager, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
