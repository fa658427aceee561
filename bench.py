#!/usr/bin/env python
"""bench.py -- expanded images/sec of the DistDiff guided-expansion hot path on B200 (+ guidance-kernel rooflines).

    python bench.py --gpus N --steps K --warmup W            # ours (torchrun launches it for N > 1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the host cores (oracle port)
    python bench.py --config 5                               # BASELINE configs[4]: C=1000, bf16, direct guidance every step
    python bench.py --ops eager                              # same host code, fused kernels swapped for the eager sequences

Workload (BASELINE.json configs[1]): Caltech-101-shaped synthetic data (100 classes),
SD-v1.4-shaped UNet/VAE + ResNet-50 guide with random-init weights (no network), 512 px, 50 DDIM steps
(--strength 1.0), CFG 7.5, transform_guidance at t = 381 over 2 sub-steps, K = 3 agglomerative group
prototypes, rho 10, L-inf radius 0.2 (scripts/exps/expand_diff.sh).  One "step" = one batch of B images
expanded once (one pass of generate_data.py:1145-1227).  Multi-GPU = the reference's image split: every rank
expands its own images, no data-path collective ("scaling": "weak").

What one JSON line carries besides the contract keys:
  * ``e2e``          the same step through the public API with HOST batches (a fresh one per step), H2D of the batch,
                     D2H of the image bytes, PNG encode + atomic publish (generate_data.py:1230-1234) inside the region;
  * ``proto_sweep``  BASELINE configs[3] at THIS world size: the 100k x 2048 per-class k-means sweep, K = 3/5/10, per
                     exchange variant (1 GPU: none; N > 1: fused NVLink peer kernel and NCCL all-reduce);
  * ``dist_parity``  (N > 1) sharded prototype stage == single-GPU result, peer exchange == NCCL bit for bit -- the run
                     exits non-zero if it fails;
  * ``ops_compare``  (N = 1) the same Expander with the fused kernels swapped for the literal eager op sequences
                     (distdiff_b200/eager_baseline.py), CUDA graphs on and off;
  * ``roofline`` / ``roofline_batched`` / ``cpu_baseline`` as the tier contract describes.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
import types

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "expanded images/sec"
UNIT = "images/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=3)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--config", type=int, default=2, choices=[2, 5],
                   help="BASELINE configs index (1-based): 2 = Caltech-101 5x expansion (default, the metric's config); "
                        "5 = C=1000 tables, bf16, direct_guidance on all 50 steps")
    p.add_argument("--ops", default="fused", choices=["fused", "eager"], help="eager: benchmark comparator, ATen launches instead of our kernels")
    p.add_argument("--batch", type=int, default=8, help="images per step and GPU (train_batch_size)")
    p.add_argument("--dtype", default=None, choices=["fp16", "bf16", "fp32"], help="UNet/VAE/latent storage type (reference: fp16)")
    p.add_argument("--no-cuda-graph", action="store_true")
    p.add_argument("--grad-ckpt", action="store_true", help="UNet gradient checkpointing in the guided step (generate_data.py:1049-1050): memory vs time")
    p.add_argument("--no-channels-last", action="store_true", help="keep UNet/VAE/guide in NCHW (PyTorch-side layout choice)")
    p.add_argument("--no-kernels", action="store_true", help="skip the batched kernel micro-benchmarks")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-proto-sweep", action="store_true")
    p.add_argument("--no-ops-compare", action="store_true")
    p.add_argument("--tiny", action="store_true", help="tiny networks (CI smoke of bench.py itself; not a valid number)")
    o = p.parse_args()
    if o.dtype is None:
        o.dtype = "bf16" if o.config == 5 else "fp16"
    return o


def canonical_args(batch, config=2):
    """scripts/exps/expand_diff.sh:3-16 with --strength 1.0 (50 DDIM steps, BASELINE configs[1]); config 5 =
    BASELINE configs[4]: ImageNet-scale tables, direct guidance on every step (--guidance_step 50 --guidance_period 50)."""
    a = types.SimpleNamespace(do_classifier_free_guidance=True, guidance_scale=7.5, gs=1.0, ls=1.0, rho=10.0,
                              guidance_type="transform_guidance", guidance_step=20, guidance_period=2, constraint_value=0.2,
                              K=3, strength=1.0, seed=42, train_batch_size=batch, cluster_method="agglomerative",
                              num_classes=100, arch="resnet50", dataset="caltech-101")
    if config == 5:
        a.guidance_type, a.guidance_step, a.guidance_period, a.num_classes, a.dataset = "direct_guidance", 50, 50, 1000, "imagenet-shaped"
    return a


def workload_config(a, batch, world, graph, dtype, cl=False, config=2, ops_mode="fused"):
    if config == 5:
        name = "ImageNet-shaped C=1000 prototype tables, ResNet-50 guide, SD v1.4 512px, 50 DDIM steps, direct_guidance on every step, bf16 (BASELINE configs[4])"
        guid = "direct_guidance on all 50 steps"
    else:
        name = "Caltech-101 5x expansion, ResNet-50 guide, SD v1.4 512px, 50 DDIM steps (BASELINE configs[1])"
        guid = "transform_guidance t=381 period 2"
    return {"workload": name, "images_per_step_per_gpu": batch, "ddim_steps": 50, "guidance": guid, "classes": a.num_classes,
            "K": a.K, "cluster_method": a.cluster_method, "cfg_scale": a.guidance_scale, "latent": "4x64x64",
            "storage_dtype": dtype, "weights": "random-init SD-v1.x UNet (859.5M) / VAE / ResNet-50", "ops": ops_mode,
            "parallelism": f"image-split x{world} (no collective)", "cuda_graph_unguided_step": bool(graph),
            "channels_last": bool(cl),
            "l2": "every step streams ~3.4 GB of UNet weights+activations (> 126 MB L2) between two launches of the same kernel; "
                  "micro-benchmarks evict L2 by reading 512 MB before each timed launch"}


# ----------------------------------------------------------------------------------------------- clocks / utilisation
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def busy(self, t0, t1):
        """mean utilization.gpu (%) of the samples taken inside the wall-clock window [t0, t1]"""
        v = [int(r[9]) for ts, r in self.rows if t0 <= ts <= t1 and len(r) > 9 and r[9].isdigit()]
        return round(sum(v) / len(v), 1) if v else None

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = [r for ts, r in self.rows if (t0 is None or ts >= t0) and (t1 is None or ts <= t1)] or [r for _, r in self.rows]
        sm = sorted(int(r[1]) for r in rows if len(r) > 8 and r[1].isdigit())
        mx = [int(r[2]) for r in rows if len(r) > 8 and r[2].isdigit()]
        reasons = set()
        for r in rows:
            if len(r) > 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "window": "the two timed regions (resident + e2e)"}


# ----------------------------------------------------------------------------------------------- models / data
def build_models(tiny, seed=0, num_classes=100):
    import torch
    from distdiff_b200 import nets
    torch.manual_seed(seed)
    if tiny:
        unet = nets.UNet2DConditionModel(block_out_channels=(32, 64, 64, 64), heads=2)
        vae = nets.AutoencoderKL(chs=(32, 32, 64, 64), with_encoder=False)
        guide = nets.create_model("resnet18", num_classes=num_classes)
    else:
        unet = nets.UNet2DConditionModel()
        vae = nets.AutoencoderKL(with_encoder=False)
        guide = nets.create_model("resnet50", num_classes=num_classes)
    for m in (unet, vae, guide):
        m.requires_grad_(False).eval()
    return unet, vae, guide


def synthetic_batch(batch, rank, step_seed=0, size=64, classes=100):
    """One batch of the Caltech-shaped workload as HOST tensors: VAE latents [B,4,64,64] (scaled like
    latent_dist.sample()*0.18215), class prompt / unconditional embeddings [B,77,768], labels, names."""
    import torch
    g = torch.Generator().manual_seed(1234 + 7919 * rank + 104729 * step_seed)
    targets = [(rank * batch + step_seed * 31 + i) % classes for i in range(batch)]
    return {"image_latents": (torch.randn(batch, 4, size, size, generator=g) * 0.18215 * 4.0),
            "input_ids": torch.randn(batch, 77, 768, generator=g), "uncond_inputs_ids": torch.randn(1, 77, 768, generator=g).expand(batch, -1, -1).contiguous(),
            "targets": targets, "class_names": [f"class {t:03d}" for t in targets],
            "image_paths": [f"synthetic/class_{t:03d}/image_{rank:02d}_{step_seed:03d}_{i:04d}.jpg" for i, t in enumerate(targets)]}


def synthetic_guide_features(n=3000, d=2048, c=100, seed=7):
    import torch
    g = torch.Generator().manual_seed(seed)
    centers = torch.randn(c, 3, d, generator=g)
    labels = torch.arange(n) % c
    which = torch.randint(0, 3, (n,), generator=g)
    return centers[labels, which] * 1.5 + torch.randn(n, d, generator=g), labels


# ----------------------------------------------------------------------------------------------- CPU oracle arm
def cpu_reference_sample(a, tiny, steps, warmup, log=lambda *_: None):
    """The reference's algorithm (oracle port: CPU fp32 restatement of generate_data.py:109-137,687-732) on the host
    cores, B = 1, full-size networks.  Bounded sample: `steps` unguided DDIM steps (after `warmup`), ONE
    transform_guidance call (2 sub-steps fwd+bwd through UNet, VAE decoder, ResNet-50) and ONE final VAE decode.
    One image = 50 unguided steps + 1 guided call + 1 decode, so a sampled DDIM step is the fraction
    t_u / (50 t_u + t_g + t_dec) of an image and images/sec = that fraction / t_u."""
    import torch
    from oracle import ddim as o_ddim, guidance as o_guid, prototypes as o_proto
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    unet, vae, guide = build_models(tiny)
    proc = __import__("distdiff_b200.nets", fromlist=["VaeImageProcessor"]).VaeImageProcessor()
    feats, labels = synthetic_guide_features(600 if tiny else 3000, 512 if tiny else 2048)
    t0 = time.time()
    fn = o_proto.l2_normalize_rows(feats.numpy())
    gl, lc, _ = o_proto.extract_prototype_from_features(fn, labels.tolist(), a.K)
    t_proto = time.time() - t0
    gp, lp = (torch.from_numpy(v) for v in o_proto.normalize_prototypes(gl, lc))
    sched = o_ddim.OracleDDIMScheduler(50)
    b = synthetic_batch(1, 0, size=8 if tiny else 64)
    prompt = torch.cat([b["uncond_inputs_ids"], b["input_ids"]])
    lat = o_ddim.add_noise(b["image_latents"], torch.randn(b["image_latents"].shape, generator=torch.Generator().manual_seed(0)),
                           o_ddim.alphas_cumprod()[981])
    ts = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.time()
            lat2, _ = o_guid.denoise_one_step(a, lat, sched, 981 - 20 * (i % 25), unet, prompt, None)
            dt = time.time() - t0
            if i >= warmup:
                ts.append(dt)
            log(f"cpu unguided step {i}: {dt:.2f}s")
    t_u = sum(ts) / len(ts)
    t0 = time.time()
    torch.manual_seed(0)
    lat3, score = o_guid.transform_guidance(a, lat, {"targets": b["targets"]}, [381, 361], sched, unet, prompt, None, vae, guide,
                                            proc, torch.float32, None, gp, lp)
    t_g = time.time() - t0
    log(f"cpu transform_guidance: {t_g:.2f}s score {float(score):.4f}")
    t0 = time.time()
    with torch.no_grad():
        vae.decode(lat3 / vae.config.scaling_factor)[0]
    t_d = time.time() - t0
    t_img = 50 * t_u + t_g + t_d
    return {"value": 1.0 / t_img, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"B=1, {steps} unguided DDIM steps (mean {t_u:.2f} s) + 1 transform_guidance call ({t_g:.2f} s) + 1 VAE decode "
                      f"({t_d:.2f} s) at full SD-v1.x/ResNet-50 size, fp32, torch CPU {cores} threads; one image = 50*t_u + t_g + t_dec "
                      f"= {t_img:.1f} s; prototype construction (sklearn agglomerative, 3000x2048) {t_proto:.2f} s not included",
            "t_unguided_s": t_u, "t_guided_s": t_g, "t_decode_s": t_d, "t_prototypes_s": t_proto,
            "image_fraction_per_step": t_u / t_img}


def run_reference(opt):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return 0
    a = canonical_args(1)
    t_wall = time.time()
    res = cpu_reference_sample(a, opt.tiny, max(1, opt.steps), max(0, opt.warmup), log=lambda m: print(m, file=sys.stderr, flush=True))
    # a reference-arm "step" is ONE sampled DDIM step (ms_per_step is its measured time, so steps x ms_per_step is what
    # actually ran); value = the fraction of an image that step is / its time = 1 / (50 t_u + t_g + t_dec)
    line = {"impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": opt.gpus, "steps": opt.steps,
            "warmup": opt.warmup, "ms_per_step": 1e3 * res["t_unguided_s"], "image_fraction_per_step": res["image_fraction_per_step"],
            "step_is": "one sampled unguided DDIM step at B=1 (a bounded sample of the workload); one guided call and one decode are "
                       "measured once after the steps and enter the per-image time",
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(a, 1, 1, False, "fp32"),
            "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": round(time.time() - t_wall, 1),
            "note": "reference has no CPU path and cannot be imported here (diffusers/timm absent): this is the oracle port of its algorithm"}
    print(json.dumps(line), flush=True)
    return 0


# ----------------------------------------------------------------------------------------------- ours
def run_ours(opt):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device; there is no CPU fallback for the hot path")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from distdiff_b200 import eager_baseline, expand, microbench, nets, ops, protobench, prototypes
    from distdiff_b200.scheduler import DDIMScheduler

    def barrier():
        if world > 1:
            dist.barrier()

    a = canonical_args(opt.batch, opt.config)
    C = a.num_classes
    wd = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[opt.dtype]
    unet, vae, guide = build_models(opt.tiny, num_classes=C)
    if opt.grad_ckpt:
        unet.enable_gradient_checkpointing()
    unet.to(dev, wd); vae.to(dev, wd); guide.to(dev)
    if not opt.no_channels_last:   # PyTorch-side: NHWC weights let cuDNN skip its per-conv nchw<->nhwc transposes (14 % of a step)
        for m in (unet, vae, guide):
            m.to(memory_format=torch.channels_last)

    # ---- multi-GPU parity of the sharded prototype stage, BEFORE anything is timed (exit non-zero on failure) ----
    dist_parity = None
    peer = nccl = None
    if world > 1:
        nccl = prototypes.NcclCollective()
        peer = prototypes.make_collective()
        if not isinstance(peer, prototypes.PeerCollective):
            peer = nccl
        ok, rep = protobench.dist_parity(peer, nccl, n_agglo=600 if opt.tiny else 3000, n_kmeans=5000 if opt.tiny else 100_000,
                                         d=512 if opt.tiny else 2048)
        dist_parity = {"status": "ok" if ok else "FAILED", "world": world, "detail": rep,
                       "checks": "sharded agglomerative (3000x2048) and k-means (100k x 2048, 10 Lloyd iterations) prototypes vs the "
                                 "unsharded single-GPU result (1e-6 global / 1e-5 group); fused peer exchange vs NCCL all-reduce + update bit-identical"}
        if not ok:
            if rank == 0:
                print(json.dumps({"dist_parity": dist_parity}), flush=True)
            raise SystemExit("bench.py: multi-GPU prototype parity FAILED")

    # ---- prototypes through K1/K2/K3' (sharded over ranks + NCCL all-reduce when world > 1); outside the timed region ----
    n_feat = (600 if opt.tiny else 3000) * (C // 100)
    feats, labels = synthetic_guide_features(n_feat, 512 if opt.tiny else 2048, c=C)
    per = -(-feats.shape[0] // world)
    sl = slice(per * rank, min(per * (rank + 1), feats.shape[0]))
    torch.cuda.synchronize()
    t0 = time.time()
    gmean, lmean = prototypes.build_prototypes(feats[sl].to(dev), labels[sl].to(dev), C, a.K, a.cluster_method, coll=nccl)
    torch.cuda.synchronize()
    t_proto = time.time() - t0
    gproto, lproto = ops.normalize_rows(gmean), ops.normalize_rows(lmean)
    guide.to(wd)

    # ---- BASELINE configs[3]: the k-means prototype sweep at this world size ----
    proto_sweep = None
    if not opt.no_proto_sweep:
        colls = {"none": None} if world == 1 else ({"peer": peer, "nccl": nccl} if peer is not nccl else {"nccl": nccl})
        proto_sweep = protobench.sweep(colls, n=5000 if opt.tiny else 100_000, d=512 if opt.tiny else 2048, ks=(3, 5, 10), iters=20, reps=3)
    if world > 1:
        if peer is not nccl:
            peer.close()
        nccl.close()

    def make_expander(graph, graph_guidance=True):
        return expand.Expander(a, unet, vae, guide, nets.VaeImageProcessor(), DDIMScheduler(), gproto, lproto, weight_dtype=wd, device=dev,
                               use_cuda_graph=graph, graph_guidance=graph_guidance)

    size = 8 if opt.tiny else 64
    n_host = max(opt.steps, 1)
    hosts = [synthetic_batch(opt.batch, rank, step_seed=i, size=size, classes=C) for i in range(n_host)]
    for h in hosts:
        for k in ("image_latents", "input_ids", "uncond_inputs_ids"):
            h[k] = h[k].pin_memory()
    resident = dict(hosts[0])
    for k in ("image_latents", "input_ids", "uncond_inputs_ids"):
        resident[k] = hosts[0][k].to(dev, wd)

    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")
    smi_index = int(vis[local]) if len(vis) > local and vis[local].strip().isdigit() else local
    sampler = ClockSampler(smi_index)

    def timed(fn, tag, steps, after=None):
        barrier(); torch.cuda.synchronize()
        torch.cuda.nvtx.range_push(tag)      # ncu --nvtx --nvtx-include "<tag>/" profiles exactly the timed region
        l0 = ops.launch_count
        w0 = time.time()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        if after is not None:
            after()
        e1.record()
        torch.cuda.synchronize()
        w1 = time.time()
        torch.cuda.nvtx.range_pop()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) * 1e-3, ops.launch_count - l0, (w0, w1)

    swap = eager_baseline.swapped() if opt.ops == "eager" else None
    if swap is not None:
        swap.__enter__()
    # (the eager sequences cannot be captured around the guided step: tensor_clamp's masked scatters synchronise)
    ex = make_expander(not opt.no_cuda_graph, graph_guidance=opt.ops == "fused")
    expand.set_seed(a.seed)
    for i in range(max(opt.warmup, 1)):
        img, lat, info = ex.expand_batch(resident)
    torch.cuda.synchronize()
    finite = bool(torch.isfinite(img.float()).all()) and bool(torch.isfinite(torch.stack(info["scores"])).all())
    if not finite:
        raise RuntimeError(f"non-finite output in {opt.dtype}: images finite={bool(torch.isfinite(img.float()).all())}, scores={info['scores']}")
    if rank == 0:
        sampler.start()

    # (1) inputs resident in HBM
    t_res, launches, win_res = timed(lambda i: ex.expand_batch(resident), "dd_timed_resident", opt.steps)
    # (2) end to end through the public API, like generate_data.py:1130-1236: a FRESH host batch per step (pinned), H2D of
    #     the batch, D2H of the [B,H,W,3] bytes (K9 output), PNG encode + atomic publish by the writer threads, and the
    #     scores the reference logs; the writer is drained inside the timed region
    out_dir = tempfile.mkdtemp(prefix=f"dd_bench_png_r{rank}_")
    writer = expand.AsyncPngWriter()
    a_out = types.SimpleNamespace(output_dir=out_dir)

    def e2e_step(i):
        h = hosts[i % n_host]
        img, _lat, info = ex.expand_batch(h, as_uint8=True)
        writer.submit(img, [expand.output_path(a_out, h, j, 0) for j in range(opt.batch)])
        float(torch.stack(info["scores"]).sum())          # the score the reference logs (D2H read, syncs)
    t_e2e, _, win_e2e = timed(e2e_step, "dd_timed_e2e", opt.steps, after=writer.close)
    n_png = sum(len(fs) for _, _, fs in os.walk(out_dir))
    shutil.rmtree(out_dir, ignore_errors=True)
    clocks = sampler.stop(win_res[0], win_e2e[1]) if rank == 0 else None
    gpu_busy = {"resident_pct": sampler.busy(*win_res), "e2e_pct": sampler.busy(*win_e2e),
                "how": "mean nvidia-smi utilization.gpu over the samples (200 ms) inside each timed region only"} if rank == 0 else None
    h2d = sum(hosts[0][k].numel() * hosts[0][k].element_size() for k in ("image_latents", "input_ids", "uncond_inputs_ids"))
    d2h = opt.batch * (size * 8) * (size * 8) * 3 + 4 * len(info["scores"])
    if swap is not None:
        swap.__exit__(None, None, None)

    # (3) per-kernel durations of OUR kernels: one instrumented eager step (inside CUDA-graph replays the K5 launches
    #     cannot be bracketed by events), same workload, CUDA events on the launching stream
    ops.profiler = []
    ex_nog = make_expander(False)
    ex_nog.expand_batch(resident)
    torch.cuda.synchronize()
    prof, ops.profiler = ops.profiler, None
    per_kernel = {}
    for name, nbytes, s0, s1 in prof:
        d = per_kernel.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0})
        d["launches"] += 1; d["ms"] += s0.elapsed_time(s1); d["bytes"] += nbytes
    peak, peak_src = microbench.hbm_peak_gbs()
    dom = max(per_kernel, key=lambda k: per_kernel[k]["ms"])
    dk = per_kernel[dom]
    achieved = dk["bytes"] / (dk["ms"] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s", "frac": round(achieved / peak, 5),
                "traffic": None, "peak_source": peak_src, "launches_per_step": dk["launches"],
                "avg_launch_us": round(1e3 * dk["ms"] / dk["launches"], 2), "algorithmic_bytes_per_launch": dk["bytes"] // dk["launches"],
                "regime": f"launch-latency-bound: at B={opt.batch} the kernel moves {dk['bytes'] // dk['launches'] // 1024} KB per launch "
                          "(HBM-bound sizes are in roofline_batched)",
                "measured": "CUDA events around each launch on the launching stream, one instrumented eager step after the timed region"}
    # dram bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same launch
    # shape (tools/instep_k5.py); null when this run's batch / dtype has no capture
    try:
        with open(os.path.join(ROOT, "profiles", "instep_traffic.json")) as fh:
            cap = json.load(fh).get(f"{dom}:B{opt.batch}:{opt.dtype}")
        if cap:
            roofline["traffic"] = cap["dram_bytes_per_launch"]
            roofline["traffic_source"] = cap["source"]
    except (OSError, ValueError):
        pass
    kernels_in_step = {k: {"launches": v["launches"], "avg_us": round(1e3 * v["ms"] / v["launches"], 2),
                           "GBps": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1)} for k, v in per_kernel.items()}
    glue_us = round(1e3 * sum(v["ms"] for v in per_kernel.values()), 1)

    # (4) fused vs eager: the SAME Expander with ops swapped for the literal reference sequences, graphs on and off
    ops_compare = None
    if world == 1 and not opt.no_ops_compare and opt.ops == "fused":
        n_cmp = min(opt.steps, 3)
        rows = {}

        def measure(tag, graph, eager):
            ctx = eager_baseline.swapped() if eager else None
            if ctx is not None:
                ctx.__enter__()
            try:
                e = ex if (graph and not eager and not opt.no_cuda_graph) else make_expander(graph, graph_guidance=not eager)
                expand.set_seed(a.seed)
                e.expand_batch(resident); e.expand_batch(resident)
                t, _, _ = timed(lambda i: e.expand_batch(resident), "dd_cmp_" + tag, n_cmp)
                rows[tag] = {"images_per_s": round(opt.batch * n_cmp / t, 4), "ms_per_step": round(1e3 * t / n_cmp, 2)}
                if e is not ex:
                    del e
                    torch.cuda.empty_cache()
            finally:
                if ctx is not None:
                    ctx.__exit__(None, None, None)
        measure("fused_graph", True, False)
        measure("fused_nograph", False, False)
        measure("eager_graph", True, True)
        measure("eager_nograph", False, True)
        # kernel launches of the glue (everything that is not UNet / VAE / guide) in ONE step, counted with the profiler
        glue = {}
        try:
            from torch.profiler import ProfilerActivity, profile
            for tag, eager in (("fused", False), ("eager", True)):
                ctx = eager_baseline.swapped() if eager else None
                if ctx is not None:
                    ctx.__enter__()
                try:
                    mod = eager_baseline if eager else ops
                    x = resident["image_latents"]; npred = torch.randn(2 * opt.batch, *x.shape[1:], device=dev, dtype=wd)
                    f = torch.randn(opt.batch, gproto.shape[1], device=dev, requires_grad=True)
                    with profile(activities=[ProfilerActivity.CUDA]) as pr:
                        with torch.no_grad():
                            mod.cfg_ddim_step(npred, x, 7.5, 0.3, 0.35)
                            mod.add_noise(x, x, 0.3)
                            mod.affine_project(x, torch.rand(opt.batch, 4, 1, 1, device=dev), torch.randn(opt.batch, 4, 1, 1, device=dev), 0.2)
                        s = mod.PrototypeEnergy.apply(f, hosts[0]["targets"], gproto, lproto, 1.0, 1.0, False)
                        torch.autograd.grad(s, f)
                        torch.cuda.synchronize()
                    ev = [e for e in pr.events() if e.device_type is not None and "cuda" in str(e.device_type).lower()]
                    glue[tag] = {"kernel_launches": len(ev), "what": "one K5 step + add_noise + affine+projection + energy fwd+bwd"}
                finally:
                    if ctx is not None:
                        ctx.__exit__(None, None, None)
        except Exception as exc:   # profiler unavailable: keep the timing rows
            glue = {"error": repr(exc)[:200]}
        fg, eg = rows["fused_graph"]["images_per_s"], rows["eager_graph"]["images_per_s"]
        ops_compare = {"rows": rows, "glue_launch_count": glue, "fused_glue_us_per_step": glue_us,
                       "e2e_delta_pct_fused_vs_eager_graph": round(100.0 * (fg - eg) / eg, 2),
                       "e2e_delta_pct_fused_vs_eager_nograph": round(100.0 * (rows["fused_nograph"]["images_per_s"] - rows["eager_nograph"]["images_per_s"])
                                                                     / rows["eager_nograph"]["images_per_s"], 2),
                       "note": f"B={opt.batch}, {n_cmp} steps each, inputs resident; eager = distdiff_b200/eager_baseline.py (the literal "
                               "generate_data.py op sequences as ATen launches) behind the same host code; UNet/VAE/guide identical in all rows"}

    line = None
    if rank == 0:
        value = world * opt.batch * opt.steps / t_res
        e2e = world * opt.batch * opt.steps / t_e2e
        line = {"metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": opt.steps, "warmup": opt.warmup,
                "ms_per_step": round(1e3 * t_res / opt.steps, 2), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                # what runs: fp32 arithmetic inside the guidance kernels; the latents / UNet / VAE / guide are stored and run by
                # PyTorch in config.storage_dtype (fp16 is the reference's own, generate_data.py:1039)
                "dtype": f"f32 arithmetic in the guidance kernels, {opt.dtype} storage and {opt.dtype} UNet/VAE/guide (PyTorch)",
                "data": "synthetic", "config": workload_config(a, opt.batch, world, not opt.no_cuda_graph, opt.dtype, not opt.no_channels_last,
                                                               opt.config, opt.ops),
                "e2e": {"value": round(e2e, 4), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": round(1e3 * t_e2e / opt.steps, 2), "png_files_written": n_png,
                        "includes": "fresh pinned host batch per step, H2D, K9 bytes D2H, PNG encode + atomic publish (drained inside the region), score read"},
                "gpu_launches": launches, "clocks": clocks, "gpu_busy_timed_regions": gpu_busy, "roofline": roofline,
                "kernels_in_step": kernels_in_step, "prototype_construction_s": round(t_proto, 4),
                "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 1), "gradient_checkpointing": bool(opt.grad_ckpt),
                "tiny": bool(opt.tiny)}
        if dist_parity is not None:
            line["dist_parity"] = dist_parity
        if proto_sweep is not None:
            line["proto_sweep"] = proto_sweep
        if ops_compare is not None:
            line["ops_compare"] = ops_compare
        if world == 1 and not opt.no_kernels:
            recs = microbench.run(iters=5, ks=(3, 5, 10), latent_dtypes=(torch.float32, torch.float16),
                                  want=lambda n: not n.startswith("agglo"))
            keep = [r for r in recs if any(s in r["kernel"] for s in ("B4096", "B65536", "K1_", "K3_", "_B128", "_B256", "eager_", "_B1_", "_B8_"))
                    or r["kernel"].endswith(("_B1", "_B8"))]
            line["roofline_batched"] = [{"kernel": r["kernel"], "achieved": r["GBps"], "peak": peak, "unit": "GB/s", "frac": r["frac"],
                                         "ms": r["ms"]} for r in keep]
    barrier()
    if rank == 0:
        if world == 1 and not opt.no_cpu_baseline:
            del ex, ex_nog
            torch.cuda.empty_cache()
            line["cpu_baseline"] = {k: v for k, v in cpu_reference_sample(canonical_args(1), opt.tiny, 2, 1).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    o = parse()
    sys.exit(run_reference(o) if o.impl == "reference" else run_ours(o))
