"""The expansion engine: the body of generate_data.py:main (reference lines 1100-1236) around the kernels.

``Expander.expand_batch(batch, image_i)`` is one pass of the reference's inner loop (:1145-1227): noise at
t_enc, the guided / unguided denoise steps, the final VAE decode and de-normalisation.  The per-step control
flow (which step is guided, literal timestep-index arithmetic) is the reference's; the tensor math between the
network calls is the fused kernels (guidance.py).  Optionally the unguided step (UNet forward + K5) is
captured once in a CUDA graph and replayed -- at the reference's batch sizes that step is launch-bound.
"""
from __future__ import annotations

import logging
import math
import os
import random
from typing import Optional

import numpy as np
import torch

from . import guidance, ops, prototypes
from .scheduler import DDIMScheduler, retrieve_timesteps

logger = logging.getLogger("distdiff_b200")

NUM_INFERENCE_STEPS = 50  # hard-coded in the reference (generate_data.py:1043); --steps is unused there too


def set_seed(seed: int) -> None:
    """accelerate.utils.set_seed (generate_data.py:861): python, numpy, torch CPU + CUDA."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)


class _GraphedStep:
    """UNet forward + K5 for one (timestep, batch shape) captured in a CUDA graph."""

    def __init__(self, expander, latents, prompt_embeds, t):
        self.static_lat = latents.clone()
        self.static_prompt = prompt_embeds.clone()
        t = (int(t), torch.full((1,), int(t), dtype=torch.int64, device=latents.device))  # no H2D copy inside the capture
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):  # warm-up outside capture (cuDNN autotune, workspace allocation)
                expander._step_eager(self.static_lat, self.static_prompt, t)
        torch.cuda.current_stream().wait_stream(side)
        # one memory pool for all 50 per-timestep graphs: they replay strictly one after the other, so the UNet
        # activations of every capture reuse the same blocks (private pools cost ~0.7 GB per graph and image)
        with torch.cuda.graph(self.graph, pool=expander._graph_pool), torch.no_grad():
            self.out_prev, self.out_x0 = expander._step_eager(self.static_lat, self.static_prompt, t)

    def __call__(self, latents, prompt_embeds):
        self.static_lat.copy_(latents)
        self.static_prompt.copy_(prompt_embeds)
        self.graph.replay()
        ops.launch_count += 1  # the K5 launch inside the replayed graph
        return self.out_prev.clone(), self.out_x0


class _GraphedGuidance:
    """transform_guidance's device work -- ChannelAffine, `period` x (UNet fwd, K5, VAE decode, K8, guide, K4), the autograd
    backward through all of it, the parameter update and the projected transform -- captured once per (timesteps, batch
    shape) in a CUDA graph.  Only the two CPU RNG draws of the channel parameters (generate_data.py:692-695, reference
    order) and three small copies into static buffers happen per call; the guided step is otherwise launch-by-launch
    (~2.5k launches at B = 8)."""

    def __init__(self, expander, latents, prompt_embeds, targets, guide_ts):
        dev = latents.device
        bs, ch = latents.shape[0], latents.shape[1]
        self.static_lat = latents.clone()
        self.static_prompt = prompt_embeds.clone()
        self.static_targets = targets.clone()
        self.static_cn = torch.zeros(bs, ch, 1, 1, device=dev)
        self.static_cb = torch.zeros(bs, ch, 1, 1, device=dev)
        ts = [int(t) for t in guide_ts]
        e = expander

        def run():
            # the timestep is a python int here: the UNet builds its embedding on the device (no H2D inside the capture
            # because nets.UNet2DConditionModel takes the device tensor below) and the scheduler tables are host floats
            return guidance.transform_guidance_core(self.static_lat, self.static_targets, self.static_cn, self.static_cb,
                                                    [self._t(t, dev) for t in ts], e.sched, e.unet, self.static_prompt, None, e.vae,
                                                    e.image_encoder, e.image_processor, None, e.gproto, e.lproto)
        self._tcache = {}
        for t in ts:
            self._t(t, dev)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):   # warm-up outside capture (cuDNN autotune for the backward convs, workspaces)
                run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()   # the warm-up's ~9 GB per image sit cached in the default pool; the capture allocates from its own
        with torch.cuda.graph(self.graph, pool=expander._graph_pool):
            self.out_lat, self.out_score = run()

    def _t(self, t, dev):
        v = self._tcache.get(t)
        if v is None:
            v = self._tcache[t] = _DevTimestep(t, dev)
        return v

    def __call__(self, latents, prompt_embeds, targets, channel_noise, channel_noise_bias):
        self.static_lat.copy_(latents)
        self.static_prompt.copy_(prompt_embeds)
        self.static_targets.copy_(targets)
        self.static_cn.copy_(channel_noise, non_blocking=True)
        self.static_cb.copy_(channel_noise_bias, non_blocking=True)
        self.graph.replay()
        return self.out_lat.clone(), self.out_score.clone()


class _DevTimestep(int):
    """A timestep that is an int for the scheduler tables (``int(t)``) and carries a device tensor for the UNet, so no
    host-to-device copy is issued while a graph is being captured."""

    def __new__(cls, t, dev):
        obj = super().__new__(cls, int(t))
        obj.dev = torch.full((1,), int(t), dtype=torch.int64, device=dev)
        return obj


class Expander:
    def __init__(self, args, unet, vae, image_encoder, image_processor, noise_scheduler: DDIMScheduler,
                 total_global_proto, total_local_proto, weight_dtype=torch.float16, device="cuda", use_cuda_graph=False,
                 graph_guidance=True):
        self.args = args
        guidance.set_args(args)
        self.unet, self.vae, self.image_encoder, self.image_processor = unet, vae, image_encoder, image_processor
        self.sched = noise_scheduler
        self.gproto, self.lproto = total_global_proto, total_local_proto
        self.weight_dtype = weight_dtype
        self.device = torch.device(device)
        self.timesteps, _ = retrieve_timesteps(noise_scheduler, NUM_INFERENCE_STEPS, "cpu")   # :1043-1044
        self.use_cuda_graph = use_cuda_graph
        self.graph_guidance = bool(graph_guidance)
        self._graphs = {}
        self._graph_pool = torch.cuda.graph_pool_handle() if use_cuda_graph else None

    # ---- one unguided step ----
    def _step_eager(self, latents, prompt_embeds, t):
        if not isinstance(t, tuple):
            return guidance.denoise_one_step(latents, self.sched, t, self.unet, prompt_embeds, None)
        # graph capture: same four lines as denoise_one_step (generate_data.py:109-121) with the timestep already on
        # the device for the UNet and as a python int for the scheduler tables
        t_host, t_dev = t
        cfg = bool(self.args.do_classifier_free_guidance)
        x = torch.cat([latents] * 2) if cfg else latents
        noise_pred = self.unet(x, t_dev, prompt_embeds, class_labels=None, return_dict=False)[0]
        a_t, a_prev = self.sched.alpha_pair(t_host)
        return ops.cfg_ddim_step(noise_pred, latents, float(self.args.guidance_scale), a_t, a_prev, cfg=cfg)

    def _step(self, latents, prompt_embeds, t):
        if not self.use_cuda_graph:
            with torch.no_grad():
                return self._step_eager(latents, prompt_embeds, t)
        key = (int(t), tuple(latents.shape), latents.dtype)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = _GraphedStep(self, latents, prompt_embeds, t)
        return g(latents, prompt_embeds)

    def _guided(self, latents, prompt_embeds, batch, guide_ts):
        """transform_guidance through a captured graph: the CPU draws stay in the reference's order, everything else replays."""
        cn, cb = guidance.draw_channel_noise(latents.shape[0], latents.shape[1])
        C_ = (self.gproto if self.gproto is not None else self.lproto).shape[0]
        targets = ops.targets_tensor(batch["targets"], C_, latents.device)
        key = ("guided", tuple(int(t) for t in guide_ts), tuple(latents.shape), latents.dtype)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = _GraphedGuidance(self, latents, prompt_embeds, targets, guide_ts)
        ops.launch_count += 2 * len(guide_ts) + 4   # K6 fwd/bwd + per sub-step (K5 fwd/bwd, K8 fwd/bwd, K4) + K6 project, replayed
        return g(latents, prompt_embeds, targets, cn.pin_memory(), cb.pin_memory())

    def expand_batch(self, batch, image_i: int = 0, decode: bool = True, as_uint8: bool = False):
        """generate_data.py:1145-1227 for one batch -> (images [B,3,H,W] in [0,1] or None, latents, info).
        as_uint8: return the [B,H,W,3] bytes save_image would write (K9) instead of the float image."""
        a = self.args
        dev, wd = self.device, self.weight_dtype
        prompt_embeds = batch["input_ids"].to(dev, dtype=wd, non_blocking=True)
        negative_prompt_embeds = batch["uncond_inputs_ids"].to(dev, dtype=wd, non_blocking=True)
        timesteps = self.timesteps
        model_input = batch["image_latents"].to(dev, dtype=wd, non_blocking=True)
        if getattr(a, "offset_noise", False):                                              # :1164-1168
            noise = torch.randn_like(model_input) + 0.1 * torch.randn(model_input.shape[0], model_input.shape[1], 1, 1,
                                                                       device=model_input.device, dtype=wd)
        else:
            noise = torch.randn_like(model_input)                                          # :1170 (CUDA default gen)
        start_index = guidance.start_index(a.strength, len(timesteps))                     # :1174
        t_enc = timesteps[start_index]
        noisy_model_input = self.sched.add_noise(model_input, noise, t_enc)                # :1176 (K7)
        guide_ts = guidance.guide_timesteps(timesteps, a.guidance_step, a.guidance_period)  # :1178-1180
        if a.do_classifier_free_guidance:
            prompt_embeds = torch.cat([negative_prompt_embeds, prompt_embeds])             # :1184
        generator = None if a.seed is None else torch.Generator(device=dev).manual_seed(a.seed)
        latents = noisy_model_input
        info = {"guide_timesteps": guide_ts, "scores": []}
        logger.info("Guidance timesteps: %s", ", ".join(str(x) for x in guide_ts))
        for t in timesteps[start_index:]:                                                  # :1199
            if t == guide_ts[0] and a.guidance_type == "transform_guidance":
                if self.use_cuda_graph and self.graph_guidance:
                    latents, score = self._guided(latents, prompt_embeds, batch, guide_ts)
                else:
                    latents, score = guidance.transform_guidance(latents, batch, guide_ts, self.sched, self.unet, prompt_embeds,
                                                                 None, self.vae, self.image_encoder, self.image_processor, wd,
                                                                 generator, self.gproto, self.lproto)
                latents, x_0 = self._step(latents, prompt_embeds, t)
                info["scores"].append(score)       # no .item(): the reference syncs here every guided step (:1208)
            elif int(t) in guide_ts and a.guidance_type == "direct_guidance":
                latents, x_0, score = guidance.direct_guidance(latents, batch, t, self.sched, self.unet, prompt_embeds, None,
                                                               self.vae, self.image_encoder, self.image_processor, wd,
                                                               generator, self.gproto, self.lproto)
                info["scores"].append(score)
            else:
                latents, x_0 = self._step(latents, prompt_embeds, t)
        image = None
        if decode:
            with torch.no_grad():                                                          # :1221-1227
                image = self.vae.decode(latents / self.vae.config.scaling_factor, return_dict=False, generator=generator)[0]
                if as_uint8:   # K9: denormalise + save_image's quantisation in one launch -> [B,H,W,C] bytes (:1227 + :1234)
                    image = ops.image_to_uint8(image, denormalize=True)
                else:
                    image = self.image_processor.postprocess(image, output_type="pt", do_denormalize=[True] * image.shape[0])
        return image, latents, info


def output_path(args, batch, i, image_i) -> str:
    """generate_data.py:1134-1135 / 1231-1232."""
    image_file_path = os.path.basename(batch["image_paths"][i]).split(".")[0]
    return f'{args.output_dir}/{batch["class_names"][i]}/{image_file_path}_expand_{image_i}.png'


class AsyncPngWriter:
    """Device->host copy and PNG encoding off the sampling thread.

    generate_data.py:1230-1234 converts, copies and PNG-encodes every image synchronously before the next batch
    starts.  Here the [B,H,W,C] bytes from K9 are copied into pinned memory asynchronously; worker threads wait on
    the copy's event, encode with PIL's default PNG settings (what torchvision.utils.save_image uses, so the files
    are byte-identical) and publish each file with an atomic rename -- an interrupted run never leaves a truncated
    PNG that the skip-if-exists resume (:1132-1143) would keep.  At most ``depth`` batches are in flight."""

    def __init__(self, workers: int = 4, depth: int = 4):
        import collections
        import concurrent.futures as cf
        self._pool = cf.ThreadPoolExecutor(max_workers=workers, thread_name_prefix="dd-png")
        self._pending = collections.deque()
        self._depth = depth

    @staticmethod
    def _encode(arr, path):
        from PIL import Image
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        tmp = f"{path}.tmp.{os.getpid()}"
        Image.fromarray(arr[:, :, 0] if arr.shape[2] == 1 else arr).save(tmp, format="PNG")
        os.replace(tmp, path)

    def _write(self, host, event, keep, paths):
        if event is not None:
            event.synchronize()
        del keep
        arr = host.numpy()
        for i, p in enumerate(paths):
            self._encode(arr[i], p)
        return len(paths)

    def submit(self, images_u8: torch.Tensor, paths) -> None:
        """images_u8: [B,H,W,C] uint8 (CUDA or CPU); paths: B output files."""
        if images_u8.dtype != torch.uint8 or images_u8.dim() != 4 or images_u8.shape[0] != len(paths):
            raise ValueError("AsyncPngWriter.submit expects uint8 [B,H,W,C] and one path per image")
        if images_u8.is_cuda:
            host = torch.empty(images_u8.shape, dtype=torch.uint8, pin_memory=True)
            host.copy_(images_u8, non_blocking=True)
            event = torch.cuda.Event()
            event.record()
        else:
            host, event = images_u8.contiguous(), None
        self._pending.append(self._pool.submit(self._write, host, event, images_u8, list(paths)))
        while len(self._pending) > self._depth:
            self._pending.popleft().result()

    def close(self) -> None:
        while self._pending:
            self._pending.popleft().result()
        self._pool.shutdown(wait=True)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def run_expansion(args, expander: Expander, train_dataloader, save: bool = True):
    """generate_data.py:1130-1236: batches x image_i with skip-if-exists resume."""
    n_done = 0
    with AsyncPngWriter() as writer:
        for _step, batch in enumerate(train_dataloader):
            for image_i in range(args.first_image_index, args.num_images_per_prompt):
                paths = [output_path(args, batch, i, image_i) for i in range(len(batch["image_paths"]))]
                if save and all(os.path.exists(p) for p in paths):                             # :1132-1143
                    for p in paths:
                        print(f"File {p} exists, so skipped.")
                    continue
                image, _lat, info = expander.expand_batch(batch, image_i, as_uint8=save)
                for s, t in zip(info["scores"], info["guide_timesteps"]):
                    logger.info("%s in %s step for %s steps period, score: %.4f", args.guidance_type, t, args.guidance_period, float(s))
                if save:
                    writer.submit(image, paths)                                                # :1230-1234
                n_done += len(paths)
    return n_done
