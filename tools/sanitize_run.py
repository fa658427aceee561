"""Reduced-size run of EVERY dd_* compute entry point, for compute-sanitizer (tools/gpu_sanitize.sh):
memcheck / racecheck / synccheck over the hand-rolled mbarrier rings (K1, K3, K4 tile), named barriers and last-CTA
tickets (K3 pair kernel, K4, partial reduce), and the flag protocol of the peer exchange (a 1-rank process group on one
GPU; `--world 2` under torchrun covers the cross-GPU path when two GPUs are available).
Sizes are small on purpose: racecheck serialises shared-memory accesses and runs ~100x slower."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from distdiff_b200 import ops, prototypes  # noqa: E402


def main():
    world = int(os.environ.get("WORLD_SIZE", 1))
    rank = int(os.environ.get("RANK", 0))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    coll = None
    if world > 1 or "--peer" in sys.argv:
        import torch.distributed as dist
        if world == 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29577")
            dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        else:
            dist.init_process_group("nccl", device_id=dev)
        coll = prototypes.PeerCollective()
    g = torch.Generator(device=dev).manual_seed(3)
    only = [a for a in sys.argv[1:] if not a.startswith("--")]
    want = lambda n: not only or n in only

    if want("K5"):   # K5 / K6 / K7 (fwd + bwd), fp32 and fp16
        for dt in (torch.float32, torch.float16):
            x = torch.randn(3, 4, 64, 64, device=dev, generator=g).to(dt)
            npred = torch.randn(6, 4, 64, 64, device=dev, generator=g).to(dt)
            xr = x.clone().requires_grad_(True); nr = npred.clone().requires_grad_(True)
            p, x0 = ops.CfgDdimStep.apply(nr, xr, 7.5, 0.3, 0.35, True)
            (p.float().sum() + x0.float().sum()).backward()
            a = torch.rand(3, 4, 1, 1, device=dev, requires_grad=True); b = torch.randn(3, 4, 1, 1, device=dev, requires_grad=True)
            y = ops.ChannelAffine.apply(x, a, b); y.float().sum().backward()
            ops.affine_project(x, a.detach(), b.detach(), 0.2); ops.add_noise(x, x, 0.3)
    if want("K4"):   # K4: per-sample kernel and every tile instantiation (exact fallback included via the on-prototype rows)
        for K in (3, 4, 6, 8, 10, 16):
            C, D, B = 7, 2048, 96
            gp = torch.nn.functional.normalize(torch.randn(C, D, device=dev, generator=g), dim=-1)
            lp = torch.nn.functional.normalize(torch.randn(C, K, D, device=dev, generator=g), dim=-1)
            f = torch.randn(B, D, device=dev, generator=g); y = torch.randint(0, C, (B,), device=dev, generator=g)
            f[:4] = gp[y[:4]]
            for mode in ("sample", "tile_cta") + (("tile_pair",) if K <= 10 else ()):   # per sample / thread groups / warp pairs
                for nf in (False, True):
                    ops.energy_fwd_bwd(f, y, gp, lp, 1.0, 1.0, nf, mode=mode)
        ops.energy_fwd_bwd(f[:, :512].contiguous(), y, gp[:, :512].contiguous(), None, 1.0, 1.0, False, mode="tile")   # generic (non-FULL) path
        # warp-pair kernel, generic D (predicated chunks), odd K (zero-padded pair), an out-of-range target, ragged last batches
        C, K, D, B = 5, 5, 520, 203
        gp = torch.nn.functional.normalize(torch.randn(C, D, device=dev, generator=g), dim=-1)
        lp = torch.nn.functional.normalize(torch.randn(C, K, D, device=dev, generator=g), dim=-1)
        f = torch.randn(B, D, device=dev, generator=g); y = torch.randint(0, C, (B,), device=dev, generator=g); y[17] = C
        f[:3] = lp[y[:3].clamp(max=C - 1), 1]
        for nf in (False, True):
            ops.energy_fwd_bwd(f, y, gp, lp, 1.0, 1.0, nf, mode="tile_pair")
    if want("K8"):   # K8 fwd / bwd, K9
        img = torch.randn(2, 3, 128, 160, device=dev, generator=g)
        out = ops.bicubic_resize(img, (56, 70)); ops.bicubic_resize_bwd(out, (128, 160))
        ops.bicubic_resize(img.half(), (224, 224)); ops.bicubic_resize_bwd(torch.randn(1, 3, 224, 224, device=dev), (512, 512))
        ops.image_to_uint8(img, True)
    if want("K1"):   # K1 / K2 / K3 (stream K=3 with and without inertia, pair kernel K=4 and K=10) / K3' / Lloyd loop (+ peer exchange)
        N, C, D = 1500, 9, 2048
        feats = torch.randn(N, D, device=dev, generator=g); labels = torch.randint(0, C, (N,), device=dev, generator=g); labels[:C] = torch.arange(C, device=dev)
        for K, method in ((3, "kmeans"), (4, "kmeans"), (10, "kmeans"), (3, "agglomerative")):
            per = -(-N // world); sl = slice(per * rank, min(per * (rank + 1), N))
            prototypes.build_prototypes(feats[sl].contiguous(), labels[sl].contiguous(), C, K, method, 3, coll=coll if method == "kmeans" else None)
        prototypes.build_prototypes(feats, labels, C, 3, "kmeans", 2, return_debug=True)     # inertia variant + partial_reduce
        # the opt-in tensor-core variant of the pass (split-fp16 mma.sync), K = 4 and 10
        perm, off = ops.sort_by_class(labels, C)
        xs, _, _ = ops.rownorm_classsum(feats, perm, off)
        for K in (4, 10):
            buf = ops.KMeansBuffers(N, D, C, K, dev)
            idx = off[:-1, None] + (torch.arange(K, device=dev)[None, :] * (off[1:] - off[:-1])[:, None]) // K
            s_, c_ = ops.kmeans_seed(xs, idx.contiguous()); ops.kmeans_update(s_, c_, buf.centroid, buf.cnorm)
            ops.kmeans_assign_accum(xs, off, buf, mma=True)
    torch.cuda.synchronize()
    if coll is not None:
        coll.close()
        torch.distributed.destroy_process_group()
    print("sanitize_run: done", flush=True)


if __name__ == "__main__":
    main()
