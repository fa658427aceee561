"""Drop-in for the reference's ``generate_data.py`` (DistDiff data expansion) on the B200-native hot path.

Same command line for everything on the path (flag names, defaults and meanings follow
/root/reference/generate_data.py:164-639; the canonical values are scripts/exps/expand_diff.sh:3-16), same
``--split/--total_split`` image sharding (:1002-1009), same output tree
``{output_dir}/{class_name}/{stem}_expand_{i}.png`` with skip-if-exists resume (:1132-1143).  DreamBooth
leftovers the reference parses but never uses are accepted and ignored.

    python generate_data.py -d caltech-101 -a resnet50 --guidance_type transform_guidance --K 3 \
        --guidance_step 20 --guidance_period 2 --constraint_value 0.2 --rho 10 --strength 0.5 \
        --optimize_targets global_prototype-local_prototype --train_batch_size 1 --total_split 4 --split 0

Without network access the SD-v1.4 / ResNet-50 weights are random-init in the reference's shapes unless
``--pretrained_model_name_or_path`` / ``--encoder_weight_path`` point at local state dicts.
"""
from __future__ import annotations

import argparse
import logging
import os
import sys

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def parse_args(input_args=None):
    p = argparse.ArgumentParser(description="DistDiff expansion (B200-native hot path)")
    # ---- reference flags on the hot path (generate_data.py line numbers in comments) ----
    p.add_argument("--pretrained_model_name_or_path", type=str, default="CompVis/stable-diffusion-v1-4")   # :167
    p.add_argument("-d", "--dataset", default="caltech-101", type=str)                                       # :187
    p.add_argument("-a", "--arch", default="resnet50", type=str)                                             # :195 (ref default open_clip_vit_b32 is not available offline)
    p.add_argument("--encoder_weight_path", type=str, default=None)                                          # :204
    p.add_argument("--guidance_type", type=str, default=None, choices=[None, "transform_guidance", "direct_guidance"])  # :212
    p.add_argument("--constraint_value", type=float, default=0.8)                                            # :216
    p.add_argument("--steps", type=int, default=50)                                                          # :217 (unused upstream too)
    p.add_argument("--K", type=int, default=3)                                                               # :218
    p.add_argument("--guidance_step", type=int, default=1)                                                   # :219
    p.add_argument("--guidance_period", type=int, default=1)                                                 # :220
    p.add_argument("--total_split", type=int, default=8)                                                     # :221
    p.add_argument("--split", type=int, default=0)                                                           # :222
    p.add_argument("--num_images_per_prompt", type=int, default=4)                                           # :223
    p.add_argument("--first_image_index", type=int, default=0)                                               # :224
    p.add_argument("--optimize_targets", type=str, default=None)                                             # :226
    p.add_argument("--rho", type=float, default=10.0)                                                        # :229
    p.add_argument("--gs", type=float, default=1.0)                                                          # :230
    p.add_argument("--ls", type=float, default=1.0)                                                          # :231
    p.add_argument("--strength", type=float, default=0.9)                                                    # :233
    p.add_argument("--language_enhance", action="store_true")                                                # :239
    p.add_argument("--cache_dir", type=str, default=None)                                                    # :261
    p.add_argument("--resolution", type=int, default=512)                                                    # :277
    p.add_argument("--output_dir", type=str, default="data_expand")                                          # :365
    p.add_argument("--seed", type=int, default=42)                                                           # :370
    p.add_argument("--train_batch_size", type=int, default=2)                                                # :377
    p.add_argument("--gradient_checkpointing", action="store_true")                                          # :439
    p.add_argument("--guidance_scale", type=float, default=7.5)                                              # :445
    p.add_argument("--do_classifier_free_guidance", type=bool, default=True)                                 # :453 (any non-empty string parses True upstream)
    p.add_argument("--offset_noise", action="store_true")                                                    # :558
    p.add_argument("--text_to_img", action="store_true")                                                     # :1150 (pure-noise start; not built)
    p.add_argument("--mixed_precision", type=str, default=None)                                              # :523 (ignored upstream: fp16 hard-coded :1039)
    p.add_argument("--dataloader_num_workers", type=int, default=0)
    # ---- additions; defaults reproduce the reference ----
    p.add_argument("--cluster_method", type=str, default="agglomerative", choices=["agglomerative", "kmeans"])
    p.add_argument("--kmeans_iters", type=int, default=20)
    p.add_argument("--dtype", type=str, default="fp16", choices=["fp16", "bf16", "fp32"])
    p.add_argument("--data_root", type=str, default="data")
    p.add_argument("--synthetic_classes", type=int, default=100)
    p.add_argument("--synthetic_per_class", type=int, default=30)
    p.add_argument("--cuda_graph", action="store_true", help="replay the unguided step (UNet + K5) from a CUDA graph")
    p.add_argument("--cache_latents", action="store_true", help="persist save/vae_embedding/.../image_latents.pt like the reference")
    p.add_argument("--shard_latents", action="store_true", help="(default behaviour now; kept for compatibility)")
    p.add_argument("--gpu_decode", action="store_true",
                   help="prototype stage: decode + resize the train JPEGs on the GPU (nvJPEG) instead of PIL; prototypes ~5e-3 from the PIL path")
    p.add_argument("--tiny_models", action="store_true", help="tiny random-init UNet/VAE/guide (tests)")
    p.add_argument("--max_batches", type=int, default=None, help="stop after this many batches (smoke runs)")
    args, unknown = p.parse_known_args(input_args)
    if unknown:
        logging.getLogger("distdiff_b200").warning("ignoring flags that are not on the expansion path: %s", unknown)
    # flags that change what is generated but are outside the hot path built here: refuse instead of silently ignoring
    if args.text_to_img:
        raise SystemExit("--text_to_img (generate_data.py:1150-1158, sampling from pure noise) is not implemented; the img2img path is")
    if args.language_enhance:
        raise SystemExit("--language_enhance (LLM-rewritten prompts, generate_data.py:239) is not implemented")
    if args.optimize_targets is not None and args.encoder_weight_path is not None and not os.path.exists(args.encoder_weight_path):
        raise SystemExit(f"--encoder_weight_path {args.encoder_weight_path!r} does not exist (the reference asserts this when "
                         "--optimize_targets is set)")
    return args


def build_guide(args, num_classes, device):
    """The guide network is sized from the DATASET (len(class_names), like the reference), so a real checkpoint loads
    strictly for any class count; missing weights raise (nets.create_model)."""
    from distdiff_b200 import nets
    if args.tiny_models:
        guide = nets.create_model("resnet18", num_classes=num_classes)
    else:
        guide = nets.create_model(args.arch, num_classes=num_classes, weight_path=args.encoder_weight_path)
        if args.encoder_weight_path is None and args.optimize_targets is not None:
            logging.getLogger("distdiff_b200").warning("no --encoder_weight_path: the guide network is RANDOM-INIT (synthetic runs only)")
    guide.requires_grad_(False)
    guide.eval()
    return guide.to(device)


def build_models(args, device, weight_dtype):
    from distdiff_b200 import nets
    if args.tiny_models:
        unet = nets.UNet2DConditionModel(block_out_channels=(32, 64, 64, 64), cross_attention_dim=768, heads=2)
        vae = nets.AutoencoderKL(chs=(32, 32, 64, 64))
    else:
        unet = nets.UNet2DConditionModel()
        vae = nets.AutoencoderKL()
    local = args.pretrained_model_name_or_path
    if local and os.path.isdir(local):
        for name, m in (("unet", unet), ("vae", vae)):
            f = os.path.join(local, name, "state_dict.pt")
            if os.path.exists(f):
                m.load_state_dict(torch.load(f, map_location="cpu"))
    if args.gradient_checkpointing:
        unet.enable_gradient_checkpointing()                                                                # :1049-1050
    for m in (unet, vae):
        m.requires_grad_(False)
        m.eval()
    return unet.to(device), vae.to(device)


def main(args):
    from torch.utils.data import DataLoader, Subset
    from distdiff_b200 import data as dd_data, expand, guidance, nets, prototypes
    from distdiff_b200.scheduler import DDIMScheduler

    logging.basicConfig(format="%(asctime)s - %(levelname)s - %(name)s - %(message)s", datefmt="%m/%d/%Y %H:%M:%S",
                        level=logging.INFO)
    if not torch.cuda.is_available():
        raise RuntimeError("generate_data.py needs a CUDA device: the guidance hot path has no CPU fallback")
    device = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(device)
    # Launched by torchrun (one process per GPU): rank r IS the reference's `--split r --total_split P` process
    # (scripts/exps/expand_diff.sh starts those by hand), and the prototype stage is sharded over the ranks -- features
    # of 1/P of the images per GPU, class sums over NCCL, k-means centroids through the fused NVLink peer exchange --
    # instead of every process recomputing all of it (generate_data.py:1104).
    world, rank = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0))
    coll = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        coll = prototypes.make_collective()   # fused peer exchange when every GPU maps every other one, else NCCL
        args.split, args.total_split = rank, world
        logging.getLogger("distdiff_b200").info("torchrun: rank %d of %d -> --split %d --total_split %d", rank, world, rank, world)
    if args.seed is not None:
        expand.set_seed(args.seed)                                                                          # :860-861
    weight_dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype]       # :1039
    noise_scheduler = DDIMScheduler.from_pretrained(args.pretrained_model_name_or_path, subfolder="scheduler")  # :863
    unet, vae = build_models(args, device, weight_dtype)

    # dataset + latents (VAE encode in fp32 like the reference, :983), then the --split block (:1001-1009).
    # LIMITATION: there is no CLIP text encoder offline -- prompt embeddings come from a fixed random embedding table
    # (deterministic per class name); images are NOT comparable with the reference's until a real text encoder is plugged
    # in here.  Everything downstream of the embeddings is the reference's computation.
    embed = dd_data.random_text_embedder()
    # Every process VAE-encodes only its own --split block (the reference re-encodes the whole set per process,
    # dataloader.py:798-811); the RNG streams are advanced through the skipped images, so the block's latents are the
    # single-process ones bit for bit.  A complete image_latents.pt (--cache_latents) is written by a single process
    # (total_split 1) or, under torchrun, by rank 0 after the blocks are gathered.
    n_images = len(dd_data.load_trainset(args, None))
    mask = guidance.split_mask(n_images, args.split, args.total_split)
    only = mask if (args.total_split > 1 and (world > 1 or not args.cache_latents)) else None
    dataset = dd_data.SDDataset(args, embed, vae, size=512 if not args.tiny_models else 64, device=device, only=only)
    if world > 1 and args.cache_latents and only is not None:
        import torch.distributed as dist

        def gather(obj):
            box = [None] * world
            dist.all_gather_object(box, obj)
            return box
        dataset.merge_blocks(gather, rank)
    loader = DataLoader(Subset(dataset, mask), batch_size=args.train_batch_size, shuffle=False,
                        collate_fn=dd_data.collate_fn, num_workers=args.dataloader_num_workers, drop_last=False)

    # prototypes (:1100-1127) -- guide in fp32 for extraction (dataloader.py:745), then cast (:1106)
    args.num_classes = len(dataset.class_names)
    image_encoder = build_guide(args, args.num_classes, device)
    global_np, local_np = prototypes.extract_prototypes_with_encoder(args, image_encoder, coll=coll)
    if args.optimize_targets is not None:
        args.optimize_targets = args.optimize_targets.split("-")                                            # :1109
    print(f"optimize strategy: {args.guidance_type}, target: {args.optimize_targets}, learning rate: {args.rho}")
    gproto, lproto = prototypes.prototypes_to_device(global_np, local_np, args.optimize_targets, device)
    if lproto is not None:
        print("local prototype shape:", lproto.shape)

    unet.to(dtype=weight_dtype); vae.to(dtype=weight_dtype); image_encoder.to(dtype=weight_dtype)          # :1057-1064,1106
    ex = expand.Expander(args, unet, vae, image_encoder, nets.VaeImageProcessor(), noise_scheduler, gproto, lproto,
                         weight_dtype=weight_dtype, device=device, use_cuda_graph=args.cuda_graph)
    if args.max_batches is not None:
        import itertools
        loader = list(itertools.islice(iter(loader), args.max_batches))
    n = expand.run_expansion(args, ex, loader, save=True)
    print(f"expanded {n} images into {args.output_dir}")
    if coll is not None:
        import torch.distributed as dist
        tot = torch.tensor([n], device=device)
        dist.all_reduce(tot)                     # also the final barrier: every rank's files are on disk
        if rank == 0:
            print(f"all {world} ranks: expanded {int(tot)} images")
        coll.close()
        dist.destroy_process_group()
    return n


if __name__ == "__main__":
    main(parse_args())
