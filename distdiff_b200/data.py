"""Datasets for the expansion driver: a Caltech-101-shaped image folder if present, else a deterministic
synthetic stand-in of the same shape (there is no network: SURVEY.md section 8d).

Reference behaviour kept: class folders -> labels in class-name order, ``_`` -> space in class names
(dataloader.py:294), BACKGROUND_Google / Faces_easy dropped (:276), prompt template ``"a photo of a {}."``
(dataloader.py:52-62, caltech-101), VAE latents computed once and cached as a python list of [1,4,64,64]
tensors in ``save/vae_embedding/{dataset}/{model--id}/image_latents.pt`` (:788-796) -- written atomically
here (the reference races when several splits start together).  Files inside a class are SORTED (the
reference uses unsorted os.listdir, :285, which makes the position-indexed latent cache host dependent).
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch
from PIL import Image
from torch.utils import data

CALTECH_DROP = ("BACKGROUND_Google", "Faces_easy")
PROMPT_TEMPLATE = "a photo of a {}."


class SyntheticCaltech(data.Dataset):
    """100 classes x ``per_class`` RGB images (~300x200), pixels from numpy default_rng(seed, index)."""

    def __init__(self, num_classes=100, per_class=30, transform=None, seed=1234, hw=(200, 300)):
        self.num_classes, self.per_class, self.transform, self.seed, self.hw = num_classes, per_class, transform, seed, hw
        self.class_names = [f"class {c:03d}" for c in range(num_classes)]
        self.targets = [c for c in range(num_classes) for _ in range(per_class)]
        self.paths = [f"synthetic/{self.class_names[c].replace(' ', '_')}/image_{i:04d}.jpg"
                      for c in range(num_classes) for i in range(per_class)]

    def __len__(self):
        return len(self.targets)

    def image(self, idx) -> Image.Image:
        rng = np.random.default_rng([self.seed, idx])
        c = self.targets[idx]
        base = np.random.default_rng([self.seed, 10_000_000 + c]).integers(0, 256, size=(1, 1, 3))
        low = rng.integers(0, 256, size=(self.hw[0] // 8, self.hw[1] // 8, 3))
        img = (0.5 * base + 0.5 * np.kron(low, np.ones((8, 8, 1)))).astype(np.uint8)   # blocky, class-tinted
        return Image.fromarray(img, "RGB")

    def size(self, idx):
        return self.hw[1], self.hw[0]                                          # (w, h) like PIL

    def __getitem__(self, idx):
        img = self.image(idx)
        if self.transform is not None:
            img = self.transform(img)
        return img, self.targets[idx]


class ImageFolderSorted(data.Dataset):
    def __init__(self, root, transform=None):
        # dataloader.py:275-286: every entry of the train dir is a category (minus the two Caltech drops) and every entry
        # of a category dir is a sample -- no extension filter; only the order differs (sorted, see the module docstring)
        classes = [d for d in sorted(os.listdir(root)) if d not in CALTECH_DROP]
        self.class_names = [c.replace("_", " ") for c in classes]
        self.paths, self.targets = [], []
        for ci, c in enumerate(classes):
            for f in sorted(os.listdir(os.path.join(root, c))):
                self.paths.append(os.path.join(root, c, f))
                self.targets.append(ci)
        self.transform = transform

    def __len__(self):
        return len(self.targets)

    def image(self, idx):
        from PIL.ImageOps import exif_transpose
        image = exif_transpose(Image.open(self.paths[idx]))                    # dataloader.py:79-82
        return image if image.mode == "RGB" else image.convert("RGB")

    def size(self, idx):
        """(w, h) as ``image(idx).size`` would report, from the file header only (no pixel decode)."""
        with Image.open(self.paths[idx]) as im:
            w, h = im.size
            if im.getexif().get(0x0112) in (5, 6, 7, 8):                       # EXIF orientations that swap the axes
                w, h = h, w
        return w, h

    def __getitem__(self, idx):
        img = self.image(idx)
        if self.transform is not None:
            img = self.transform(img)
        return img, self.targets[idx]


def load_trainset(args, transform=None):
    root = os.path.join(getattr(args, "data_root", "data"), str(args.dataset), "train")
    if os.path.isdir(root):
        return ImageFolderSorted(root, transform)
    return SyntheticCaltech(getattr(args, "synthetic_classes", 100), getattr(args, "synthetic_per_class", 30), transform)


class SDDataset(data.Dataset):
    """dataloader.py:750-852: per item the cached VAE latent, the class prompt embedding and the unconditional
    embedding (stored under the reference's key names), the label, class name and image path."""

    def __init__(self, args, text_embed_fn, vae, size=512, device="cuda", latent_dtype=torch.float32, only=None):
        """``only``: indices this process will read (its ``--split`` block): the VAE encode then covers just those images
        (the reference encodes the whole set in every process, dataloader.py:798-811), while the two RNG
        streams are advanced through the skipped images exactly as a full encode would -- the block's latents equal the
        single-process ones bit for bit."""
        from torchvision import transforms
        self.args = args
        self.base = load_trainset(args, None)
        self.class_names = self.base.class_names
        self.size = size
        center_crop = self.center_crop = bool(getattr(args, "center_crop", False))   # dataloader.py:757-764
        self.tf = transforms.Compose([transforms.Resize(size, interpolation=transforms.InterpolationMode.BILINEAR),
                                      transforms.CenterCrop(size) if center_crop else transforms.RandomCrop(size),
                                      transforms.ToTensor(), transforms.Normalize([0.5], [0.5])])
        self.prompt_embeds = [text_embed_fn(PROMPT_TEMPLATE.format(n)) for n in self.class_names]
        self.uncond_embeds = text_embed_fn("")
        self.image_latents = self._latents(vae, device, latent_dtype, only)

    def _cache_path(self):
        model_id = str(getattr(self.args, "pretrained_model_name_or_path", "random-init")).replace("/", "--")
        return os.path.join("save", "vae_embedding", str(self.args.dataset), model_id, "image_latents.pt")

    # ---- the reference's draw order (dataloader.py:798-811), image by image: the crop position from the CPU global
    # generator (RandomCrop), then the posterior noise randn([1,4,h,w]) from the DEVICE's default generator
    # (latent_dist.sample()).  Both streams are consumed for EVERY image of the set in order; an image outside this
    # process's block only advances the two generators (header-only size probe, no decode, no encode), so the latents of
    # a block are bit-identical to the ones a single process computes for the whole set.
    def _resized_hw(self, j):
        w, h = self.base.size(j)                                               # PIL convention (w, h), EXIF orientation applied
        short, long_ = (w, h) if w <= h else (h, w)
        new_short, new_long = self.size, int(self.size * long_ / short)        # transforms.Resize(int): smaller edge -> size
        return (new_long, new_short) if w <= h else (new_short, new_long)      # (h, w)

    def _skip_draws(self, j, device, wdt):
        from torchvision import transforms
        if not self.center_crop:
            h, w = self._resized_hw(j)
            transforms.RandomCrop.get_params(torch.empty(3, h, w), (self.size, self.size))
        torch.randn(1, 4, self.size // 8, self.size // 8, device=device, dtype=wdt)

    @torch.no_grad()
    def _latents(self, vae, device, dtype, only=None) -> List[Optional[torch.Tensor]]:
        path = self._cache_path()
        cache = bool(getattr(self.args, "cache_latents", False))
        if cache and os.path.exists(path):
            return torch.load(path, map_location="cpu")
        n = len(self.base)
        want = None if only is None else set(int(j) for j in only)
        out: List[Optional[torch.Tensor]] = [None] * n
        wdt = next(vae.parameters()).dtype
        sf = vae.config.scaling_factor
        pend = []

        def flush():
            if not pend:
                return
            dist_ = vae.encode(torch.stack([im for _, im, _ in pend]).to(device, wdt)).latent_dist     # batched encoder forward
            for k, (j, _im, noise) in enumerate(pend):
                out[j] = ((dist_.mean[k:k + 1] + dist_.std[k:k + 1] * noise) * sf).to("cpu", dtype)      # :808-809
            pend.clear()

        for j in range(n):
            if want is not None and j not in want:
                self._skip_draws(j, device, wdt)
                continue
            img = self.tf(self.base.image(j))                                   # crop draw (CPU generator)
            noise = torch.randn(1, 4, img.shape[1] // 8, img.shape[2] // 8, device=device, dtype=wdt)   # posterior draw (device generator)
            pend.append((j, img, noise))
            if len(pend) == 16:
                flush()
        flush()
        if cache and want is None:
            self.save_cache(out)
        return out

    def save_cache(self, latents) -> None:
        """Atomic write of the complete list in the reference's layout (dataloader.py:788-796)."""
        if any(t is None for t in latents):
            raise ValueError("image_latents.pt must hold every image of the set")
        path = self._cache_path()
        os.makedirs(os.path.dirname(path), exist_ok=True)
        tmp = f"{path}.tmp{os.getpid()}"
        torch.save(list(latents), tmp)
        os.replace(tmp, path)

    def merge_blocks(self, gather_object, rank: int):
        """torchrun: every rank encoded its own block; ``gather_object(obj) -> [obj per rank]`` (e.g.
        torch.distributed.all_gather_object) merges them, rank 0 publishes the cache file.  Returns the full list."""
        mine = {j: t for j, t in enumerate(self.image_latents) if t is not None}
        full = list(self.image_latents)
        for blk in gather_object(mine):
            for j, t in blk.items():
                full[j] = t
        if rank == 0 and getattr(self.args, "cache_latents", False):
            self.save_cache(full)
        return full

    def __len__(self):
        return len(self.base)

    def __getitem__(self, idx):
        y = self.base.targets[idx]
        if self.image_latents[idx] is None:
            raise IndexError(f"image {idx} is outside the block this process encoded (its --split block)")
        return {"image_latents": self.image_latents[idx], "instance_prompt_ids": self.prompt_embeds[y],
                "uncond_inputs_ids": self.uncond_embeds, "targets": y, "class_names": self.class_names[y],
                "image_paths": self.base.paths[idx]}


def collate_fn(examples):
    """generate_data.py:642-684 -- ``targets`` stays a python list (:650,671)."""
    return {"input_ids": torch.cat([e["instance_prompt_ids"] for e in examples], 0),
            "uncond_inputs_ids": torch.cat([e["uncond_inputs_ids"] for e in examples], 0),
            "image_latents": torch.cat([e["image_latents"] for e in examples], 0),
            "targets": [e["targets"] for e in examples],
            "class_names": [e["class_names"] for e in examples],
            "image_paths": [e["image_paths"] for e in examples]}


def random_text_embedder(seed=0, tokens=77, dim=768, dtype=torch.float32):
    """Stand-in for CLIPTextModel (no tokenizer vocabulary offline): a fixed random [1,77,768] per prompt string."""
    import zlib

    def embed(prompt: str) -> torch.Tensor:
        g = torch.Generator().manual_seed(seed * 1_000_003 + zlib.crc32(prompt.encode()))
        return torch.randn(1, tokens, dim, generator=g).to(dtype)

    return embed
