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
        """``only``: indices this process will read (its ``--split`` block).  Used only with ``args.shard_latents``:
        the VAE encode then covers just those images, each with its own generator seeded by (seed, index), so the
        latents do not depend on how the set is split -- but they are NOT the reference's draws (the reference encodes
        the whole set in every process from one sequential RNG stream, dataloader.py:798-811), hence opt-in."""
        from torchvision import transforms
        self.args = args
        self.base = load_trainset(args, None)
        self.class_names = self.base.class_names
        self.size = size
        center_crop = bool(getattr(args, "center_crop", False))                # dataloader.py:757-764
        self.tf = transforms.Compose([transforms.Resize(size, interpolation=transforms.InterpolationMode.BILINEAR),
                                      transforms.CenterCrop(size) if center_crop else transforms.RandomCrop(size),
                                      transforms.ToTensor(), transforms.Normalize([0.5], [0.5])])
        self.prompt_embeds = [text_embed_fn(PROMPT_TEMPLATE.format(n)) for n in self.class_names]
        self.uncond_embeds = text_embed_fn("")
        if getattr(args, "shard_latents", False):
            self.image_latents = self._latents_sharded(vae, device, latent_dtype, range(len(self.base)) if only is None else only)
        else:
            self.image_latents = self._latents(vae, device, latent_dtype)

    def _cache_path(self):
        model_id = str(getattr(self.args, "pretrained_model_name_or_path", "random-init")).replace("/", "--")
        return os.path.join("save", "vae_embedding", str(self.args.dataset), model_id, "image_latents.pt")

    @torch.no_grad()
    def _latents(self, vae, device, dtype) -> List[torch.Tensor]:
        path = self._cache_path()
        if getattr(self.args, "cache_latents", False) and os.path.exists(path):
            return torch.load(path, map_location="cpu")
        out = []
        bs = 16
        for i in range(0, len(self.base), bs):
            imgs = torch.stack([self.tf(self.base.image(j)) for j in range(i, min(i + bs, len(self.base)))]).to(device)
            z = vae.encode(imgs.to(next(vae.parameters()).dtype)).latent_dist.sample() * vae.config.scaling_factor  # :808
            out.extend(t[None].to("cpu", dtype) for t in z)
        if getattr(self.args, "cache_latents", False):
            os.makedirs(os.path.dirname(path), exist_ok=True)
            tmp = f"{path}.tmp{os.getpid()}"
            torch.save(out, tmp)
            os.replace(tmp, path)
        return out

    @torch.no_grad()
    def _latents_sharded(self, vae, device, dtype, indices) -> List[Optional[torch.Tensor]]:
        """Encode only ``indices`` (no P-fold redundant encode under --total_split P).  Crop position
        and posterior noise of image j come from generators seeded by (seed, j): split-invariant by construction."""
        out: List[Optional[torch.Tensor]] = [None] * len(self.base)
        seed = int(getattr(self.args, "seed", 0) or 0)
        idx = [int(j) for j in indices]
        bs = 16
        wdt = next(vae.parameters()).dtype
        for i in range(0, len(idx), bs):
            chunk = idx[i:i + bs]
            imgs = []
            for j in chunk:
                state = torch.random.get_rng_state()
                torch.manual_seed(seed * 1_000_003 + j)                       # RandomCrop draws from the global CPU generator
                imgs.append(self.tf(self.base.image(j)))
                torch.random.set_rng_state(state)
            moments = vae.encode(torch.stack(imgs).to(device, wdt)).latent_dist
            for k, j in enumerate(chunk):
                g = torch.Generator(device="cpu").manual_seed(seed * 1_000_003 + 7919 + j)
                noise = torch.randn(moments.mean[k].shape, generator=g).to(device, moments.mean.dtype)
                z = (moments.mean[k] + moments.std[k] * noise) * vae.config.scaling_factor
                out[j] = z[None].to("cpu", dtype)
        return out

    def __len__(self):
        return len(self.base)

    def __getitem__(self, idx):
        y = self.base.targets[idx]
        if self.image_latents[idx] is None:
            raise IndexError(f"image {idx} is outside the block this process encoded (--shard_latents)")
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
