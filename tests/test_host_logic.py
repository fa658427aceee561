"""CPU: host-side logic of the product (no kernels run): scheduler tables and index arithmetic against the oracle,
dataset ordering / class-name rules / latent-cache layout, output path layout, CLI flag surface."""
import os
import types

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import ddim as o_ddim


def test_scheduler_tables_and_pairs_match_oracle():
    from distdiff_b200.scheduler import DDIMScheduler, retrieve_timesteps
    s = DDIMScheduler.from_pretrained("CompVis/stable-diffusion-v1-4", subfolder="scheduler")
    assert torch.equal(s.alphas_cumprod, o_ddim.alphas_cumprod())
    ts, n = retrieve_timesteps(s, 50, "cpu")
    assert n == 50 and ts.dtype == torch.int64 and ts.device.type == "cpu"
    assert torch.equal(ts, o_ddim.timesteps(50)) and ts[0] == 981 and ts[-1] == 1
    for t in ts.tolist():
        a, p = s.alpha_pair(t)
        ra, rp = o_ddim.alpha_pair(t)
        assert a == float(ra) and p == float(rp)
    assert s.alpha_pair(1)[1] == float(o_ddim.alphas_cumprod()[0])            # prev_t < 0 -> final_alpha_cumprod = abar[0]
    assert s.init_noise_sigma == 1.0
    x = torch.randn(2, 3)
    assert s.scale_model_input(x, 981) is x


@pytest.mark.parametrize("strength,want", [(0.9, 4), (0.8, 9), (0.5, 25), (1.0, 0), (0.0, 50)])
def test_start_index_truncation(strength, want):
    """generate_data.py:1174: int((1 - strength) * len(timesteps)) truncates a float product (0.9 -> 4, not 5)."""
    from distdiff_b200 import guidance
    assert guidance.start_index(strength, 50) == want == o_ddim.start_index(strength, 50)


def test_guide_timestep_window():
    from distdiff_b200 import guidance
    ts = o_ddim.timesteps(50)
    for step, period in [(20, 2), (50, 50), (1, 1), (30, 5)]:
        g = guidance.guide_timesteps(ts, step, period)
        assert g == o_ddim.guide_timesteps(ts, step, period) and len(g) == period
    assert guidance.guide_timesteps(ts, 20, 2) == [381, 361]
    assert guidance.guide_timesteps(ts, 50, 50) == ts.tolist()                  # "guidance every step" (BASELINE configs[4])
    with pytest.raises(AssertionError):
        guidance.guide_timesteps(ts, 1, 2)                                      # window runs past the last step


def test_split_mask_matches_reference_except_documented_clamp():
    from distdiff_b200 import guidance
    for total, P in [(10, 2), (3030, 8), (7, 7), (100, 3)]:
        got = [guidance.split_mask(total, r, P) for r in range(P)]
        assert sum(got, []) == list(range(total))
        for r in range(P):
            assert got[r] == o_ddim.split_mask(total, r, P)
    # N=5, 4 splits: ceil = 2 -> blocks [0,1] [2,3] [4] and an EMPTY last split; the reference's slice arithmetic
    # (generate_data.py:1004-1007) mis-sizes a non-last split that overshoots -- clamped here, documented in DESIGN.md
    assert [guidance.split_mask(5, r, 4) for r in range(4)] == [[0, 1], [2, 3], [4], []]


def test_image_folder_rules(tmp_path):
    from distdiff_b200 import data
    root = tmp_path / "data" / "caltech-101" / "train"
    for cls, files in {"sea_horse": ["b.jpg", "a.png", "notes.txt"], "ant": ["z.jpeg"], "BACKGROUND_Google": ["x.jpg"],
                       "Faces_easy": ["y.jpg"]}.items():
        (root / cls).mkdir(parents=True)
        for f in files:
            if f.endswith(".txt"):
                (root / cls / f).write_text("x")
            else:
                Image.fromarray(np.zeros((8, 8, 3), np.uint8)).save(root / cls / f)
    ds = data.load_trainset(types.SimpleNamespace(data_root=str(tmp_path / "data"), dataset="caltech-101"), None)
    assert ds.class_names == ["ant", "sea horse"]                               # sorted, '_' -> ' ', the two dropped classes gone
    # files sorted inside a class; like the reference (dataloader.py:284-286) EVERY directory entry is a sample -- a stray
    # non-image file is a dataset error there too (PIL raises when it is opened), not something to filter silently
    assert [os.path.basename(p) for p in ds.paths] == ["z.jpeg", "a.png", "b.jpg", "notes.txt"]
    assert ds.targets == [0, 1, 1, 1]


def test_synthetic_set_is_deterministic():
    from distdiff_b200 import data
    a, b = data.SyntheticCaltech(5, 3), data.SyntheticCaltech(5, 3)
    assert len(a) == 15 and a.targets == [c for c in range(5) for _ in range(3)]
    assert np.array_equal(np.asarray(a.image(7)), np.asarray(b.image(7)))
    assert not np.array_equal(np.asarray(a.image(7)), np.asarray(a.image(8)))
    assert a.paths[4] == "synthetic/class_001/image_0001.jpg" and a.class_names[1] == "class 001"


def test_latent_cache_layout_and_collate(tmp_path, monkeypatch):
    """dataloader.py:788-796 layout (`save/vae_embedding/{dataset}/{model--id}/image_latents.pt`, a list of [1,4,h,w]);
    the second construction must load the file instead of re-encoding; collate keeps `targets` a python list."""
    from distdiff_b200 import data, nets
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    vae = nets.AutoencoderKL(chs=(32, 32, 64, 64)).eval()
    args = types.SimpleNamespace(dataset="caltech-101", data_root="nowhere", synthetic_classes=2, synthetic_per_class=2,
                                 pretrained_model_name_or_path="CompVis/stable-diffusion-v1-4", cache_latents=True, center_crop=True)
    embed = data.random_text_embedder()
    ds = data.SDDataset(args, embed, vae, size=32, device="cpu")
    path = os.path.join("save", "vae_embedding", "caltech-101", "CompVis--stable-diffusion-v1-4", "image_latents.pt")
    assert os.path.exists(path) and not [f for f in os.listdir(os.path.dirname(path)) if ".tmp" in f]
    stored = torch.load(path)
    assert isinstance(stored, list) and len(stored) == 4 and stored[0].shape == (1, 4, 4, 4)

    class Boom:                                                                 # a second run must not touch the VAE
        def __getattr__(self, name):
            raise AssertionError("latents were re-encoded although the cache exists")
    ds2 = data.SDDataset(args, embed, Boom(), size=32, device="cpu")
    assert all(torch.equal(x, y) for x, y in zip(ds.image_latents, ds2.image_latents))
    batch = data.collate_fn([ds[0], ds[3]])
    assert batch["image_latents"].shape == (2, 4, 4, 4) and batch["input_ids"].shape == (2, 77, 768)
    assert batch["targets"] == [0, 1] and isinstance(batch["targets"], list)
    assert batch["class_names"] == ["class 000", "class 001"]
    assert torch.equal(embed("a photo of a class 000."), ds[0]["instance_prompt_ids"])   # dataloader.py:52-62 template


def test_block_sharded_latents_equal_the_single_process_ones(tmp_path, monkeypatch):
    """each --split process encodes only its block, advancing the crop / posterior RNG streams through the
    skipped images exactly as the full encode does (dataloader.py:798-811 order), so block latents == single-process
    latents BIT FOR BIT, images outside the block are refused, and the merged list is the reference-layout cache file."""
    from distdiff_b200 import data, guidance, nets
    monkeypatch.chdir(tmp_path)
    torch.manual_seed(0)
    vae = nets.AutoencoderKL(chs=(32, 32, 64, 64)).eval()
    args = types.SimpleNamespace(dataset="caltech-101", data_root="nowhere", synthetic_classes=3, synthetic_per_class=3, seed=42,
                                 pretrained_model_name_or_path="x", cache_latents=False, center_crop=False)
    embed = data.random_text_embedder()
    torch.manual_seed(42)
    full = data.SDDataset(args, embed, vae, size=32, device="cpu")
    assert all(t is not None and t.shape == (1, 4, 4, 4) for t in full.image_latents)
    blocks = []
    for split in range(2):
        mask = guidance.split_mask(len(full), split, 2)
        torch.manual_seed(42)                                                   # same --seed in every process (set_seed)
        part = data.SDDataset(args, embed, vae, size=32, device="cpu", only=mask)
        blocks.append({j: t for j, t in enumerate(part.image_latents) if t is not None})
        for j in range(len(full)):
            if j in mask:
                assert torch.equal(part.image_latents[j], full.image_latents[j])
            else:
                assert part.image_latents[j] is None
                with pytest.raises(IndexError):
                    part[j]
    # merge (torchrun: all_gather_object of the blocks) -> rank 0 publishes the complete cache file, atomically
    args.cache_latents = True
    merged = part.merge_blocks(lambda mine: blocks, rank=0)
    assert all(torch.equal(a, b) for a, b in zip(merged, full.image_latents))
    stored = torch.load(part._cache_path())
    assert len(stored) == len(full) and all(torch.equal(a, b) for a, b in zip(stored, full.image_latents))
    torch.manual_seed(43)                                                       # a different --seed gives different draws
    other = data.SDDataset(types.SimpleNamespace(**{**vars(args), "cache_latents": False}), embed, vae, size=32, device="cpu", only=[0])
    assert not torch.equal(other.image_latents[0], full.image_latents[0])


def test_image_size_probe_matches_decoded_size(tmp_path):
    from distdiff_b200 import data
    root = tmp_path / "data" / "caltech-101" / "train" / "ant"
    root.mkdir(parents=True)
    Image.fromarray(np.zeros((20, 31, 3), np.uint8)).save(root / "a.png")
    im = Image.fromarray(np.zeros((20, 31, 3), np.uint8))
    ex = im.getexif(); ex[0x0112] = 6                                           # rotated 90 degrees: axes swap
    im.save(root / "b.jpg", exif=ex)
    ds = data.load_trainset(types.SimpleNamespace(data_root=str(tmp_path / "data"), dataset="caltech-101"), None)
    for j in range(len(ds)):
        assert ds.size(j) == ds.image(j).size


def test_output_path_layout():
    from distdiff_b200 import expand
    args = types.SimpleNamespace(output_dir="data_expand")
    batch = {"class_names": ["sea horse", "ant"], "image_paths": ["x/sea_horse/image_0007.v2.jpg", "y/ant/a.png"]}
    # generate_data.py:1134-1135: basename.split('.')[0] -- everything after the FIRST dot is dropped
    assert expand.output_path(args, batch, 0, 3) == "data_expand/sea horse/image_0007_expand_3.png"
    assert expand.output_path(args, batch, 1, 0) == "data_expand/ant/a_expand_0.png"


def test_cli_surface_matches_reference_flags():
    import generate_data as gd
    a = gd.parse_args([])
    # reference defaults (generate_data.py:167-453)
    assert (a.total_split, a.split, a.num_images_per_prompt, a.K, a.guidance_scale, a.seed, a.train_batch_size) == (8, 0, 4, 3, 7.5, 42, 2)
    assert (a.strength, a.rho, a.gs, a.ls, a.constraint_value, a.guidance_step, a.guidance_period) == (0.9, 10.0, 1.0, 1.0, 0.8, 1, 1)
    assert a.guidance_type is None and a.cluster_method == "agglomerative"
    b = gd.parse_args(["--guidance_type", "transform_guidance", "--optimize_targets", "global_prototype-local_prototype", "--K", "5",
                       "--report_to", "wandb", "--some_training_only_flag"])     # flags off the expansion path are ignored, not fatal
    assert b.guidance_type == "transform_guidance" and b.K == 5 and b.optimize_targets == "global_prototype-local_prototype"
    with pytest.raises(SystemExit):
        gd.parse_args(["--guidance_type", "not_a_mode"])


def test_gpu_decode_loader_logic_on_cpu(tmp_path):
    """prototypes.GpuDecodeLoader (the opt-in --gpu_decode loader) run with device='cpu': torchvision decodes on the host
    then, but the batching, the JPEG sniffing, the per-file PIL fallback (a PNG) and the resize / normalise arithmetic are the
    same code.  Against the reference's PIL transform (dataloader.py:736-742) the tensors agree to a few grey levels."""
    import numpy as np
    import torch
    from PIL import Image
    from torchvision import transforms
    from distdiff_b200 import data as dd_data, prototypes
    syn = dd_data.SyntheticCaltech(2, 5, hw=(120, 160))
    root = tmp_path / "train"
    for i in range(len(syn)):
        d = root / syn.class_names[syn.targets[i]].replace(" ", "_")
        d.mkdir(parents=True, exist_ok=True)
        img = syn.image(i).resize((160 + 3 * i, 120 + 5 * (i % 3)), Image.BICUBIC)
        img.save(d / (f"image_{i:04d}.png" if i == 2 else f"image_{i:04d}.jpg"), quality=95)
    tf = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                             transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    base = dd_data.ImageFolderSorted(str(root), tf)
    idx = list(range(len(base)))
    loader = prototypes.GpuDecodeLoader([base.paths[i] for i in idx], [base.targets[i] for i in idx], torch.device("cpu"), batch_size=4,
                                        fallback=base, fallback_index=idx, transform=tf)
    xs, ys = zip(*list(loader))
    x = torch.cat(xs); y = torch.cat(ys)
    assert len(xs) == len(loader) == 3 and x.shape == (10, 3, 224, 224) and y.tolist() == base.targets
    assert loader.decoded_on_gpu == 9                                        # everything but the PNG
    ref = torch.stack([base[i][0] for i in idx])
    assert torch.equal(x[2], ref[2])                                         # the PNG went through the reference's own path
    err = (x - ref).abs() * 0.226 * 255                                      # back to grey levels (std ~0.226)
    assert float(err.max()) <= 6.0 and float(err.mean()) <= 0.6
