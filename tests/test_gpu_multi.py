"""GPU, >= 2 devices: sharded prototype construction over NCCL == single GPU (tools/dist_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_prototypes_nccl(cuda_device):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "[dist_check] PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_generate_data_under_torchrun(cuda_device, tmp_path):
    """`torchrun --nproc-per-node 2 generate_data.py ...`: rank r acts as `--split r --total_split 2`, the prototype stage
    is sharded over the ranks (NCCL class sums, peer-memory k-means exchange / class-sharded agglomerative).  The output
    tree must equal the single-process run's and the cached prototypes must agree to 1e-5."""
    import numpy as np
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    common = ["-d", "caltech-101", "-a", "resnet18", "--tiny_models", "--synthetic_classes", "4", "--synthetic_per_class", "5",
              "--K", "2", "--guidance_type", "transform_guidance", "--guidance_step", "20", "--guidance_period", "2",
              "--constraint_value", "0.2", "--rho", "10", "--strength", "0.5", "--optimize_targets", "global_prototype-local_prototype",
              "--train_batch_size", "2", "--num_images_per_prompt", "1", "--dtype", "fp32"]
    gd = os.path.join(ROOT, "generate_data.py")

    def files(root):
        out = []
        for b, _, fs in os.walk(root):
            out += [os.path.relpath(os.path.join(b, f), root) for f in fs]
        return sorted(out)

    for method in ("agglomerative", "kmeans"):
        d1, d2 = tmp_path / f"single_{method}", tmp_path / f"torchrun_{method}"
        d1.mkdir(); d2.mkdir()
        r = subprocess.run([sys.executable, gd] + common + ["--cluster_method", method, "--output_dir", "out", "--total_split", "1", "--split", "0"],
                           cwd=d1, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                            "--master-port", "29523", gd] + common + ["--cluster_method", method, "--output_dir", "out"],
                           cwd=d2, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
        assert "all 2 ranks: expanded 20 images" in r.stdout
        assert files(d1 / "out") == files(d2 / "out") and len(files(d1 / "out")) == 20
        if method == "agglomerative":
            p = os.path.join("save", "prototypes", "resnet18", "caltech-101", "class_wise_prototype_K2.npz")
            a, b = np.load(d1 / p), np.load(d2 / p)
            for k in ("global_prototypes", "local_prototypes"):
                assert np.allclose(a[k], b[k], rtol=1e-5, atol=1e-7)
