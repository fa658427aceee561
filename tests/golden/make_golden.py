"""Generate golden vectors by running the REFERENCE'S OWN function bodies on CPU.

Run once in the build container (``python tests/golden/make_golden.py``); the outputs
(``tests/golden/*.pt``) are committed, this script is committed, and nothing at test /
bench / smoke time reads /root/reference (it does not exist on the GPU box).

How: the reference modules cannot be imported (diffusers/accelerate/timm/open_clip/matplotlib
are absent), but the functions on the hot path are pure torch/numpy/sklearn.  We pull their
``FunctionDef`` nodes out of the reference sources with ``ast`` and exec them UNMODIFIED in a
namespace that supplies the few free names they use (``torch``, ``Variable``, ``args``,
``np``, ``cluster``, ``AverageMeter``, ``Bar``); ``Tensor.cuda`` / ``Module.cuda`` are patched
to identity because the reference hard-codes ``.cuda()`` (generate_data.py:692,695;
dataloader.py:674).  The third-party objects they call (UNet, VAE, guide encoder, image
processor) are the tiny stand-ins of tests/standins.py; the scheduler is the restated
DDIMScheduler of oracle/ddim.py (diffusers is absent -- that part stays "restated").
"""
from __future__ import annotations

import ast
import os
import sys
import types

import numpy as np
import torch
from torch.autograd import Variable

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle.ddim import OracleDDIMScheduler  # noqa: E402
import standins  # noqa: E402

REF = "/root/reference"


def extract_functions(path, names, namespace):
    src = open(path).read()
    tree = ast.parse(src)
    found = {}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, path, "exec"), namespace)
            found[node.name] = (node.lineno, node.end_lineno)
    missing = set(names) - set(found)
    assert not missing, missing
    return found


def main(out_path=None):
    torch.Tensor.cuda = lambda self, *a, **k: self          # reference hard-codes .cuda()
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.empty_cache = lambda: None

    args = types.SimpleNamespace(do_classifier_free_guidance=True, guidance_scale=7.5, gs=1.0, ls=1.0,
                                 rho=10.0, guidance_period=2, constraint_value=0.2, K=3)
    gd = {"torch": torch, "Variable": Variable, "args": args}
    lines = extract_functions(os.path.join(REF, "generate_data.py"),
                              ["denoise_one_step", "tensor_clamp", "linfball_proj",
                               "transform_guidance", "direct_guidance"], gd)

    class _Meter:
        val = 0.0

    class _Bar:
        def __init__(self, *a, **k): self.suffix = ""
        def next(self): pass
        def finish(self): pass

    from sklearn import cluster
    dl = {"torch": torch, "np": np, "cluster": cluster, "AverageMeter": _Meter, "Bar": _Bar, "print": lambda *a, **k: None}
    lines.update(extract_functions(os.path.join(REF, "dataloader.py"), ["extract_prototype"], dl))

    out = {"reference_lines": lines, "torch_version": torch.__version__}
    sched = OracleDDIMScheduler(50)
    proc = standins.IdentityProcessor()

    # ---------------------------------------------------------------- guidance goldens
    for name, feat_dim, C, B in [("small", 64, 5, 2), ("d2048", 2048, 7, 3), ("b1", 64, 5, 1)]:
        unet, vae, enc = standins.make_nets(seed=11, feat_dim=feat_dim)
        g = torch.Generator().manual_seed(5)
        latents = torch.randn(B, 4, 8, 8, generator=g)
        prompt = torch.randn(2 * B, 6, 8, generator=g)
        targets = torch.randint(0, C, (B,), generator=g).tolist()
        gp = torch.randn(C, feat_dim, generator=g)
        lp = torch.randn(C, args.K, feat_dim, generator=g)
        gp = gp / gp.norm(dim=-1, keepdim=True)
        lp = lp / lp.norm(dim=-1, keepdim=True)
        batch = {"targets": targets}
        case = dict(latents=latents.clone(), prompt_embeds=prompt.clone(), targets=targets, global_proto=gp.clone(),
                    local_proto=lp.clone(), unet=unet.state_dict(), vae=vae.state_dict(), enc=enc.state_dict(),
                    feat_dim=feat_dim, args=dict(vars(args)))

        # denoise_one_step (generate_data.py:109-121) at t=381
        t = sched.timesteps[30]
        with torch.no_grad():
            noise_pred = unet(torch.cat([latents] * 2), t, prompt)[0]
            prev, x0 = gd["denoise_one_step"](latents.clone(), sched, t, unet, prompt, None)
        case["denoise"] = dict(t=int(t), noise_pred=noise_pred, prev=prev, x0=x0)

        # transform_guidance (:687-732) -- channel noise comes from the CPU global RNG: seed it
        torch.manual_seed(77)
        lat_out, score = gd["transform_guidance"](latents.clone(), batch, [381, 361], sched, unet, prompt, None,
                                                  vae, enc, proc, torch.float32, None, gp, lp)
        case["transform"] = dict(seed=77, sub_timesteps=[381, 361], latents_out=lat_out.detach().clone(),
                                 score=float(score))
        # global-only / local-only variants (optimize_targets parsing, generate_data.py:1112-1127)
        torch.manual_seed(78)
        lat_g, score_g = gd["transform_guidance"](latents.clone(), batch, [381, 361], sched, unet, prompt, None,
                                                  vae, enc, proc, torch.float32, None, gp, None)
        case["transform_global_only"] = dict(seed=78, latents_out=lat_g.detach().clone(), score=float(score_g))

        # direct_guidance (:735-767)
        lat_d, x0_d, score_d = gd["direct_guidance"](latents.clone(), batch, 381, sched, unet, prompt, None,
                                                     vae, enc, proc, torch.float32, None, gp, lp)
        case["direct"] = dict(t=381, latents_out=lat_d.clone(), x0=x0_d.clone(), score=float(score_d))
        lat_l, x0_l, score_l = gd["direct_guidance"](latents.clone(), batch, 381, sched, unet, prompt, None,
                                                     vae, enc, proc, torch.float32, None, None, lp)
        case["direct_local_only"] = dict(t=381, latents_out=lat_l.clone(), x0=x0_l.clone(), score=float(score_l))
        out["guidance_" + name] = case

    # ---------------------------------------------------------------- linfball_proj (:124-137)
    g = torch.Generator().manual_seed(9)
    center = torch.randn(2, 4, 8, 8, generator=g)
    tt = center + 0.5 * torch.randn(2, 4, 8, 8, generator=g)
    res = gd["linfball_proj"](center.clone(), 0.2, tt.clone(), in_place=True)
    out["linfball"] = dict(center=center, radius=0.2, t=tt, out=res.clone())

    # ---------------------------------------------------------------- extract_prototype (dataloader.py:664-731)
    for name, C, D, K, nmin, nmax in [("caltech_like", 10, 64, 3, 6, 30), ("d2048", 4, 2048, 3, 5, 14),
                                      ("k5", 8, 128, 5, 8, 40)]:
        rng = np.random.default_rng(1234)
        counts = rng.integers(nmin, nmax + 1, size=C)
        labels = np.repeat(np.arange(C), counts)
        rng.shuffle(labels)                                   # dataset order interleaves classes
        centers = rng.normal(size=(C, 4, D)).astype(np.float32)
        feats = (centers[labels, rng.integers(0, 4, size=len(labels))] * 1.5
                 + rng.normal(size=(len(labels), D)).astype(np.float32)).astype(np.float32)

        class _Model:
            def eval(self): return self
            def encode_image(self, x): return x       # the "images" already are the raw features

        bs = 64
        loader = [(torch.from_numpy(feats[i:i + bs]), torch.from_numpy(labels[i:i + bs]))
                  for i in range(0, len(labels), bs)]
        gl, lc = dl["extract_prototype"](types.SimpleNamespace(K=K), loader, _Model())
        out["proto_" + name] = dict(features=torch.from_numpy(feats), labels=torch.from_numpy(labels), K=K,
                                    global_prototypes=torch.from_numpy(np.asarray(gl, dtype=np.float32)),
                                    local_prototypes=torch.from_numpy(np.asarray(lc, dtype=np.float32)))

    path = out_path or os.path.join(HERE, "reference_golden.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes; reference line spans:", lines)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)   # optional output path (tests regenerate into a temp file)
