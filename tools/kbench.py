"""CLI for distdiff_b200.microbench: `python tools/kbench.py [K1 K3 ...] [--short]`; KBENCH_KS=3,10 limits the K sweep."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from distdiff_b200 import microbench  # noqa: E402

only = [a for a in sys.argv[1:] if not a.startswith("-")]
ks = tuple(int(v) for v in os.environ.get("KBENCH_KS", "3,4,5,6,7,8,9,10").split(","))
microbench.run(want=lambda n: not only or any(o in n for o in only), iters=3 if "--short" in sys.argv else 10, ks=ks,
               latent_dtypes=(torch.float32, torch.bfloat16, torch.float16), emit=lambda r: print(json.dumps(r), flush=True))
