"""Raw tcgen05.mma rate of every operand form (single CTA, no other traffic): cycles per K=16 instruction."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ring_flash_attn_b200.ops import cuda_ext  # noqa: E402
from probe_descriptors import MODES  # noqa: E402

C = cuda_ext.load()
for name, (a_kind, b_kind, n, kdim, ash, bsh, _ref) in MODES.items():
    a = torch.randn(ash, device="cuda").to(torch.bfloat16)
    b = torch.randn(bsh, device="cuda").to(torch.bfloat16)
    reps = 200
    cyc = C.probe(a, b, [a_kind, b_kind, n, kdim, -1, -1, -1, -1, -1, -1, reps])
    torch.cuda.synchronize()
    tot, iss = int(cyc[0]), int(cyc[1])
    n_mma = reps * kdim // 16
    ideal = 128 * n / 256
    print(f"{name:20s} n={n:3d}  {tot / n_mma:7.1f} cyc/MMA total, {iss / n_mma:7.1f} cyc/MMA issue  (ideal {ideal:.0f})")
