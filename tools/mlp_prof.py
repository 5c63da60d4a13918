"""Role-level cycle breakdown of the fused projection + MLP kernel (MST_MLP_PROF=1 turns on the in-kernel clock64 counters)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops
for (M, C, ln) in [(262144, 128, True), (65536, 256, True), (32768, 256, False)]:
    A = torch.randn(M, C, device="cuda").bfloat16()
    r = lambda *s: torch.randn(*s, device="cuda")
    pm = ops.pack_mlp(r(4 * C, C) / C ** 0.5, r(4 * C), r(C, 4 * C) / (4 * C) ** 0.5, r(C), wpre=r(C, C) / C ** 0.5, bpre=r(C))
    x = r(M, C)
    g, b = (r(C), r(C)) if ln else (None, None)
    print(f"--- M={M} C={C} ln={ln}", flush=True)
    for _ in range(2):
        ops.mlp_fused(A, pm, M, res=x, out_f32=x, pre=True, ln_g=g, ln_b=b)
        torch.cuda.synchronize()
