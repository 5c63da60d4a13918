"""One eager forward of the bench workload for ncu (after two warm-up forwards).  Run under gpurun:
   ncu --set full -k regex:'gemm_tc|mlp_fused|conv_band|window_attn' -s 98 -c 49 ... python tools/prof_forward.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import MasterStyleTransferModel, synthetic
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 256
m = MasterStyleTransferModel(); synthetic.fill_state_dict_(m, 0); m = m.eval().cuda()
c, s = synthetic.synthetic_images(B, S, seed=0)
c, s = c.cuda(), s.cuda()
with torch.no_grad():
    for _ in range(3):
        out = m(c, s, 1)
torch.cuda.synchronize()
print("ok", float(out.float().mean()))
