import sys, json
sys.path.insert(0, "/root/repo")
import torch, bench
r = bench.loss_reduction_rates(torch.device("cuda"), 32, 256, 6545.9)
for k, v in r.items(): print(k, v)
r = bench.loss_reduction_rates(torch.device("cuda"), 8, 256, 6545.9)
for k, v in r.items():
    if "all" in k: print("B8", k, v)
