"""B200 check of the fused attention-block kernel (csrc/attn_fused.cu) against the unfused kernels it replaces (QKV GEMM +
window_attn_kernel) and against the CPU oracle, over the geometries the path uses; then timings of both.
    python tools/attn_fused_check.py            (MST_ATTN_TMA=0 for the cp.async-only producers)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mastermetastyletransfer_b200 import ops
from oracle import master_oracle as O

torch.manual_seed(0)
dev = "cuda"
bad = 0


def run_case(B, H, C, ws, shift, seed=0, timing=False):
    global bad
    heads = C // 32
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, H, H, C, generator=g)
    wq, wk, wv = (torch.randn(C, C, generator=g) * (0.7 / C ** 0.5) for _ in range(3))
    bq, bk, bv = (torch.randn(C, generator=g) * 0.2 for _ in range(3))
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5
    T = B * H * H
    x16 = x.to(dev).bfloat16().view(T, C).contiguous()
    d = lambda t: t.to(dev).contiguous()
    # unfused: QKV GEMM + attention kernel
    pm = ops.pack_linear(torch.cat([d(wq), d(wk), d(wv)], 0), torch.cat([d(bq), d(bk), d(bv)], 0))
    qkv = torch.empty(T, 3 * C, dtype=torch.bfloat16, device=dev)
    o_ref = torch.zeros(T, C, dtype=torch.bfloat16, device=dev)
    def unfused():
        ops.gemm(x16, pm, T, out_bf16=qkv)
        ops.window_attention(qkv, qkv[:, C:], qkv[:, 2 * C:], o_ref, d(table), B, H, H, heads, ws, shift, 3 * C, 3 * C, 3 * C, C,
                             pad_q=d(bq), pad_k=d(bk), pad_v=d(bv))
    tb = d(table)
    unfused()
    pk = ops.pack_attn_qkv(d(wq), d(wk), d(wv), d(bq), d(bk), d(bv), heads)
    o = torch.zeros(T, C, dtype=torch.bfloat16, device=dev)
    dbg = torch.zeros(T, 3 * C, dtype=torch.bfloat16, device=dev)
    ops.attn_block(x16, pk, tb, o, B, H, H, ws, shift, dbg_qkv=dbg)
    torch.cuda.synchronize()
    e_qkv = (dbg.float() - qkv.float()).abs().max().item()
    e_o = (o.float() - o_ref.float()).abs().max().item()
    # oracle (fp32 on the bf16-rounded x and weights) without the output projection
    xr = x16.float().cpu().view(B, H, H, C)
    r = lambda t: t.bfloat16().float()
    q = torch.nn.functional.linear(O._to_windows(xr, ws, shift), r(wq), bq)
    k = torch.nn.functional.linear(O._to_windows(xr, ws, shift), r(wk), bk)
    v = torch.nn.functional.linear(O._to_windows(xr, ws, shift), r(wv), bv)
    pr = O._softmax_probs(q, k, heads, O._bias_from_table(table, ws), O.shift_mask(H, H, ws, shift), B)
    oo = O._from_windows(O._apply_probs(pr, v, heads), B, H, H, ws, shift).reshape(T, C)
    e_or = (o.float().cpu() - oo).abs().max().item()
    e_or_ref = (o_ref.float().cpu() - oo).abs().max().item()
    scale = oo.abs().max().item()
    ok = e_qkv <= 0.04 and e_or <= max(2.5 * e_or_ref, 0.02 * scale)
    bad += not ok
    print(f"{'OK ' if ok else 'BAD'} B={B} H={H} C={C} ws={ws} shift={shift}: qkv max|diff| {e_qkv:.4f}  out vs unfused {e_o:.4f}  "
          f"vs oracle {e_or:.4f} (unfused kernels vs oracle {e_or_ref:.4f}, |out|max {scale:.3f})", flush=True)
    if not ok:
        dd = (o.float() - o_ref.float()).abs().view(B, H, H, heads, 32).amax(-1)
        print("   worst per (b, head):", dd.amax((1, 2)).cpu().tolist())
        print("   per-row(y) max, image 0:", [round(v, 3) for v in dd[0].amax((1, 2)).cpu().tolist()])
        print("   per-col(x) max, image 0:", [round(v, 3) for v in dd[0].amax((0, 2)).cpu().tolist()])
        dq = (dbg.float() - qkv.float()).abs().view(T, 3, heads, 32).amax(-1)
        print("   qkv diff per (q/k/v, head):", dq.amax(0).cpu().tolist())
    if timing:
        def t(fn, n=20):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(n): fn()
            b_.record(); torch.cuda.synchronize()
            return a.elapsed_time(b_) / n * 1e3
        tu = t(unfused)
        tf = t(lambda: ops.attn_block(x16, pk, tb, o, B, H, H, ws, shift))
        fl = 6.0 * T * C * C + 4.0 * T * ws * ws * C
        print(f"   time: unfused (gemm + attn) {tu:.1f} us, fused {tf:.1f} us ({fl / tf / 1e6:.0f} TFLOP/s reference-algorithm FLOPs)", flush=True)


print("TMA:", os.environ.get("MST_ATTN_TMA", "1"))
for case in [(1, 8, 256, 8, 0), (1, 16, 256, 8, 4), (2, 32, 256, 8, 4), (1, 24, 256, 8, 4), (2, 32, 256, 7, 4), (1, 16, 256, 7, 3),
             (2, 32, 128, 7, 0), (2, 32, 128, 7, 3), (1, 64, 128, 7, 3), (1, 16, 128, 8, 4), (3, 64, 256, 8, 4), (3, 64, 256, 7, 4)]:
    run_case(*case)
if bad == 0 or os.environ.get("TIME_ANYWAY"):
    for case in [(64, 64, 128, 7, 0), (64, 64, 128, 7, 3), (64, 32, 256, 7, 0), (64, 32, 256, 7, 3), (32, 32, 256, 8, 4), (16, 64, 256, 8, 4)]:
        run_case(*case, timing=True)
print("FAILED" if bad else "ALL OK", bad)
