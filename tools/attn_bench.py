"""Pieces of the SelfAttention block at the network's size (scale s: C = 128 on the 80x80 map, batch 64), CUDA events."""
import ctypes as C, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from transparent_object_detection_b200 import lib
from transparent_object_detection_b200._lib import AttentionDesc, check
from transparent_object_detection_b200.attention import _gemm
from transparent_object_detection_b200.engine import pack_conv_weight

B, H, W, Cc, d16 = int(os.environ.get("B", 64)), 80, 80, 128, 16
N = H * W
g = torch.Generator().manual_seed(0)
x = torch.randn((B, H, W, Cc), generator=g).to(torch.bfloat16).cuda()
wq = pack_conv_weight(torch.randn((d16, Cc, 1, 1), generator=g) * 0.08).cuda()
bq = torch.zeros(d16).cuda()
wv = (torch.randn((Cc, Cc), generator=g) * 0.1).to(torch.bfloat16).cuda()
bv = torch.zeros(Cc).cuda()
q = torch.empty((B, N, d16), dtype=torch.bfloat16, device="cuda"); k = torch.empty_like(q)
vT = torch.empty((B, Cc, N), dtype=torch.bfloat16, device="cuda")
out = torch.empty_like(x)
L = lib(); st = torch.cuda.current_stream().cuda_stream

def qk():
    _gemm(L, st, x.data_ptr(), B * H, W, Cc, Cc, wq.data_ptr(), d16, q.data_ptr(), d16, bias_ptr=bq.data_ptr())
    _gemm(L, st, x.data_ptr(), B * H, W, Cc, Cc, wq.data_ptr(), d16, k.data_ptr(), d16, bias_ptr=bq.data_ptr())
def vt():
    for i in range(B):
        _gemm(L, st, wv.data_ptr(), 1, Cc, Cc, Cc, x.data_ptr() + i * N * Cc * 2, N, vT.data_ptr() + i * Cc * N * 2, N)
def fused():
    a = AttentionDesc()
    a.d_q, a.d_k, a.d_vt, a.d_bias, a.d_x, a.d_out = q.data_ptr(), k.data_ptr(), vT.data_ptr(), bv.data_ptr(), x.data_ptr(), out.data_ptr()
    a.batch, a.n, a.c, a.d16, a.x_pitch, a.out_pitch = B, N, Cc, d16, Cc, Cc
    check(L.tod_attention_fused(C.byref(a), st), "fused")
for name, fn in (("q + k projections", qk), ("V^T, one GEMM per image (the path before the batched projection + transpose)", vt), ("fused attention kernel", fused)):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); fn(); fn(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    extra = ""
    if name.startswith("fused"):
        fl = B * (2.0 * 2 * N * N * d16 + 2.0 * N * N * Cc)
        extra = f"  {fl / ms / 1e9:.0f} TFLOP/s, {B * N * N / ms / 1e6:.2f} G exp/s"
    print(f"{name:28s} {ms * 1e3:8.1f} us{extra}")
