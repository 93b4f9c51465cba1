"""Developer: how precise is the fp32 accumulation of tcgen05.mma kind::f16?  Inputs that are exactly representable in bf16
(lo planes = 0) make every product exact, so the error of the conv probe against fp64 is pure accumulation error."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tinyvc_b200 import _lib

dev = torch.device("cuda")
for K in (128, 968, 3072):
    g = torch.Generator().manual_seed(K)
    B, T, Cout = 2, 256, 128
    x = torch.randn(B, K, T, generator=g).bfloat16().float()
    w = (torch.randn(Cout, K, 1, generator=g) / K ** 0.5).bfloat16().float()
    b = torch.zeros(Cout)
    ref = torch.einsum("bkt,ok->bot", x.double(), w[:, :, 0].double())
    ref32 = torch.einsum("bkt,ok->bot", x, w[:, :, 0])
    y = torch.empty(B, Cout, T, device=dev)
    ptr = lambda t: ctypes.c_void_p(t.data_ptr()) if t is not None else None
    xd, wc, bc = x.to(dev).contiguous(), w.contiguous(), b.contiguous()
    rc = _lib.lib().tvc_tc_conv_probe(ptr(xd), ptr(wc), ptr(bc), B, T, K, Cout, 1, 1, None, None, None, 0, 0, None, 0, 0, 128, ptr(y), None, None)
    _lib.check(rc, "probe")
    torch.cuda.synchronize()
    got = y.cpu().double()
    s = ref.pow(2).mean().sqrt()
    print(f"K={K}: tensor core rel_rms {float((got-ref).pow(2).mean().sqrt()/s):.3e} rel_max {float((got-ref).abs().max()/s):.3e} | "
          f"torch fp32 CPU rel_rms {float((ref32.double()-ref).pow(2).mean().sqrt()/s):.3e} rel_max {float((ref32.double()-ref).abs().max()/s):.3e}")
