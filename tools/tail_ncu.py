"""One launch of the fused DPT tail at the product shape, for ncu (GPU box)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from omnivggt_official_b200 import ops
dev = torch.device("cuda", 0)
Fr, h, w, H, W, Cin, outc = 8, 296, 296, 518, 518, 128, 4
dt = torch.float16
xp = torch.zeros(Fr, h + 2, w + 2, Cin, device=dev, dtype=dt)
xp[:, 1:-1, 1:-1] = torch.randn(Fr, h, w, Cin, device=dev).to(dt)
wb = (torch.randn(32, 9 * Cin, device=dev) * (9 * Cin) ** -0.5).to(dt)
b1, w2, b2 = torch.randn(32, device=dev) * 0.1, torch.randn(outc, 32, device=dev) * 32 ** -0.5, torch.randn(outc, device=dev) * 0.1
tx, ty = torch.randn(W, 64, device=dev) * 0.1, torch.randn(H, 64, device=dev) * 0.1
for _ in range(2):
    ops.dpt_tail(xp, tx, ty, wb, b1, w2, b2, 1, Fr, h, w, H, W)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.dpt_tail(xp, tx, ty, wb, b1, w2, b2, 1, Fr, h, w, H, W)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
# timeline of CTA 0 (clock64 stamps; tools/call_tail_prof.sh)
if os.environ.get("TAIL_PROF"):
    import ctypes
    from omnivggt_official_b200 import _lib
    lib = _lib.lib()
    lib.ovg_debug_set_tail_profile.argtypes = [ctypes.c_void_p]
    buf = torch.zeros(192 * 8, device=dev, dtype=torch.int64)
    lib.ovg_debug_set_tail_profile(buf.data_ptr())
    ops.dpt_tail(xp, tx, ty, wb, b1, w2, b2, 1, Fr, h, w, H, W)
    torch.cuda.synchronize()
    lib.ovg_debug_set_tail_profile(None)
    t = buf.view(192, 8).cpu()
    t0 = int(t[0, 0])
    print("row  p_start p_sync1 p_aempty p_done p_sync2 | m_afull m_issued | e_ofull   (clocks since row 0 start)")
    for i in list(range(0, 8)) + list(range(46, 60)) + list(range(96, 110)) + list(range(140, 160)):
        print(i, " ".join(f"{int(x) - t0:8d}" for x in t[i]))
