"""Per-phase cycle counters of the attention softmax warps (needs a -DOVG_ATT_PROFILE build: OVG_LIB_PATH=.../libovg_prof.so)."""
import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from omnivggt_official_b200 import ops, _lib
lib = _lib.lib()
lib.ovg_debug_set_attn_profile.argtypes = [ctypes.c_void_p]
lib.ovg_debug_set_attn_profile.restype = None
BF16 = torch.bfloat16
for (b, h, n) in ((1, 16, 10992), (8, 16, 1374)):
    q = torch.randn(b, h, n, 64, device="cuda").to(BF16) * 0.2
    k = torch.randn(b, h, n, 64, device="cuda").to(BF16)
    v = torch.randn(b, h, n, 64, device="cuda").to(BF16)
    o = torch.empty(b, n, h * 64, device="cuda", dtype=BF16)
    prof = torch.zeros(16, dtype=torch.int64, device="cuda")
    for _ in range(2):
        ops.attention(q, k, v, o, b, h, n)
    lib.ovg_debug_set_attn_profile(prof.data_ptr())
    ops.attention(q, k, v, o, b, h, n)
    torch.cuda.synchronize()
    lib.ovg_debug_set_attn_profile(None)
    p = prof.cpu().tolist()
    names = ["wait S", "tmem read", "wait o_ready", "max+exp+st", "st wait+arrive", "whole step"]
    for t in range(2):
        nk = max(p[t * 8 + 7], 1)
        print(f"n={n} tile {t}: " + "  ".join(f"{nm}={p[t*8+i]/nk:.0f}" for i, nm in enumerate(names)) + f"  (cycles per KV step, {nk} steps)")
