"""In-kernel timeline of the attention kernel (the experiment behind DESIGN.md §4 "single MMA-issuing thread").

Build a variant of the library with the trace points compiled in, then run this script against it:

    cd vision-language-model-framework_b200/csrc && VLMFW_EXTRA_FLAGS="-DVLMFW_ATT_TRACE" \
        VLMFW_OUT=$PWD/../../tools/_variants/lib_trace.so VLMFW_BUILD_DIR=/tmp/build_trace bash build.sh
    VLMFW_B200_LIB=$PWD/tools/_variants/lib_trace.so python tools/att_trace.py [siglip|clip]

CTA 0 logs (clock64 << 8 | event) per role (loader, S issuer, softmax warp 0 lane 0, P.V issuer); the script prints, per
role, the mean / median cycles between consecutive events over the steady-state half of the kernel and their share of
that role's timeline. Event numbers are the TR(role, n) points in attention_tc.cu. VLMFW_ATTN_DBG masks combine with it.
"""
import ctypes as C, os, sys, collections
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlmfw_b200 import ops, _lib
lib = _lib.load()
# VLMFW_ATTN=tc traces the round-1 kernel (attention_tc.cu), otherwise the wide-block kernel (attention_wb.cu)
fetch = lib.vlmfw_debug_att_trace if os.environ.get("VLMFW_ATTN", "wb").startswith("t") else lib.vlmfw_debug_wb_trace
fetch.argtypes = [C.c_void_p]
which = sys.argv[1] if len(sys.argv) > 1 else "siglip"
bsz, s, hd, dh = {"siglip": (256, 256, 16, 72), "clip": (128, 577, 16, 64), "eva": (128, 257, 16, 88)}[which]
qkv = (torch.randn(bsz * s, 3 * hd * dh, device="cuda") * 0.5).to(torch.float16)
for _ in range(3):
    ops.attention(qkv, bsz, s, hd, dh)
torch.cuda.synchronize()
buf = np.zeros((4, 8192), np.int64)
assert fetch(buf.ctypes.data) == 0
names = {0: "loader", 1: "s-issuer", 2: "softmax", 3: "pv-issuer"}
for role in range(4):
    ev = buf[role]
    ev = ev[ev != 0]
    code, clk = ev & 0xFF, ev >> 8
    n = len(ev)
    lo, hi = n // 4, 3 * n // 4  # steady state
    d = collections.defaultdict(list)
    for i in range(lo + 1, hi):
        d[(int(code[i - 1]), int(code[i]))].append(int(clk[i] - clk[i - 1]))
    total = int(clk[hi - 1] - clk[lo])
    print(f"== {names[role]}: {n} events, steady window {total} clk")
    for k in sorted(d):
        v = np.array(d[k])
        print(f"   {k[0]}->{k[1]}: n={len(v):4d} mean={v.mean():8.1f} p50={np.median(v):8.1f} max={v.max():7d} share={v.sum() / total * 100:5.1f}%")
