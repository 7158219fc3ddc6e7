"""Per-op timeline of the chain kernel's CTA 0 (debug build):
    NRC_EXTRA_NVCC_FLAGS=-DNRC_CHAIN_TRACE python -m neural_radiance_caching_b200.build --force
    python tools/trace_chain.py [spec] [P] [fwd|bwd]
Prints, per tile iteration and context, the cycles each op took (LOAD / GEMM+wait / EPI / SAVE)."""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import tests.test_chain_gpu as T
from neural_radiance_caching_b200 import mlp_chain as mc, _lib
from tests.util import gen, f32

name = sys.argv[1] if len(sys.argv) > 1 else "slf"
P = int(sys.argv[2]) if len(sys.argv) > 2 else 524288
mode = sys.argv[3] if len(sys.argv) > 3 else "fwd"
use_img = len(sys.argv) > 4 and sys.argv[4] == "img"   # sources / destinations as bf16 tile images
dev = torch.device("cuda:0")
g = gen(1)
spec = mc.ChainSpec(**T.SPECS[name])
p = {k: {a: b.to(dev) for a, b in v.items()} for k, v in T.make_params(g, spec).items()}
srcs = [f32(g.normal(size=(P, w))).to(dev) for w in spec.in_widths]
packed = mc.pack_weights(spec, p)
progs = []
orig = _lib.call
def spy(fn, *a):
    if fn == "nrc_chain_run":
        progs.append(a[1]._obj)
    return orig(fn, *a)
mc._lib.call = spy
_lib.call = spy
if use_img:
    n_in = len(spec.in_atoms)
    img = mc.new_image(P, n_in, dev)
    img.copy_(torch.randn(img.shape, device=dev).to(torch.bfloat16))
    fsrcs, a0 = [], 0
    col = 0
    for i, w in enumerate(spec.in_widths):
        wpad = (spec.in_pad - col) if i == len(spec.in_widths) - 1 else w
        na = len(mc._atoms_of(wpad))
        fsrcs.append(mc.ImgRef(img, a0, na, n_in))
        a0 += na
        col += w
else:
    fsrcs = srcs
bufs, outs, act = mc.run_forward(spec, p, fsrcs, packed, save=(mode != "fwd"), P=P)
if mode != "fwd":
    gh = [torch.randn_like(b) for b in bufs]
    d_src = [(torch.empty((P, w), device=dev), False) for w in spec.in_widths]
    if use_img:
        dimg = mc.new_image(P, n_in, dev)
        d_src[0] = mc.ImgRef(dimg, 0, fsrcs[0].natoms, n_in)
    progs.clear()
    mc.run_backward_data(spec, p, gh, act, packed, P, d_src)
torch.cuda.synchronize()
prog = progs[-1]
lib = _lib.load()
lib.nrc_chain_trace_dump.argtypes = [C.POINTER(C.c_longlong)]
buf = np.zeros((2, 48, 32), dtype=np.int64)
rc = lib.nrc_chain_trace_dump(buf.ctypes.data_as(C.POINTER(C.c_longlong)))
assert rc == 0
kinds = {0: "LOAD", 1: "GEMM", 2: "EPI", 3: "SAVE", 4: "GATH", 5: "LIMG"}
ops = [(kinds[prog.ops[i].kind], prog.ops[i].n if prog.ops[i].kind == 1 else prog.ops[i].npad) for i in range(prog.num_ops)]
print("ops:", ops)
t00 = buf[0, 0, 31]
for it in range(0, 8):
    for c in range(2):
        t0 = buf[c, it, 31]
        if t0 == 0:
            continue
        prev, parts = t0, []
        for i, (k, n) in enumerate(ops):
            t = buf[c, it, i]
            if t == 0:
                continue   # all but the last GEMM of a run leave no stamp
            parts.append(f"{k}{n}:{t - prev}")
            prev = t
        print(f"it{it} ctx{c} start@{t0 - t00:8d} total {prev - t0:7d} | " + " ".join(parts))

marks = np.zeros(16, dtype=np.int64)
lib.nrc_chain_marks_dump.argtypes = [C.POINTER(C.c_longlong)]
lib.nrc_chain_marks_dump(marks.ctypes.data_as(C.POINTER(C.c_longlong)))
print("v2 phases (cycles from kernel entry): decode %d, prologue done %d, weights resident %d, tiles done %d, before exit sync %d; "
      "entry->exit %.2f us" % tuple([int(marks[i] - marks[8]) for i in (9, 10, 11, 12, 13)] + [(marks[15] - marks[14]) * 1e-3]))
print("last EPI of ctx0 thread 0 (begin, args, ld issued, ld waited, chunk done, end):", [int(m - marks[0]) for m in marks[:6]])

try:
    mma = np.zeros((48, 8, 2), dtype=np.int64)
    lib.nrc_chain_mma_dump.argtypes = [C.POINTER(C.c_longlong)]
    lib.nrc_chain_mma_dump(mma.ctypes.data_as(C.POINTER(C.c_longlong)))
    for it in range(0, 4):
        if mma[it, 0, 0] == 0:
            continue
        print(f"it{it} UMMA issuer (v2, ctx0): " + " ".join(
            f"g{g}: ready@{mma[it, g, 0] - buf[0, it, 31]} issue {mma[it, g, 1] - mma[it, g, 0]}" for g in range(8) if mma[it, g, 0]))
except AttributeError:
    pass
