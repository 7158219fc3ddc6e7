"""Device time of the tcgen05 chain kernels alone (forward, data gradient, weight gradient) on the
cache shader's stacks at P points; prints TFLOP/s of each launch (CUDA events, warm L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tests.test_chain_gpu as T
from neural_radiance_caching_b200 import mlp_chain as mc, _lib
from tests.util import gen, f32

P = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
dev = torch.device("cuda:0")
peak = 1603.9

def timeit(fn, iters=20):
    """GPU time of fn's launches: captured once in a CUDA graph, replayed back to back."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(2):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3  # us

only = sys.argv[2].split(',') if len(sys.argv) > 2 else None
for name in T.SPECS:
    if only and name not in only:
        continue
    g = gen(1)
    spec = mc.ChainSpec(**T.SPECS[name])
    p = {k: {a: b.to(dev) for a, b in v.items()} for k, v in T.make_params(g, spec).items()}
    srcs = [f32(g.normal(size=(P, w))).to(dev) for w in spec.in_widths]
    macs = 0
    for li, (nm, w, _) in enumerate(spec.hidden):
        macs += sum(x[1] for x in spec.x_parts[li]) * w
    macs += sum(x[1] for x in spec.x_last) * sum(spec.head_widths)
    packed = mc.pack_weights(spec, p)
    bufs, outs, act = mc.run_forward(spec, p, srcs, packed, save=True)
    t_f = timeit(lambda: mc.run_forward(spec, p, srcs, packed, save=True))
    t_f0 = timeit(lambda: mc.run_forward(spec, p, srcs, packed, save=False))
    gh = [torch.randn_like(b) for b in bufs]
    d_src = [(torch.empty((P, w), device=dev), False) for w in spec.in_widths]
    dy = mc.run_backward_data(spec, p, gh, act, packed, P, d_src)
    t_b = timeit(lambda: mc.run_backward_data(spec, p, gh, act, packed, P, d_src))
    sinks = {k: (torch.zeros_like(v["kernel"]), torch.zeros_like(v["bias"])) for k, v in p.items()}
    def wg():
        wp = mc._Ptrs()
        mc.wgrad_launch(mc.wgrad_layers(spec, act, dy, sinks, wp), wp, P)
    t_w = timeit(wg)
    t_p = timeit(lambda: mc.pack_weights(spec, p))
    fl = 2.0 * macs * P
    print(f"{name:9s} P={P} MAC/pt={macs:7d} | fwd(save) {t_f:7.1f} us {fl/t_f/1e6:7.1f} TF/s ({100*fl/t_f/1e6/peak:4.1f}%) | "
          f"fwd(nosave) {t_f0:7.1f} us {fl/t_f0/1e6:7.1f} TF/s | bwd-data {t_b:7.1f} us {fl/t_b/1e6:7.1f} TF/s | "
          f"wgrad {t_w:7.1f} us {fl/t_w/1e6:7.1f} TF/s | pack {t_p:5.1f} us", flush=True)

# the three stacks behind the shader's mid stage: separate launches vs. one multi-program launch
if only is None or "multi" in only:
    names = ["int_brdf", "env", "slf"]
    st = {}
    for name in names:
        g = gen(1)
        spec = mc.ChainSpec(**T.SPECS[name])
        p = {k: {a: b.to(dev) for a, b in v.items()} for k, v in T.make_params(g, spec).items()}
        srcs = [f32(g.normal(size=(P, w))).to(dev) for w in spec.in_widths]
        packed = mc.pack_weights(spec, p)
        bufs, outs, act = mc.run_forward(spec, p, srcs, packed, save=True)
        gh = [torch.randn_like(b) for b in bufs]
        d_src = [(torch.empty((P, w), device=dev), False) for w in spec.in_widths]
        st[name] = (spec, p, srcs, packed, act, gh, d_src)
    def fwd(batched):
        b = mc.Batch() if batched else None
        for name in names:
            spec, p, srcs, packed, act, gh, d_src = st[name]
            mc.run_forward(spec, p, srcs, packed, save=(name != "env"), batch=b)
        if b: b.flush()
    def bwd(batched):
        b = mc.Batch() if batched else None
        for name in ("int_brdf", "slf"):
            spec, p, srcs, packed, act, gh, d_src = st[name]
            mc.run_backward_data(spec, p, gh, act, packed, P, d_src, batch=b)
        if b: b.flush()
    print(f"shader stacks P={P}: fwd separate {timeit(lambda: fwd(False)):.1f} us, one launch {timeit(lambda: fwd(True)):.1f} us | "
          f"bwd separate {timeit(lambda: bwd(False)):.1f} us, one launch {timeit(lambda: bwd(True)):.1f} us", flush=True)
