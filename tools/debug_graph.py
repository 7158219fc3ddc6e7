import sys, os, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neural_radiance_caching_b200 import workload, _lib
dev = torch.device("cuda:0")
st = workload.CacheSamplerStep(dev)
g = np.random.Generator(np.random.PCG64(1))
R = 1024
rn = workload.make_rays_np(g, R)
u = [g.uniform(size=(R, 1)).astype(np.float32) for _ in range(3)]
tgt = g.uniform(size=(R, 1)).astype(np.float32)
dbuf = torch.from_numpy(workload.pack_rays(rn, u, tgt)).to(dev)

def unpack():
    return workload.unpack_rays(dbuf)

def infer():
    rays, u01, extra = unpack()
    return st.forward(rays, u01, train=False)[2]["weights"]

def train_fwd():
    rays, u01, extra = unpack()
    return st.forward(rays, u01, train=True)[2]["weights"]

def full():
    rays, u01, extra = unpack()
    return st.step(rays, u01, extra[:, 0])

def enc_only():
    rays, u01, extra = unpack()
    m = st.sampler.mlps[0]
    return m.grid(st.params["MLP_0"]["density_grid"], rays["origins"] * 0.2)

def enc_bwd():
    st.zero_grad()
    rays, u01, extra = unpack()
    m = st.sampler.mlps[0]
    out = m.grid(st.params["MLP_0"]["density_grid"], rays["origins"] * 0.2)
    out.sum().backward()
    return out

for name, fn in [("enc_only", enc_only), ("enc_bwd", enc_bwd), ("infer", infer), ("train_fwd", train_fwd), ("full", full)]:
    try:
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(2):
                fn()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            out = fn()
        gr.replay(); torch.cuda.synchronize()
        print(name, "OK", float(out.float().sum()))
    except Exception as e:
        print(name, "FAILED:", repr(e)[:300])
        traceback.print_exc(limit=6)
        break
