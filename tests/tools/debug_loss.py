import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from oracle import loss_utils as oloss
from neural_radiance_caching_b200 import loss_utils as nloss
from tests.test_loss_gpu import _stepfun
from tests.util import gen
dev = torch.device("cuda:0")
for blur in (0.03, 0.003, 0.0):
    for m, nq in ((32, 64), (8, 5), (64, 128)):
        g = gen(700 + m + nq)
        R = 300
        c, w = _stepfun(g, R, m, 0.8)
        cp, wp = _stepfun(g, R, nq, 0.7)
        want = oloss.blur_and_resample_weights(cp, c, w, blur)
        got = nloss.blur_and_resample_weights(cp.to(dev), c.to(dev), w.to(dev), blur).cpu()
        err = (got - want).abs()
        i = int(err.argmax()); r, j = divmod(i, nq)
        print(blur, m, nq, "max err", float(err.max()), "rel", float(err.max() / want.abs().max()), "at ray", r, "bin", j, "want", float(want[r, j]), "got", float(got[r, j]), "rows bad", int((err.max(1).values > 1e-6).sum()))
