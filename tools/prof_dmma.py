import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import dune_functions_b200 as dfx
b = dfx.makeBasis(dfx.YaspGrid([96, 96, 96]), dfx.lagrangeDG(4))
x = b.newVector(0.0).uniform_(-1, 1)
rng = np.random.default_rng(1)
pts = rng.random((125, 3))
f = dfx.makeDiscreteGlobalBasisFunction(b, x)
for _ in range(2):
    v, j = f.evaluate(pts, jacobians=True)
torch.cuda.synchronize()
print("ok")
