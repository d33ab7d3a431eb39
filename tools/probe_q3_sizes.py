import json, sys
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import dune_functions_b200 as dfx
PEAK = json.load(open('/root/repo/MEASURED_PEAKS.json'))["hbm_gbs"]
def timeit(fn, reps=7, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))
for n in (128, 256):
    b = dfx.makeBasis(dfx.YaspGrid([n,n,n]), dfx.lagrange(3))
    x = b.newVector(0.0)
    ms = timeit(lambda: dfx.interpolate(b, x, dfx.gaussian(1.0)))
    print(json.dumps({"probe":"Q3 interpolate, f at every node","n":n,"ms":ms,"frac":b.dimension()*8/ms/1e6/PEAK}),flush=True)
    pts,_ = dfx.tensor_gauss(3,4)
    v = torch.empty((b.numElements(),64,1),dtype=torch.float64,device='cuda')
    f = dfx.makeDiscreteGlobalBasisFunction(b,x)
    ms = timeit(lambda: f.evaluate(pts,out_values=v))
    print(json.dumps({"probe":"Q3 evaluate at 4^3","n":n,"ms":ms,"frac":(b.dimension()*8+v.numel()*8)/ms/1e6/PEAK}),flush=True)
    del b,x,v,f; torch.cuda.empty_cache()
