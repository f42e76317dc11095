"""How many host threads make the CPU oracle fastest on this box? (informs bench.py's cpu_baseline thread count)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import ref_model as M, ref_train as T
spatial = [128, 96]
tab, _, _ = M.generator_layout(spatial + [1])
var = M.init_variables(tab, 123)
x, y = T.synthetic_batch(4, spatial, seed=1)
for nt in [8, 16, 32, 64, 128]:
    if nt > (os.cpu_count() or 1):
        break
    torch.set_num_threads(nt)
    T.generator_loss_and_grads(y, x, var)
    t0 = time.time()
    T.generator_loss_and_grads(y, x, var)
    print("threads", nt, "sec/step(batch4)", round(time.time() - t0, 3), flush=True)
