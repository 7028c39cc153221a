import sys, os, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from montecosmo_b200 import nbody as nb
o = nb.ops()
dev = o.A.device
for n in (64, 128, 256):
    x = torch.randn(n, n, n, device=dev)
    k = o.rfftn(x); torch.cuda.synchronize(); print(n, "main thread ok", flush=True)
    res = {}
    def work(setdev):
        try:
            if setdev: torch.cuda.set_device(0)
            k2 = o.rfftn(x); torch.cuda.synchronize()
            res[setdev] = "ok err=%g" % float((k2 - k).abs().max())
        except Exception as e:
            res[setdev] = "FAIL " + str(e)
    for sd in (False, True):
        t = threading.Thread(target=work, args=(sd,)); t.start(); t.join()
        print(n, "thread set_device=%s:" % sd, res[sd], flush=True)
    # autograd thread
    xx = x.clone().requires_grad_()
    try:
        y = nb.irfftn(nb.rfftn(xx)); y.sum().backward(); torch.cuda.synchronize(); print(n, "autograd ok", flush=True)
    except Exception as e:
        print(n, "autograd FAIL", e, flush=True)
