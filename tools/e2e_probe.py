import sys, time, numpy as np, torch
sys.path.insert(0, "/root/repo")
from primalcr_b200 import api
from primalcr_b200.data import synth_dataset
ds = synth_dataset("netflix", scale=1.0, device="cuda", test_per_user=0); torch.cuda.empty_cache()
k=100
U = api.reference_init(ds.d1, k); V = U[:ds.d2].copy()
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
rp, it, ra, hU, hV = pin(ds.train.row_ptr), pin(ds.train.item), pin(ds.train.rating), pin(U), pin(V)
for rep in range(2):
    torch.cuda.synchronize(); t0=time.perf_counter()
    e = api.Engine(api.Parameter(solver_type=2, k=k, lambda_=5000.0, maxiter=3, do_predict=0)); t1=time.perf_counter()
    e.set_levels(np.arange(1,6)); e.set_train_raw(ds.d1, ds.d2, ds.train.nnz, rp, it, ra); t2=time.perf_counter()
    e.set_factors(hU, hV); t3=time.perf_counter()
    e.run(log=None); t4=time.perf_counter()
    oU=torch.empty_like(hU); oV=torch.empty_like(hV); e.get_factors(oU, oV); t5=time.perf_counter()
    e.close(); t6=time.perf_counter()
    print("create %.3f set_train %.3f set_factors %.3f run(3 it + obj0) %.3f get %.3f close %.3f total %.3f" % (t1-t0,t2-t1,t3-t2,t4-t3,t5-t4,t6-t5,t6-t0), flush=True)
