import cProfile, pstats, sys, time, io, os
sys.path.insert(0, '.')
import numpy as np, torch, torch.distributed as dist
lr = int(os.environ.get('LOCAL_RANK', '0')); torch.cuda.set_device(lr)
dist.init_process_group(backend='nccl', device_id=torch.device('cuda', lr))
from optiml_b200.configs import make_config
from optiml_b200.ml.svm import DualSVC
from optiml_b200.ml.svm.kernels import GaussianKernel
spec, X, y = make_config('C4')
for i in range(3):
    m = DualSVC(kernel=GaussianKernel(), C=1).fit(X, y); m.obj.release()
dist.barrier()
m = DualSVC(kernel=GaussianKernel(), C=1)
pr = cProfile.Profile()
t = time.perf_counter(); pr.enable(); m.fit(X, y); pr.disable(); dt = time.perf_counter() - t
if dist.get_rank() == 0:
    print('fit wall', dt, m.fit_times_, 'pg device ms', m.optimizer.device_ms)
    s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:5000])
dist.destroy_process_group()
