"""Where does the host time of one training step go?  (run on the GPU box)"""
import cProfile, pstats, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, scipy.sparse as sp
import bench
from sslrec_b200.config import default_config, load_config
from sslrec_b200.data_handler import DataHandlerGeneralCF
from sslrec_b200.optim import FusedAdam
from sslrec_b200.general_cf.simgcl import SimGCL

model_name, graph, hp = bench.WORKLOADS['simgcl-amazon']
rows, cols, n_user, n_item = bench.graph_arrays(graph)
cfg = default_config(model_name, **hp); cfg['train']['batch_size'] = 4096
load_config(base=cfg, device='cuda')
dh = DataHandlerGeneralCF(sp.coo_matrix((np.ones(len(rows), dtype=np.float32), (rows, cols)), shape=(n_user, n_item))); dh.load_data()
model = SimGCL(dh).cuda(); opt = FusedAdam(model.parameters(), lr=1e-3)
batches = [torch.from_numpy(b).cuda() for b in bench.make_batches(rows, cols, n_item, 40)]
def step(i):
    opt.zero_grad(); b = batches[i % 40]
    loss, parts = model.cal_loss([b[0], b[1], b[2]]); loss.backward(); opt.step(); return loss
for i in range(5): step(i)
torch.cuda.synchronize()
t0 = time.perf_counter()
for i in range(20): step(i)
t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f'enqueue {1e3*(t1-t0)/20:.2f} ms/step, drained after {1e3*(t2-t0)/20:.2f} ms/step')
# phases
def phase(name, fn, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n): fn(i)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f'{name}: host {1e3*(t1-t0)/n:.2f} ms, total {1e3*(t2-t0)/n:.2f} ms')
keep = {}
def fwd(i):
    opt.zero_grad(); b = batches[i % 40]; keep['l'] = model.cal_loss([b[0], b[1], b[2]])[0]
phase('cal_loss', fwd)
def fb(i):
    fwd(i); keep['l'].backward()
phase('cal_loss+backward', fb)
pr = cProfile.Profile(); pr.enable()
for i in range(20): step(i)
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(35); print(s.getvalue()[:6000])
import subprocess
print(subprocess.run(['nvidia-smi', '--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap', '--format=csv,noheader,nounits'], capture_output=True, text=True))
print(subprocess.run(['nvidia-smi', '--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.hw_slowdown,clocks_throttle_reasons.sw_power_cap', '--format=csv,noheader,nounits'], capture_output=True, text=True))
