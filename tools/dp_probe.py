"""2-GPU probe of the data-parallel step: ms/step of the graphed step with overlapped collectives,
without collectives, with serial collectives, and of each all-reduce alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/dp_probe.py
    ELIMREC_DP_SINGLE=1 timeout 120 python -m torch.distributed.run ... tools/dp_probe.py      # experimental one-graph step
"""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import numpy as np
from elimrec_b200 import synth
from elimrec_b200.data import Config, Dataset
from elimrec_b200.model import EliMRec
from elimrec_b200.sampler import PairwiseSamplerV2
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
inter, feats = synth.make_shape("tiktok")
ds = Dataset(None, interactions=inter, features=feats, name="tiktok_shape")
single = bool(int(os.environ.get("ELIMREC_DP_SINGLE", "0")))     # 1: the experimental one-graph step with captured collectives
conf = Config(**{"data.input.dataset": "tiktok_shape", "topks": [20], "device": dev, "alpha": 0.5, "batch_size": 2048,
                 "dp_single_graph": single})
torch.manual_seed(2022)
model = EliMRec(conf, ds).to(dev); model.make_optimizer(); model.enable_data_parallel()
B = 2048
s = PairwiseSamplerV2(ds, batch_size=B, mode="device", device=dev, seed=2022 + rank)
u, p, n = s.sample_epoch_device(60 * B)
bt = [(u[i*B:(i+1)*B], p[i*B:(i+1)*B], n[i*B:(i+1)*B]) for i in range(60)]
run = model.make_graphed_step()
import elimrec_b200.model as M
def timeit(fn, label):
    for b in bt[:5]: fn(b)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for b in bt[5:55]: fn(b)
    e1.record(); torch.cuda.synchronize()
    print(f"rank {rank} {label}: {e0.elapsed_time(e1)/50:.4f} ms/step", flush=True)
timeit(lambda b: run(*b), "full dp step" + (" (single graph)" if single else ""))
if single:      # the variants below patch the eager collectives, which the single graph no longer calls
    dist.destroy_process_group()
    sys.exit(0)
bucket = model._ws["cache"]["bucket"]
# variants: patch the bucket's collective
orig = bucket.all_reduce_mean_part
bucket.all_reduce_mean_part = lambda part, async_op=False: None
timeit(lambda b: run(*b), "graphs only, no collectives")
bucket.all_reduce_mean_part = lambda part, async_op=False: orig(part, async_op=False)
timeit(lambda b: run(*b), "synchronous collectives (no overlap)")
bucket.all_reduce_mean_part = orig
def only_comm(b):
    w = orig(0, async_op=True); w.wait()
timeit(only_comm, "all-reduce of the table part alone")
def only_comm2(b):
    w = orig(1, async_op=True); w.wait()
timeit(only_comm2, "all-reduce of the small part alone")
dist.destroy_process_group()
