"""Run W warm-up steps, then ONE training step between cudaProfilerStart/Stop (use with
`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...`)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gangealing_b200.training import TrainConfig, Trainer
B = int(os.environ.get("B", "8"))
torch.backends.cudnn.benchmark = True
tr = Trainer(TrainConfig(batch=B, channels_last=os.environ.get("CL", "1") == "1", dtype=os.environ.get("DT", "f32")), "cuda")
for _ in range(int(os.environ.get("W", "4"))):
    tr.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
tr.step()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
