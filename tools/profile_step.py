"""One eager training step of the bench workload between cudaProfilerStart/Stop, for ncu:

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_step.py [--batch 32]
  python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from npp_b200 import engine  # noqa: E402
from npp_b200 import functional as F_  # noqa: E402
from npp_b200.core.criterion import Criterion_par, Criterion_pose  # noqa: E402
from npp_b200.models.model_augment import Network  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--size", type=int, default=384)
ap.add_argument("--layers", type=int, default=16)
ap.add_argument("--channels", type=int, default=64)
args = ap.parse_args()

F_.set_compute_dtype(torch.bfloat16)
torch.manual_seed(0)
model = Network(engine.make_cfg(layers=args.layers, init_channels=args.channels)).cuda().train()
cpose, cpar = Criterion_pose(out_len=2).cuda(), Criterion_par(out_len=2).cuda()
opt = engine.build_optimizer(model, cpose, cpar)
step = engine.TrainStep(model, cpose, cpar, opt, args.batch, args.size, use_graph=False)
step.load(*engine.synthetic_batch(args.batch, args.size, seed=1))
step.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
step.run()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("loss", float(step.loss))
