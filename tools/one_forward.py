"""One eager forward of the bench workload between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ['DI_B200_GRAPH'] = '0'
import torch
import bench

torch.set_grad_enabled(False)
device = torch.device('cuda:0')
neck, head = bench.build_models(device)
fr = bench.h2d(bench.host_frame(1, 'lidar', bench.SEED), device)
for _ in range(2):
    bench.forward(neck, head, fr)
torch.cuda.synchronize()
torch.cuda.profiler.start()
bench.forward(neck, head, fr)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
