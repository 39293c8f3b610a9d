"""Dev tool for ncu: one full-size eager UNet3D call (config #2: [2,4,16,64,64], text ctx [2,77,768]) bracketed by
cudaProfilerStart/Stop.  Usage (under gpurun):
  ncu --profile-from-start off --clock-control none --csv --log-file gpurun_out/unet_traffic.csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum python scripts/one_unet_call.py
then  python scripts/summarize_profiles.py traffic gpurun_out/unet_traffic.csv profiles/rNN_unet_call_dram_traffic.json"""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from util_models import FULL_CFG, rerandomise_zero_inits  # noqa: E402
from emote_hack_b200.unet3d import UNet3DConditionModel  # noqa: E402

dev = torch.device("cuda")
torch.manual_seed(0)
with torch.device(dev):
    unet = UNet3DConditionModel(**FULL_CFG).eval()
rerandomise_zero_inits(unet)
x = torch.randn(2, 4, 16, 64, 64, device=dev); ctx = torch.randn(2, 77, 768, device=dev)
for _ in range(2):
    unet(x, 981, ctx)
torch.cuda.synchronize()
torch.cuda.profiler.start()
unet(x, 981, ctx)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
