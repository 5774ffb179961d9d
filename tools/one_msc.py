import sys, torch
sys.path.insert(0, "spectrogram-yolov11_b200")
from specyolo import ops
from specyolo.nn.modules import MSCSpatialAttention
m = MSCSpatialAttention(64).cuda()
x = ops.new_act(64, 64, 80, 80, "cuda").normal_()
o = ops.new_act(64, 64, 80, 80, "cuda")
for _ in range(6): m(x, out=o)
torch.cuda.synchronize()
print("ok")
