import sys, torch
sys.path.insert(0,'/root/repo')
from tetris_reinforcement_learning_b200 import architectures as arch, trunk, _native
net = arch.AlphaSame(arch.AlphaSameConfig()).to('cuda:0').eval()
w = trunk.pack_alphasame_heads(net)
G=4096
feats = torch.randn(2*G,400,device='cuda:0').to(torch.bfloat16)
extras = torch.randn(G,105,device='cuda:0').to(torch.bfloat16)
x = torch.empty(G,528,dtype=torch.bfloat16,device='cuda:0'); v=torch.empty(G,dtype=torch.bfloat16,device='cuda:0')
L=_native.lib(); st=torch.cuda.current_stream().cuda_stream
for _ in range(3): L.trl_alphasame_heads(feats.data_ptr(), extras.data_ptr(), G, w.data_ptr(), 0, x.data_ptr(), v.data_ptr(), st)
torch.cuda.synchronize()
e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): L.trl_alphasame_heads(feats.data_ptr(), extras.data_ptr(), G, w.data_ptr(), 0, x.data_ptr(), v.data_ptr(), st)
e1.record(); torch.cuda.synchronize()
print('heads kernel', e0.elapsed_time(e1)/50*1e3, 'us')
