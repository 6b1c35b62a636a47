import sys, copy, torch
sys.path.insert(0,'/root/repo')
import bench
import denseclip_vit_multimodal_b200 as D
torch.manual_seed(0)
m=D.DenseCLIP(**copy.deepcopy(bench.model_kwargs()))
bench.init_uninitialised(m)
m=m.eval().cuda()
img=torch.randn(16,3,512,1024,device='cuda')
with torch.no_grad():
    for _ in range(2): m(img,return_loss=False)
    torch.cuda.synchronize()
    m(img,return_loss=False)
    torch.cuda.synchronize()
print("ok")
