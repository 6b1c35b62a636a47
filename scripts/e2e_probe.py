import sys, copy, time, torch
sys.path.insert(0,'/root/repo')
import bench
import denseclip_vit_multimodal_b200 as D
torch.manual_seed(0)
m=D.DenseCLIP(**copy.deepcopy(bench.model_kwargs())); bench.init_uninitialised(m); m=m.eval().cuda(); m.enable_cuda_graph(True)
B,H,W=16,512,1024
host=[torch.randn(B,3,H,W).pin_memory() for _ in range(2)]
dev=[torch.empty(B,3,H,W,device='cuda') for _ in range(2)]
segh=torch.empty(B,H,W,dtype=torch.uint8).pin_memory(); dph=torch.empty(B,1,H,W).pin_memory()
def t(fn,n=20):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
with torch.no_grad():
    for _ in range(3): o=m.predict(dev[0])
    print('compute only  ms', t(lambda: m.predict(dev[0])))
    print('h2d only      ms', t(lambda: dev[1].copy_(host[1],non_blocking=True)))
    print('d2h only      ms', t(lambda: (segh.copy_(o['seg'],non_blocking=True), dph.copy_(o['depth'],non_blocking=True))))
    cs=torch.cuda.Stream()
    def both():
        with torch.cuda.stream(cs):
            dev[1].copy_(host[1],non_blocking=True)
        m.predict(dev[0])
    print('compute || h2d ms', t(both))
    def both2():
        with torch.cuda.stream(cs):
            dev[1].copy_(host[1],non_blocking=True); segh.copy_(o['seg'],non_blocking=True); dph.copy_(o['depth'],non_blocking=True)
        m.predict(dev[0])
    print('compute || h2d+d2h ms', t(both2))
    from denseclip_vit_multimodal_b200.pipeline import PipelinedPredictor
    pipe=PipelinedPredictor(m,(B,3,H,W),'cuda')
    def run(n):
        for i in range(n):
            if i>=2: pipe.collect()
            pipe.submit(host[i%2])
        while pipe.n_collected<pipe.n_submitted: pipe.collect()
    run(4); torch.cuda.synchronize(); t0=time.perf_counter(); run(30); torch.cuda.synchronize(); print('pipeline ms/step', (time.perf_counter()-t0)/30*1e3)
