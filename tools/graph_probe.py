import sys, torch, time
sys.path.insert(0,'/root/repo')
import mcgvc_loader; pkg=mcgvc_loader.load(); eng=pkg.engine
torch.manual_seed(0); G=pkg.Generator().cuda()
x=torch.randn(1,80,64,device='cuda'); m=torch.ones_like(x)
with torch.no_grad():
    for i in range(6):
        y=G(x,m); torch.cuda.synchronize(); print(i, eng.graph_stats(), eng.lib().mcgvc_last_error())
    t0=time.perf_counter()
    for i in range(50): y=G(x,m)
    torch.cuda.synchronize(); print('ms/iter', (time.perf_counter()-t0)/50*1e3, eng.graph_stats())
    eng.set_graphs(False)
    for i in range(5): y=G(x,m)
    torch.cuda.synchronize(); t0=time.perf_counter()
    for i in range(50): y=G(x,m)
    torch.cuda.synchronize(); print('eager ms/iter', (time.perf_counter()-t0)/50*1e3)
