import sys, cProfile, pstats, io, time
ROOT='/root/repo'
for p in (ROOT, ROOT+'/2d-gaussian-splatting_b200'): sys.path.insert(0,p)
import torch
import surfel_scenes as S
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer
dev='cuda'
P,W,H=S.CONFIGS['config1']
scene,cam=S.named('config1')
rs=GaussianRasterizationSettings(image_height=H,image_width=W,tanfovx=cam['tanfovx'],tanfovy=cam['tanfovy'],bg=torch.zeros(3,device=dev),scale_modifier=1.0,viewmatrix=cam['viewmatrix'].to(dev),projmatrix=cam['projmatrix'].to(dev),sh_degree=3,campos=cam['campos'].to(dev),prefiltered=False,debug=False)
leaf={k:v.to(dev).requires_grad_(True) for k,v in scene.items()}
m2d=torch.zeros(P,3,device=dev,requires_grad=True)
gc,go=S.make_cotangents(W,H,1); gc,go=gc.to(dev),go.to(dev)
rast=GaussianRasterizer(rs)
def step():
    for t in list(leaf.values())+[m2d]: t.grad=None
    color,radii,allmap=rast(means3D=leaf['means3D'],means2D=m2d,shs=leaf['shs'],opacities=leaf['opacities'],scales=leaf['scales'],rotations=leaf['rotations'])
    torch.autograd.backward([color,allmap],[gc,go])
for _ in range(20): step()
torch.cuda.synchronize()
t0=time.perf_counter()
for _ in range(300): step()
torch.cuda.synchronize()
print('ms/step', (time.perf_counter()-t0)/300*1e3)
pr=cProfile.Profile(); pr.enable()
for _ in range(300): step()
torch.cuda.synchronize(); pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('tottime').print_stats(18); print(s.getvalue()[:3500])
