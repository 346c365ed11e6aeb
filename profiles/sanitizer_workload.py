import sys, os
ROOT='/root/repo'
for p in (ROOT, ROOT+'/2d-gaussian-splatting_b200', ROOT+'/tests'):
    sys.path.insert(0,p)
import numpy as np, torch
import surfel_scenes as S
from diff_surfel_rasterization import GaussianRasterizationSettings, GaussianRasterizer, _cabi
from diff_surfel_rasterization.postprocess import surface_outputs
from diff_surfel_rasterization.loss import l1_ssim_loss
import types
dev='cuda'
lib=_cabi.load()
def run(P,W,H,variant=None):
    if variant: lib.surfel_set_variant(variant[0].encode(), variant[1].encode())
    cam=S.make_camera(W,H); scene=S.make_scene(P,W,H,3,depth_complexity=25)
    rs=GaussianRasterizationSettings(image_height=H,image_width=W,tanfovx=cam['tanfovx'],tanfovy=cam['tanfovy'],bg=torch.zeros(3,device=dev),scale_modifier=1.0,viewmatrix=cam['viewmatrix'].to(dev),projmatrix=cam['projmatrix'].to(dev),sh_degree=3,campos=cam['campos'].to(dev),prefiltered=False,debug=False)
    leaf={k:v.to(dev).requires_grad_(True) for k,v in scene.items()}
    m2d=torch.zeros(P,3,device=dev,requires_grad=True)
    for it in range(2):
        color,radii,allmap=GaussianRasterizer(rs)(means3D=leaf['means3D'],means2D=m2d,shs=leaf['shs'],opacities=leaf['opacities'],scales=leaf['scales'],rotations=leaf['rotations'])
        view=types.SimpleNamespace(world_view_transform=cam['viewmatrix'].to(dev),full_proj_transform=cam['projmatrix'].to(dev),image_width=W,image_height=H)
        o=surface_outputs(allmap,view,0.5)
        loss=l1_ssim_loss(color,torch.rand(3,H,W,device=dev),0.2)+(o['rend_normal']*o['surf_normal']).sum()+o['surf_depth'].sum()+allmap.sum()
        loss.backward()
    torch.cuda.synchronize()
    if variant: lib.surfel_set_variant(variant[0].encode(), variant[2].encode())
    print('ok',P,W,H,variant, float(loss.detach()))
run(3000,200,136)
run(700,100,70,("sort","radix","bucket"))
run(9000,160,120)      # ~500-entry tile lists: the render kernels take a second staging round
