import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from conftest import CASE_SPECS, load_golden
from gdb_nerf_b200 import ops
from gdb_nerf_b200.config import make_cfg
DEV='cuda'
def md(a,b): return float((a.double().cpu()-b.double().cpu()).abs().max())
for case in ['dtu_b2','nerf_b4','train_b2']:
  for prefix in ['', 'inj_']:
    g=load_golden(case); spec=CASE_SPECS[case]; cfg=make_cfg(spec['recipe']); b=cfg.nerf.bundle_size
    adaptive = True if prefix else cfg.nerf.is_adaptive
    tex_ref=g.t('tex_nchw'); B,V,F,Hb,Wb=tex_ref.shape; fd=F-3
    cam=ops.camera_block(g.t("in_tar_exts").to(DEV), g.t("in_tar_ints").to(DEV), g.t("in_src_exts").to(DEV), g.t("in_src_ints").to(DEV), g.t("in_near_far").to(DEV), b, cfg.nerf.global_num_depth, cfg.mvs.inv_depth[-1])
    dr,vr=g.t(prefix+'depth_range').to(DEV), g.t(prefix+'vol_range').to(DEV)
    sl=ops.sample_bundles(dr,vr,cam,b,cfg.nerf.max_num_samples,cfg.mvs.inv_depth[-1],adaptive,want_rays=False)
    src=ops.prepare_sources(tex_ref[:,:,:fd].contiguous().to(DEV), g.t('in_rgb').to(DEV), b, cfg.nerf.max_mipmap_level)
    vol=ops.to_channels_last(g.t('feat_volume').to(DEV),8)
    mlp=ops.pack_mlp(g.mlp(), fd, device=DEV)
    args=(src,vol,dr,vr,cam,mlp,B,V,spec['H'],spec['W'],b,cfg.nerf.max_num_samples,cfg.mvs.inv_depth[-1],adaptive)
    tc=ops.render_fused(*args,taps=sl,precision=1); torch.cuda.synchronize()
    print(case,prefix,'sigma',md(tc['sigma'],g.t(prefix+'sigma')),'feat',md(tc['sample_feat'],g.t(prefix+'feat')),'w',md(tc['weights'],g.t(prefix+'weights')),
      'bfeat', md(tc['feat'], g.t(prefix+'bundle_feat').view(B,Hb,Wb,-1).permute(0,3,1,2)), 'depth', md(tc['depth'].reshape(-1), g.t(prefix+'bundle_depth'))/(spec['far']-spec['near']),
      'geo-head', md(tc['sample_feat'][:,-8:], g.t(prefix+'feat')[:,-8:]))
