import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time, copy
from vpho_b200 import synthetic as syn, capi
from vpho_b200.head_mano import HeadMano
from vpho_b200.aggregation import Assets, HOI_Aggregator, HeadObject, HeadPhysics
from oracle import vpho_oracle as O
lib = capi.Library(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/emu/_build/libvpho_emu.so'), strict=False) if not torch.cuda.is_available() else capi.lib()
dev = 'cuda' if torch.cuda.is_available() else 'cpu'
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 100
Kh = int(sys.argv[3]) if len(sys.argv) > 3 else 30
Ko = int(sys.argv[4]) if len(sys.argv) > 4 else 10
mano = syn.make_mano_model(); anch = syn.make_anchor_assets(mano); objs = syn.make_object_tables()
batch = syn.make_eval_batch(bs, seed=5, sample_num=S)
T_ = lambda k: torch.from_numpy(np.asarray(batch[k]))
g = torch.Generator().manual_seed(3)
# hand candidates: noisy versions of a hidden pose
true_pose = torch.cat([T_('true_wrist'), torch.randn(bs, 45, generator=g)*0.2], 1)
pose_diff = (true_pose[:, None] + torch.randn(bs, S, 48, generator=g)*0.25).reshape(-1, 48).float()
shape = T_('pd_mano_shape')[:, None].repeat(1, S, 1).reshape(-1, 10)
rot = torch.randn(bs, S, 6, generator=g, dtype=torch.float64)
tr = T_('true_obj_rot').double()[:, None, :2, :].reshape(bs, 1, 6)
rot[:, ::2] = tr + 0.15*rot[:, ::2]
tt = T_('true_obj_trans').double()[:, None] + 0.02*torch.randn(bs, S, 3, generator=g, dtype=torch.float64)
OBJP = torch.cat([rot, tt], -1)
kw = dict(cam_intrinsic=T_('cam_intr_crop_flip'), root_joint_flip=T_('root_joint_flip'), root_joint=T_('root_joint'),
          is_right=T_('is_right'), force_local=T_('force_local'), is_grasped=T_('is_grasped'),
          hand_pose_diff=pose_diff.clone(), hand_pose_regression=T_('pd_mano_pose'), hand_shape=shape,
          hand_heatmap=T_('hm_hand'), hand_bbox=T_('bbox_hand'), hand_topk=Kh,
          obj_pose6d=OBJP, obj_heatmap=T_('hm_obj'), obj_bbox=T_('bbox_obj_rect'), obj_topk=Ko, obj_name=batch['obj_name'])
om = O.OracleMano(mano)
t0=time.time(); oo = O.hoi_aggregate(om, O.OracleObject(objs), O.OracleAnchors(anch), **copy.deepcopy(kw)); t1=time.time()
hm = HeadMano(mano, lib=lib); assets = Assets(anch, objs, lib=lib)
agg = HOI_Aggregator(hm, assets, debug=True)
kwd = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in kw.items()}
t2=time.time(); r = agg(**kwd); 
if dev == 'cuda': torch.cuda.synchronize()
t3=time.time()
print('oracle', t1-t0, 'ours', t3-t2)
d = agg.last_debug; od = oo['_dbg']
for lv in range(4):
    L = od['cascade']['levels'][lv]
    sc = L['score']; sc = sc[..., None] if sc.dim() == 2 else sc
    nf = sc.shape[-1]
    ours = d['hand_score'][lv, :, :, :nf].cpu()
    print('level', lv, 'score maxdiff', (ours - sc).abs().max().item(), 'mag', sc.abs().mean().item())
    tk = L['topk']; tk = tk[..., None] if tk.dim() == 2 else tk   # (bs,K,nf)
    ot = d['hand_topk'][lv].cpu()[:, :nf].permute(0, 2, 1)
    print('   topk equal', (ot == tk).float().mean().item())
print('cascade pose', (d['cascade_pose'].cpu() - od['cascade']['agg_hand_mano'][:, :48]).abs().max().item())
print('force_point', (d['force_point'].cpu() - od['force_point']).abs().max().item(), 'force_global', (d['force_global'].cpu() - od['force_global']).abs().max().item())
omax = max(S, Ko*Ko)
for i, (nm, C) in enumerate([('obj_transl_score', S), ('obj_rot_score', S), ('phys_score', Ko*Ko), ('heat5_score', Ko*Ko)]):
    ours = d['obj_score'][i, :bs*C].reshape(bs, C).cpu()
    print(nm, (ours - od[nm]).abs().max().item(), 'rel', ((ours-od[nm]).abs()/od[nm].abs().clamp(min=1e-12)).max().item())
for i, nm in enumerate(['obj_transl_topk', 'obj_rot_topk', 'phys_topk', 'heat5_topk']):
    k = od[nm].shape[1]
    print(nm, (d['obj_topk'][i, :, :k].cpu() == od[nm]).float().mean().item())
print('finger_score', (d['finger_score'].cpu() - od['finger_score']).abs().max().item(), 'topk', (d['finger_topk'].cpu() == od['finger_topk']).float().mean().item())
for k in ['obj_agg_6d', 'pose6d_candidate', 'agg_obj_vert', 'hand_agg_mano', 'hand_agg_vert', 'hand_agg_joint']:
    print(k, r[k].dtype, oo[k].dtype, tuple(r[k].shape), (r[k].cpu()-oo[k]).abs().max().item())
