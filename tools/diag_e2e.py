import sys; import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, time, copy
from vpho_b200 import synthetic as syn, capi
from vpho_b200.vpho import VphoHotPath, to_device
from vpho_b200.score_based_model import ve_prior_std
from oracle import vpho_oracle as O
cuda = torch.cuda.is_available()
lib = capi.lib() if cuda else capi.Library(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests/emu/_build/libvpho_emu.so'), strict=False)
dev = 'cuda' if cuda else 'cpu'
bs = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 16
Kh = int(sys.argv[3]) if len(sys.argv) > 3 else 6
Ko = int(sys.argv[4]) if len(sys.argv) > 4 else 4
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 10
mano = syn.make_mano_model(); anch = syn.make_anchor_assets(mano); objs = syn.make_object_tables()
batch = syn.make_eval_batch(bs, seed=5, sample_num=S)
st_h = syn.make_denoiser_state('mano_pose', 0); st_o = syn.make_denoiser_state('obj', 0)
hp = VphoHotPath(mano, anch, objs, st_h, st_o, sample_num=S, sampling_steps=steps, topk_hand=Kh, topk_obj=Ko, lib=lib, debug=True)
g = torch.Generator().manual_seed(11)
ph = torch.randn(bs*S, 96, generator=g) * ve_prior_std(0.65); po = torch.randn(bs*S, 9, generator=g) * ve_prior_std(0.65)
t0=time.time(); pd = hp.predict(to_device(batch, dev), prior_hand=ph, prior_obj=po)
if cuda: torch.cuda.synchronize()
t1=time.time()
oo = O.oracle_predict(batch, O.OracleDenoiser(st_h), O.OracleDenoiser(st_o), O.OracleMano(mano), O.OracleObject(objs), O.OracleAnchors(anch),
                      init_x_hand=ph, init_x_obj=po, sample_num=S, sampling_steps=steps, topk_hand=Kh, topk_obj=Ko, with_inprocess=True)
t2=time.time()
print('ours', t1-t0, 'oracle', t2-t1, hp.last_info, oo['_info'])
for k in sorted(oo):
    if k.startswith('_'): continue
    if k not in pd: print('MISSING', k); continue
    a, b = pd[k].cpu(), oo[k]
    print(k, tuple(a.shape), tuple(b.shape), a.dtype, b.dtype, (a.double()-b.double()).abs().max().item())
a, b = pd['agg_hand_mano'].cpu(), oo['agg_hand_mano']
print((a-b).abs()[0].reshape(-1)[:48].reshape(16,3))
print(b[0,:48].reshape(16,3))
d = hp.hoi_aggregator.last_debug; od = oo['_sel']['_dbg']
print('cascade pose diff', (d['cascade_pose'].cpu() - od['cascade']['agg_hand_mano'][:, :48]).abs().reshape(16,3))
for lv in range(4):
    L = od['cascade']['levels'][lv]
    sc = L['score']; sc = sc[..., None] if sc.dim() == 2 else sc
    nf = sc.shape[-1]
    ours = d['hand_score'][lv, :, :, :nf].cpu()
    tk = L['topk']; tk = tk[..., None] if tk.dim() == 2 else tk
    ot = d['hand_topk'][lv].cpu()[:, :nf].permute(0, 2, 1)
    print('level', lv, 'score maxdiff', (ours - sc).abs().max().item(), 'topk equal', (ot == tk).float().mean().item())
for i, nm in enumerate(['obj_transl_topk', 'obj_rot_topk', 'phys_topk', 'heat5_topk']):
    k = od[nm].shape[1]
    print(nm, (d['obj_topk'][i, :, :k].cpu() == od[nm]).float().mean().item())
print('finger topk', (d['finger_topk'].cpu() == od['finger_topk']).float().mean().item(), 'finger score', (d['finger_score'].cpu() - od['finger_score']).abs().max().item())
print('hand_cand_pose -> fused', (pd['agg_hand_mano'].cpu() - oo['agg_hand_mano']).abs().max().item())
