import sys, os, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import multimodal_transformer_b200 as mtb
from oracle import fill, mt_oracle as O
from oracle.dropout_rng import Dropper
from tests import util
MODS = ['acoustic','image','linguistic']; DEV='cuda:0'; t=torch.from_numpy
res={}
def rel(a,b): 
    a=a.detach().double().cpu(); b=b.detach().double().cpu()
    return ((a-b).abs().max()/b.abs().max().clamp_min(1e-30)).item()
# 1. medium batch bf16 grads
N,B,T,seed=1,40,128,4711
dims={'acoustic':88,'image':256,'linguistic':300}
sd=util.filled_sd(util.mods_shapes('MFT.MultiTransformer',N),23)
inputs,mask,target,lengths=fill.make_batch(B,T,dims,23)
sdr={k:v.double().requires_grad_(True) for k,v in sd.items()}
predr=O.multi_transformer(sdr,'',{k:t(v).double() for k,v in inputs.items()},t(mask).double(),MODS,N=N,drop=Dropper(seed))
O.train_loss(predr,t(target).double(),lengths).backward()
gmax=max(v.grad.abs().max().item() for v in sdr.values() if v.grad is not None)
mtb.set_compute_dtype('bf16')
model=mtb.MultiTransformer(MODS,dims,N=N).train(); model.load_state_dict(sd)
mtb.fix_seed(seed)
pred=model({k:t(v).to(DEV) for k,v in inputs.items()},t(mask).to(DEV),lengths)
(((pred-t(target).to(DEV))**2).sum()/sum(lengths)).backward()
worst=[]
for k,p in model.named_parameters():
    w=sdr[k].grad
    if w is None: continue
    e=(p.grad.double().cpu()-w).abs().max().item()
    worst.append((e/max(w.abs().max().item(),1e-3*gmax), e/gmax, k))
worst.sort(reverse=True)
res['medium_bf16_worst_rel']=worst[:6]
mtb.fix_seed(None)
# 2. C4 T=1024 B=3 grads
t0=time.time()
dims4={'acoustic':256,'image':256,'linguistic':300}
T4,B4=1024,3
sd4=util.filled_sd(util.mods_shapes('B3.MultiTransformer'),12)
in4,mk4,tg4,len4=fill.make_batch(B4,T4,dims4,12)
sdr4={k:v.double().requires_grad_(True) for k,v in sd4.items()}
pr4=O.multi_transformer(sdr4,'',{k:t(v).double() for k,v in in4.items()},t(mk4).double(),MODS,use_encoder=False)
O.train_loss(pr4,t(tg4).double(),len4).backward()
res['c4_oracle_s']=time.time()-t0
g4max=max(v.grad.abs().max().item() for v in sdr4.values() if v.grad is not None)
for mode in ('fp32','bf16'):
    mtb.set_compute_dtype(mode)
    m4=mtb.B3MultiTransformer(MODS,dims4).eval(); m4.load_state_dict(sd4)
    p4=m4({k:t(v).to(DEV) for k,v in in4.items()},t(mk4).to(DEV),len4)
    (((p4-t(tg4).to(DEV))**2).sum()/sum(len4)).backward()
    w=[]
    for k,p in m4.named_parameters():
        ww=sdr4[k].grad
        if ww is None: continue
        e=(p.grad.double().cpu()-ww).abs().max().item()
        w.append((e/max(ww.abs().max().item(),1e-3*g4max), k))
    w.sort(reverse=True)
    res['c4_'+mode]={'pred_abs':(p4.detach().double().cpu()-pr4.detach()).abs().max().item(),'worst':w[:5]}
# 3. C5 d512 T=4096 B=1 N=1
t0=time.time()
from multimodal_transformer_b200.multiTransformer import _make_encoder as mk
def shapes(N,d,dff):
    s={}
    for l in range(N):
        p=f'layers.{l}.'
        for i in range(4): s[p+f'self_attn.linears.{i}.weight']=(d,d); s[p+f'self_attn.linears.{i}.bias']=(d,)
        s[p+'feed_forward.w_1.weight']=(dff,d); s[p+'feed_forward.w_1.bias']=(dff,)
        s[p+'feed_forward.w_2.weight']=(d,dff); s[p+'feed_forward.w_2.bias']=(d,)
        for j in range(2): s[p+f'sublayer.{j}.norm.a_2']=(d,); s[p+f'sublayer.{j}.norm.b_2']=(d,)
    s['norm.a_2']=(d,); s['norm.b_2']=(d,)
    return s
d,dff,N5,B5,T5=512,256,1,1,4096
enc=mk(d,dff,8,0.1,N5).to(DEV).eval()
sd5=util.filled_sd(shapes(N5,d,dff),33); enc.load_state_dict(sd5)
x=t(fill.fill_array('c5_x',(B5,T5,d),33))*3.0
mask=torch.ones(B5,T5,1); mask[:,3*T5//4:]=0
w=t(fill.fill_array('c5_w',(B5,T5,d),34))
sdr5={'e.'+k:v.double().requires_grad_(True) for k,v in sd5.items()}
xr=x.double().requires_grad_(True)
yr=O.encoder(sdr5,'e',xr,mask.double(),N5,8)
(yr*w.double()).sum().backward()
res['c5_oracle_s']=time.time()-t0
mtb.set_compute_dtype('bf16')
xd=x.to(DEV).requires_grad_(True)
y=enc(xd,mask.to(DEV))
(y.float()*w.to(DEV)).sum().backward()
res['c5_bf16']={'y_rel':rel(y,yr),'dx_rel':rel(xd.grad,xr.grad)}
gw=[]
for k,p in enc.named_parameters():
    ww=sdr5['e.'+k].grad
    gw.append((rel(p.grad,ww),k))
gw.sort(reverse=True); res['c5_bf16']['worst_param_grads']=gw[:5]
print(json.dumps(res,indent=1,default=str))
