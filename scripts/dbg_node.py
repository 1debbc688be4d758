import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import sake_b200
from tests import synth
B,N,S=2,9,4
h,x,mask,am=synth.molecules(5,B,N,S,False,0)
T=lambda a: torch.tensor(a,device='cuda')
layer=sake_b200.DenseSAKELayer(64,64,update=True,engine=os.environ.get('ENG','auto'))
hh=torch.randn(B,N,64,device='cuda',generator=torch.Generator(device='cuda').manual_seed(1))
p=layer.init(0,hh,T(x))['params']
hh=hh.requires_grad_(True); xx=T(x).requires_grad_(True)
ho,xo,vo=layer.apply({'params':p},hh,xx,None,None)
s=(ho**2).sum()+(xo*0.3).sum()
gx,gh=torch.autograd.grad(s,[xx,hh])
torch.save({'ho':ho.detach().cpu(),'xo':xo.detach().cpu(),'gx':gx.cpu(),'gh':gh.cpu()}, 'gpurun_out/dbg_%s.pt'%os.environ.get('SAKE_NODE_TC','1'))
print('done', float(gh.abs().max()), float(gx.abs().max()))
