import torch, numpy as np, sys
sys.path.insert(0,'.')
import mlsp_b200 as M
from mlsp_b200 import synth
dev=torch.device('cuda:0')
for (B,C,N,k) in [(1,64,256,20),(2,64,512,20),(2,128,640,20)]:
    x=synth.smooth_features(B,C,N,55)
    idx,v,stats=M.knn_tensor_debug(x.to(dev),k)
    torch.cuda.synchronize()
    xd=x.double(); xx=(xd**2).sum(1)
    exact=xx[:,None,:]-2*torch.einsum('bci,bcj->bij',xd,xd)
    err=(v.cpu().double()-exact).abs()
    bound=2.0**-11*xx.sqrt()[:,:,None]*xx.sqrt().amax(dim=1)[:,None,None]
    print((B,C,N,k),'nan',int(torch.isnan(v).sum()),'maxerr',float(err.max()),'max err/bound',float((err/bound).max()),stats, 'xx mean', float(xx.mean()))
    ex=M.knn(x.to(dev),k,flags=1)
    print('  idx equal exact kernel:',bool(torch.equal(ex,idx)), 'mismatch rows', int((ex!=idx).any(-1).sum()))
