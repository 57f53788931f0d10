import torch, sys
sys.path.insert(0,'/root/repo')
import mlsp_b200 as M
from mlsp_b200 import synth, _lib
dev=torch.device('cuda:0')
for C in (64,128):
    x=synth.smooth_features(32,C,1024,1244+C).to(dev)
    idx,st=M.knn(x,20,flags=_lib.KNN_TENSOR_ONLY,return_stats=True)
    print(C,st, st["candidates"]/max(st["certified_rows"],1), st["exact_recomputed"]/max(st["certified_rows"],1))
