import ctypes, os, sys
ROOT='/root/repo'
for p in (ROOT, ROOT+'/tests', ROOT+'/oracle'): sys.path.insert(0,p)
import numpy as np, torch, scenes, util, warnings
warnings.simplefilter('ignore')
from diffvg_b200 import _native as n
sys.path.insert(0, ROOT+'/tools')
import measure_configs as mc
s = mc.Scene(scenes.blobs())
W=H=2048
img=torch.empty(H,W,4,device='cuda'); dimg=torch.empty_like(img)
for _ in range(2): s.step(W,H,2,2,0,1,img,dimg)
torch.cuda.synchronize()
n.profile_enable(True)
for _ in range(3): s.step(W,H,2,2,0,1,img,dimg)
torch.cuda.synchronize()
rep=n.profile_report()
for k,(c,ms) in sorted(rep.items(), key=lambda kv:-kv[1][1]): print('%-36s %3d launches %9.3f ms each'%(k,c,ms/c))
