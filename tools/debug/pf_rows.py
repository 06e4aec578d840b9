import sys
for p in ('/root/repo','/root/repo/tests','/root/repo/oracle'): sys.path.insert(0,p)
import numpy as np, torch, scenes, util, warnings
warnings.simplefilter('ignore')
topo, params = util.pack(scenes.blobs(num_paths=64, canvas=96))
W=H=192
rng=np.random.RandomState(0)
d=(rng.rand(H,W,4).astype(np.float32)-0.5)
whole=util.gpu_render(topo,params,W,H,2,2,7,use_prefiltering=True,d_render_image=d)['d_params'].astype(np.float64)
for bands in ([(0,96),(96,192)], [(0,192)], [(0,64),(64,128),(128,192)]):
    parts=util.gpu_render_rows(topo,params,W,H,2,2,7,bands,d_render_image=d,use_prefiltering=True)
    print(bands, 'rel', util.rel_l2(whole, parts['d_params']))
# band-only d_image (zeros elsewhere)
tot=np.zeros_like(whole)
for (a,b) in [(0,96),(96,192)]:
    dd=np.zeros_like(d); dd[a:b]=d[a:b]
    tot+=util.gpu_render_rows(topo,params,W,H,2,2,7,[(a,b)],d_render_image=dd,use_prefiltering=True)['d_params']
print('band-only d_image rel', util.rel_l2(whole, tot))
import oracle_check
ref=oracle_check.render(topo,params,W,H,2,2,7,use_prefiltering=True,d_render_image=d)['d_params']
print('whole vs oracle', util.rel_l2(ref, whole))
