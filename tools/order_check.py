import sys, contextlib, io
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import style_transfer_b200 as stb
from oracle import st_oracle as O
H,W=384,512
wts=O.make_vgg_weights(1234)
content, style = O.synth_image(1, 16, W, H), O.synth_image(2, 32, W // 2 + 40, H // 2 + 24)
st = stb.StyleTransfer(devices=['cuda:0'], pooling='max', vgg_weights=wts)
tr=[]
with contextlib.redirect_stdout(io.StringIO()):
    st.stylize(content,[style],min_scale=512,end_scale=512,initial_iterations=5,callback=lambda it: tr.append(it.loss))
print('LOSSES', ' '.join(f'{x:.8f}' for x in tr))
