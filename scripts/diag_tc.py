import numpy as np, sys, time
sys.path.insert(0,'.')
from imagestitch_b200 import gpu, synth
A,B,off=synth.pair(1234,2048,205,1)
L=409
kA,dA=gpu.surf_detect_and_describe(A[2048-L:]); kB,dB=gpu.surf_detect_and_describe(B[:L])
print(len(dA),len(dB))
gpu.set_matcher("tc")
m=gpu.match_descriptors(dA,dB,2,0.75); print("matches",len(m),"fallbacks",gpu.last_match_fallbacks())
