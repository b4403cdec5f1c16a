import sys; sys.path.insert(0,'/root/repo')
import numpy as np, hvo_b200 as hvo
from hvo_b200 import synth
g,d = synth.frame('S1',0); c = synth.CONFIGS['S1']
K = np.array([[c['fx'],0,c['cx']],[0,c['fy'],c['cy']],[0,0,1]],np.float32)
pd = hvo.PlaneDetection(640,480,max_batch=1)
pd.readDepthImage(d,K,np.float32(1.0/c['factor']))
pd.runPlaneDetection(480,640)
pd.runPlaneDetection(480,640)
print(pd.phase_cycles(0))
