import sys, re, collections
rows=[]
for l in sys.stdin:
    m=re.match(r"DBG n (\d+) created (\d+) maxlive (\d+) nodes (\d+) requeue (\d+)", l)
    if m: rows.append(tuple(int(x) for x in m.groups()))
import numpy as np
a=np.array(rows)
if len(a):
    n=a[:,0]; 
    for lo,hi in [(65,128),(129,256),(257,512),(513,1024)]:
        s=a[(n>=lo)&(n<=hi)]
        if len(s):
            print(lo,hi,len(s),'created/n mean %.2f max %.2f'%((s[:,1]/s[:,0]).mean(),(s[:,1]/s[:,0]).max()),'maxlive/n mean %.2f max %.2f'%((s[:,2]/s[:,0]).mean(),(s[:,2]/s[:,0]).max()),'nodes/n mean %.2f max %.2f'%((s[:,3]/s[:,0]).mean(),(s[:,3]/s[:,0]).max()), 'requeue', s[:,4].sum())
