"""Reads a verbose K3T timeline (`python tools/tensor_timeline.py 75600 v > file`, profiling build stamps per op) and prints
(1) the hand-off latencies between the MMA and EPI streams, (2) the tile's critical path walked back from its last op along
the binding dependency (previous op of the same stream, or the op whose event it waited for), by category, and (3) a
least-squares model of an MMA op's period.  What profiles/r2d_k_solve_tc_timeline.txt quotes.  CPU only (device -1 plan).
usage: python tools/tensor_critical_path.py gpurun_out/timeline.txt"""
import os, sys
import numpy as np
ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'sdfa-2019_b200')]
fn=sys.argv[1]
E=[];M=[]
for l in open(fn):
    p=l.split()
    if p and p[0]=='E': E.append([int(x) for x in p[1:]])
    if p and p[0]=='M': M.append([int(x) for x in p[1:]])
E=np.array(E);M=np.array(M)
import deformation as D
from deformation import workloads as W
from tests import tplan_emulator as T
V,F,nfv,nft=W.load_flame()
rec=D.Reconstructor(V,F,cnsts=nfv,device=-1,solver="tensor")
pl=T.plan(rec); e=pl["epi"]; m=pl["mma"]
Et=E[:,3:]; Mt=M[:,3:]
# event -> producer op
mma_of_evt={int(m["commit_mma"][i]):i for i in range(len(m)) if m["commit_mma"][i]>=0}
epi_of_evt={}
for i in range(len(e)):
    if e["signal_epi"][i]>=0: epi_of_evt[int(e["signal_epi"][i])]=(i,'end')
    if e["signal_read"][i]>=0: epi_of_evt[int(e["signal_read"][i])]=(i,'read')
# latencies M commit issued (t4) -> E evwait done (t2) when E actually waited (t2-t1 > 150)
lat=[]
for i in range(len(e)):
    w=int(e["wait_mma"][i])
    if w>=0 and Et[i,2]-Et[i,1]>150:
        j=mma_of_evt[w]; lat.append(Et[i,2]-Mt[j,4])
print("M commit-issued -> E woke: n",len(lat),"median",np.median(lat),"mean",np.mean(lat), "p10/p90",np.percentile(lat,[10,90]))
lat2=[]
for i in range(len(m)):
    for w in (int(m["wait_epi"][i]),int(m["wait_epi2"][i])):
        if w>=0 and Mt[i,1]-Mt[i,0]>150:
            j,kind=epi_of_evt[w]
            t_sig = Et[j,5] if kind=='end' else Et[j,3]
            lat2.append((Mt[i,1]-t_sig,kind))
a=[x for x,k in lat2 if k=='end']; b=[x for x,k in lat2 if k=='read']
print("E signal(end stamp) -> M woke: n",len(a),"median",np.median(a) if a else None, " read-signal (ld-done stamp) -> M woke: n",len(b),"median",np.median(b) if b else None)
# critical path walk backwards from last op end
# nodes: ('E',i) / ('M',i). finish times: E: t5 ; M: t4. predecessor candidates: previous op in same stream, and waited ops. choose the one that finished last before our 'work start'.
def prev_in_stream_E(i):
    s=e["stream"][i]
    for j in range(i-1,-1,-1):
        if e["stream"][j]==s: return j
    return None
cur=('E',int(np.argmax(Et[:,5])))
path=[]
acc={}
def add(k,v): acc[k]=acc.get(k,0)+v
while cur is not None:
    kind,i=cur
    if kind=='E':
        t=Et[i]
        # candidates
        cands=[]
        pj=prev_in_stream_E(i)
        if pj is not None: cands.append((Et[pj,5],('E',pj),'stream'))
        w=int(e["wait_mma"][i])
        if w>=0: cands.append((Mt[mma_of_evt[w],4],('M',mma_of_evt[w]),'mma'))
        if not cands: break
        best=max(cands)
        # time attributed: from best finish to our end
        if best[2]=='mma':
            add('M->E latency (commit issued to E awake)', t[2]-best[0]); add('E body after wake', t[5]-t[2])
        else:
            add('E body (stream-bound)', t[5]-best[0])
        cur=best[1]
    else:
        t=Mt[i]
        cands=[]
        if i>0: cands.append((Mt[i-1,4],('M',i-1),'stream'))
        for w in (int(m["wait_epi"][i]),int(m["wait_epi2"][i])):
            if w>=0:
                j,k=epi_of_evt[w]; ts=Et[j,5] if k=='end' else Et[j,3]
                cands.append((ts,('E',j),'epi-'+k))
        if not cands: break
        best=max(cands)
        if best[2].startswith('epi'):
            add('E->M latency (signal to M awake)', t[1]-best[0]); add('M body after wake', t[4]-t[1])
            # E op: attribute only up to its signal time next iteration: hack: if read-signal, next E node contributes up to t3
        else:
            add('M body (stream-bound)', t[4]-best[0])
        cur=best[1]
tot=sum(acc.values())
for k,v in sorted(acc.items(), key=lambda x:-x[1]): print(f"{k:45s} {v:8d} {100*v/tot:5.1f}%")
print("total",tot)

# ---- regression of an MMA op's period on what it does
gap=Mt[1:,0]-Mt[:-1,4]
nc=((m["flags"][:-1]&4)>0).astype(int)+(m["commit_mma"][:-1]>=0).astype(int)
for k in (0,1,2): print("gap after op with",k,"commits: n",(nc==k).sum(),"median",np.median(gap[nc==k]) if (nc==k).any() else None)
work=(m["k8"][:-1]*3*m["n"][:-1]).astype(float)
print("corr(gap, n*k)",np.corrcoef(gap,work)[0,1], "corr(gap,k8)",np.corrcoef(gap,m["k8"][:-1])[0,1])
iss=Mt[:,3]-Mt[:,2]
print("corr(issue, 3k8)",np.corrcoef(iss,m["k8"]*3)[0,1],"corr(issue, n*k8)",np.corrcoef(iss,m["k8"]*m["n"])[0,1])
A=np.stack([np.ones(len(m)),m["k8"]*3.0,m["k8"]*3.0*m["n"]],1)
co,res,_,_=np.linalg.lstsq(A,iss,rcond=None)
print("issue ~ %.0f + %.1f*nMMA + %.3f*nMMA*N"%tuple(co))
tot=Mt[1:,0]-Mt[:-1,0]
noev=(m["wait_epi"][:-1]<0)&(m["wait_epi2"][:-1]<0)
A2=np.stack([np.ones(noev.sum()),(m["k8"][:-1]*3.0)[noev],(m["k8"][:-1]*3.0*m["n"][:-1])[noev], nc[noev], ((m["flags"][:-1]&2)>0)[noev]],1)
co2,_,_,_=np.linalg.lstsq(A2,tot[noev],rcond=None)
print("period(no event) ~ %.0f + %.1f*nMMA + %.3f*nMMA*N + %.0f*commits + %.0f*chunkfirst"%tuple(co2))
