import sys, os
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT,"tests"))
import numpy as np, torch
import rnabloom_b200 as rb
from oracle.binding import Oracle, OracleGraph, MODE_CANON
from rnabloom_b200.sharded import GpuBackend, ShardedGraph
from test_gpu_parity import DevReads
from parity_util import all_bases, counters_that_may_differ
orc=Oracle(); ctx=rb.Context(0)
k,hd,hc,dbg_bits,cbf_bytes=25,3,3,(1<<30)+77,(1<<28)+13
reads=orc.synth_reads(61,40000,0,1200,150,6000); seqs=[bytes(r).decode() for r in reads]
be=GpuBackend(ctx,1,0,dbg_bits,cbf_bytes,hd,hc,k,False,80000); sg=ShardedGraph(be,0,1)
og=OracleGraph(orc,dbg_bits,cbf_bytes,64,hd,hc,1,k,False,False)
def cmp(tag):
    torch.cuda.synchronize()
    d=sg.gather_filter(0,(dbg_bits+7)//8); c=sg.gather_filter(1,cbf_bytes)
    diff=np.nonzero(c!=og.cbf())[0]
    allowed,frac=counters_that_may_differ(all_bases(orc,seqs,k,[MODE_CANON]),k,hc,cbf_bytes)
    bad=[x for x in diff.tolist() if x not in allowed]
    print(tag,"dbg equal",(d==og.dbgbf()).all(),"cbf diffs",len(diff),"unexplained",len(bad),"sum gpu",int(c.sum()),"sum orc",int(og.cbf().sum()))
    if bad: print("  gpu",c[bad[:12]],"orc",og.cbf()[bad[:12]])
for r in range(3):
    chunk=seqs[r*400:(r+1)*400]; dr=DevReads(ctx,rb.pack_reads(chunk)); sg.add_round(dr.args,0); torch.cuda.synchronize(); dr.free()
    for s in chunk: og.add_read(s)
    cmp("round%d"%r)
chunk=seqs[:400]+seqs[:200]; dr=DevReads(ctx,rb.pack_reads(chunk)); sg.add_round(dr.args,0); torch.cuda.synchronize(); dr.free()
for s in chunk: og.add_read(s)
cmp("dups")
dr=DevReads(ctx,rb.pack_reads(seqs[100:300])); sg.add_round(dr.args,2); torch.cuda.synchronize()
for s in seqs[100:300]: og.add_read(s,flags=2)
cmp("countifpresent")
