import torch, sys
sys.path.insert(0, "/root/repo")
from echo_tts_b200 import ops
def rnd(shape, seed): return torch.randn(shape, generator=torch.Generator().manual_seed(seed)).to("cuda", torch.bfloat16)
for (b,S,H,D,L2) in [(1,160,2,128,900),(2,100,2,64,900)]:
    q=rnd((b,S,H,D),1); k=rnd((b,S,H,D),2); v=rnd((b,S,H,D),3); k2=rnd((1,L2,H,D),4); v2=rnd((1,L2,H,D),5)
    segs=[dict(k=k,v=v),dict(k=k2,v=v2,batch_mod=1)]
    ref=torch.empty(b,S,H*D,device="cuda",dtype=torch.bfloat16); ops.attention(q,segs,ref,nsplit=1)
    for ns in range(2,9):
        out=torch.empty_like(ref)
        try:
            ops.attention(q,segs,out,nsplit=ns); torch.cuda.synchronize()
            d=(out.float()-ref.float())
            bad_rows=(d.abs().amax(-1)>0.05).nonzero()
            print(D,"nsplit",ns,"rel",(d.norm()/ref.float().norm()).item(),"bad rows",bad_rows[:,1].unique().tolist()[:40])
        except Exception as e:
            print(D,"nsplit",ns,"ERR",e)
