import torch, sys
sys.path.insert(0, '/root/repo')
from gameplay_vision_llm_b200 import ops
DEV='cuda:0'
def run(B,T,H,hd,scale_in=1.5):
    g = torch.Generator().manual_seed(B*100+T)
    D=H*hd
    qkv=(torch.randn(B*T,3*D,generator=g)*scale_in).to(torch.bfloat16).to(DEV)
    out=ops.attention(qkv,B,T,H,hd); torch.cuda.synchronize()
    q,k,v=qkv.float().view(B,T,3,H,hd).permute(2,0,3,1,4)
    s=(q@k.transpose(-1,-2))*hd**-0.5
    ref=(torch.softmax(s,-1)@v).transpose(1,2).reshape(B*T,D)
    err=(out.float()-ref).abs()
    rowerr=err.view(B,T,H,hd).amax(-1)   # B,T,H
    bad=(rowerr>0.05).nonzero()
    print(f"B={B} T={T} H={H} hd={hd} scale={scale_in}: max err {err.max().item():.4f}, bad rows {bad.shape[0]} / {B*T*H}")
    if bad.shape[0]:
        print(' first bad (b,t,h):', bad[:8].tolist())
        print(' bad t mod 128 histogram:', torch.bincount(bad[:,1]%128, minlength=128).nonzero().flatten().tolist()[:40])
        # column pattern
        b0,t0,h0=bad[0].tolist()
        e=err.view(B,T,H,hd)[b0,t0,h0]
        print(' err by d:', [round(x,3) for x in e.tolist()][:hd])
for args in [(1,1568,12,64),(1,1568,12,64,0.5),(1,729,16,72,3.0),(1,1024,4,64),(1,640,4,64),(1,1568,2,72)]:
    run(*args)
