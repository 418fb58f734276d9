import sys, numpy as np, torch, time
sys.path.insert(0, '.')
from facet_graph_convolution_b200 import ops
from oracle import closed_form as cf
dev = torch.device('cuda:0')
def run(B,N,K,seed,holes=True):
    rs = np.random.RandomState(seed)
    Cin=Cout=64; M=8
    x = rs.randn(B,N,Cin).astype(np.float32)
    adj = rs.randint(0 if holes else 1, N+1, size=(B,N,K)).astype(np.int32); adj[:,:,0]=np.arange(1,N+1)
    if holes: adj[0,3]=0
    W0=(rs.randn(M,Cout,Cin)*0.05).astype(np.float32); b=(rs.randn(Cout)*0.01).astype(np.float32)
    u=(rs.randn(M,Cin)*0.05).astype(np.float32); v=(rs.randn(M,Cin)*0.05).astype(np.float32); c=(rs.randn(M)*0.05).astype(np.float32)
    t=lambda a: torch.from_numpy(a).to(dev)
    y = ops.conv_fwd(t(x),t(adj),t(W0),t(b),t(u),t(v),t(c))
    torch.cuda.synchronize()
    yr = cf.conv_fwd(x,adj,W0,b,u,v,c)
    err = np.abs(y.cpu().numpy()-yr).max()
    print('B',B,'N',N,'K',K,'max err',err, 'ymax', np.abs(yr).max(), flush=True)
    return err
errs=[run(1,64,16,0), run(1,100,16,1), run(2,256,16,2), run(1,5000,16,3), run(1,777,23,4), run(3,130,7,5)]
assert max(errs) < 1e-5, errs
print("TC OK")
