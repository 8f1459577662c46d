"""ief_attn_fwd at the small-batch shapes of the DDIM inversion forwards (B=1/2): checks the pair / split-KV / hybrid dispatch
(IEF_TC_SPLITKV=0|1|2 forces a mode; unset = the launcher's estimate). CUDA events, L2 flushed, GPU kept busy before the bracket."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from image_editing_framework_b200 import ops
dev = torch.device("cuda:0")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def time_call(fn, reps=20, warm=5):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        flush.zero_(); torch.cuda._sleep(200000)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); e.synchronize(); ts.append(s.elapsed_time(e))
    ts.sort(); return ts[len(ts)//2]
for (B,H,N,d) in ((1,8,4096,40),(2,8,4096,40),(1,8,1024,80),(1,5,9216,64),(1,10,4096,64)):
    q,k,v = (torch.randn(B,N,H*d,device=dev).to(torch.bfloat16) for _ in range(3))
    ms = time_call(lambda: ops.attention(q,k,v,H,d**-0.5, impl=ops.IEF_IMPL_TCGEN05))
    print(json.dumps(dict(B=B,H=H,N=N,d=d,ms=round(ms,4),tflops=round(4*B*H*N*N*d/ms/1e9,1), mode=os.environ.get("IEF_TC_SPLITKV","auto"))), flush=True)
