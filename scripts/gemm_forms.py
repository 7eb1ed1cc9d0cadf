"""Which operand layouts does cuBLAS run fastest for the fc6 GEMMs (TF32)?  python scripts/gemm_forms.py"""
import torch, json
torch.backends.cuda.matmul.allow_tf32 = True
M, K, N = 8000, 25088, 4096
x = torch.randn(M, K, device="cuda"); W = torch.randn(N, K, device="cuda") * 0.01; Wt = W.t().contiguous()
g = torch.randn(M, N, device="cuda")
def t(fn, n=6):
    fn(); fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
fl = 2.0 * M * K * N
out = {}
for name, fn in {
    "fwd  x @ W.t()   (W [N,K])": lambda: torch.mm(x, W.t()),
    "fwd  x @ Wt      (Wt [K,N])": lambda: torch.mm(x, Wt),
    "dgrad g @ W      (W [N,K])": lambda: torch.mm(g, W),
    "dgrad g @ Wt.t() (Wt [K,N])": lambda: torch.mm(g, Wt.t()),
    "wgrad g.t() @ x  -> [N,K]": lambda: torch.mm(g.t(), x),
    "wgrad x.t() @ g  -> [K,N]": lambda: torch.mm(x.t(), g),
}.items():
    ms = t(fn); out[name] = (round(ms, 3), round(fl / ms / 1e9, 1))
    print(name, out[name])
