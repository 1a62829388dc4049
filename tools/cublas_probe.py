import torch
for (M,N,K) in [(147456,6144,768),(147456,768,3072)]:
    a=torch.randn(M,K,device="cuda").bfloat16(); w=torch.randn(N,K,device="cuda").bfloat16()
    for _ in range(2): torch.matmul(a,w.t())
    torch.cuda.synchronize()
