# cuBLAS DGEMM peak (roofline denominator for the K1 expansion), same recipe as MEASURED_PEAKS.json's bf16 figure.
import torch, json, time
n=8192
a=torch.randn(n,n,dtype=torch.float64,device='cuda'); b=torch.randn(n,n,dtype=torch.float64,device='cuda')
for _ in range(3): c=a@b
torch.cuda.synchronize()
best=1e9
for _ in range(10):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); c=a@b; e1.record(); torch.cuda.synchronize()
    best=min(best,e0.elapsed_time(e1))
burst=2*n**3/best*1e-9
t0=time.time(); k=0
e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
e0.record()
while time.time()-t0<4.0:
    c=a@b; k+=1
    if k%5==0: torch.cuda.synchronize()
e1.record(); torch.cuda.synchronize()
sus=2*n**3*k/e0.elapsed_time(e1)*1e-9
# tall-skinny shape like K1: M=1M, K=100, N=512
m=1<<20; kk=100; nn=512
A=torch.randn(m,kk,dtype=torch.float64,device='cuda'); B=torch.randn(kk,nn,dtype=torch.float64,device='cuda')
for _ in range(3): C=A@B
torch.cuda.synchronize()
bt=1e9
for _ in range(10):
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record(); C=A@B; e1.record(); torch.cuda.synchronize()
    bt=min(bt,e0.elapsed_time(e1))
print(json.dumps({"dgemm_tflops_burst":burst,"dgemm_tflops_sustained":sus,"dgemm_8192_ms":best,
  "tallskinny_1Mx100x512_tflops":2*m*kk*nn/bt*1e-9,"tallskinny_ms":bt}))
