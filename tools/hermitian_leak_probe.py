"""What a C2R that does not ignore Im X(0) / Im X(N/2) does to the interlaced final mesh (CPU, numpy): the even-length
real-transform trick (N/2-point complex FFT of Z[k] = (X[k] + conj X[N/2-k]) + i w^k (X[k] - conj X[N/2-k])) applied
along z after full complex transforms along x and y, against numpy's irfftn (which returns the real part of the full
inverse).  On Hermitian input the two agree to 1e-15; on the interlaced, deconvolved spectrum of a displaced lattice
(nbody.py:513-577) they differ by 9e-2 of the rms of delta at 64^3.  Supports tests/test_zz_cross_check_256.py."""
import numpy as np, sys
sys.path.insert(0,'/root/repo')
from oracle import cpu_port
ops = cpu_port.cpu_ops()
import torch
n=64
rng=np.random.default_rng(0)
ax=[np.arange(n,dtype=np.float32)]*3
q=np.stack(np.meshgrid(*ax,indexing='ij'),-1).reshape(-1,3)
# smooth displacement field ~1 cell rms
kk=np.sqrt(sum(np.meshgrid(np.fft.fftfreq(n)**2,np.fft.fftfreq(n)**2,np.fft.rfftfreq(n)**2,indexing='ij'))); kk[0,0,0]=1
disp=[np.fft.irfftn(np.fft.rfftn(rng.normal(size=(n,n,n)))*kk**-1.5) for _ in range(3)]
disp=np.stack([d/d.std() for d in disp],-1).reshape(-1,3).astype(np.float32)*1.0
pos=torch.tensor(q+disp)
gk=ops.nufft_paint(pos,(n,n,n),None,1.0,None,2,2,True).numpy().astype(np.complex128)
proper=np.fft.irfftn(gk,s=(n,n,n))
# leaky: c2c inverse over x,y then even-length real trick on z without ignoring Im(X0), Im(XN/2)
Y=np.fft.ifftn(gk,axes=(0,1))
N=n; h=N//2
k=np.arange(h)
Xk=Y[...,:h]; Xc=np.conj(Y[..., h-k])  # X[N/2-k]
w=np.exp(2j*np.pi*k/N)
Z=(Xk+Xc)+1j*w*(Xk-Xc)
z=np.fft.ifft(Z,axis=-1)*0.5   # scaling: x = irfft
leaky=np.empty((n,n,n)); leaky[...,0::2]=z.real; leaky[...,1::2]=z.imag
# check the trick equals proper on a Hermitian input
H=np.fft.rfftn(proper); Yh=np.fft.ifftn(H,axes=(0,1)); Xk=Yh[...,:h]; Xc=np.conj(Yh[...,h-k]); Zh=(Xk+Xc)+1j*w*(Xk-Xc); zh=np.fft.ifft(Zh,axis=-1)*0.5
chk=np.empty((n,n,n)); chk[...,0::2]=zh.real; chk[...,1::2]=zh.imag
print("trick sanity", np.abs(chk-proper).max())
d=proper-1
print("rel L2 diff of delta (leaky vs proper):", np.linalg.norm(leaky-proper)/np.linalg.norm(d), " delta rms", d.std())
