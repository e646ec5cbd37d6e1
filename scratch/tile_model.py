"""numpy model of the tile algorithm the CUDA kernel implements (dev aid, not shipped)."""
import numpy as np, torch, math
from oracle import *
from pgmuvi_b200 import synthetic as S
torch.set_default_dtype(torch.float64)
TS=64
def run(n=200,Q=4):
    bt=S.make_batch_1d(1,n,Q=Q,learn_noise=True)
    spec=ModelSpec(d=1,Q=Q,kind=0,learn_noise=True)
    t=lambda a: torch.tensor(a)
    x,y,nz,raw,lb,ub,kinds=t(bt['x'][0]),t(bt['y'][0]),t(bt['noise'][0]),t(bt['raw'][0]),t(bt['lb'][0]),t(bt['ub'][0]),t(bt['kinds'])
    mll_o,g_o,_=mll_and_grad_analytic(x,y,nz,raw,kinds,lb,ub,spec)
    theta=constrain(raw,kinds,lb,ub)
    mean,w,mu,sg,noise=unpack_params(theta,spec)
    K=sm_kernel_dense(x,x,w,mu,sg,0).numpy()+np.diag(nz.numpy()+float(noise))
    N=(n+TS-1)//TS; npad=N*TS
    Kp=np.eye(npad); Kp[:n,:n]=K
    rhs=np.zeros(npad); rhs[:n]=y.numpy()-float(mean)
    T={}  # tiles
    z=np.zeros(npad); logdet=0.0
    sl=lambda i: slice(i*TS,(i+1)*TS)
    for j in range(N):
        for i in range(j,N):
            acc=np.zeros((TS,TS))
            for k in range(j): acc+=T[(i,k)]@T[(j,k)].T
            C=Kp[sl(i),sl(j)]-acc
            if i==j:
                L=np.linalg.cholesky(np.tril(C)+np.tril(C,-1).T)
                logdet+=2*np.log(np.diag(L)).sum()
                Xd=np.linalg.inv(L); Xd=np.tril(Xd)
                T[(j,j)]=Xd
                u=np.zeros(TS)
                for k in range(j): u+=T[(j,k)]@z[sl(k)]
                z[sl(j)]=Xd@(rhs[sl(j)]-u)
            else:
                T[(i,j)]=C@T[(j,j)].T
    invquad=z@z
    mll=-0.5*(invquad+logdet+n*math.log(2*math.pi))/n
    print('mll',mll,float(mll_o))
    # T phase
    for j in range(N-1):
        for i in range(j+1,N):
            acc=np.zeros((TS,TS))
            for k in range(j,i): acc+=T[(i,k)]@T[(k,j)]
            T[(i,j)]=-T[(i,i)]@acc
    alpha=np.zeros(npad)
    for j in range(N):
        for i in range(j,N): alpha[sl(j)]+=T[(i,j)].T@z[sl(i)]
    # G phase
    Kinv=np.zeros((npad,npad))
    for i in range(N):
        for j in range(i+1):
            acc=np.zeros((TS,TS))
            for k in range(i,N): acc+=T[(k,i)].T@T[(k,j)]
            Kinv[sl(i),sl(j)]=acc
    Kinv_full=np.tril(Kinv)+np.tril(Kinv,-1).T
    ref=np.linalg.inv(Kp)
    print('Kinv err',np.abs(Kinv_full-ref).max()/np.abs(ref).max())
    print('alpha err',np.abs(alpha[:n]-np.linalg.solve(K,rhs[:n])).max())
    # gradient by lower-tri weighting
    tt=x.numpy()[:,0]; tp=np.zeros(npad); tp[:n]=tt
    wv,muv,sgv=w.numpy(),mu.numpy()[:,0],sg.numpy()[:,0]
    gw=np.zeros(Q);gm=np.zeros(Q);gs=np.zeros(Q);trW=0.0
    for i in range(n):
        for j in range(i+1):
            Wij=alpha[i]*alpha[j]-Kinv[i,j]
            wt=1.0 if i==j else 2.0
            tau=tp[i]-tp[j]
            if i==j: trW+=Wij
            for q in range(Q):
                E=math.exp(-2*math.pi**2*sgv[q]**2*tau*tau); ph=2*math.pi*muv[q]*tau
                gw[q]+=wt*Wij*E*math.cos(ph); gm[q]+=wt*Wij*tau*E*math.sin(ph); gs[q]+=wt*Wij*tau*tau*E*math.cos(ph)
    h=0.5/n
    g=np.zeros(spec.P)
    g[0]=alpha[:n].sum()/n
    g[1:1+Q]=h*gw; g[1+Q:1+2*Q]=h*(-2*math.pi*wv)*gm; g[1+2*Q:1+3*Q]=h*(-4*math.pi**2*sgv*wv)*gs
    g[-1]=h*trW
    from oracle.sm_gp import constraint_jacobian
    g=g*constraint_jacobian(raw,kinds,lb,ub).numpy()
    print('grad rel err',np.abs(g-g_o.numpy()).max()/np.abs(g_o.numpy()).max())
run(200,4); run(130,2)
