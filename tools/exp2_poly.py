import numpy as np
from numpy.polynomial import chebyshev as C, polynomial as P
# fit q(f) such that 2^f ~= 1 + f*q(f) on [-0.5,0.5]; minimise max relative error via iterative reweighted LS (Lawson)
f = np.cos(np.linspace(0, np.pi, 20001))*0.5
f = f[np.abs(f)>1e-9]
for deg in (3,4,5):
    target = (np.exp2(f)-1)/f
    w = np.ones_like(f)
    for it in range(300):
        # weights: error in 2^f relative = f*(q-target)/2^f
        scale = np.abs(f)/np.exp2(f)
        V = np.vander(f, deg, increasing=True)
        sw = np.sqrt(w)*scale
        c,*_ = np.linalg.lstsq(V*sw[:,None], target*sw, rcond=None)
        err = np.abs((V@c-target)*scale)
        w = w*err; w/=w.sum()
    c32 = c.astype(np.float32)
    # evaluate in fp32 Horner
    ff = np.linspace(-0.5,0.5,400001).astype(np.float32)
    p = np.full_like(ff, c32[-1])
    for k in range(deg-2,-1,-1):
        p = (p*ff + c32[k]).astype(np.float32)
    p = (p*ff + np.float32(1)).astype(np.float32)
    ref = np.exp2(ff.astype(np.float64))
    rel = np.abs(p.astype(np.float64)-ref)/ref
    # error relative to (1-a)
    m = np.abs(ff)>1e-4
    rel1 = np.abs(p.astype(np.float64)-ref)[m]/np.abs(ref[m]-1)
    print(deg+0, "total degree", deg, "max rel", rel.max(), "max rel to |a-1|", rel1.max(), [float(x) for x in c32])
