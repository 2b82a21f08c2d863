import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
import __graft_entry__ as ge
pkg = ge.load_package(); orc = ge.load_oracle(); abi = pkg.abi
eng = pkg.Engine(0); OL = orc.lib(abi)
OL.orc_debug_primitives.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_size_t]
rng = np.random.default_rng(1)
n = 20000
def gen(kind):
    x = rng.standard_normal((n, 8)) * np.exp(rng.uniform(-20, 20, (n, 8)))
    if kind == 'ival':   # ordered pairs
        for a in (0, 2, 4, 6):
            lo = np.minimum(x[:, a], x[:, a+1]); hi = np.maximum(x[:, a], x[:, a+1]); x[:, a] = lo; x[:, a+1] = hi
        x[:, 2] = np.abs(x[:, 2]) if False else x[:, 2]
    if kind == 'mt':
        for a in (0, 2, 4, 6):
            x[:, a+1] = np.abs(x[:, a]) * (1 + np.abs(rng.standard_normal(n)) * 1e-10)
    if kind == 'k':
        for a in (0,):
            lo = np.minimum(x[:, a], x[:, a+1]); hi = np.maximum(x[:, a], x[:, a+1]); x[:, a] = lo; x[:, a+1] = hi
        x[:, 2] = rng.integers(1, 21, n)
    if kind == 'horner':
        for a in (0, 4):
            lo = np.minimum(x[:, a], x[:, a+1]); hi = np.maximum(x[:, a], x[:, a+1]); x[:, a] = lo; x[:, a+1] = hi
        t = np.abs(rng.standard_normal((n, 2))) * 0.03; x[:, 2] = t.min(1); x[:, 3] = t.max(1)
    return np.ascontiguousarray(x)
kinds = {0:'r',1:'r',2:'r',3:'r',4:'r',5:'r',6:'r',7:'r',10:'ival',11:'ival',12:'mt',13:'mt',14:'k',15:'horner',16:'ival',17:'ival',18:'k',19:'horner',20:'r'}
for op, kind in kinds.items():
    x = gen(kind)
    og = np.zeros((n, 4)); oc = np.zeros((n, 4))
    rc = eng.L.ccp_debug_primitives(op, x.ctypes.data_as(C.POINTER(C.c_double)), og.ctypes.data_as(C.POINTER(C.c_double)), n)
    OL.orc_debug_primitives(op, x.ctypes.data_as(C.POINTER(C.c_double)), oc.ctypes.data_as(C.POINTER(C.c_double)), n)
    bad = np.argwhere(og != oc)
    print("op", op, "rc", rc, "mismatches", len(bad), bad[:3].tolist())
    if len(bad):
        i = bad[0][0]; print("   in", x[i].tolist()); print("   gpu", og[i].tolist()); print("   cpu", oc[i].tolist())
